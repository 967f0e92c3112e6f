"""Drop-in check against the reference's OWN model code (only where /root/reference exists, i.e. in the build
container): quant_transformer/model/quant_bert.py is imported unmodified on top of
outlier_suppression_b200.install_as_reference_backend() and must construct, expose the reference's quantizer
census, obey the state togglers and run an FP forward.  (GPU-side numerics of the same modules are covered by
tests/test_gpu_*.py; the reference tree does not travel to the GPU box.)"""
import os
import sys
import types

import pytest
import torch

from oracle import make_ref

REF = make_ref.root() or "/root/reference"
pytestmark = pytest.mark.skipif(make_ref.root() is None, reason="reference tree not present (neither /root/reference nor oracle/_ref)")


class QC:
    def __init__(self, quantizer, observer, bit, symmetric, ch_axis):
        self.quantizer, self.observer, self.bit, self.symmetric, self.ch_axis = quantizer, observer, bit, symmetric, ch_axis


def _compat_shims():
    """transformers 4.18 symbols the reference model files import (shim 3)."""
    import transformers
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    for name in ("apply_chunking_to_forward", "prune_linear_layer", "find_pruneable_heads_and_indices"):
        if not hasattr(mu, name):
            setattr(mu, name, getattr(pu, name, lambda *a, **k: (set(), None)))
    if "transformers.generation_utils" not in sys.modules:
        g = types.ModuleType("transformers.generation_utils")
        g.GenerationMixin = transformers.generation.GenerationMixin
        sys.modules["transformers.generation_utils"] = g


def _fp_bert():
    from transformers import BertConfig, BertForSequenceClassification
    cfg = BertConfig(num_hidden_layers=2, hidden_size=128, num_attention_heads=2, intermediate_size=512, vocab_size=100,
                     max_position_embeddings=64)
    fp = BertForSequenceClassification(cfg).eval()
    fp.bert.embeddings.position_embedding_type = "absolute"
    fp.bert.encoder.gradient_checkpointing = False
    for layer in fp.bert.encoder.layer:
        layer.attention.self.position_embedding_type = "absolute"
        layer.attention.pruned_heads = set()
    return fp


def test_reference_quant_bert_runs_on_this_backend():
    for m in ("seaborn", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from oracle import ref_shim
    ref_shim.purge()
    import outlier_suppression_b200
    backend = outlier_suppression_b200.install_as_reference_backend()
    _compat_shims()
    from quant_transformer.model import quant_bert  # the reference's file, unmodified
    assert quant_bert.Quantizer is backend.Quantizer

    from outlier_suppression_b200.quantization import quantized_module as qm
    from outlier_suppression_b200.quantization.fake_quant import FixedFakeQuantize, LSQPlusFakeQuantize, QuantizeBase
    a_cfg = QC("LSQPlusFakeQuantize", "AvgPruneMinMaxObserver", 6, False, -1)
    w_cfg = QC("FixedFakeQuantize", "MinMaxObserver", 6, True, 0)
    model = quant_bert.QuantizedBertForSequenceClassification(_fp_bert(), w_cfg, a_cfg, qoutput=False)
    quantizers = {n: m for n, m in model.named_modules() if isinstance(m, QuantizeBase)}
    act = [n for n in quantizers if "act_fake_quant" in n]
    wgt = [n for n in quantizers if "weight_fake_quant" in n]
    # census formula of SURVEY.md section 2: 1 + 8L - 1 + 2 activation, 3 + 6L + 2 weight quantizers
    layers = 2
    assert len(act) == 1 + 8 * layers - 1 + 2 and len(wgt) == 3 + 6 * layers + 2
    assert all(isinstance(quantizers[n], LSQPlusFakeQuantize) for n in act)
    assert all(isinstance(quantizers[n], FixedFakeQuantize) for n in wgt)
    assert sum(isinstance(m, qm.QLinear) for m in model.modules()) == 6 * layers + 2

    backend.enable_calibration_woquantization(model, quantizer_type="weight_fake_quant")
    assert all(quantizers[n].observer_enabled == 1 for n in wgt) and all(quantizers[n].observer_enabled == 0 for n in act)
    backend.disable_all(model)
    from outlier_suppression_b200.quantization.state import set_observer_name
    set_observer_name(model)
    assert any("attention_probs" in quantizers[n].observer.name for n in act)
    # the state_dict carries the reference's keys for every quantizer
    sd = model.state_dict()
    for n in quantizers:
        assert n + ".scale" in sd and n + ".zero_point" in sd and n + ".observer.min_val" in sd


def test_staged_reference_tree_is_verbatim():
    """oracle/make_ref.py stages a byte-identical copy of the reference package for the GPU box (git-ignored)."""
    root = make_ref.build()
    assert root is not None and make_ref.staged() and make_ref.verify()
    if make_ref.source_available():
        import json
        assert json.load(open(make_ref.MANIFEST))["files"] == make_ref._tree(make_ref.SRC_ROOT)
    tracked = os.popen("git -C %s ls-files oracle/_ref 2>/dev/null" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))).read()
    assert tracked.strip() == "", "reference sources must never be committed"


def test_fp_forward_of_unmodified_quant_bert_on_this_backend_cpu():
    """With every quantizer disabled the model is plain torch: the reference's model file, bound to this backend's
    Quantizer / QLinear / QEmbedding classes, must reproduce the FP HuggingFace model (no CUDA needed)."""
    from oracle import ref_model as RM
    ns = RM.load_stack("b200")
    qcfg = RM.quant_config()
    fp = RM.fp_bert()
    model = RM.build_model(ns, fp, qcfg, "cpu")
    ns.quantization.disable_all(model)
    batch = RM.synth_batches(1, 4, 32, 100, "cpu")[0]
    with torch.no_grad():
        got = model(**batch)
        want = fp(**batch).logits
    got = got[0] if isinstance(got, tuple) else got.logits
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6), (got, want)


def test_lockstep_harness_self_check_reference_vs_reference():
    """tests/lockstep.py (used by the -m gpu model-level test) replays our calls into the reference's modules; run with
    the reference on BOTH sides it must find zero differences and must have visited every quantizer / operator."""
    from oracle import ref_model as RM
    from tests import lockstep
    r = lockstep.run("reference", "cpu", RM.quant_config())
    assert r["n_act"] == 18 and r["checked"]["q"] == 18 * 7 and r["checked"]["op"] == 17 * 7
    assert r["scale_drift"] == 0.0
    for a, b in zip(r["logits"], r["ref_logits"]):
        assert torch.equal(a, b)
