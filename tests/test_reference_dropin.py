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

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "quant_transformer")), reason="reference tree not present")


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
    import outlier_suppression_b200
    backend = outlier_suppression_b200.install_as_reference_backend()
    _compat_shims()
    sys.modules.pop("quant_transformer.model.quant_bert", None)
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
