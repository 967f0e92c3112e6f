"""TEST INFRASTRUCTURE ONLY -- the reference's OWN model / solver code (quant_transformer/model/quant_bert.py,
solver/gamma_migration.py, solver/token_wise_clipping.py, imported unmodified from /root/reference or the staged
oracle/_ref) bound to either backend:

    stack = load_stack("reference")   # quant_transformer.quantization = the reference's package  (CPU oracle)
    stack = load_stack("b200")        # quant_transformer.quantization = outlier_suppression_b200.quantization (CUDA)

and the PTQ schedule of solver/ptq_glue_quant.py:228-252 restated as ``run_schedule`` (the solver file itself needs HF
datasets / Trainer / checkpoints, which do not exist offline; the calls it makes into the path are kept one for one).
"""
from __future__ import annotations

import sys
import types

import torch

from . import make_ref, ref_shim


class Cfg(dict):
    """6-line EasyDict stand-in."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def quant_config(a_bit=6, w_bit=6, a_quantizer="LSQPlusFakeQuantize", a_observer="AvgPruneMinMaxObserver",
                 w_observer="MinMaxObserver", delay=True):
    """exp/bert_ptq/twc_fine_gamma/*/config.yaml:1-17 (defaults) or minmax/cola (Fixed + AvgMinMax, delay False)."""
    return Cfg(a_qconfig=Cfg(quantizer=a_quantizer, observer=a_observer, bit=a_bit, symmetric=False, ch_axis=-1),
               w_qconfig=Cfg(quantizer="FixedFakeQuantize", observer=w_observer, bit=w_bit, symmetric=True, ch_axis=0),
               ln=Cfg(delay=delay))


def fp_bert(layers=2, hidden=128, heads=2, inter=512, vocab=100, max_pos=64, num_labels=2, seed=0):
    """Random-init FP BertForSequenceClassification (no checkpoints offline) with the 4.18-era attributes the
    reference model file reads (shim 3).  LayerNorm gammas are randomised so that the gamma
    migration is not a no-op."""
    from transformers import BertConfig, BertForSequenceClassification
    torch.manual_seed(seed)
    cfg = BertConfig(num_hidden_layers=layers, hidden_size=hidden, num_attention_heads=heads, intermediate_size=inter,
                     vocab_size=vocab, max_position_embeddings=max_pos, num_labels=num_labels)
    fp = BertForSequenceClassification(cfg).eval()
    g = torch.Generator().manual_seed(seed + 1)
    for m in fp.modules():
        if isinstance(m, torch.nn.LayerNorm):
            m.weight.data = torch.rand(m.weight.shape, generator=g) * 1.6 + 0.3
            m.bias.data = torch.randn(m.bias.shape, generator=g) * 0.1
            m.weight.data[:2] *= 4.0  # outlier channels (the phenomenon the reference exists for)
    fp.bert.embeddings.position_embedding_type = "absolute"
    fp.bert.encoder.gradient_checkpointing = False
    for layer in fp.bert.encoder.layer:
        layer.attention.self.position_embedding_type = "absolute"
        layer.attention.pruned_heads = set()
    return fp


def load_stack(backend: str):
    """Imports the reference's model + solver modules with ``quant_transformer.quantization`` resolving to
    ``backend`` ("reference" or "b200").  Returns a namespace: quant_bert, util_layernorm, gamma_migration,
    token_wise_clipping, quantization."""
    root = make_ref.root()
    if root is None:
        raise RuntimeError("reference tree not present (run `python -m oracle.make_ref` where /root/reference exists)")
    ref_shim.purge()
    ref_shim._stub_plot_modules()
    ref_shim.compat_transformers()
    if root not in sys.path:
        sys.path.insert(0, root)
    if backend == "reference":
        q = ref_shim.load()
    elif backend == "b200":
        import outlier_suppression_b200
        import quant_transformer  # noqa: F401  (the reference's top-level package: model/ and solver/ come from it)
        q = outlier_suppression_b200.install_as_reference_backend()
    else:
        raise ValueError(backend)
    import importlib
    ns = types.SimpleNamespace(backend=backend, quantization=q)
    ns.quant_bert = importlib.import_module("quant_transformer.model.quant_bert")
    ns.util_layernorm = importlib.import_module("quant_transformer.model.util_layernorm")
    # the solver files import each other as top-level modules (ptq_glue_quant.py:21-24 runs from solver/)
    ns.gamma_migration = importlib.import_module("quant_transformer.solver.gamma_migration")
    ns.token_wise_clipping = importlib.import_module("quant_transformer.solver.token_wise_clipping")
    for mod in (ns.quant_bert, ns.util_layernorm, ns.gamma_migration, ns.token_wise_clipping):
        assert mod.__file__.startswith(root), mod.__file__
    assert ns.quant_bert.Quantizer is q.Quantizer
    # transformers >= 4.3x changed ModuleUtilsMixin.get_extended_attention_mask / get_head_mask; pin the 4.18 formulas
    def _ext_mask(self, attention_mask, input_shape, device=None):
        return (1.0 - attention_mask[:, None, None, :].to(torch.float32)) * -10000.0
    ns.quant_bert.QuantizedBertModel.get_extended_attention_mask = _ext_mask
    ns.quant_bert.QuantizedBertModel.get_head_mask = lambda self, head_mask, n, *a, **k: [None] * n
    return ns


def build_model(ns, fp, qcfg, device):
    model = ns.quant_bert.QuantizedBertForSequenceClassification(fp, qcfg.w_qconfig, qcfg.a_qconfig, qoutput=False,
                                                                 is_remove_padding=True)
    return model.to(device).eval()


def synth_batches(n_batches, batch, seq, vocab, device, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_batches):
        ids = torch.randint(1, vocab, (batch, seq), generator=g)
        lens = torch.randint(max(1, seq // 4), seq + 1, (batch,), generator=g)
        lens[0] = seq
        mask = (torch.arange(seq)[None, :] < lens[:, None]).long()
        out.append({"input_ids": (ids * mask).to(device), "attention_mask": mask.to(device)})
    return out


def run_schedule(ns, model, qcfg, batches, ratio=0.99, on_forward=None):
    """solver/ptq_glue_quant.py:228-252 for a token-wise-clipping config at one fixed ratio, or the plain
    calibration branch (:247-249) for the other observers.  Returns the logits of every batch under quantization."""
    Q = ns.quantization
    model_cfg = Cfg(model_type="bert")

    def forward(batch):
        with torch.no_grad():
            out = model(**batch)
        logits = out[0] if isinstance(out, tuple) else out.logits
        if on_forward is not None:
            on_forward()
        return logits

    if qcfg.ln.delay:
        model = ns.gamma_migration.delay_ln(model, qcfg, model_cfg)                 # :230-232
    Q.enable_calibration_woquantization(model, quantizer_type="weight_fake_quant")  # :234
    forward(batches[0])                                                             # :235
    if "PruneMinMaxObserver" in qcfg.a_qconfig.observer:
        Q.disable_all(model)                                                        # :237
        from_state = getattr(Q, "set_observer_name", None) or Q.state.set_observer_name
        from_state(model)                                                           # :238
        ns.token_wise_clipping.set_ratio(model, ratio)                              # token_wise_clipping.py:64-65
        for b in batches:
            forward(b)
    else:
        Q.enable_calibration_woquantization(model, quantizer_type="act_fake_quant") # :248
        for b in batches:
            forward(b)                                                              # :249
    Q.enable_quantization(model)                                                    # :253
    return model, [forward(b) for b in batches]
