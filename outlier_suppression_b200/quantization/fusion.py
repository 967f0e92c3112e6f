"""Model-level fusion hooks: the NEXT activation quantizer -- and the GELU in front of it -- fused into
the epilogue of the producing QLinear, behind unchanged module calls.

The reference wires the feed-forward block as three module calls (model/quant_bert.py:277-280, quant_roberta.py likewise):

    hidden_states = self.dense(hidden_states)                                            # QLinear
    hidden_states = self.intermediate_act_fn(hidden_states)                              # GELU
    hidden_states = self.intermediate_act_fn_post_act_fake_quantize(hidden_states, observation_mask, 1)

A Linear cannot know what will consume its output, so the fusion is attached to the ONE module that owns all three steps:
``fuse_ffn_activation(model)`` finds every module with exactly that attribute triple (duck typing: no import of the
reference's classes) and wraps its ``forward``.  In the quantized inference state (weight and activation fake-quant on,
observers off, no autograd) the wrapper runs the block as TWO launches instead of three: the fused Linear, then ONE
elementwise pass that applies GELU and the quantizer and writes the quantizer's uint8 bins (osq_act_fq_per_tensor_bins_f32:
9 B / element instead of 8 + 9), so the following QLinear (``output.dense``) runs bins-in.  The returned tensor carries the
quantizer's tag exactly as if the three modules had been called.  In every other state the original forward runs unchanged
(calibration, learn_scale, CPU).

``OSQ_EPILOGUE_STAGE=1`` selects the one-launch form instead (GELU and the quantizer inside the Linear's epilogue,
osq_fused_fq_linear's output stage): bit-identical, but measured 3-4x SLOWER at BERT-base size (768 -> 3072, M = 16384:
366 us against 55 + 62 + 122 us for the three separate launches) -- erf and the exact division run on the 8 epilogue
warps of a 1-CTA-per-SM kernel, far from the ALU rate a full-occupancy elementwise kernel reaches.

The state togglers call this (like ``group_sibling_linears``), so an unmodified reference driver gets it for free;
``OSQ_DISABLE_EPILOGUE_FUSION=1`` turns it off.

Second hook, same mechanism: ``fuse_layernorm_output(model)`` wraps every ``dense -> dropout -> before_LayerNorm_residual ->
LayerNorm`` module (model/quant_bert.py:197-217 QuantizedBertSelfOutput, :283-303 QuantizedBertOutput; quant_roberta.py
likewise).  In the quantized inference state the residual (optionally times gamma, util_layernorm.py:41-52), the LayerNorm
(util_layernorm.py:14-15, or the split form :34-36 after gamma migration) and the LayerNorm's output quantizer (:16-17) run
as ONE kernel (osq_residual_layernorm_fq_f32, 13 B / element instead of 12 + 8 + 9) that also writes the quantizer's uint8
bins for the following QLinear(s).  ``OSQ_DISABLE_LN_FUSION=1`` turns it off.
"""
from __future__ import annotations

import os
import types

import torch

from .. import ops
from .fake_quant import LSQPlusFakeQuantize, QuantizeBase, _qparam_stamp  # noqa: F401
from .quantized_module import QLinear, _take_bins, stats

_ACT_ATTR = "intermediate_act_fn"
_Q_ATTR = "intermediate_act_fn_post_act_fake_quantize"


def _is_erf_gelu(fn) -> bool:
    """transformers' ACT2FN["gelu"] (GELUActivation -> nn.functional.gelu) or torch's own gelu, exact (erf) form only."""
    if fn is torch.nn.functional.gelu:
        return True
    if isinstance(fn, torch.nn.GELU):
        return getattr(fn, "approximate", "none") == "none"
    name = type(fn).__name__
    if name == "GELUActivation":
        return getattr(fn, "act", None) in (torch.nn.functional.gelu, None) or getattr(fn, "act").__name__ == "gelu"
    return False


def _fusable(mod, x):
    if os.environ.get("OSQ_DISABLE_EPILOGUE_FUSION") == "1" or torch.is_grad_enabled() and x.requires_grad:
        return None
    dense, q = mod.dense, getattr(mod, _Q_ATTR, None)
    if not isinstance(dense, QLinear) or not isinstance(q, QuantizeBase) or not getattr(mod, "qoutput", True):
        return None
    if q.fake_quant_enabled != 1 or q.observer_enabled != 0 or q.ch_axis != -1 or q.quant_max - q.quant_min > 255:
        return None
    if torch.is_grad_enabled() and (q.scale.requires_grad or dense.weight.requires_grad):
        return None
    if not _is_erf_gelu(getattr(mod, _ACT_ATTR, None)):
        return None
    aq = dense._fusable_producer(x)
    if aq is None or not x.is_contiguous():
        return None
    return aq, q


def _fused_forward(self, hidden_states, observation_mask=None):
    pair = _fusable(self, hidden_states)
    if pair is None:
        return self._osq_unfused_forward(hidden_states, observation_mask=observation_mask)
    aq, q = pair
    dense = self.dense
    if os.environ.get("OSQ_EPILOGUE_STAGE") != "1":
        y = dense(hidden_states)                                   # fused fake-quant + Linear (tcgen05), plain fp32 output
        n_out = y.numel()
        g_out = (1.0 / (n_out * q.quant_max) ** 0.5 if q.use_grad_scaling else 1.0) if isinstance(q, LSQPlusFakeQuantize) else 0.0
        want_bins = q._emit_bins and y.shape[-1] % 128 == 0
        r = ops.fq_per_tensor(y, q.scale.detach(), q.zero_point.detach(), q.quant_min, q.quant_max, lsq_grad_factor=g_out,
                              want_bins=want_bins, act="gelu")
        stats["epilogue_fused"] = stats.get("epilogue_fused", 0) + 1
        out = r[0] if want_bins else r
        q._tag(out)
        if want_bins:
            try:
                out._osq_bins = (r[1], out._version)
            except Exception:  # pragma: no cover
                pass
        return out
    codes, rowsum, w_scale = dense._packed_weight()
    g_in = aq.grad_factor(hidden_states) if isinstance(aq, LSQPlusFakeQuantize) else 0.0
    n_out = hidden_states.numel() // hidden_states.shape[-1] * dense.out_features
    g_out = (1.0 / (n_out * q.quant_max) ** 0.5 if q.use_grad_scaling else 1.0) if isinstance(q, LSQPlusFakeQuantize) else 0.0
    y, bins = ops.fused_fq_linear(hidden_states, aq.scale.detach(), aq.zero_point.detach(), aq.quant_min, aq.quant_max, codes,
                                  w_scale, rowsum, dense.bias, lsq_grad_factor=g_in, a_bins=_take_bins(hidden_states, aq),
                                  out_q=dict(scale=q.scale.detach(), zp=q.zero_point.detach(), qmin=q.quant_min, qmax=q.quant_max,
                                             g=g_out, act="gelu", bins=True))
    stats["fused"] += 1
    stats["epilogue_fused"] = stats.get("epilogue_fused", 0) + 1
    q._tag(y)
    try:
        y._osq_bins = (bins, y._version)
    except Exception:  # pragma: no cover
        pass
    return y


def fuse_ffn_activation(model) -> int:
    """Wraps the forward of every ``dense -> intermediate_act_fn -> ..._post_act_fake_quantize`` module.  Idempotent;
    returns the number of wrapped modules."""
    n = 0
    for mod in model.modules():
        if not (hasattr(mod, "dense") and hasattr(mod, _ACT_ATTR) and hasattr(mod, _Q_ATTR)):
            continue
        n += 1
        if getattr(mod, "_osq_unfused_forward", None) is not None:
            continue
        mod._osq_unfused_forward = mod.forward
        mod.forward = types.MethodType(_fused_forward, mod)
    return n


# ---------------------------------------------------------------------------------------------------------------------------
# dense -> dropout -> GammaResidual -> LayerNorm -> quantizer  (quant_bert.py:211-217, :296-303)
# ---------------------------------------------------------------------------------------------------------------------------
def _ln_fusable(mod, h, input_tensor):
    if os.environ.get("OSQ_DISABLE_LN_FUSION") == "1" or torch.is_grad_enabled():
        return None
    if getattr(mod, "backend", "academic") == "tensorrt":      # an extra quantizer sits between the dense and the residual
        return None
    drop = getattr(mod, "dropout", None)
    if drop is not None and getattr(drop, "training", False) and getattr(drop, "p", 0.0) > 0:
        return None
    lnm = mod.LayerNorm
    inner = getattr(lnm, "layernorm", None)
    q = getattr(lnm, "layernorm_post_act_fake_quantize", None)
    if not isinstance(inner, torch.nn.LayerNorm) or not isinstance(q, QuantizeBase) or not getattr(lnm, "qoutput", True):
        return None
    if q.fake_quant_enabled != 1 or q.observer_enabled != 0 or q.ch_axis != -1:
        return None
    H = h.shape[-1]
    if tuple(inner.normalized_shape) != (H,) or H % 4 != 0:
        return None
    if not (h.is_cuda and h.dtype == torch.float32 and h.is_contiguous() and input_tensor.shape == h.shape
            and input_tensor.dtype == torch.float32 and input_tensor.is_contiguous()):
        return None
    return lnm, inner, q


def _fused_ln_forward(self, hidden_states, input_tensor, observation_mask=None):
    # the dense is an ordinary module call (fused fake-quant + Linear when its producer is tagged)
    h = self.dense(hidden_states)
    trio = _ln_fusable(self, h, input_tensor)
    if trio is None:
        # reference order from here on (quant_bert.py:212-217)
        h = self.dropout(h)
        if getattr(self, "backend", "academic") == "tensorrt":
            h = self.output_post_act_fake_quantize(h, observation_mask, 1)
        h = self.before_LayerNorm_residual(input_tensor, h)
        return self.LayerNorm(h, observation_mask)
    lnm, inner, q = trio
    res_mod = self.before_LayerNorm_residual
    gamma = res_mod.gamma.detach() if getattr(res_mod, "mul_gamma", False) else None
    weight = inner.weight.detach() if inner.weight is not None else None
    bias = inner.bias.detach() if inner.bias is not None else None
    split_bias = getattr(lnm, "bias", None)                    # QuantizedSplitLayerNorm: non-affine LayerNorm + beta / gamma
    if split_bias is not None:
        if bias is not None or weight is not None:
            h = self.before_LayerNorm_residual(input_tensor, h)
            return self.LayerNorm(h, observation_mask)
        bias = split_bias.detach()
    g = (1.0 / (h.numel() * q.quant_max) ** 0.5 if q.use_grad_scaling else 1.0) if isinstance(q, LSQPlusFakeQuantize) else 0.0
    want_bins = q._emit_bins and h.shape[-1] % 128 == 0 and q.quant_max - q.quant_min <= 255
    r = ops.residual_layernorm_fq(h, input_tensor, gamma, weight, bias, inner.eps, q.scale.detach(), q.zero_point.detach(),
                                  q.quant_min, q.quant_max, lsq_grad_factor=g, want_bins=want_bins)
    stats["ln_fused"] = stats.get("ln_fused", 0) + 1
    out = r[0] if want_bins else r
    q._tag(out)
    if want_bins:
        try:
            out._osq_bins = (r[1], out._version)
        except Exception:  # pragma: no cover
            pass
    return out


def fuse_layernorm_output(model) -> int:
    """Wraps the forward of every ``dense -> dropout -> before_LayerNorm_residual -> LayerNorm`` module.  Idempotent; returns
    the number of wrapped modules."""
    n = 0
    for mod in model.modules():
        if not (hasattr(mod, "dense") and hasattr(mod, "before_LayerNorm_residual") and hasattr(mod, "LayerNorm")
                and hasattr(getattr(mod, "LayerNorm"), "layernorm")):
            continue
        n += 1
        if getattr(mod, "_osq_unfused_ln_forward", None) is not None:
            continue
        mod._osq_unfused_ln_forward = mod.forward
        mod.forward = types.MethodType(_fused_ln_forward, mod)
    return n
