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

Third hook: ``fuse_self_attention(model)`` wraps every self-attention module that owns the query / key / value projections
and the five attention-side quantizers (model/quant_bert.py:100-195, quant_roberta.py likewise).  In the quantized
inference state its forward becomes: grouped q | k | v launch -> osq_attn_scores_fq_f32 (query and key quantizers as the
prologue of q @ k^T, 1 / sqrt(d) and the attention mask in its epilogue) -> softmax -> osq_attn_context_fq_f32 (probability and
value quantizers as the prologue of probs @ v, the permute + view and the context quantizer + bins in its epilogue):
four fake-quant launches, two scale / mask passes and a permute copy less per layer, and the exact integer contraction
instead of an fp32 GEMM over dequantised values.  ``OSQ_DISABLE_ATTN_FUSION=1`` turns it off.
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
        if (want_bins and q._lazy_ok and y.is_contiguous() and y.data_ptr() % 16 == 0 and os.environ.get("OSQ_DISABLE_LAZY_FQ") != "1"):
            # the block's output normally feeds output.dense alone: bins only (5 B / element instead of 9), fp32 values on demand
            stats["epilogue_fused"] = stats.get("epilogue_fused", 0) + 1
            return q._fq_deferred(y, q.scale.detach(), q.zero_point.detach(), g_out, act="gelu")
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
def _ln_fusable(mod, hidden_states, input_tensor):
    """Decided BEFORE anything runs (from shapes and module state), so that the unfused case is the module's own forward."""
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
    H = getattr(mod.dense, "out_features", None)
    if H is None or tuple(inner.normalized_shape) != (H,) or H % 4 != 0:
        return None
    split_bias = getattr(lnm, "bias", None)                    # QuantizedSplitLayerNorm: non-affine LayerNorm + beta / gamma
    if split_bias is not None and (inner.weight is not None or inner.bias is not None):
        return None
    if not (hidden_states.is_cuda and tuple(input_tensor.shape) == tuple(hidden_states.shape[:-1]) + (H,)
            and input_tensor.dtype == torch.float32 and input_tensor.is_contiguous()):
        return None
    return lnm, inner, q


def _fused_ln_forward(self, hidden_states, input_tensor, observation_mask=None):
    trio = _ln_fusable(self, hidden_states, input_tensor)
    if trio is None:
        return self._osq_unfused_ln_forward(hidden_states, input_tensor, observation_mask=observation_mask)
    lnm, inner, q = trio
    h = self.dense(hidden_states)           # an ordinary module call (fused fake-quant + Linear when its producer is tagged)
    if not (h.dtype == torch.float32 and h.is_contiguous()):
        h = h.float().contiguous()
    res_mod = self.before_LayerNorm_residual
    gamma = res_mod.gamma.detach() if getattr(res_mod, "mul_gamma", False) else None
    weight = inner.weight.detach() if inner.weight is not None else None
    bias = inner.bias.detach() if inner.bias is not None else None
    split_bias = getattr(lnm, "bias", None)
    if split_bias is not None:
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


# ---------------------------------------------------------------------------------------------------------------------------
# self-attention: q @ k^T and probs @ v with their quantizers  (quant_bert.py:134-195)
# ---------------------------------------------------------------------------------------------------------------------------
_ATTN_Q = ("query_permute_post_act_fake_quantize", "key_transpose_post_act_fake_quantize", "value_permute_post_act_fake_quantize",
           "attention_probs_post_act_fake_quantize")


def _q_ready(q) -> bool:
    return (isinstance(q, QuantizeBase) and q.fake_quant_enabled == 1 and q.observer_enabled == 0 and q.ch_axis == -1
            and q.quant_max - q.quant_min <= 255)


def _q_args(q, numel):
    g = (1.0 / (numel * q.quant_max) ** 0.5 if q.use_grad_scaling else 1.0) if isinstance(q, LSQPlusFakeQuantize) else 0.0
    return dict(scale=q.scale.detach(), zp=q.zero_point.detach(), qmin=q.quant_min, qmax=q.quant_max, g=g)


def _attn_fusable(mod, hidden_states, attention_mask, head_mask, output_attentions):
    if os.environ.get("OSQ_DISABLE_ATTN_FUSION") == "1" or torch.is_grad_enabled():
        return False
    if head_mask is not None or output_attentions or getattr(mod, "position_embedding_type", "absolute") != "absolute":
        return False
    drop = getattr(mod, "dropout", None)
    if drop is not None and getattr(drop, "training", False) and getattr(drop, "p", 0.0) > 0:
        return False
    if not all(_q_ready(getattr(mod, n, None)) for n in _ATTN_Q):
        return False
    if getattr(mod, "qoutput", True) and not _q_ready(getattr(mod, "context_view_post_act_fake_quantize", None)):
        return False
    if not (hidden_states.is_cuda and hidden_states.dtype == torch.float32 and hidden_states.dim() == 3):
        return False
    B, S, _ = hidden_states.shape
    if mod.attention_head_size not in (32, 64, 128) or S % 4 != 0:
        return False
    if attention_mask is not None and not (attention_mask.dtype == torch.float32 and attention_mask.numel() == B * S
                                           and attention_mask.shape[-1] == S and attention_mask.shape[0] == B):
        return False
    return True


def _fused_attn_forward(self, hidden_states, attention_mask=None, head_mask=None, output_attentions=False, observation_mask=None):
    if not _attn_fusable(self, hidden_states, attention_mask, head_mask, output_attentions):
        return self._osq_unfused_attn_forward(hidden_states, attention_mask, head_mask, output_attentions, observation_mask)
    q3 = self.query(hidden_states)          # one grouped launch for the three projections (QLinearGroup)
    k3 = self.key(hidden_states)
    v3 = self.value(hidden_states)
    # (a grouped launch returns the three projections as column slices of one [.., 3 * hidden] tensor: token stride 3 * hidden, which
    #  the kernels take as is -- only the channels must be contiguous)
    if not (q3.stride(-1) == 1 and k3.stride(-1) == 1 and v3.stride(-1) == 1):
        return self._osq_unfused_attn_forward(hidden_states, attention_mask, head_mask, output_attentions, observation_mask)
    qh, kh, vh = self.transpose_for_scores(q3), self.transpose_for_scores(k3), self.transpose_for_scores(v3)
    d = self.attention_head_size
    # `attention_scores / math.sqrt(d)` (quant_bert.py:169): ATen multiplies by the fp32 reciprocal of the scalar
    inv = float(torch.tensor(1.0, dtype=torch.float32) / torch.tensor(d ** 0.5, dtype=torch.float32))
    mask = attention_mask.contiguous() if attention_mask is not None else None
    scores = ops.attn_scores_fq(qh, kh, _q_args(self.query_permute_post_act_fake_quantize, q3.numel()),
                                _q_args(self.key_transpose_post_act_fake_quantize, k3.numel()), out_mul=inv, mask=mask)
    probs = torch.nn.functional.softmax(scores, dim=-1)
    del scores
    pq = self.attention_probs_post_act_fake_quantize
    vq = self.value_permute_post_act_fake_quantize
    oq = self.context_view_post_act_fake_quantize if getattr(self, "qoutput", True) else None
    stats["attn_fused"] = stats.get("attn_fused", 0) + 1
    if oq is None:
        return (ops.attn_context_fq(probs, vh, _q_args(pq, probs.numel()), _q_args(vq, v3.numel())),)
    want_bins = oq._emit_bins and q3.shape[-1] % 128 == 0
    r = ops.attn_context_fq(probs, vh, _q_args(pq, probs.numel()), _q_args(vq, v3.numel()), oq=_q_args(oq, q3.numel()), want_bins=want_bins)
    ctx = r[0] if want_bins else r
    oq._tag(ctx)
    if want_bins:
        try:
            ctx._osq_bins = (r[1], ctx._version)
        except Exception:  # pragma: no cover
            pass
    return (ctx,)


def fuse_self_attention(model) -> int:
    """Wraps the forward of every self-attention module with query / key / value projections and the attention-side quantizers.
    Idempotent; returns the number of wrapped modules."""
    n = 0
    for mod in model.modules():
        if not (all(hasattr(mod, a) for a in ("query", "key", "value", "transpose_for_scores", "attention_head_size")) and
                all(hasattr(mod, a) for a in _ATTN_Q)):
            continue
        n += 1
        if getattr(mod, "_osq_unfused_attn_forward", None) is not None:
            continue
        mod._osq_unfused_attn_forward = mod.forward
        mod.forward = types.MethodType(_fused_attn_forward, mod)
    return n
