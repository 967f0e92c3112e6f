"""TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

CPU (torch fp32) restatement of the reference's fake-quant / observer / QLinear
arithmetic, written as pure functions + small state records instead of nn.Modules.
Every function cites the reference file:line (relative to /root/reference/quant_transformer)
it follows.  Pinned against the reference's own outputs by tests/test_oracle_golden.py.

Arithmetic notes (all verified against the reference in tests/golden):
  * ``x / scale`` is a true IEEE fp32 division, ``round`` is round-half-to-even.
  * op order is  rint(x/s) + zp -> clamp -> (q - zp) * s   (never reassociated).
  * symmetric scale divides by (qmax-qmin)/2 (31.5 for 6 bit), not by qmax.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import torch

F32_EPS_QPARAM = 1e-8  # quantization/observer.py:30


# --------------------------------------------------------------------------------------
# ranges + qparams
# --------------------------------------------------------------------------------------
def quant_range(bit: int, symmetric: bool) -> Tuple[int, int]:
    """quantization/observer.py:31-36."""
    if symmetric:
        return -(1 << (bit - 1)), (1 << (bit - 1)) - 1
    return 0, (1 << bit) - 1


def qparams_from_minmax(min_val: torch.Tensor, max_val: torch.Tensor, qmin: int, qmax: int,
                        symmetric: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """quantization/observer.py:100-119 (ObserverBase.calculate_qparams)."""
    lo = torch.minimum(min_val, torch.zeros_like(min_val))
    hi = torch.maximum(max_val, torch.zeros_like(max_val))
    eps = torch.tensor(F32_EPS_QPARAM, dtype=torch.float32)
    if symmetric:
        hi = torch.maximum(-lo, hi)
        scale = torch.maximum(hi / (float(qmax - qmin) / 2), eps)
        zero_point = torch.zeros(lo.size(), dtype=torch.int32)
    else:
        scale = torch.maximum((hi - lo) / float(qmax - qmin), eps)
        zero_point = torch.clamp(qmin - torch.round(lo / scale), qmin, qmax)
    return scale, zero_point


# --------------------------------------------------------------------------------------
# fake-quant arithmetic
# --------------------------------------------------------------------------------------
def round_ste_value(t: torch.Tensor) -> torch.Tensor:
    """Forward value of quantization/util_quant.py:4-8: (round(t) - t) + t.  Equal to rint(t) for every finite t;
    NaN for t = +-inf (inf - inf), which is how the reference propagates overflowing quotients."""
    return (torch.round(t) - t) + t


def fq_bins(x: torch.Tensor, scale, zero_point, qmin: int, qmax: int) -> torch.Tensor:
    """Clamped bin index (fp32 tensor holding q) -- quantization/util_quant.py:12-13 / :23-24."""
    return torch.clamp(round_ste_value(x / scale) + zero_point, qmin, qmax)


def fq_per_tensor(x: torch.Tensor, scale: float, zero_point, qmin: int, qmax: int) -> torch.Tensor:
    """quantization/util_quant.py:11-15 with Python-scalar scale / zero_point
    (the way FixedFakeQuantize calls it, quantization/fake_quant.py:123-125)."""
    q = fq_bins(x, scale, zero_point, qmin, qmax)
    return (q - zero_point) * scale


def _bcast(v: torch.Tensor, ndim: int, ch_axis: int) -> torch.Tensor:
    shape = [1] * ndim
    shape[ch_axis] = v.numel()
    return v.reshape(shape)


def fq_per_channel(x: torch.Tensor, scale: torch.Tensor, zero_point: torch.Tensor, ch_axis: int,
                   qmin: int, qmax: int) -> torch.Tensor:
    """quantization/util_quant.py:18-26 (tensor / tensor division, zero_point int32)."""
    s = _bcast(scale, x.dim(), ch_axis)
    z = _bcast(zero_point, x.dim(), ch_axis)
    q = fq_bins(x, s, z, qmin, qmax)
    return (q - z) * s


def fq_bins_per_channel(x, scale, zero_point, ch_axis, qmin, qmax):
    return fq_bins(x, _bcast(scale, x.dim(), ch_axis), _bcast(zero_point, x.dim(), ch_axis), qmin, qmax)


def grad_scale_value(t: torch.Tensor, g: float) -> torch.Tensor:
    """Forward value of quantization/util_quant.py:70-71: (t - t*g) + t*g in fp32 (may differ from t by 1 ulp)."""
    return (t - (t * g)) + (t * g)


def lsqplus_effective_qparams(scale: torch.Tensor, zero_point: torch.Tensor, numel_per_scale: int,
                              qmax: int) -> Tuple[torch.Tensor, torch.Tensor, float]:
    """Effective (s', z') used by LSQPlusFakeQuantize's forward.

    quantization/fake_quant.py:193-209 picks grad_factor = 1/sqrt(numel*qmax) (per tensor) or
    1/sqrt(numel/shape[ch]*qmax) (per channel); quantization/util_quant.py:48-51 then applies
    round_ste to the zero point and grad_scale to both parameters.
    """
    g = 1.0 / (numel_per_scale * qmax) ** 0.5
    z = torch.round(zero_point)
    return grad_scale_value(scale, g), grad_scale_value(z, g), g


def fq_lsqplus_per_tensor(x: torch.Tensor, scale: torch.Tensor, zero_point: torch.Tensor,
                          qmin: int, qmax: int) -> torch.Tensor:
    """quantization/util_quant.py:48-55 forward value; scale / zero_point are fp32 tensors of shape [1]."""
    s, z, _ = lsqplus_effective_qparams(scale, zero_point, x.numel(), qmax)
    q = torch.clamp(round_ste_value(x / s) + z, qmin, qmax)
    return (q - z) * s


def fq_lsqplus_per_channel(x, scale, zero_point, ch_axis, qmin, qmax):
    """quantization/util_quant.py:58-67 forward value."""
    s, z, _ = lsqplus_effective_qparams(scale, zero_point, x.numel() // x.shape[ch_axis], qmax)
    s = _bcast(s, x.dim(), ch_axis)
    z = _bcast(z, x.dim(), ch_axis)
    q = torch.clamp(round_ste_value(x / s) + z, qmin, qmax)
    return (q - z) * s


def lsqplus_sanitize(scale: torch.Tensor, zero_point: torch.Tensor, qmin: int, qmax: int):
    """In-place parameter clean-up done on every non-observing forward, quantization/fake_quant.py:188-191."""
    eps = float(torch.finfo(torch.float32).eps)
    scale.abs_().clamp_(min=eps)
    zero_point.clamp_(qmin, qmax)
    return scale, zero_point


# --------------------------------------------------------------------------------------
# token geometry (pad removal) + reductions
# --------------------------------------------------------------------------------------
def token_matrix(x: torch.Tensor, lens: Optional[Sequence[int]], seq_pos: int) -> torch.Tensor:
    """[T, F] matrix of valid tokens, batch-major.

    quantization/observer.py:72-84 (remove_padding) when ``lens`` is given and :86-98
    (reshape_batch_embedding) when it is None: the sequence axis is moved to dim 1, the remaining
    non-batch axes are flattened into F, rows ``s < lens[b]`` are kept.  ``zip`` semantics: only the
    first ``min(len(lens), x.shape[0])`` batch entries are visited (BART 3-D probs quirk).
    """
    rest = [d for d in range(x.dim()) if d != seq_pos]
    if len(rest) == 3:
        y = x.permute(rest[0], seq_pos, rest[1], rest[2]).reshape(x.shape[rest[0]], x.shape[seq_pos], -1)
    elif len(rest) == 2:
        y = x.permute(rest[0], seq_pos, rest[1])
    else:
        raise ValueError("token_matrix expects a 3-D or 4-D activation")
    if lens is None:
        return y.reshape(-1, y.shape[-1])
    rows = [y[b, : int(n)] for b, n in zip(range(y.shape[0]), lens)]
    if not rows:
        return y.new_zeros((0, y.shape[-1]))
    return torch.cat(rows, 0)


def global_minmax(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """torch._aminmax(x) as used at quantization/observer.py:139,193,227."""
    return torch.aminmax(x)


def token_minmax(tokens: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-token extrema over the feature axis, quantization/observer.py:64-65."""
    return tokens.min(1).values, tokens.max(1).values


def prune_bounds(tmin: torch.Tensor, tmax: torch.Tensor, percentile: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """(lower, upper) clipping bounds of quantization/observer.py:50-69.

    upper = max{tmax[t] : tmax[t] <=  quantile(|tmax|, p)},
    lower = min{tmin[t] : tmin[t] >= -quantile(|tmin|, p)}.
    The reference then clips the activation to [lower, upper] and takes its global min/max,
    which is exactly (lower, upper) (re-checked in tests/golden).
    """
    up_thr = torch.quantile(tmax.abs(), percentile)
    lo_thr = -torch.quantile(tmin.abs(), percentile)
    upper = tmax[tmax <= up_thr].max()
    lower = tmin[tmin >= lo_thr].min()
    return lower, upper


def prune_minmax(tokens: torch.Tensor, percentile: float, name: str = "") -> Tuple[torch.Tensor, torch.Tensor]:
    """Global (min, max) after token pruning, quantization/observer.py:61-70 + :227."""
    if "attention_probs" in name:
        return global_minmax(tokens)
    tmin, tmax = token_minmax(tokens)
    lower, upper = prune_bounds(tmin, tmax, percentile)
    clipped = torch.clip(tokens, max=upper, min=lower)
    return global_minmax(clipped)


# --------------------------------------------------------------------------------------
# observer state machines
# --------------------------------------------------------------------------------------
@dataclass
class ObserverState:
    """min_val / max_val buffers (quantization/observer.py:38-39) + the un-checkpointed cnt."""
    min_val: torch.Tensor = field(default_factory=lambda: torch.tensor(float("inf")))
    max_val: torch.Tensor = field(default_factory=lambda: torch.tensor(float("-inf")))
    cnt: int = 0
    one_side_dist: Optional[str] = None


def running_average(st: ObserverState, cur_min: torch.Tensor, cur_max: torch.Tensor) -> None:
    """m <- (m*cnt + cur)/(cnt+1) in fp32, quantization/observer.py:194-202."""
    if st.max_val.numel() <= 1 and bool(st.max_val.isinf()):
        st.min_val, st.max_val = cur_min, cur_max
    else:
        st.min_val = st.min_val * st.cnt + cur_min
        st.max_val = st.max_val * st.cnt + cur_max
    st.cnt += 1
    st.min_val = st.min_val / st.cnt
    st.max_val = st.max_val / st.cnt


def running_extrema(st: ObserverState, cur_min: torch.Tensor, cur_max: torch.Tensor) -> None:
    """quantization/observer.py:143-144 (MinMaxObserver) and :535-536 (MSEFastObserver)."""
    st.min_val = torch.minimum(st.min_val, cur_min)
    st.max_val = torch.maximum(st.max_val, cur_max)


def _prep(x: torch.Tensor, lens, seq_pos: int) -> torch.Tensor:
    x = x.detach().to(torch.float32)
    if lens is not None:
        x = token_matrix(x, lens, seq_pos)
    return x


def observe_minmax(st: ObserverState, x: torch.Tensor, ch_axis: int = -1, lens=None, seq_pos: int = -1) -> None:
    """MinMaxObserver.forward, quantization/observer.py:130-145."""
    if x.numel() == 0:
        return
    x = _prep(x, lens, seq_pos)
    if ch_axis == -1:
        mn, mx = global_minmax(x)
    else:
        y = x.transpose(0, ch_axis).flatten(1) if ch_axis != 0 else x.flatten(1)
        mn, mx = torch.aminmax(y, dim=1)
    running_extrema(st, mn, mx)


def observe_avg_minmax(st: ObserverState, x: torch.Tensor, lens=None, seq_pos: int = -1) -> None:
    """AvgMinMaxObserver.forward, quantization/observer.py:184-203."""
    if x.numel() == 0:
        return
    mn, mx = global_minmax(_prep(x, lens, seq_pos))
    running_average(st, mn, mx)


def observe_avg_prune_minmax(st: ObserverState, x: torch.Tensor, percentile: float, name: str = "",
                             lens=None, seq_pos: int = -1) -> None:
    """AvgPruneMinMaxObserver.forward, quantization/observer.py:214-237."""
    if x.numel() == 0:
        return
    x = x.detach().to(torch.float32)
    if lens is not None:
        mn, mx = prune_minmax(token_matrix(x, lens, seq_pos), percentile, name)
    elif seq_pos != -1:
        mn, mx = prune_minmax(token_matrix(x, None, seq_pos), percentile, name)
    else:
        mn, mx = global_minmax(x)
    running_average(st, mn, mx)


# --------------------------------------------------------------------------------------
# MSEFast (SciPy bounded Brent driven search)
# --------------------------------------------------------------------------------------
def _as_tensor(v) -> torch.Tensor:
    # torch.tensor(np.float64) is an fp64 tensor, torch.tensor(python float) an fp32 one: the reference's
    # qparams inherit that dtype (quantization/observer.py:425-427), so keep the distinction.
    return v.detach().clone() if isinstance(v, torch.Tensor) else torch.tensor(v)


def mse_loss(x: torch.Tensor, new_min, new_max, qmin: int, qmax: int, symmetric: bool) -> torch.Tensor:
    """quantization/observer.py:420-432: fq with qparams from (new_min,new_max), mean squared error."""
    scale, zp = qparams_from_minmax(_as_tensor(new_min), _as_tensor(new_max), qmin, qmax, symmetric)
    xq = fq_per_tensor(x, scale.item(), int(zp.item()), qmin, qmax)
    return (xq - x).abs().pow(2.0).mean()


def _brent(fn, lo, hi):
    from scipy.optimize import minimize_scalar  # third-party: scipy 1.18.1 (_minimize_scalar_bounded)
    return minimize_scalar(fn, bounds=(lo, hi), method="Bounded")


def mse_search_1d(x: torch.Tensor, qmin: int, qmax: int, symmetric: bool, one_side: str,
                  counter: Optional[list] = None) -> Tuple[float, float]:
    """quantization/observer.py:483-494 (+ :453-456) for one vector / tensor."""
    x_min, x_max = global_minmax(x)
    xrange = torch.max(x_min.abs(), x_max).item()

    def loss(r):
        if counter is not None:
            counter[0] += 1
        lo = 0.0 if one_side == "pos" else -r
        hi = 0.0 if one_side == "neg" else r
        return mse_loss(x, lo, hi, qmin, qmax, symmetric).numpy()

    res = _brent(loss, min(0.1, 0.01 * xrange), xrange)
    r = res.x
    return (0.0 if one_side == "pos" else -r), (0.0 if one_side == "neg" else r)


def mse_search_2d(x: torch.Tensor, qmin: int, qmax: int, symmetric: bool,
                  counter: Optional[list] = None):
    """quantization/observer.py:434-481: outer Brent over the range, inner Brent over the shift."""
    x_min_t, x_max_t = global_minmax(x)
    x_min, x_max = x_min_t, x_max_t  # kept as tensors: python max()/min() mix like the reference
    span = float(qmax - qmin)

    def shift_loss(shift, xrange):
        if counter is not None:
            counter[0] += 1
        new_min = max(0.0 - shift, x_min)
        new_max = min(xrange - shift, x_max)
        return mse_loss(x, new_min, new_max, qmin, qmax, symmetric).numpy()

    def range_loss(xrange):
        delta = xrange / span
        return _brent(lambda s: shift_loss(s, xrange), delta * qmin, delta * qmax).fun

    total = (x_max - x_min).item()
    final_range = _brent(range_loss, min(0.1, 0.01 * total), total).x
    delta = final_range / span
    final_shift = _brent(lambda s: shift_loss(s, final_range), delta * qmin, delta * qmax).x
    best_min = max(0.0 - final_shift, x_min)
    best_max = min(final_range - final_shift, x_max)
    return best_min, best_max


def decide_one_side(x: torch.Tensor) -> str:
    """quantization/observer.py:528-529."""
    return "pos" if x.min() >= 0.0 else "neg" if x.max() <= 0.0 else "no"


def mse_fast_minmax(st: ObserverState, x: torch.Tensor, qmin: int, qmax: int, symmetric: bool,
                    ch_axis: int = -1, counter: Optional[list] = None):
    """best (min,max) of MSEFastObserver for one batch, quantization/observer.py:496-533."""
    if st.one_side_dist is None:
        st.one_side_dist = decide_one_side(x)
    one_d = st.one_side_dist != "no" or symmetric
    if ch_axis == -1:
        lo, hi = (mse_search_1d(x, qmin, qmax, symmetric, st.one_side_dist, counter) if one_d
                  else mse_search_2d(x, qmin, qmax, symmetric, counter))
        return _as_tensor(lo), _as_tensor(hi)
    y = x.transpose(0, ch_axis).flatten(1) if ch_axis != 0 else x.flatten(1)
    mins, maxs = torch.aminmax(y, dim=1)
    mins, maxs = mins.clone(), maxs.clone()
    for ch in range(y.shape[0]):
        lo, hi = (mse_search_1d(y[ch], qmin, qmax, symmetric, st.one_side_dist, counter) if one_d
                  else mse_search_2d(y[ch], qmin, qmax, symmetric, counter))
        mins[ch], maxs[ch] = lo, hi
    return mins, maxs


def observe_mse_fast(st: ObserverState, x, qmin, qmax, symmetric, ch_axis=-1, lens=None, seq_pos=-1, counter=None):
    """MSEFastObserver.forward, quantization/observer.py:520-536 (running min/max of per-batch optima)."""
    if x.numel() == 0:
        return
    x = _prep(x, lens, seq_pos)
    lo, hi = mse_fast_minmax(st, x, qmin, qmax, symmetric, ch_axis, counter)
    running_extrema(st, lo, hi)


def observe_avg_mse_fast(st: ObserverState, x, qmin, qmax, symmetric, lens=None, seq_pos=-1, counter=None):
    """AvgMSEFastObserver.forward, quantization/observer.py:545-567."""
    if x.numel() == 0:
        return
    x = _prep(x, lens, seq_pos)
    lo, hi = mse_fast_minmax(st, x, qmin, qmax, symmetric, -1, counter)
    running_average(st, lo, hi)


# --------------------------------------------------------------------------------------
# QLinear and the fused reference
# --------------------------------------------------------------------------------------
def weight_qparams_minmax(weight: torch.Tensor, bit: int, symmetric: bool = True):
    """MinMaxObserver(ch_axis=0) + calculate_qparams on a [N, K] weight (every shipped config)."""
    qmin, qmax = quant_range(bit, symmetric)
    st = ObserverState()
    observe_minmax(st, weight, ch_axis=0)
    scale, zp = qparams_from_minmax(st.min_val, st.max_val, qmin, qmax, symmetric)
    return scale, zp.to(torch.int32), qmin, qmax


def qlinear(x_fq: torch.Tensor, weight: torch.Tensor, w_scale: torch.Tensor, w_zp: torch.Tensor,
            w_qmin: int, w_qmax: int, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """QLinear.forward, quantization/quantized_module.py:71-72: F.linear on the fake-quantized weight."""
    w_fq = fq_per_channel(weight, w_scale, w_zp, 0, w_qmin, w_qmax)
    return torch.nn.functional.linear(x_fq, w_fq, bias)


def fused_fq_linear(a: torch.Tensor, a_scale: torch.Tensor, a_zp: torch.Tensor, a_qmin: int, a_qmax: int,
                    lsqplus: bool, weight: torch.Tensor, w_scale: torch.Tensor, w_zp: torch.Tensor,
                    w_qmin: int, w_qmax: int, bias: Optional[torch.Tensor]):
    """activation quantizer followed by QLinear exactly as the model chains them
    (model/quant_bert.py:216 -> :142 ; quantization/fake_quant.py:107-126,178-209).
    Returns (Y, activation bins, weight bins); bins are integer-valued fp32 (nearest integer of the
    reference's q, which is non-integer only through LSQ+'s 1-ulp z')."""
    if lsqplus:
        s, z, _ = lsqplus_effective_qparams(a_scale, a_zp, a.numel(), a_qmax)
        qa = torch.clamp(round_ste_value(a / s) + z, a_qmin, a_qmax)
        a_fq = (qa - z) * s
    else:
        qa = fq_bins(a, a_scale.item(), a_zp.item(), a_qmin, a_qmax)
        a_fq = (qa - a_zp.item()) * a_scale.item()
    qw = fq_bins_per_channel(weight, w_scale, w_zp, 0, w_qmin, w_qmax)
    y = qlinear(a_fq, weight, w_scale, w_zp, w_qmin, w_qmax, bias)
    return y, torch.round(qa), qw


# --------------------------------------------------------------------------------------
# model-level blocks around the quantizers: restated from model/quant_bert.py and
# model/util_layernorm.py, pinned by tests/golden/blocks.npz (generated from the reference's unmodified modules)
# --------------------------------------------------------------------------------------
def act_fq(x: torch.Tensor, scale: torch.Tensor, zero_point: torch.Tensor, qmin: int, qmax: int, lsqplus: bool) -> torch.Tensor:
    """An activation quantizer in the quantized state: LSQPlusFakeQuantize (fake_quant.py:188-195) or FixedFakeQuantize
    (fake_quant.py:123-125); scale / zero_point are one-element tensors."""
    if lsqplus:
        return fq_lsqplus_per_tensor(x, scale.clone(), zero_point.clone(), qmin, qmax)
    return fq_per_tensor(x, float(scale), int(zero_point), qmin, qmax)


def attention_block(q3: torch.Tensor, k3: torch.Tensor, v3: torch.Tensor, mask: Optional[torch.Tensor], heads: int,
                    qq, kq, pq, vq, oq, qmin: int, qmax: int, lsqplus: bool, dtype=torch.float32):
    """model/quant_bert.py:141-193 (QuantizedBertSelfAttention.forward) after the three projections, dropout inactive:
    q3 / k3 / v3 are the [B, S, H] outputs of query / key / value; qq .. oq = (scale, zero_point) of the query_permute,
    key_transpose, attention_probs, value_permute and context_view quantizers (oq None: qoutput=False).
    ``dtype``: arithmetic type of the two matmuls and the softmax (fp32 = the reference; fp64 = the exact value of the
    contraction of the fake-quantised operands).  Returns (scores, probs, context)."""
    B, S, H = q3.shape
    d = H // heads
    tfs = lambda x: x.view(B, -1, heads, d).permute(0, 2, 1, 3)                                  # :128-132
    ql = act_fq(tfs(q3), *qq, qmin, qmax, lsqplus)                                               # :148
    kt = act_fq(tfs(k3).transpose(-1, -2), *kq, qmin, qmax, lsqplus)                             # :150
    scores = torch.matmul(ql.to(dtype), kt.to(dtype)) / (d ** 0.5)                               # :150, :169
    if mask is not None:
        scores = scores + mask.to(dtype)                                                         # :172
    probs = torch.nn.functional.softmax(scores, dim=-1)                                          # :175
    pf = act_fq(probs.float(), *pq, qmin, qmax, lsqplus)                                         # :185
    vl = act_fq(tfs(v3), *vq, qmin, qmax, lsqplus)                                               # :186
    ctx = torch.matmul(pf.to(dtype), vl.to(dtype))                                               # :187
    ctx = ctx.permute(0, 2, 1, 3).contiguous().view(B, S, H)                                     # :189-191
    if oq is not None:
        ctx = act_fq(ctx.float(), *oq, qmin, qmax, lsqplus)                                      # :192-193
    return scores, probs, ctx


def residual_layernorm_fq(h: torch.Tensor, res: torch.Tensor, gamma: Optional[torch.Tensor], ln_weight: Optional[torch.Tensor],
                          ln_bias: Optional[torch.Tensor], eps: float, q, qmin: int, qmax: int, lsqplus: bool, dtype=torch.float32):
    """model/quant_bert.py:214-216 behind the dense: GammaResidual (util_layernorm.py:49-52), then QuantizedLayerNorm
    (util_layernorm.py:14-17, affine) or QuantizedSplitLayerNorm (util_layernorm.py:34-37: non-affine LayerNorm, then
    ``+= bias``), then the LayerNorm's output quantizer q = (scale, zero_point).  Returns (LayerNorm output, its fake-quant)."""
    u = (res * gamma if gamma is not None else res) + h
    u = u.to(dtype)
    ln = torch.nn.functional.layer_norm(u, (u.shape[-1],), None if ln_weight is None else ln_weight.to(dtype),
                                        None if ln_weight is None or ln_bias is None else ln_bias.to(dtype), eps)
    if ln_weight is None and ln_bias is not None:
        ln = ln + ln_bias.to(dtype)
    return ln, act_fq(ln.float(), *q, qmin, qmax, lsqplus)


# --------------------------------------------------------------------------------------
# synthetic inputs -- shared by tests and bench so CPU and GPU see identical bits
# --------------------------------------------------------------------------------------
def synth_activation(b: int, s: int, h: int, seed: int = 0, outlier_channels: int = 6, outlier_gain: float = 30.0):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(b, s, h, generator=g, dtype=torch.float32)
    idx = torch.randperm(h, generator=g)[:outlier_channels]
    a[..., idx] *= outlier_gain
    lens = torch.randint(max(1, s // 4), s + 1, (b,), generator=g)
    lens[0] = s
    return a, lens


def synth_linear(n: int, k: int, seed: int = 1, gamma: bool = False):
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(n, k, generator=g, dtype=torch.float32) * 0.05
    bias = torch.randn(n, generator=g, dtype=torch.float32) * 0.02
    if gamma:
        w = w * (torch.rand(k, generator=g) * 2.0 + 0.2)[None, :]
    return w, bias
