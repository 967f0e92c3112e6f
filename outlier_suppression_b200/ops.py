"""Torch-facing wrappers over the C ABI (include/osq.h).

PyTorch is plumbing here: it owns device memory and streams; every arithmetic step of the hot path
runs in libosq_b200.so.  All functions require CUDA tensors and raise otherwise -- there is no CPU
fallback (the CPU restatement lives in oracle/ and is test infrastructure only).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import FusedLinearArgs, ReplayTarget, StatEpilogue, Tokens, check

_workspaces = {}

STAT_NONE, STAT_AVERAGE, STAT_EXTREMA = 0, 1, 2


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("outlier_suppression_b200 runs on CUDA tensors only (got a %s tensor); "
                               "there is no CPU fallback" % t.device)


def plain(t):
    """A deferred fake-quant tensor (quantization.fake_quant.LazyFakeQuant: uint8 bins now, fp32 values on demand) -> its fp32
    tensor; anything else unchanged.  Every entry point that takes an activation's data pointer goes through this."""
    m = getattr(t, "_osq_materialize", None)
    return t if m is None else m()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def workspace(device) -> torch.Tensor:
    """Caller-owned reduction scratch (zeroed ticket counter), one per (device, stream)."""
    key = (torch.device(device).index, _stream())
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.zeros(int(_lib.load().osq_workspace_bytes()), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _is_dense(x: torch.Tensor) -> bool:
    """non-overlapping and dense: some permutation of the dims is contiguous"""
    if x.is_contiguous():
        return True
    expect = 1
    for size, stride in sorted(((sz, st) for sz, st in zip(x.shape, x.stride()) if sz != 1), key=lambda p: p[1]):
        if stride != expect:
            return False
        expect *= size
    return True


def _dense_like(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(x', y') such that both cover the same memory layout densely; elementwise kernels run over
    the flat storage so permuted-but-dense views (q / k^T / v, quant_bert.py:148-186) need no copy."""
    if x.dtype != torch.float32:
        x = x.float()
    if not (_is_dense(x)):
        x = x.contiguous()
    return x, torch.empty_like(x)  # preserve_format keeps the dense strides


# ------------------------------------------------------------------------------------------------
# K1 / K2
# ------------------------------------------------------------------------------------------------
def fq_per_tensor(x: torch.Tensor, scale: torch.Tensor, zero_point: torch.Tensor, qmin: int, qmax: int,
                  lsq_grad_factor: float = 0.0, want_codes: bool = False, want_bins: bool = False, act: Optional[str] = None):
    """util_quant.py:11-15 / :48-55 with device-resident qparams. Returns y (and int16 bins with want_codes, or
    uint8 ``bin - qmin`` in the fused Linear's operand format with want_bins)."""
    x = plain(x)
    _require_cuda(x, scale, zero_point)
    x, y = _dense_like(x)
    if act not in (None, "none"):
        # activation + quantizer as one pass (quant_bert.py:278-280); GELU bit-identical to torch's CUDA kernel
        if act != "gelu" or want_codes:
            raise ValueError("act must be None or 'gelu' (and is incompatible with want_codes)")
        if scale.dtype != torch.float32 or zero_point.dtype not in (torch.float32, torch.int32):
            raise TypeError("scale must be float32, zero_point float32 or int32")
        bins = torch.empty_like(x, dtype=torch.uint8) if want_bins else None
        if x.numel() > 0:
            check(_lib.load().osq_act_fq_per_tensor_bins_f32(x.data_ptr(), y.data_ptr(), _ptr(bins), x.numel(), 1, scale.data_ptr(),
                                                             zero_point.data_ptr(), int(zero_point.dtype == torch.int32),
                                                             float(lsq_grad_factor), int(qmin), int(qmax), _stream()),
                  "osq_act_fq_per_tensor_bins_f32")
        return (y, bins) if want_bins else y
    if want_bins and x.numel() > 0:
        if scale.dtype != torch.float32:
            raise TypeError("scale must be float32")
        bins = torch.empty_like(x, dtype=torch.uint8)  # same (dense) strides as x / y: the kernel runs over the flat storage
        check(_lib.load().osq_fq_per_tensor_bins_f32(x.data_ptr(), y.data_ptr(), bins.data_ptr(), x.numel(), scale.data_ptr(),
                                                     zero_point.data_ptr(), int(zero_point.dtype == torch.int32),
                                                     float(lsq_grad_factor), int(qmin), int(qmax), _stream()),
              "osq_fq_per_tensor_bins_f32")
        return y, bins
    if x.numel() == 0:
        return (y, torch.empty_like(x, dtype=torch.int16)) if (want_codes or want_bins) else y
    zp_is_int = zero_point.dtype == torch.int32
    if not zp_is_int and zero_point.dtype != torch.float32:
        raise TypeError("zero_point must be int32 or float32")
    if scale.dtype != torch.float32:
        raise TypeError("scale must be float32")
    codes = torch.empty_like(x, dtype=torch.int16) if want_codes else None
    check(_lib.load().osq_fq_per_tensor_f32(x.data_ptr(), y.data_ptr(), _ptr(codes), x.numel(), scale.data_ptr(),
                                            zero_point.data_ptr(), int(zp_is_int), float(lsq_grad_factor),
                                            int(qmin), int(qmax), _stream()), "osq_fq_per_tensor_f32")
    return (y, codes) if want_codes else y


def residual_layernorm_fq(h: torch.Tensor, res: Optional[torch.Tensor], res_gamma: Optional[torch.Tensor],
                          ln_weight: Optional[torch.Tensor], ln_bias: Optional[torch.Tensor], eps: float, scale: torch.Tensor,
                          zero_point: torch.Tensor, qmin: int, qmax: int, lsq_grad_factor: float = 0.0, want_bins: bool = False,
                          want_ln: bool = False):
    """GammaResidual + LayerNorm + its output quantizer in one pass (util_layernorm.py:14-17 / :34-37, :41-52).
    Returns y, or (y, bins, ln_out) with the optional outputs as None when not requested."""
    h, res = plain(h), plain(res)
    _require_cuda(h, scale, zero_point)
    if h.dtype != torch.float32 or not h.is_contiguous():
        raise TypeError("h must be a contiguous float32 tensor")
    H = h.shape[-1]
    rows = h.numel() // max(H, 1)
    for name, t, n in (("res", res, h.numel()), ("res_gamma", res_gamma, H), ("ln_weight", ln_weight, H), ("ln_bias", ln_bias, H)):
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != n or not t.is_cuda):
            raise TypeError("%s must be a contiguous float32 CUDA tensor of %d elements" % (name, n))
    if scale.dtype != torch.float32 or zero_point.dtype not in (torch.float32, torch.int32):
        raise TypeError("scale must be float32, zero_point float32 or int32")
    y = torch.empty_like(h)
    bins = torch.empty_like(h, dtype=torch.uint8) if want_bins else None
    ln = torch.empty_like(h) if want_ln else None
    if h.numel() > 0:
        check(_lib.load().osq_residual_layernorm_fq_f32(h.data_ptr(), _ptr(res), _ptr(res_gamma), _ptr(ln_weight), _ptr(ln_bias), float(eps),
                                                        rows, H, scale.data_ptr(), zero_point.data_ptr(),
                                                        int(zero_point.dtype == torch.int32), float(lsq_grad_factor), int(qmin),
                                                        int(qmax), y.data_ptr(), _ptr(bins), _ptr(ln), _stream()),
              "osq_residual_layernorm_fq_f32")
    return (y, bins, ln) if (want_bins or want_ln) else y


def _quantizer_args(q: dict):
    """dict(scale=, zp=, qmin=, qmax=, g=) -> osq_quantizer_t (the tensors must stay alive until the launch is issued)."""
    scale, zp = q["scale"], q["zp"]
    if scale.dtype != torch.float32 or zp.dtype not in (torch.float32, torch.int32) or not scale.is_cuda or not zp.is_cuda:
        raise TypeError("quantizer scale must be a float32 CUDA tensor, zero_point float32 or int32")
    a = _lib.QuantizerArgs()
    a.scale = scale.data_ptr(); a.zero_point = zp.data_ptr(); a.zp_is_int32 = int(zp.dtype == torch.int32)
    a.lsq_grad_factor = float(q.get("g", 0.0)); a.qmin = int(q["qmin"]); a.qmax = int(q["qmax"])
    return a


def _head_view(x: torch.Tensor, what: str):
    """[B, h, S, d] view with unit channel stride -> (strides {batch, head, token})."""
    if x.dim() != 4 or x.dtype != torch.float32 or not x.is_cuda or x.stride(3) != 1:
        raise TypeError("%s must be a float32 CUDA [batch, heads, tokens, d] view with contiguous channels" % what)
    import ctypes as C
    return (C.c_int64 * 3)(x.stride(0), x.stride(1), x.stride(2))


def attn_scores_fq(q: torch.Tensor, k: torch.Tensor, qq: dict, kq: dict, out_mul: float = 1.0, mask: Optional[torch.Tensor] = None):
    """fq_q(q) @ fq_k(k)^T [* out_mul] [+ mask] in one launch (quant_bert.py:148-150, :169-172).  q [B, h, Sq, d], k [B, h, Sk, d]
    (transpose_for_scores views); mask additive, broadcastable from [B, 1, 1, Sk].  Returns scores [B, h, Sq, Sk]."""
    q, k = plain(q), plain(k)
    qs, ks = _head_view(q, "q"), _head_view(k, "k")
    B, h, sq, d = q.shape
    sk = k.shape[2]
    if k.shape[0] != B or k.shape[1] != h or k.shape[3] != d:
        raise ValueError("q / k shape mismatch")
    if mask is not None:
        if mask.dtype != torch.float32 or mask.numel() != B * sk or not mask.is_contiguous() or not mask.is_cuda:
            raise TypeError("mask must be a contiguous float32 CUDA tensor of batch * sk elements")
    out = torch.empty(B, h, sq, sk, device=q.device, dtype=torch.float32)
    qa, ka = _quantizer_args(qq), _quantizer_args(kq)
    check(_lib.load().osq_attn_scores_fq_f32(q.data_ptr(), k.data_ptr(), B, h, sq, sk, d, qs, ks, qa, ka, float(out_mul), _ptr(mask),
                                             out.data_ptr(), _stream()), "osq_attn_scores_fq_f32")
    return out


def attn_context_fq(probs: torch.Tensor, v: torch.Tensor, pq: dict, vq: dict, oq: Optional[dict] = None, want_bins: bool = False):
    """fq_p(probs) @ fq_v(v), written as [B, Sq, h * d] (quant_bert.py:185-191); with ``oq`` the context quantizer (:192-193) runs in
    the epilogue and ``want_bins`` adds its uint8 bins.  probs [B, h, Sq, Sk] contiguous, v [B, h, Sk, d] view."""
    probs, v = plain(probs), plain(v)
    vs = _head_view(v, "v")
    B, h, sk, d = v.shape
    if probs.dim() != 4 or probs.dtype != torch.float32 or not probs.is_contiguous() or probs.shape[0] != B or probs.shape[1] != h or probs.shape[3] != sk:
        raise TypeError("probs must be a contiguous float32 [batch, heads, sq, sk] tensor matching v")
    sq = probs.shape[2]
    out = torch.empty(B, sq, h * d, device=v.device, dtype=torch.float32)
    bins = torch.empty(B, sq, h * d, device=v.device, dtype=torch.uint8) if (want_bins and oq is not None) else None
    pa, va = _quantizer_args(pq), _quantizer_args(vq)
    oa = _quantizer_args(oq) if oq is not None else None
    check(_lib.load().osq_attn_context_fq_f32(probs.data_ptr(), v.data_ptr(), B, h, sq, sk, d, vs, pa, va, oa, out.data_ptr(), _ptr(bins),
                                              _stream()), "osq_attn_context_fq_f32")
    return (out, bins) if want_bins else out


def fq_bins_only(x: torch.Tensor, scale: torch.Tensor, zero_point: torch.Tensor, qmin: int, qmax: int, lsq_grad_factor: float = 0.0,
                 act: Optional[str] = None):
    """K1d: the uint8 bins of the per-tensor fake-quantize without its fp32 output (5 B / element).  Returns (bins, eff) with
    eff = device float[2] holding the effective (scale, zero_point) of the launch, for ``dequant_bins``."""
    _require_cuda(x, scale, zero_point)
    if x.dtype != torch.float32 or not x.is_contiguous() or x.numel() % 4 != 0:
        raise TypeError("x must be a contiguous float32 tensor with a multiple of 4 elements")
    if scale.dtype != torch.float32 or zero_point.dtype not in (torch.float32, torch.int32):
        raise TypeError("scale must be float32, zero_point float32 or int32")
    bins = torch.empty_like(x, dtype=torch.uint8)
    eff = torch.empty(2, dtype=torch.float32, device=x.device)
    if x.numel() > 0:
        check(_lib.load().osq_fq_per_tensor_bins_only_f32(x.data_ptr(), bins.data_ptr(), x.numel(), {None: 0, "none": 0, "gelu": 1}[act],
                                                          scale.data_ptr(), zero_point.data_ptr(),
                                                          int(zero_point.dtype == torch.int32), float(lsq_grad_factor), int(qmin), int(qmax),
                                                          eff.data_ptr(), _stream()), "osq_fq_per_tensor_bins_only_f32")
    return bins, eff


def dequant_bins(bins: torch.Tensor, eff: torch.Tensor, qmin: int, qmax: int) -> torch.Tensor:
    """(bins, eff) of ``fq_bins_only`` -> the fp32 tensor the fake-quantize would have written (util_quant.py:14)."""
    _require_cuda(bins, eff)
    if bins.dtype != torch.uint8 or not bins.is_contiguous() or bins.numel() % 4 != 0:
        raise TypeError("bins must be a contiguous uint8 tensor with a multiple of 4 elements")
    y = torch.empty(bins.shape, dtype=torch.float32, device=bins.device)
    if bins.numel() > 0:
        check(_lib.load().osq_dequant_bins_f32(bins.data_ptr(), eff.data_ptr(), int(qmin), int(qmax), y.data_ptr(), bins.numel(), _stream()),
              "osq_dequant_bins_f32")
    return y


def fq_per_channel(x: torch.Tensor, scale: torch.Tensor, zero_point: torch.Tensor, qmin: int, qmax: int,
                   want_codes: bool = False):
    """util_quant.py:18-26 for ch_axis = 0; x is viewed as [rows, cols]."""
    _require_cuda(x, scale, zero_point)
    x = x.float().contiguous()
    rows = x.shape[0]
    cols = x.numel() // max(rows, 1)
    y = torch.empty_like(x)
    codes = torch.empty_like(x, dtype=torch.int16) if want_codes else None
    if x.numel():
        scale = scale.float().contiguous()
        zp = zero_point.to(torch.int32).contiguous()
        if scale.numel() != rows or zp.numel() != rows:
            raise ValueError("per-channel scale / zero_point must have one entry per row")
        check(_lib.load().osq_fq_per_channel_f32(x.data_ptr(), y.data_ptr(), _ptr(codes), rows, cols, scale.data_ptr(),
                                                 zp.data_ptr(), int(qmin), int(qmax), _stream()), "osq_fq_per_channel_f32")
    return (y, codes) if want_codes else y


# ------------------------------------------------------------------------------------------------
# token geometry (observer.py:72-98)
# ------------------------------------------------------------------------------------------------
def token_geometry(x: torch.Tensor, seq_pos: int) -> Tokens:
    """[B, S, F1, F2] view of a 3-D / 4-D activation with the sequence axis at ``seq_pos``."""
    nd = x.dim()
    if seq_pos < 0:
        seq_pos += nd
    rest = [d for d in range(nd) if d != seq_pos]
    if len(rest) not in (2, 3):
        raise ValueError("expected a 3-D or 4-D activation, got %d-D" % nd)
    size, stride = x.shape, x.stride()
    b = rest[0]
    if len(rest) == 2:
        f1s, f1st = 1, 0
        f2 = rest[1]
    else:
        f1s, f1st = size[rest[1]], stride[rest[1]]
        f2 = rest[2]
    F1, F2, sf1, sf2 = f1s, size[f2], f1st, stride[f2]
    if F1 > 1 and sf2 == 1 and sf1 == F2:  # token row is one contiguous run (q / k^T / v views)
        F1, F2, sf1 = 1, F1 * F2, 0
    if F1 > 1 and sf1 == 1 and sf2 == F1:  # transposed but still contiguous
        F1, F2, sf1, sf2 = 1, F1 * F2, 0, 1
    return Tokens(size[b], size[seq_pos], F1, F2, stride[b], stride[seq_pos], sf1, sf2)


def _lens_arg(lens: Optional[torch.Tensor], device):
    if lens is None:
        return None, 0
    if not torch.is_tensor(lens):
        lens = torch.as_tensor(lens)
    lens = lens.to(device=device, dtype=torch.int64).contiguous()
    return lens, lens.numel()


def _epilogue(mode: int, cnt: int, state_min, state_max, scale_out, zp_out, qmin, qmax, symmetric) -> StatEpilogue:
    e = StatEpilogue()
    e.mode, e.cnt = mode, int(cnt)
    e.state_min, e.state_max = _ptr(state_min), _ptr(state_max)
    e.scale_out, e.zp_out = _ptr(scale_out), _ptr(zp_out)
    e.zp_out_is_int32 = int(zp_out is not None and zp_out.dtype == torch.int32)
    e.qmin, e.qmax, e.symmetric = int(qmin), int(qmax), int(bool(symmetric))
    return e


def _prep_act(x: torch.Tensor) -> torch.Tensor:
    _require_cuda(x)
    x = x.detach()
    return x if x.dtype == torch.float32 else x.float()


def _cur_out(out, device):
    if out is None:
        return torch.empty(2, dtype=torch.float32, device=device)
    assert out.is_cuda and out.dtype == torch.float32 and out.numel() == 2 and out.is_contiguous()
    return out


def observe_minmax(x, lens, seq_pos, *, mode=STAT_NONE, cnt=0, state_min=None, state_max=None, scale_out=None,
                   zp_out=None, qmin=0, qmax=255, symmetric=False, out=None) -> torch.Tensor:
    """Masked global (min,max) + running statistic + qparams in ONE launch (observer.py:184-203)."""
    x = _prep_act(x)
    cur = _cur_out(out, x.device)
    epi = _epilogue(mode, cnt, state_min, state_max, scale_out, zp_out, qmin, qmax, symmetric)
    ws = workspace(x.device)
    lib = _lib.load()
    if lens is None and (seq_pos == -1 or x.dim() < 3):
        xc = x if (_is_dense(x)) else x.contiguous()
        check(lib.osq_minmax_flat_f32(xc.data_ptr(), xc.numel(), cur.data_ptr(), C.byref(epi), ws.data_ptr(), _stream()),
              "osq_minmax_flat_f32")
        return cur
    if lens is None and (_is_dense(x)):
        check(lib.osq_minmax_flat_f32(x.data_ptr(), x.numel(), cur.data_ptr(), C.byref(epi), ws.data_ptr(), _stream()),
              "osq_minmax_flat_f32")
        return cur
    tok = token_geometry(x, seq_pos)
    lens_t, n_lens = _lens_arg(lens, x.device)
    check(lib.osq_minmax_masked_f32(x.data_ptr(), C.byref(tok), _ptr(lens_t), n_lens, cur.data_ptr(), C.byref(epi),
                                    ws.data_ptr(), _stream()), "osq_minmax_masked_f32")
    return cur


def token_minmax(x, lens, seq_pos):
    """Per-token extrema (observer.py:64-65) with pad removal fused in. Returns (tmin, tmax, n_valid)."""
    x = _prep_act(x)
    tok = token_geometry(x, seq_pos)
    n = tok.B * tok.S
    tmin = torch.empty(n, dtype=torch.float32, device=x.device)
    tmax = torch.empty(n, dtype=torch.float32, device=x.device)
    n_valid = torch.empty(1, dtype=torch.int32, device=x.device)
    lens_t, n_lens = _lens_arg(lens, x.device)
    check(_lib.load().osq_token_minmax_f32(x.data_ptr(), C.byref(tok), _ptr(lens_t), n_lens, tmin.data_ptr(),
                                           tmax.data_ptr(), n_valid.data_ptr(), _stream()), "osq_token_minmax_f32")
    return tmin, tmax, n_valid


_token_scratch = {}


def _token_vectors(device, n: int):
    """per-(device, stream) scratch for the per-token extrema of one observer call: (tmin[n], tmax[n], n_valid[1]).
    Grown, never shrunk: the observer runs ~100 k times per calibration and must not hit the allocator."""
    key = (torch.device(device).index, _stream())
    buf = _token_scratch.get(key)
    n4 = (n + 3) & ~3   # both vectors 16-byte aligned (the select tail reads them with 128-bit loads)
    if buf is None or buf[0].numel() < 2 * n4:
        cap = max(2 * n4, 1 << 16)
        buf = (torch.empty(cap, dtype=torch.float32, device=device), torch.empty(1, dtype=torch.int32, device=device))
        _token_scratch[key] = buf
    return buf[0][:n], buf[0][n4:n4 + n], buf[1]


def observe_prune_minmax(x, lens, seq_pos, percentile, *, mode=STAT_NONE, cnt=0, state_min=None, state_max=None,
                         scale_out=None, zp_out=None, qmin=0, qmax=255, symmetric=False, out=None, use_sort=False,
                         legacy_select=False, percentile_dev=None) -> torch.Tensor:
    """AvgPruneMinMaxObserver's token pruning (observer.py:50-70,214-237): one pass over the activation (per-token
    extrema) and, parked behind it by programmatic dependent launch, one thread-block cluster that keeps the [T]
    vectors in shared memory for the exact radix select + clip selection + running statistics
    (osq_prune_observe_f32).  ``legacy_select`` / ``use_sort`` run the earlier L2-resident select paths (cross-checks).
    ``percentile_dev`` (device fp32[1]) overrides ``percentile`` at run time (CUDA-graph replays of a calibration forward)."""
    if not (use_sort or legacy_select):
        x = _prep_act(x)
        tok = token_geometry(x, seq_pos)
        tmin, tmax, n_valid = _token_vectors(x.device, tok.B * tok.S)
        cur = _cur_out(out, x.device)
        epi = _epilogue(mode, cnt, state_min, state_max, scale_out, zp_out, qmin, qmax, symmetric)
        lens_t, n_lens = _lens_arg(lens, x.device)
        check(_lib.load().osq_prune_observe_f32(x.data_ptr(), C.byref(tok), _ptr(lens_t), n_lens, float(percentile),
                                                _ptr(percentile_dev), tmin.data_ptr(), tmax.data_ptr(), n_valid.data_ptr(), cur.data_ptr(),
                                                C.byref(epi), workspace(x.device).data_ptr(), _stream()), "osq_prune_observe_f32")
        return cur
    tmin, tmax, n_valid = token_minmax(x, lens, seq_pos)
    cur = _cur_out(out, tmin.device)
    epi = _epilogue(mode, cnt, state_min, state_max, scale_out, zp_out, qmin, qmax, symmetric)
    if use_sort:
        # first version: invalid tokens hold (+inf, -inf): |.| maps both to +inf so they sort behind the T valid entries
        abs_tmin_sorted = torch.sort(tmin.abs()).values
        abs_tmax_sorted = torch.sort(tmax.abs()).values
        check(_lib.load().osq_prune_select_f32(tmin.data_ptr(), tmax.data_ptr(), abs_tmin_sorted.data_ptr(),
                                               abs_tmax_sorted.data_ptr(), tmin.numel(), n_valid.data_ptr(),
                                               float(percentile), cur.data_ptr(), C.byref(epi),
                                               workspace(tmin.device).data_ptr(), _stream()), "osq_prune_select_f32")
    else:
        check(_lib.load().osq_prune_select_unsorted_f32(tmin.data_ptr(), tmax.data_ptr(), tmin.numel(), n_valid.data_ptr(),
                                                        float(percentile), cur.data_ptr(), C.byref(epi),
                                                        workspace(tmin.device).data_ptr(), _stream()),
              "osq_prune_select_unsorted_f32")
    return cur


_many_scratch = {}


def observe_prune_minmax_many(xs, lens, seq_pos, percentile, epilogues, outs=None, percentile_dev=None):
    """``observe_prune_minmax`` for a list of calibration batches of one geometry in ONE call (osq_prune_observe_many_f32): the same
    results as calling it once per batch, with batch i + 1's per-token pass running next to batch i's one-CTA select tail.
    ``epilogues``: one dict per batch with the keyword arguments of ``_epilogue`` (mode, cnt, state_min, ...); ``outs``: optional
    list of float32[2] outputs (e.g. slots of a sharded calibration table).  Returns the list of per-batch (min, max) tensors."""
    xs = [_prep_act(x) for x in xs]
    tok = token_geometry(xs[0], seq_pos)
    for x in xs[1:]:
        if x.shape != xs[0].shape or x.stride() != xs[0].stride():
            raise ValueError("all batches of one call must share shape and strides")
    n, n_slots = len(xs), tok.B * tok.S
    dev = xs[0].device
    n4 = (n_slots + 3) & ~3
    key = (dev.index, _stream())
    buf = _many_scratch.get(key)
    if buf is None or buf[0].numel() < 4 * n4:
        buf = (torch.empty(max(4 * n4, 1 << 17), dtype=torch.float32, device=dev), torch.empty(2, dtype=torch.int32, device=dev))
        _many_scratch[key] = buf
    tmin2, tmax2 = buf[0][:2 * n4], buf[0][2 * n4:4 * n4]
    curs = [_cur_out(None if outs is None else outs[i], dev) for i in range(n)]
    epis = (StatEpilogue * n)(*[_epilogue(**e) for e in epilogues])
    xp = (C.c_void_p * n)(*[x.data_ptr() for x in xs])
    cp = (C.c_void_p * n)(*[c.data_ptr() for c in curs])
    lens_t, n_lens = _lens_arg(lens, dev)
    check(_lib.load().osq_prune_observe_many_f32(xp, n, C.byref(tok), _ptr(lens_t), n_lens, float(percentile), _ptr(percentile_dev),
                                                 tmin2.data_ptr(), tmax2.data_ptr(), buf[1].data_ptr(), cp, epis, workspace(dev).data_ptr(),
                                                 _stream()), "osq_prune_observe_many_f32")
    return curs


_hist_scratch = {}


def token_minmax_hist(x, lens, seq_pos):
    """Per-token extrema of one activation plus the first-digit table of the select, into FRESH buffers (a cache entry of the
    token-wise-clipping sweep, osq_token_minmax_hist_f32).  Returns (tmin, tmax, n_valid, hist0) or None when the token count is
    beyond the cached select."""
    x = _prep_act(x)
    tok = token_geometry(x, seq_pos)
    n = tok.B * tok.S
    if n >= (1 << 21):
        return None
    tmin = torch.empty(n, dtype=torch.float32, device=x.device)
    tmax = torch.empty(n, dtype=torch.float32, device=x.device)
    n_valid = torch.zeros(1, dtype=torch.int32, device=x.device)
    hist0 = torch.zeros(2 * 2048 + 64, dtype=torch.int32, device=x.device)
    lens_t, n_lens = _lens_arg(lens, x.device)
    check(_lib.load().osq_token_minmax_hist_f32(x.data_ptr(), C.byref(tok), _ptr(lens_t), n_lens, tmin.data_ptr(), tmax.data_ptr(),
                                                n_valid.data_ptr(), hist0.data_ptr(), _stream()), "osq_token_minmax_hist_f32")
    return tmin, tmax, n_valid, hist0


def select_problems(entries, device) -> torch.Tensor:
    """Device copy of an osq_select_problem_t array.  entries: ((tmin, tmax, n_valid, hist0), cur) with cur a float32[2] view."""
    arr = (_lib.SelectProblem * len(entries))()
    for t, ((tmin, tmax, n_valid, hist0), cur) in zip(arr, entries):
        t.tmin, t.tmax, t.n_slots, t.n_valid, t.hist0, t.cur = tmin.data_ptr(), tmax.data_ptr(), tmin.numel(), n_valid.data_ptr(), hist0.data_ptr(), cur.data_ptr()
    return torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)


def prune_select_cached(problems: torch.Tensor, n_problems: int, percentile: float, percentile_dev: Optional[torch.Tensor] = None) -> None:
    """observer.py:50-70 on ``n_problems`` recorded (observer, batch) pairs in one launch (osq_prune_select_cached_f32)."""
    _require_cuda(problems, percentile_dev)
    check(_lib.load().osq_prune_select_cached_f32(problems.data_ptr(), int(n_problems), float(percentile), _ptr(percentile_dev), _stream()),
          "osq_prune_select_cached_f32")


def observe_quantile(x, lens, seq_pos, bins, threshold, *, mode=STAT_NONE, cnt=0, state_min=None, state_max=None,
                     scale_out=None, zp_out=None, qmin=0, qmax=255, symmetric=False, out=None) -> torch.Tensor:
    """AvgQuantileObserver.forward (observer.py:253-282): masked min/max, |x| histogram, cumulative-threshold clip and
    the running average in two launches (osq_quantile_observe_f32)."""
    x = _prep_act(x)
    if x.dim() >= 3 and seq_pos != -1:
        tok = token_geometry(x, seq_pos)
    else:
        x = x if x.is_contiguous() else x.contiguous()
        tok = Tokens(1, 1, 1, x.numel(), 0, 0, 0, 1)
        lens = None
    key = (x.device.index, _stream(), int(bins))
    hist = _hist_scratch.get(key)
    if hist is None:
        hist = _hist_scratch[key] = torch.zeros(int(bins), dtype=torch.int32, device=x.device)
    cur = _cur_out(out, x.device)
    epi = _epilogue(mode, cnt, state_min, state_max, scale_out, zp_out, qmin, qmax, symmetric)
    lens_t, n_lens = _lens_arg(lens, x.device)
    check(_lib.load().osq_quantile_observe_f32(x.data_ptr(), C.byref(tok), _ptr(lens_t), n_lens, int(bins), float(threshold),
                                               hist.data_ptr(), cur.data_ptr(), C.byref(epi), workspace(x.device).data_ptr(),
                                               _stream()), "osq_quantile_observe_f32")
    return cur


def replay_targets(entries, device) -> torch.Tensor:
    """Device copy of an osq_replay_target_t array.  entries: (state_min, state_max, scale_out, zp_out, qmin, qmax, symmetric)
    per observer (tensors; scale_out / zp_out may be None)."""
    arr = (ReplayTarget * len(entries))()
    for t, (mn, mx, s_out, z_out, qmin, qmax, sym) in zip(arr, entries):
        t.state_min, t.state_max, t.scale_out, t.zp_out = _ptr(mn), _ptr(mx), _ptr(s_out), _ptr(z_out)
        t.zp_out_is_int32 = int(z_out is not None and z_out.dtype == torch.int32)
        t.qmin, t.qmax, t.symmetric = int(qmin), int(qmax), int(bool(sym))
    raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    return raw.to(device)


def replay_average(table: torch.Tensor, cnt0: int, targets: torch.Tensor) -> None:
    """observer.py:194-202 replayed over the all-reduced slot table [n_obs, n_batches, 2] for every observer in one launch."""
    _require_cuda(table, targets)
    n_obs, n_batches = table.shape[0], table.shape[1]
    assert table.is_contiguous() and table.dtype == torch.float32
    check(_lib.load().osq_replay_average_f32(table.data_ptr(), n_obs, n_batches, int(cnt0), targets.data_ptr(), _stream()),
          "osq_replay_average_f32")


def replay_average_peer(peer_ptrs: torch.Tensor, world: int, n_obs: int, n_batches: int, cnt0: int, targets: torch.Tensor) -> None:
    """The replay with every slot loaded from the table of the rank that owns it (peer-mapped pointers, no collective)."""
    _require_cuda(peer_ptrs, targets)
    assert peer_ptrs.dtype == torch.int64 and peer_ptrs.numel() == world
    check(_lib.load().osq_replay_average_peer_f32(peer_ptrs.data_ptr(), int(world), int(n_obs), int(n_batches), int(cnt0),
                                                  targets.data_ptr(), _stream()), "osq_replay_average_peer_f32")


def replay_exchange(local_table: torch.Tensor, region_ptrs: torch.Tensor, rank: int, world: int, n_obs: int, n_batches: int, cnt0: int,
                    targets: torch.Tensor, pass_counter: torch.Tensor, err_flag: torch.Tensor) -> None:
    """Exchange + replay in one launch over peer memory (osq_replay_exchange_f32): publish this rank's slots into its symmetric
    region, signal every peer, wait for every peer, replay with each slot loaded from its owner's region."""
    _require_cuda(local_table, region_ptrs, targets, pass_counter, err_flag)
    assert region_ptrs.dtype == torch.int64 and region_ptrs.numel() == world and local_table.dtype == torch.float32
    assert pass_counter.dtype == torch.int32 and err_flag.dtype == torch.int32
    check(_lib.load().osq_replay_exchange_f32(local_table.data_ptr(), region_ptrs.data_ptr(), int(rank), int(world), int(n_obs), int(n_batches),
                                              int(cnt0), targets.data_ptr(), pass_counter.data_ptr(), err_flag.data_ptr(), _stream()),
          "osq_replay_exchange_f32")


def rowwise_minmax_qparams(w, first, state_min, state_max, scale_out, zp_out, qmin, qmax, symmetric):
    """MinMaxObserver(ch_axis=0) + calculate_qparams on a [N, K] weight in one launch."""
    _require_cuda(w, state_min, state_max)
    w = w.detach().float().contiguous()
    rows = w.shape[0]
    cols = w.numel() // rows
    check(_lib.load().osq_rowwise_minmax_qparams_f32(w.data_ptr(), rows, cols, int(bool(first)), state_min.data_ptr(),
                                                     state_max.data_ptr(), _ptr(scale_out), _ptr(zp_out), int(qmin),
                                                     int(qmax), int(bool(symmetric)), _stream()),
          "osq_rowwise_minmax_qparams_f32")


def calc_qparams(min_val, max_val, qmin, qmax, symmetric):
    """observer.py:100-119 on device; returns (scale fp32, zero_point: int32 if symmetric else fp32)."""
    _require_cuda(min_val, max_val)
    shape = min_val.shape
    mn = min_val.detach().float().contiguous().reshape(-1)
    mx = max_val.detach().float().contiguous().reshape(-1)
    scale = torch.empty_like(mn)
    zp_f = None if symmetric else torch.empty_like(mn)
    zp_i = torch.empty(mn.shape, dtype=torch.int32, device=mn.device) if symmetric else None
    check(_lib.load().osq_calc_qparams_f32(mn.data_ptr(), mx.data_ptr(), mn.numel(), int(qmin), int(qmax),
                                           int(bool(symmetric)), scale.data_ptr(), _ptr(zp_f), _ptr(zp_i), _stream()),
          "osq_calc_qparams_f32")
    zp = zp_i if symmetric else zp_f
    return scale.reshape(shape), zp.reshape(shape)


# ------------------------------------------------------------------------------------------------
# K5
# ------------------------------------------------------------------------------------------------
def mse_multi(x, lens, seq_pos, cand_scale, cand_zp, qmin, qmax):
    """sum of squared fq error per candidate + number of valid elements (device tensors)."""
    x = _prep_act(x)
    if x.dim() >= 3 and seq_pos != -1:
        tok = token_geometry(x, seq_pos)
    else:
        xc = x if x.is_contiguous() else x.contiguous()
        x = xc
        tok = Tokens(1, 1, 1, x.numel(), 0, 0, 0, 1)
        lens = None
    lens_t, n_lens = _lens_arg(lens, x.device)
    cand_scale = cand_scale.to(device=x.device, dtype=torch.float32).contiguous()
    cand_zp = cand_zp.to(device=x.device, dtype=torch.float32).contiguous()
    n = cand_scale.numel()
    loss = torch.empty(n, dtype=torch.float64, device=x.device)
    n_valid = torch.empty(1, dtype=torch.int64, device=x.device)
    check(_lib.load().osq_mse_multi_f32(x.data_ptr(), C.byref(tok), _ptr(lens_t), n_lens, cand_scale.data_ptr(),
                                        cand_zp.data_ptr(), n, int(qmin), int(qmax), loss.data_ptr(), n_valid.data_ptr(),
                                        _stream()), "osq_mse_multi_f32")
    return loss, n_valid


_mse_scratch = {}


def mse_brent_tensor(x, lens, seq_pos, qmin, qmax, symmetric, one_side_state):
    """(Avg)MSEFastObserver per-tensor search (observer.py:434-494) in ONE cooperative launch, no host round trip.
    one_side_state: device int32[1] (-1 undecided / 0 no / 1 pos / 2 neg), decided on the first call and kept.
    Returns (out4: device float64[4] = best_min, best_max, x_min, x_max; evals: device int32[1])."""
    x = _prep_act(x)
    if x.dim() >= 3 and seq_pos != -1:
        tok = token_geometry(x, seq_pos)
    else:
        x = x if x.is_contiguous() else x.contiguous()
        tok = Tokens(1, 1, 1, x.numel(), 0, 0, 0, 1)
        lens = None
    lens_t, n_lens = _lens_arg(lens, x.device)
    key = (x.device.index, _stream())
    scratch = _mse_scratch.get(key)
    if scratch is None:
        scratch = _mse_scratch[key] = torch.zeros(int(_lib.load().osq_mse_tensor_scratch_bytes()), dtype=torch.uint8, device=x.device)
    out4 = torch.empty(4, dtype=torch.float64, device=x.device)
    evals = torch.empty(1, dtype=torch.int32, device=x.device)
    check(_lib.load().osq_mse_brent_tensor_f32(x.data_ptr(), C.byref(tok), _ptr(lens_t), n_lens, int(qmin), int(qmax),
                                               int(bool(symmetric)), one_side_state.data_ptr(), out4.data_ptr(), evals.data_ptr(),
                                               scratch.data_ptr(), _stream()), "osq_mse_brent_tensor_f32")
    return out4, evals


def mse_brent_rows(w, qmin, qmax, one_side: str, want_evals=False):
    """MSEFastObserver per-channel 1-D search (observer.py:483-517) entirely on-chip."""
    _require_cuda(w)
    w = w.detach().float().contiguous()
    rows = w.shape[0]
    cols = w.numel() // rows
    out_min = torch.empty(rows, dtype=torch.float32, device=w.device)
    out_max = torch.empty(rows, dtype=torch.float32, device=w.device)
    evals = torch.empty(rows, dtype=torch.int32, device=w.device) if want_evals else None
    side = {"no": 0, "pos": 1, "neg": 2}[one_side]
    check(_lib.load().osq_mse_brent_rows_f32(w.data_ptr(), rows, cols, int(qmin), int(qmax), side, out_min.data_ptr(),
                                             out_max.data_ptr(), _ptr(evals), _stream()), "osq_mse_brent_rows_f32")
    return (out_min, out_max, evals) if want_evals else (out_min, out_max)


# ------------------------------------------------------------------------------------------------
# K6
# ------------------------------------------------------------------------------------------------
def pack_weight(w, scale, zero_point, qmin, qmax):
    """bins (q - zp) as int8 [N, K] + per-row sums (int32 [N])."""
    _require_cuda(w, scale, zero_point)
    w = w.detach().float().contiguous()
    n, k = w.shape[0], w.numel() // w.shape[0]
    scale = scale.detach().float().contiguous()
    zp = zero_point.detach().to(torch.int32).contiguous()
    codes = torch.empty((n, k), dtype=torch.int8, device=w.device)
    rowsum = torch.empty(n, dtype=torch.int32, device=w.device)
    check(_lib.load().osq_pack_weight_s8(w.data_ptr(), n, k, scale.data_ptr(), zp.data_ptr(), int(qmin), int(qmax),
                                         codes.data_ptr(), rowsum.data_ptr(), _stream()), "osq_pack_weight_s8")
    return codes, rowsum


def fused_linear_supported(k: int, n: int, a: Optional[torch.Tensor] = None) -> bool:
    """The shape / alignment contract of osq_fused_fq_linear (include/osq.h): K % 128 == 0, K <= 32768 (int32
    accumulators), N % 16 == 0, N <= 2^20, 16-byte aligned operands."""
    if not (128 <= k <= 32768 and k % 128 == 0 and 16 <= n <= (1 << 20) and n % 16 == 0):
        return False
    if a is not None and not getattr(a, "_osq_lazy", False) and a.is_contiguous() and a.data_ptr() % 16 != 0:
        return False
    return True


def fused_fq_linear(a, a_scale, a_zp, a_qmin, a_qmax, w_codes, w_scale, w_rowsum, bias, lsq_grad_factor=0.0,
                    mma_kind=0, want_codes=False, out=None, use_code_cache=True, trace=None, a_bins=None, out_q=None):
    """activation fq + weight fq + Linear in one tcgen05 kernel. a: [..., K] fp32 -> [..., N] fp32.
    a_bins (uint8, same shape as a, contiguous; from fq_per_tensor(want_bins=True) with the same qparams): bins-in launch,
    the fp32 tensor is not read at all (only its shape is used); the result is bit-identical.
    out_q (dict: scale, zp, qmin, qmax, g, act in {None, "gelu"}, bins: bool): fuse the NEXT activation quantizer into the
    epilogue -- the call returns (fq(act(Linear)), uint8 bins or None) instead of the Linear's output."""
    if a_bins is None:
        a = plain(a)
    _require_cuda(a, a_scale, a_zp, w_codes, w_scale, w_rowsum, bias)
    if a_bins is not None:
        if a_bins.dtype != torch.uint8 or a_bins.numel() != a.numel() or not a_bins.is_contiguous() or not a_bins.is_cuda:
            raise ValueError("a_bins must be a contiguous CUDA uint8 tensor with as many elements as a")
        m, k = a.numel() // a.shape[-1], a.shape[-1]
        a_ptr = None
    else:
        if a.dtype != torch.float32:
            a = a.float()
        a2 = a.reshape(-1, a.shape[-1])
        if not a2.is_contiguous():
            a2 = a2.contiguous()
        m, k = a2.shape
        a_ptr = a2.data_ptr()
    n = w_codes.shape[0]
    y = out if out is not None else torch.empty((m, n), dtype=torch.float32, device=a.device)
    # K > 1024: the bins do not fit in shared memory -> give the kernel an (L2 resident) code cache
    need_cache = use_code_cache and k > 1024 and n > 256
    dbg = a_bins if a_bins is not None else (
        torch.empty((m, k), dtype=torch.uint8, device=a.device) if (want_codes or need_cache) else None)
    args = FusedLinearArgs()
    args.A, args.M, args.K = a_ptr, m, k
    args.a_scale, args.a_zp = a_scale.data_ptr(), a_zp.data_ptr()
    args.a_zp_is_int32 = int(a_zp.dtype == torch.int32)
    args.lsq_grad_factor = float(lsq_grad_factor)
    args.a_qmin, args.a_qmax = int(a_qmin), int(a_qmax)
    args.w_codes, args.w_scale, args.w_rowsum = w_codes.data_ptr(), w_scale.data_ptr(), w_rowsum.data_ptr()
    args.bias = _ptr(bias)
    args.Y, args.N = y.data_ptr(), n
    args.mma_kind = int(mma_kind)
    args.a_codes = _ptr(dbg)
    args.debug_trace = _ptr(trace)
    out_bins = None
    if out_q is not None:
        _require_cuda(out_q["scale"], out_q["zp"])
        if out_q["scale"].dtype != torch.float32 or out_q["zp"].dtype not in (torch.float32, torch.int32):
            raise TypeError("output quantizer: scale must be float32, zero_point float32 or int32")
        args.out_act = {None: 0, "none": 0, "gelu": 1}[out_q.get("act")]
        args.out_scale, args.out_zp = out_q["scale"].data_ptr(), out_q["zp"].data_ptr()
        args.out_zp_is_int32 = int(out_q["zp"].dtype == torch.int32)
        args.out_lsq_grad_factor = float(out_q.get("g", 0.0))
        args.out_qmin, args.out_qmax = int(out_q["qmin"]), int(out_q["qmax"])
        if out_q.get("bins", True):
            out_bins = torch.empty((m, n), dtype=torch.uint8, device=a.device)
            args.out_bins = out_bins.data_ptr()
    if m > 0:
        check(_lib.load().osq_fused_fq_linear(C.byref(args), _stream()), "osq_fused_fq_linear")
    y = y.reshape(*a.shape[:-1], n)
    if out_q is not None:
        return y, (None if out_bins is None else out_bins.reshape(*a.shape[:-1], n))
    return (y, dbg) if want_codes else y


def fused_fq_linear_multi(sites):
    """One C call for a list of INDEPENDENT fused sites (osq_fused_fq_linear_multi): compatible runs share one persistent
    launch.  ``sites``: dicts with the keyword arguments of ``fused_fq_linear`` (a, a_scale, a_zp, a_qmin, a_qmax, w_codes,
    w_scale, w_rowsum, bias, lsq_grad_factor, out).  fp32-in, resident / code-cache plans only; returns the outputs."""
    arr = (FusedLinearArgs * len(sites))()
    outs, keep = [], []
    for args, st in zip(arr, sites):
        a = plain(st["a"])
        _require_cuda(a, st["a_scale"], st["a_zp"], st["w_codes"], st["w_scale"], st["w_rowsum"], st.get("bias"))
        a2 = a.reshape(-1, a.shape[-1])
        if a2.dtype != torch.float32 or not a2.is_contiguous():
            a2 = a2.float().contiguous()
        m, k = a2.shape
        n = st["w_codes"].shape[0]
        y = st.get("out")
        if y is None:
            y = torch.empty((m, n), dtype=torch.float32, device=a.device)
        cache = torch.empty((m, k), dtype=torch.uint8, device=a.device) if (k > 1024 and n > 256) else st.get("cache")
        if st.get("cache") is not None:
            cache = st["cache"]
        args.A, args.M, args.K = a2.data_ptr(), m, k
        args.a_scale, args.a_zp = st["a_scale"].data_ptr(), st["a_zp"].data_ptr()
        args.a_zp_is_int32 = int(st["a_zp"].dtype == torch.int32)
        args.lsq_grad_factor = float(st.get("lsq_grad_factor", 0.0))
        args.a_qmin, args.a_qmax = int(st["a_qmin"]), int(st["a_qmax"])
        args.w_codes, args.w_scale, args.w_rowsum = st["w_codes"].data_ptr(), st["w_scale"].data_ptr(), st["w_rowsum"].data_ptr()
        args.bias = _ptr(st.get("bias"))
        args.Y, args.N = y.data_ptr(), n
        args.mma_kind = 0
        args.a_codes = _ptr(cache)
        args.debug_trace = None
        outs.append(y.reshape(*a.shape[:-1], n))
        keep.append((a2, cache))
    check(_lib.load().osq_fused_fq_linear_multi(arr, len(sites), _stream()), "osq_fused_fq_linear_multi")
    return outs


def lsqplus_backward(x, dy, scale, zero_point, lsq_grad_factor, qmin, qmax):
    """gradients of util_quant.py:48-55: returns (dx, dscale[1], dzero_point[1])."""
    _require_cuda(x, dy, scale, zero_point)
    x = x.detach().float().contiguous()
    dy = dy.detach().float().contiguous()
    dx = torch.empty_like(x)
    acc = torch.zeros(2, dtype=torch.float64, device=x.device)
    check(_lib.load().osq_lsqplus_backward_f32(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), x.numel(), scale.data_ptr(),
                                               zero_point.data_ptr(), float(lsq_grad_factor), int(qmin), int(qmax),
                                               acc.data_ptr(), _stream()), "osq_lsqplus_backward_f32")
    g = acc.float()
    return dx, g[0:1], g[1:2]
