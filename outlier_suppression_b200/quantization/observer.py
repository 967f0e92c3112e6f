"""Calibration observers with the reference's class names, ctor signatures, buffers and
``forward(x_orig, observation_mask=None, seq_pos=-1)`` contract (quantization/observer.py), backed
by the sm_100a reduction kernels.

Differences that are deliberate and invisible to callers:
  * ``forward`` reads the activation ONCE in place -- no clone / permute / cat (observer.py:188-193,
    72-84) -- with the pad-token mask evaluated in-kernel as ``s < lens[b]``;
  * running statistics live in the ``min_val`` / ``max_val`` buffers and are updated in place by the
    kernel epilogue (the reference rebinds the attributes to new tensors);
  * ``AvgPruneMinMaxObserver`` never materialises the clipped tensor: the clip + global min/max of
    observer.py:69,227 equals the (lower, upper) pair selected from the per-token vectors.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import ops


def _transform_to_ch_axis(x, ch_axis):
    """observer.py:11-21: [C, rest] view with the channel axis first."""
    if ch_axis == -1:
        return x
    order = list(range(x.dim()))
    order[ch_axis], order[0] = 0, ch_axis
    return torch.flatten(x.permute(order), start_dim=1)


def _qparams_host(min_val: torch.Tensor, max_val: torch.Tensor, quant_min: int, quant_max: int, symmetric: bool):
    """observer.py:100-119 for HOST scalars (search bookkeeping of the MSE observers); same dtype
    promotion as the reference: an fp64 candidate stays fp64."""
    lo = torch.min(min_val, torch.zeros_like(min_val))
    hi = torch.max(max_val, torch.zeros_like(max_val))
    eps = torch.tensor(1e-8, dtype=torch.float32)
    if symmetric:
        hi = torch.max(-lo, hi)
        scale = torch.max(hi / (float(quant_max - quant_min) / 2), eps)
        zero_point = torch.zeros(lo.size(), dtype=torch.int)
    else:
        scale = torch.max((hi - lo) / float(quant_max - quant_min), eps)
        zero_point = torch.clamp(quant_min - torch.round(lo / scale), quant_min, quant_max)
    return scale, zero_point


class ObserverBase(nn.Module):
    """observer.py:24-119."""

    def __init__(self, bit=8, symmetric=False, ch_axis=-1):
        super().__init__()
        self.bit, self.symmetric, self.ch_axis = bit, symmetric, ch_axis
        self.eps = torch.tensor(1e-8, dtype=torch.float32)
        if symmetric:
            self.quant_min, self.quant_max = -2 ** (bit - 1), 2 ** (bit - 1) - 1
        else:
            self.quant_min, self.quant_max = 0, 2 ** bit - 1
        self.register_buffer("min_val", torch.tensor(float("inf")))
        self.register_buffer("max_val", torch.tensor(float("-inf")))

    _owner_hint = None   # the quantizer on whose behalf __call__ runs (set by QuantizeBase._run_observer)
    _last_fused = False

    # --- attribute pokes used by state.py / token_wise_clipping.py ---
    def set_name(self, name):
        self.name = name

    def set_batch(self, batch):
        self.batch = batch

    _percentile_dev = None  # device fp32[1] mirror of `percentile` (twc.GraphedFindRatio: the ratio of a replayed graph is data)

    def set_percentile(self, percentile):
        self.percentile = percentile
        if self._percentile_dev is not None:
            self._percentile_dev.fill_(float(percentile))

    @torch.jit.export
    def calculate_qparams(self, min_val, max_val):
        """observer.py:100-119.  Device tensors go through the kernel (true fp32 division -- ATen's CUDA
        div by a Python scalar multiplies by the reciprocal and is NOT bit-identical); host tensors use
        the same formula on the host."""
        if min_val.is_cuda:
            if min_val.dtype == torch.float64:  # MSEFast results are fp64 in the reference (np.float64)
                s, z = _qparams_host(min_val.cpu(), max_val.cpu(), self.quant_min, self.quant_max, self.symmetric)
                return s.to(min_val.device), z.to(min_val.device)
            return ops.calc_qparams(min_val, max_val, self.quant_min, self.quant_max, self.symmetric)
        return _qparams_host(min_val, max_val, self.quant_min, self.quant_max, self.symmetric)

    # --- helpers shared by the concrete observers ---
    def _ensure_scalar_state(self, device):
        if self.min_val.device != device:
            raise RuntimeError("observer buffers live on %s but the activation is on %s; move the model first"
                               % (self.min_val.device, device))
        if self.min_val.dtype != torch.float32 or self.min_val.numel() != 1:
            self.min_val = self.min_val.detach().float().reshape(-1)[:1].reshape(()).clone()
            self.max_val = self.max_val.detach().float().reshape(-1)[:1].reshape(()).clone()

    def _fused_targets(self, quantizer):
        """(scale_out, zp_out) of the owning quantizer when the kernel epilogue may write them."""
        if quantizer is None:
            return None, None
        return quantizer._per_tensor_qparam_targets()

    def _observe(self, x, observation_mask, seq_pos, quantizer=None) -> bool:
        """Updates min_val/max_val from ``x``; returns True when the quantizer's (scale, zero_point)
        were refreshed by the same launch."""
        raise NotImplementedError

    def forward(self, x_orig, observation_mask=None, seq_pos=-1):
        if x_orig.numel() == 0:
            return x_orig
        self._last_fused = self._observe(x_orig, observation_mask, seq_pos, self._owner_hint)
        return x_orig


class MinMaxObserver(ObserverBase):
    """Running min/max over the calibration set (observer.py:122-145)."""

    def _observe(self, x, observation_mask, seq_pos, quantizer=None):
        x = x.detach()
        if self.ch_axis == -1:
            self._ensure_scalar_state(x.device)
            s_out, z_out = self._fused_targets(quantizer)
            ops.observe_minmax(x, observation_mask, seq_pos, mode=ops.STAT_EXTREMA, state_min=self.min_val,
                               state_max=self.max_val, scale_out=s_out, zp_out=z_out, qmin=self.quant_min,
                               qmax=self.quant_max, symmetric=self.symmetric)
            return s_out is not None
        assert observation_mask is None
        y = _transform_to_ch_axis(x, self.ch_axis if self.ch_axis >= 0 else self.ch_axis + x.dim())
        rows = y.shape[0]
        first = self.min_val.numel() != rows
        if first:
            self.min_val = torch.empty(rows, dtype=torch.float32, device=x.device)
            self.max_val = torch.empty(rows, dtype=torch.float32, device=x.device)
        s_out = z_out = None
        if quantizer is not None:
            s_out, z_out = quantizer._per_channel_qparam_targets(rows)
        ops.rowwise_minmax_qparams(y, first, self.min_val, self.max_val, s_out, z_out, self.quant_min, self.quant_max,
                                   self.symmetric)
        return s_out is not None


class AvgMinMaxObserver(ObserverBase):
    """Average of per-batch min/max (observer.py:176-203)."""

    def __init__(self, bit=8, symmetric=False, ch_axis=-1):
        super().__init__(bit=bit, symmetric=symmetric, ch_axis=ch_axis)
        self.cnt = 0

    def _observe(self, x, observation_mask, seq_pos, quantizer=None):
        assert self.ch_axis == -1
        self._ensure_scalar_state(x.device)
        if getattr(self, "_shard", None) is not None:  # rank-sharded pass (dist.py): record this batch's pair only
            ctl, idx = self._shard
            ops.observe_minmax(x, observation_mask, seq_pos, out=ctl.table.slot(idx, ctl.batch))
            return True
        s_out, z_out = self._fused_targets(quantizer)
        ops.observe_minmax(x, observation_mask, seq_pos, mode=ops.STAT_AVERAGE, cnt=self.cnt, state_min=self.min_val,
                           state_max=self.max_val, scale_out=s_out, zp_out=z_out, qmin=self.quant_min,
                           qmax=self.quant_max, symmetric=self.symmetric)
        self.cnt += 1
        return s_out is not None


class AvgPruneMinMaxObserver(ObserverBase):
    """Token-wise clipping observer (observer.py:206-237, 50-70)."""

    def __init__(self, bit=8, symmetric=False, ch_axis=-1):
        super().__init__(bit=bit, symmetric=symmetric, ch_axis=ch_axis)
        self.cnt = 0

    def _observe(self, x, observation_mask, seq_pos, quantizer=None):
        assert self.ch_axis == -1
        self._ensure_scalar_state(x.device)
        s_out, z_out = self._fused_targets(quantizer)
        kw = dict(mode=ops.STAT_AVERAGE, cnt=self.cnt, state_min=self.min_val, state_max=self.max_val, scale_out=s_out,
                  zp_out=z_out, qmin=self.quant_min, qmax=self.quant_max, symmetric=self.symmetric)
        shard = getattr(self, "_shard", None)
        if shard is not None:  # rank-sharded pass (dist.py): record this batch's pair only
            kw = dict(out=shard[0].table.slot(shard[1], shard[0].batch))
        tokenwise = observation_mask is not None or seq_pos != -1
        record = getattr(self, "_twc_record", None)   # twc.GraphedFindRatio: keep what this call depends on besides the ratio
        if tokenwise and "attention_probs" not in self.name:  # observer.py:62-63
            if record is not None:
                record.append(("prune", ops.token_minmax_hist(x, observation_mask, seq_pos)))
            ops.observe_prune_minmax(x, observation_mask, seq_pos, self.percentile, percentile_dev=self._percentile_dev, **kw)
        else:
            cur = ops.observe_minmax(x, observation_mask, seq_pos, **kw)
            if record is not None:
                record.append(("plain", cur if shard is None else None))
        if shard is not None:
            return True
        self.cnt += 1
        return s_out is not None


def _observe_many_prune(self, xs, observation_mask, seq_pos, quantizer=None, batch_indices=None) -> bool:
    """``_observe`` for a list of calibration batches of one geometry in ONE call (osq_prune_observe_many_f32): same state as calling
    it once per batch in order, with the one-CTA select tails overlapped by the next batch's per-token pass.  ``batch_indices``:
    the batches' positions in a rank-sharded pass (dist.sharded_calibration)."""
    assert self.ch_axis == -1
    xs = [x for x in xs if x.numel() > 0]
    if not xs:
        return False
    tokenwise = (observation_mask is not None or seq_pos != -1) and "attention_probs" not in self.name
    same = all(x.shape == xs[0].shape and x.stride() == xs[0].stride() for x in xs)
    shard = getattr(self, "_shard", None)
    if not (tokenwise and same and len(xs) > 1 and getattr(self, "_twc_record", None) is None):
        fused = False
        for i, x in enumerate(xs):   # anything else: one call per batch
            if shard is not None and batch_indices is not None:
                shard[0].set_batch(batch_indices[i])
            fused = self._observe(x, observation_mask, seq_pos, quantizer)
        return fused
    self._ensure_scalar_state(xs[0].device)
    s_out, z_out = self._fused_targets(quantizer)
    if shard is not None:
        idx = batch_indices if batch_indices is not None else list(range(shard[0].batch, shard[0].batch + len(xs)))
        outs = [shard[0].table.slot(shard[1], b) for b in idx]
        epis = [dict(mode=ops.STAT_NONE, cnt=0, state_min=None, state_max=None, scale_out=None, zp_out=None, qmin=self.quant_min,
                     qmax=self.quant_max, symmetric=self.symmetric) for _ in xs]
        ops.observe_prune_minmax_many(xs, observation_mask, seq_pos, self.percentile, epis, outs=outs, percentile_dev=self._percentile_dev)
        return True
    epis = [dict(mode=ops.STAT_AVERAGE, cnt=self.cnt + i, state_min=self.min_val, state_max=self.max_val, scale_out=s_out, zp_out=z_out,
                 qmin=self.quant_min, qmax=self.quant_max, symmetric=self.symmetric) for i in range(len(xs))]
    ops.observe_prune_minmax_many(xs, observation_mask, seq_pos, self.percentile, epis, percentile_dev=self._percentile_dev)
    self.cnt += len(xs)
    return s_out is not None


AvgPruneMinMaxObserver._observe_many = _observe_many_prune


class MSEFastObserver(ObserverBase):
    """Golden-section / Brent search of the clipping range that minimises the fake-quant MSE
    (observer.py:412-536).  Per-channel 1-D searches (config 3 weights) run entirely on-chip
    (osq_mse_brent_rows_f32); per-tensor searches keep SciPy's bounded Brent on the host, with each
    loss evaluation a single masked pass over the activation on the GPU."""

    def __init__(self, bit=8, symmetric=False, ch_axis=-1):
        super().__init__(bit=bit, symmetric=symmetric, ch_axis=ch_axis)
        self.p = 2.0
        self.num = 100
        self._one_side = None      # 'pos', 'neg', 'no' once known on the host
        self._one_side_dev = None  # device int32[1]: -1 undecided / 0 no / 1 pos / 2 neg (decided by the search kernel)
        self._evals_dev = []       # device counters of the per-tensor searches (summed lazily by `loss_evals`)
        self._evals_host = 0

    # one_side_dist is decided on the first batch (observer.py:528-529).  The per-tensor search decides it on the device;
    # reading the attribute synchronises once and caches the answer.
    @property
    def one_side_dist(self):
        if self._one_side is None and self._one_side_dev is not None:
            v = int(self._one_side_dev.item())
            if v >= 0:
                self._one_side = ("no", "pos", "neg")[v]
        return self._one_side

    @one_side_dist.setter
    def one_side_dist(self, value):
        self._one_side = value
        if self._one_side_dev is not None:
            self._one_side_dev.fill_({None: -1, "no": 0, "pos": 1, "neg": 2}[value])

    @property
    def loss_evals(self):
        if self._evals_dev:
            self._evals_host += int(torch.stack(self._evals_dev).sum().item())
            self._evals_dev = []
        return self._evals_host

    @loss_evals.setter
    def loss_evals(self, value):
        self._evals_host, self._evals_dev = int(value), []

    host_search = False  # True: SciPy on the host drives the per-tensor search (round-1 path, kept as a cross-check)

    # ---- loss on the GPU (observer.py:420-432) ----
    def _loss(self, x, mask, seq_pos, new_min, new_max):
        new_min = new_min if torch.is_tensor(new_min) else torch.tensor(new_min)
        new_max = new_max if torch.is_tensor(new_max) else torch.tensor(new_max)
        scale, zero_point = _qparams_host(new_min.cpu(), new_max.cpu(), self.quant_min, self.quant_max, self.symmetric)
        s = torch.tensor([float(scale)], dtype=torch.float32)       # `x / scale.item()` rounds the scalar to fp32
        z = torch.tensor([float(int(zero_point))], dtype=torch.float32)
        loss_sum, n_valid = ops.mse_multi(x, mask, seq_pos, s, z, self.quant_min, self.quant_max)
        self._evals_host += 1
        both = torch.stack([loss_sum[0], n_valid[0].double()]).tolist()   # one synchronisation per evaluation, not two
        return np.float32(both[0] / max(int(both[1]), 1))

    def _search_1d(self, x, mask, seq_pos, x_min, x_max):
        from scipy.optimize import minimize_scalar
        xrange = max(abs(x_min), x_max)

        def f(r):
            lo = 0.0 if self.one_side_dist == "pos" else -r
            hi = 0.0 if self.one_side_dist == "neg" else r
            return self._loss(x, mask, seq_pos, lo, hi)

        r = minimize_scalar(f, bounds=(min(0.1, 0.01 * xrange), xrange), method="Bounded").x
        return (0.0 if self.one_side_dist == "pos" else -r), (0.0 if self.one_side_dist == "neg" else r)

    def _search_2d(self, x, mask, seq_pos, x_min, x_max):
        from scipy.optimize import minimize_scalar
        span = float(self.quant_max - self.quant_min)
        t_min, t_max = torch.tensor(x_min, dtype=torch.float32), torch.tensor(x_max, dtype=torch.float32)

        def shift_loss(shift, xrange):
            new_min = max(0.0 - shift, t_min)
            new_max = min(xrange - shift, t_max)
            return self._loss(x, mask, seq_pos, new_min, new_max)

        def range_loss(xrange):
            d = xrange / span
            return minimize_scalar(shift_loss, args=(xrange,), bounds=(d * self.quant_min, d * self.quant_max),
                                   method="Bounded").fun

        total = float(np.float32(x_max) - np.float32(x_min))
        final_range = minimize_scalar(range_loss, bounds=(min(0.1, 0.01 * total), total), method="Bounded").x
        d = final_range / span
        final_shift = minimize_scalar(shift_loss, args=(final_range,), bounds=(d * self.quant_min, d * self.quant_max),
                                      method="Bounded").x
        return float(max(0.0 - final_shift, t_min)), float(min(final_range - final_shift, t_max))

    def _best_minmax(self, x, observation_mask, seq_pos):
        """(best_min, best_max) device tensors for this batch (observer.py:496-533)."""
        x = x.detach()
        dev = x.device
        if self.ch_axis == -1 and not self.host_search:
            # the whole (1-D or nested 2-D) bounded-Brent search in ONE cooperative launch, nothing read back
            if self._one_side_dev is None or self._one_side_dev.device != dev:
                init = {None: -1, "no": 0, "pos": 1, "neg": 2}[self._one_side]
                self._one_side_dev = torch.full((1,), init, dtype=torch.int32, device=dev)
            out4, evals = ops.mse_brent_tensor(x, observation_mask, seq_pos, self.quant_min, self.quant_max, self.symmetric,
                                               self._one_side_dev)
            self._evals_dev.append(evals)
            return out4[0], out4[1]
        if self.ch_axis == -1:  # first version (host-driven SciPy, one GPU pass + sync per evaluation): cross-check
            cur = ops.observe_minmax(x, observation_mask, seq_pos).tolist()
            x_min, x_max = cur
            if self.one_side_dist is None:
                self.one_side_dist = "pos" if x_min >= 0.0 else "neg" if x_max <= 0.0 else "no"
            if self.one_side_dist != "no" or self.symmetric:
                lo, hi = self._search_1d(x, observation_mask, seq_pos, x_min, x_max)
            else:
                lo, hi = self._search_2d(x, observation_mask, seq_pos, x_min, x_max)
            return (torch.tensor(lo, dtype=torch.float64, device=dev), torch.tensor(hi, dtype=torch.float64, device=dev))
        assert observation_mask is None
        y = _transform_to_ch_axis(x, self.ch_axis if self.ch_axis >= 0 else self.ch_axis + x.dim()).contiguous()
        if self.one_side_dist is None:
            cur = ops.observe_minmax(y, None, -1).tolist()
            self.one_side_dist = "pos" if cur[0] >= 0.0 else "neg" if cur[1] <= 0.0 else "no"
        if self.one_side_dist != "no" or self.symmetric:
            mins, maxs, ev = ops.mse_brent_rows(y, self.quant_min, self.quant_max, self.one_side_dist, want_evals=True)
            self._row_evals = ev
            return mins, maxs
        # asymmetric two-sided per-channel rows (no shipped config): host-driven 2-D search per row
        mins = torch.empty(y.shape[0], dtype=torch.float32, device=dev)
        maxs = torch.empty_like(mins)
        for ch in range(y.shape[0]):
            row = y[ch]
            rmin, rmax = ops.observe_minmax(row, None, -1).tolist()
            mins[ch], maxs[ch] = self._search_2d(row, None, -1, rmin, rmax)
        return mins, maxs

    def _observe(self, x, observation_mask, seq_pos, quantizer=None):
        best_min, best_max = self._best_minmax(x, observation_mask, seq_pos)
        self.min_val = torch.min(self.min_val.to(best_min.device), best_min)  # observer.py:535-536
        self.max_val = torch.max(self.max_val.to(best_max.device), best_max)
        return False


class AvgMSEFastObserver(MSEFastObserver):
    """observer.py:539-567."""

    def __init__(self, bit=8, symmetric=False, ch_axis=-1):
        super().__init__(bit=bit, symmetric=symmetric, ch_axis=ch_axis)
        self.cnt = 0
        assert self.ch_axis == -1

    def _observe(self, x, observation_mask, seq_pos, quantizer=None):
        best_min, best_max = self._best_minmax(x, observation_mask, seq_pos)
        # observer.py:556-566 without reading the state back: `first batch` (max_val still -inf) selected on the device
        first = torch.isinf(self.max_val.to(best_max.device)).reshape(())
        mn = torch.where(first, best_min, self.min_val.to(best_min.device) * self.cnt + best_min)
        mx = torch.where(first, best_max, self.max_val.to(best_max.device) * self.cnt + best_max)
        self.cnt += 1
        self.min_val = mn / self.cnt
        self.max_val = mx / self.cnt
        return False


# ------------------------------------------------------------------------------------------------
# Observers no shipped twc/minmax/mse config reaches on the hot path (out of
# the CUDA scope).  Registry keys must resolve, so they are kept as thin torch restatements that
# share the in-kernel pad removal where a reduction is all that is needed.
# ------------------------------------------------------------------------------------------------
def _valid_tokens(x, observation_mask, seq_pos):
    """[T, F] copy of the valid tokens (only used by the out-of-scope observers below)."""
    if observation_mask is None:
        return x
    rest = [d for d in range(x.dim()) if d != seq_pos]
    y = x.permute(rest[0], seq_pos, *rest[1:]).reshape(x.shape[rest[0]], x.shape[seq_pos], -1)
    lens = torch.as_tensor(observation_mask, device=x.device)[: y.shape[0]]
    keep = torch.arange(y.shape[1], device=x.device)[None, :] < lens[:, None]
    return y[: lens.numel()][keep]


class LSQPlusObserver(ObserverBase):
    """mean +- 3 std initialisation (observer.py:148-173); weights only."""

    def __init__(self, bit=8, symmetric=False, ch_axis=-1):
        super().__init__(bit=bit, symmetric=symmetric, ch_axis=ch_axis)
        assert self.symmetric is True
        self.mean = self.std = None

    def _observe(self, x, observation_mask, seq_pos, quantizer=None):
        x = x.detach().float()
        if self.ch_axis == -1:
            self.mean, self.std = x.mean(), x.std()
        else:
            y = _transform_to_ch_axis(x, self.ch_axis)
            self.mean, self.std = y.mean(1), y.std(1)
        self.min_val = self.mean - 3 * self.std
        self.max_val = self.mean + 3 * self.std
        return False


class AvgQuantileObserver(ObserverBase):
    """Histogram-quantile clipping averaged over batches (observer.py:240-282): masked min/max, the |x| histogram, the
    cumulative-threshold clip and the running average run on the device in two launches (osq_quantile_observe_f32)."""

    def __init__(self, bit=8, symmetric=False, ch_axis=-1, ema_ratio=0.9, threshold=0.99999, bins=2048):
        super().__init__(bit=bit, symmetric=symmetric, ch_axis=ch_axis)
        assert self.ch_axis == -1, "Quantile observer only support in per-tensor scheme."
        self.ema_ratio, self.threshold, self.bins, self.cnt = ema_ratio, threshold, bins, 0

    def _observe(self, x, observation_mask, seq_pos, quantizer=None):
        self._ensure_scalar_state(x.device)
        s_out, z_out = self._fused_targets(quantizer)
        ops.observe_quantile(x.detach(), observation_mask, seq_pos, self.bins, self.threshold, mode=ops.STAT_AVERAGE,
                             cnt=self.cnt, state_min=self.min_val, state_max=self.max_val, scale_out=s_out, zp_out=z_out,
                             qmin=self.quant_min, qmax=self.quant_max, symmetric=self.symmetric)
        self.cnt += 1
        return s_out is not None


class MSEObserver(ObserverBase):
    """Exhaustive 100-candidate (x 2^bit zero points) MSE grid (observer.py:285-383): candidates are
    evaluated eight per pass by the multi-candidate kernel instead of one full pass each."""

    def __init__(self, bit=8, symmetric=False, ch_axis=-1):
        super().__init__(bit=bit, symmetric=symmetric, ch_axis=ch_axis)
        self.p, self.num, self.one_side_dist = 2.0, 100, None

    def _grid(self, x, observation_mask, seq_pos):
        assert self.ch_axis == -1, "per-channel MSEObserver is outside the CUDA scope"
        x_min, x_max = ops.observe_minmax(x, observation_mask, seq_pos).cpu().unbind()
        if self.one_side_dist is None:
            self.one_side_dist = "pos" if x_min >= 0.0 else "neg" if x_max <= 0.0 else "no"
        zero = torch.zeros_like(x_min)
        cands = []
        if self.one_side_dist != "no" or self.symmetric:  # observer.py:348-366
            xrange = torch.max(x_min.abs(), x_max)
            for i in range(1, self.num + 1):
                thres = xrange / self.num * i
                cands.append((zero if self.one_side_dist == "pos" else -thres, zero if self.one_side_dist == "neg" else thres))
        else:  # observer.py:318-346
            xrange = x_max - x_min
            for i in range(1, self.num + 1):
                tmp_max = xrange / self.num * i
                delta = tmp_max / float(self.quant_max - self.quant_min)
                for zp in range(self.quant_min, self.quant_max + 1):
                    cands.append((torch.max(zero - zp * delta, x_min), torch.min(tmp_max - zp * delta, x_max)))
        mins = torch.stack([c[0] for c in cands])
        maxs = torch.stack([c[1] for c in cands])
        scale, zp = _qparams_host(mins, maxs, self.quant_min, self.quant_max, self.symmetric)
        loss, _ = ops.mse_multi(x, observation_mask, seq_pos, scale.float(), zp.float(), self.quant_min, self.quant_max)
        best = int(torch.argmin(loss))  # first minimum, like the strict `<` update of the reference
        return mins[best].to(x.device), maxs[best].to(x.device)

    def _observe(self, x, observation_mask, seq_pos, quantizer=None):
        self._ensure_scalar_state(x.device)
        lo, hi = self._grid(x.detach(), observation_mask, seq_pos)
        self.min_val = torch.min(self.min_val, lo)
        self.max_val = torch.max(self.max_val, hi)
        return False


class AvgMSEObserver(MSEObserver):
    """observer.py:386-409."""

    def __init__(self, bit=8, symmetric=False, ch_axis=-1):
        super().__init__(bit=bit, symmetric=symmetric, ch_axis=ch_axis)
        self.cnt = 0
        assert self.ch_axis == -1

    def _observe(self, x, observation_mask, seq_pos, quantizer=None):
        self._ensure_scalar_state(x.device)
        lo, hi = self._grid(x.detach(), observation_mask, seq_pos)
        if bool(self.max_val.isinf()):
            self.min_val, self.max_val = lo, hi
        else:
            self.min_val = self.min_val * self.cnt + lo
            self.max_val = self.max_val * self.cnt + hi
        self.cnt += 1
        self.min_val = self.min_val / self.cnt
        self.max_val = self.max_val / self.cnt
        return False
