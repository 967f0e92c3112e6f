"""Token-wise clipping, coarse stage, with every model forward replayed from CUDA graphs.

The reference's ``solver/token_wise_clipping.py:50-66`` (``find_ratio``) walks ``iters`` clipping ratios; for each one it
re-calibrates every activation observer over all calibration batches (``set_ratio`` + ``calibrate``), switches to the
quantized state (``enable_quantization``) and accumulates the MSE between the quantized and the FP logits over the same
batches -- 120 iterations x 2 x 8 forwards of the whole model at seq 512, all of it issued from Python: ~100 observer
calls and as many quantizer / Linear launches per forward.

Nothing in that loop changes shape or control flow from one ratio to the next; only the ratio itself changes.  Here the
ratio is DATA: every ``AvgPruneMinMaxObserver`` reads it from a device scalar (``percentile_dev`` of
``osq_prune_observe_f32``), so one calibration forward per batch and one quantized forward + loss per batch are captured
once (2 x n_batches graphs sharing one memory pool) and replayed for every ratio.  The batch index fixes the running-average
count that is baked into a calibration graph (``cnt`` = position of the batch), exactly as the eager loop would pass it.

Results are bit-identical to the eager loop on this backend (``tests/test_gpu_twc.py``): same per-ratio losses, same best
ratio, same final ``(scale, zero_point)`` of every quantizer.

Beyond the graphs (``cache_vectors=True``, the default): ``set_ratio`` switches activation fake-quant OFF for the calibration
forwards (token_wise_clipping.py:12-19), so those forwards compute the same activations for every ratio -- only the quantile
the observers cut at moves.  The per-token extrema of every (observer, batch) pair (and the plain min / max of the
observers that do not prune: ``attention_probs``, pooled inputs) are therefore recorded ONCE, and "re-calibrating at another
ratio" becomes one launch of the cached select over all (observer, batch) problems (osq_prune_select_cached_f32) plus one replay
launch of the running averages and qparams (osq_replay_average_f32) instead of a model forward per batch.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch

from .quantization.fake_quant import QuantizeBase


def _logits(out):
    return out[0] if isinstance(out, (tuple, list)) else out.logits


class GraphedFindRatio:
    """``find_ratio`` of token_wise_clipping.py:50-66 over CUDA graphs.

    model      the quantized model (the reference's model classes on this backend), on CUDA, in eval mode
    fp_input   list of keyword-argument dicts of device tensors (the calibration batches; kept alive and static)
    fp_output  list of FP-model logits, one per batch (token_wise_clipping.calibrate's ``fp_output``, task_type 'glue')
    """

    def __init__(self, model, fp_input: Sequence[Dict[str, torch.Tensor]], fp_output: Sequence[torch.Tensor],
                 loss_fn: Optional[Callable] = None, cache_vectors: bool = True):
        self.model = model
        self.batches = list(fp_input)
        self.targets = [t.detach() for t in fp_output]
        self.loss_fn = loss_fn or torch.nn.MSELoss()
        self.act_q = [(n, m) for n, m in model.named_modules() if isinstance(m, QuantizeBase) and "act" in n]
        if not self.act_q:
            raise ValueError("no activation quantizers found")
        dev = self.targets[0].device
        if dev.type != "cuda":
            raise RuntimeError("GraphedFindRatio needs the model on a CUDA device (no CPU fallback)")
        for _, q in self.act_q:
            q.observer._percentile_dev = torch.full((1,), float(getattr(q.observer, "percentile", 1.0)), dtype=torch.float32, device=dev)
        self._cal: List[torch.cuda.CUDAGraph] = []
        self._quant: List[torch.cuda.CUDAGraph] = []
        self._loss: List[torch.Tensor] = []
        self._pool = None
        self._use_cache = bool(cache_vectors)
        self._cache = None   # recorded per-token vectors + slot table + replay targets (cache_vectors)

    # ---- the reference's two state switches (token_wise_clipping.py:12-26), verbatim in effect ----
    def set_ratio(self, ratio: float) -> None:
        for _, q in self.act_q:
            q.observer.set_percentile(ratio)   # also refreshes the device scalar the graphs read
            q.observer.cnt = 0
            q.disable_fake_quant()
            q.enable_observer()

    def enable_quantization(self) -> None:
        for _, q in self.act_q:
            q.disable_observer()
            q.enable_fake_quant()

    # ---- eager forms (warm-up, and the cross-check of the tests) ----
    def calibrate_eager(self) -> None:
        with torch.no_grad():
            for b in self.batches:
                self.model(**b)

    def loss_eager(self) -> torch.Tensor:
        loss = 0
        with torch.no_grad():
            for b, t in zip(self.batches, self.targets):
                loss = loss + self.loss_fn(_logits(self.model(**b)), t)
        return loss

    def capture(self, ratio: float = 1.0) -> None:
        """Two eager warm-up rounds (lazy caches: packed weights, bins hand-off, workspaces), then the capture."""
        for _ in range(2):
            self.set_ratio(ratio)
            self.calibrate_eager()
            self.enable_quantization()
            self.loss_eager()
        torch.cuda.synchronize()
        self._pool = torch.cuda.graph_pool_handle()
        self._cal, self._quant, self._loss = [], [], []
        self.set_ratio(ratio)
        self._cache = self._record_vectors() if self._use_cache else None
        with torch.no_grad():
            if self._cache is None:
                for b in self.batches:                 # observer.cnt advances while capturing: graph i carries cnt = i
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=self._pool):
                        self.model(**b)
                    self._cal.append(g)
            self.enable_quantization()
            for b, t in zip(self.batches, self.targets):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=self._pool):
                    loss = self.loss_fn(_logits(self.model(**b)), t)
                self._quant.append(g)
                self._loss.append(loss)
        torch.cuda.synchronize()

    def evaluate(self, ratio: float) -> torch.Tensor:
        """calibrate at ``ratio`` + quantized loss, all replays; returns the summed loss as a device scalar."""
        if not self._quant:
            self.capture(ratio)
        self._calibrate(ratio)
        for g in self._quant:
            g.replay()
        total = self._loss[0].clone()
        for l in self._loss[1:]:
            total = total + l
        return total

    def find_ratio(self, iters: int, step: float):
        """token_wise_clipping.py:50-66.  Returns (best ratio, [loss per iteration])."""
        best_i, best = 0, 10000000.0
        losses = []
        for i in range(iters):
            cur = float(self.evaluate(1.0 - step * i))    # one synchronisation per ratio (the reference: one per batch)
            losses.append(cur)
            if best > cur:
                best, best_i = cur, i
        ratio = 1.0 - step * best_i
        # final calibration at the best ratio (:64-65); the quantizers are left in the calibration state like the reference's
        self.set_ratio(ratio)
        self._calibrate(ratio)
        for _, q in self.act_q:
            q.observer.cnt = len(self.batches)
            q.qparam_epoch += 1
        return ratio, losses

    # ---- calibration at one ratio: replayed forwards, or (cache_vectors) the cached select + one replay launch ----
    def _calibrate(self, ratio: float) -> None:
        for _, q in self.act_q:
            q.observer.set_percentile(ratio)
        c = self._cache
        if c is None:
            for g in self._cal:
                g.replay()
            return
        from . import ops
        if c["n"]:
            ops.prune_select_cached(c["problems"], c["n"], float(ratio), percentile_dev=self.act_q[0][1].observer._percentile_dev)
        ops.replay_average(c["table"], 0, c["targets"])     # set_ratio restarted the running averages: cnt0 = 0

    def _record_vectors(self):
        """One eager calibration pass that keeps, per (observer, batch), what the observer's result depends on besides the ratio.
        Returns None (-> replayed forwards) when an observer cannot be cached: another observer class, a call pattern other than
        once per batch, or a token count beyond the cached select."""
        from . import ops
        from .quantization.observer import AvgPruneMinMaxObserver
        obs = [q.observer for _, q in self.act_q]
        if not all(type(o) is AvgPruneMinMaxObserver for o in obs):
            return None
        for o in obs:
            o._twc_record = []
        try:
            self.calibrate_eager()
        finally:
            recs = [o._twc_record for o in obs]
            for o in obs:
                o._twc_record = None
        n_b = len(self.batches)
        if any(len(r) != n_b or any(payload is None for _, payload in r) for r in recs):
            return None
        dev = self.targets[0].device
        table = torch.zeros(len(obs), n_b, 2, dtype=torch.float32, device=dev)
        problems = []
        for i, r in enumerate(recs):
            for b, (kind, payload) in enumerate(r):
                if kind == "plain":
                    table[i, b].copy_(payload.reshape(2))
                else:
                    problems.append((payload, table[i, b]))
        entries = []
        for (_, q), o in zip(self.act_q, obs):
            o._ensure_scalar_state(dev)
            s_out, z_out = q._per_tensor_qparam_targets()
            entries.append((o.min_val, o.max_val, s_out, z_out, o.quant_min, o.quant_max, o.symmetric))
        return {"table": table, "n": len(problems), "problems": ops.select_problems(problems, dev) if problems else None,
                "targets": ops.replay_targets(entries, dev), "keep": recs}
