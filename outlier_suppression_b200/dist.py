"""Rank-sharded calibration: one process per GPU, calibration batches dealt round-robin to ranks,
ONE packed all-reduce per calibration pass.

The reference has no distributed code at all (single V100); its only cross-batch coupling is the
running average ``m <- (m*cnt + cur)/(cnt+1)`` of the Avg* observers (observer.py:194-202).  Here
every rank writes the per-batch (min, max) of every observer into a zero-initialised slot table
``[n_observers, n_batches, 2]`` at ITS batch indices, the table is summed across ranks over
NCCL/NVLink (x + 0 is exact), and every rank then replays the recurrence in batch order -- so the
result is bit-identical to the 1-GPU / CPU-oracle result for any number of ranks.
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import List, Optional

import torch
import torch.distributed as tdist


class SlotTable:
    """[n_obs, n_batches, 2] fp32 per-batch statistics + the packed collective + the replay."""

    def __init__(self, n_obs: int, n_batches: int, device="cpu", peer: bool = False, group=None):
        """peer=True (CUDA, NCCL world > 1): every rank owns a CUDA symmetric-memory region mapped by all ranks over NVLink, and the
        exchange happens INSIDE the replay launch -- publish own slots, one remote flag store per peer, wait for every peer's
        flag, peer loads -- instead of an all-reduce (falls back to the all-reduce if symmetric memory cannot be set up)."""
        self.n_obs, self.n_batches = n_obs, n_batches
        self.hdl = self.peer_ptrs = None
        self.world = 1
        dev = torch.device(device)
        if peer and dev.type == "cuda" and tdist.is_available() and tdist.is_initialized() and tdist.get_world_size(group) > 1:
            try:
                import torch.distributed._symmetric_memory as symm
                self.world = tdist.get_world_size(group)
                self.rank = tdist.get_rank(group)
                n_f = n_obs * n_batches * 2
                # this rank's region: two generations of the slot table + one flag per rank (see osq_replay_exchange_f32)
                flat = symm.empty(2 * n_f + self.world, dtype=torch.float32, device=dev)
                flat.zero_()
                self.hdl = symm.rendezvous(flat, group if group is not None else tdist.group.WORLD)
                self.region = flat
                self.peer_ptrs = torch.tensor([int(p) for p in self.hdl.buffer_ptrs], dtype=torch.int64, device=dev)
                self.pass_counter = torch.zeros(1, dtype=torch.int32, device=dev)
                self.err_flag = torch.zeros(1, dtype=torch.int32, device=dev)
                self.hdl.barrier(channel=0)          # once: every region is zeroed before anybody signals into it
                self.buf = torch.zeros(n_obs, n_batches, 2, dtype=torch.float32, device=dev)   # private: the observers write here
                return
            except Exception as ex:  # pragma: no cover  (no P2P / fabric support: the collective path is always there)
                self.hdl = self.peer_ptrs = None
                self.world = 1
                self.peer_error = repr(ex)
        self.buf = torch.zeros(n_obs, n_batches, 2, dtype=torch.float32, device=device)

    def slot(self, obs: int, batch: int) -> torch.Tensor:
        return self.buf[obs, batch]

    def all_reduce(self, group=None) -> None:
        """the single collective of a calibration pass (6.3 kB for BERT-base x 8 batches: latency bound)"""
        if tdist.is_available() and tdist.is_initialized() and tdist.get_world_size(group) > 1:
            tdist.all_reduce(self.buf, op=tdist.ReduceOp.SUM, group=group)

    def replay(self, cnt0: int = 0, state_min: Optional[torch.Tensor] = None, state_max: Optional[torch.Tensor] = None):
        """observer.py:194-202 replayed in batch order on the host in fp32 (true division), for all
        observers at once.  ``state_*`` ([n_obs], may hold +-inf = `no batch seen yet`) and ``cnt0``
        continue an earlier pass.  Returns (min[n_obs], max[n_obs], cnt)."""
        t = self.buf.detach().cpu()
        mn = torch.full((self.n_obs,), float("inf")) if state_min is None else state_min.detach().cpu().float().clone()
        mx = torch.full((self.n_obs,), float("-inf")) if state_max is None else state_max.detach().cpu().float().clone()
        cnt = cnt0
        for b in range(self.n_batches):
            cur_min, cur_max = t[:, b, 0], t[:, b, 1]
            first = torch.isinf(mx)
            mn = torch.where(first, cur_min, mn * cnt + cur_min)
            mx = torch.where(first, cur_max, mx * cnt + cur_max)
            cnt += 1
            mn = mn / cnt
            mx = mx / cnt
        return mn, mx, cnt


def my_batches(n_batches: int, rank: int, world: int) -> List[int]:
    """batch i belongs to rank i mod world; a single batch is never split (the prune quantile is per batch)."""
    return list(range(rank, n_batches, world))


class _Controller:
    def __init__(self, table: SlotTable, observers, owners):
        self.table, self.observers, self.owners = table, observers, owners
        self.batch = 0

    def set_batch(self, batch_index: int) -> None:
        self.batch = int(batch_index)


def _avg_observers(model):
    from .quantization.fake_quant import QuantizeBase
    from .quantization.observer import AvgMinMaxObserver, AvgPruneMinMaxObserver
    obs, owners = [], []
    for m in model.modules():
        if isinstance(m, QuantizeBase) and isinstance(m.observer, (AvgMinMaxObserver, AvgPruneMinMaxObserver)) \
                and m.observer_enabled == 1:
            obs.append(m.observer)
            owners.append(m)
    return obs, owners


def _targets(table, entries, device):
    """device copy of the replay pointer table, rebuilt only when a pointer changed (one small upload per model, not per pass)"""
    from . import ops
    key = tuple((None if t is None else t.data_ptr()) for e in entries for t in e[:4]) + tuple(e[4:] for e in entries)
    cached = getattr(table, "_targets_cache", None)
    if cached is None or cached[0] != key:
        cached = (key, ops.replay_targets(entries, device))
        table._targets_cache = cached
    return cached[1]


@contextmanager
def sharded_calibration(model, n_batches: int, group=None, table: Optional[SlotTable] = None):
    """Usage (every rank):

        with sharded_calibration(model, n_batches) as ctl:
            for i in my_batches(n_batches, rank, world):
                ctl.set_batch(i); model(**batch[i])

    On exit the slot table is all-reduced once and every enabled Avg* observer (and its quantizer's
    scale / zero_point) holds exactly what a sequential pass over all batches would have produced.
    ``table``: an existing SlotTable of the right shape to reuse (zeroed in place) -- keeps the slot addresses stable,
    so the per-batch launches of a pass can be captured once in a CUDA graph and replayed."""
    from . import ops
    observers, owners = _avg_observers(model)
    device = next((o.min_val.device for o in observers), torch.device("cpu"))
    if table is None:
        table = SlotTable(len(observers), n_batches, device)
    else:
        assert table.n_obs == len(observers) and table.n_batches == n_batches and table.buf.device == torch.device(device)
        if table.hdl is None:   # the all-reduce sums: slots of other ranks must be zero.  The peer exchange publishes own slots only.
            table.buf.zero_()
    ctl = _Controller(table, observers, owners)
    for i, o in enumerate(observers):
        o._shard = (ctl, i)
    try:
        yield ctl
    finally:
        for o in observers:
            o._shard = None
    if not observers:
        return
    cnt0 = observers[0].cnt
    if table.hdl is not None:
        # no collective and no host-side barrier: ONE launch publishes this rank's slots, signals and awaits its peers over NVLink
        # and replays the recurrence with every slot loaded from its owner's region (osq_replay_exchange_f32)
        entries = []
        for o, q in zip(observers, owners):
            o._ensure_scalar_state(device)
            s_out, z_out = q._per_tensor_qparam_targets()
            entries.append((o.min_val, o.max_val, s_out, z_out, o.quant_min, o.quant_max, o.symmetric))
        ops.replay_exchange(table.buf, table.peer_ptrs, table.rank, table.world, table.n_obs, table.n_batches, cnt0,
                            _targets(table, entries, device), table.pass_counter, table.err_flag)
        for o, q in zip(observers, owners):
            o.cnt = cnt0 + n_batches
            q.qparam_epoch += 1
        return
    table.all_reduce(group)
    if table.buf.is_cuda:
        # ONE launch replays the recurrence for every observer and rewrites every quantizer's (scale, zero_point)
        # through a pointer table: no .cpu(), no per-observer launches, no synchronisation
        entries = []
        for o, q in zip(observers, owners):
            o._ensure_scalar_state(device)
            s_out, z_out = q._per_tensor_qparam_targets()
            entries.append((o.min_val, o.max_val, s_out, z_out, o.quant_min, o.quant_max, o.symmetric))
        ops.replay_average(table.buf, cnt0, _targets(table, entries, device))
        for o, q in zip(observers, owners):
            o.cnt = cnt0 + n_batches
            q.qparam_epoch += 1
        return
    # host tensors (gloo world, CPU-side tests of the sharding logic): same recurrence in torch ops
    state_min = torch.stack([o.min_val.detach().float().reshape(()) for o in observers])
    state_max = torch.stack([o.max_val.detach().float().reshape(()) for o in observers])
    mn, mx, cnt = table.replay(cnt0, state_min, state_max)
    for i, (o, q) in enumerate(zip(observers, owners)):
        o.min_val = mn[i].clone()
        o.max_val = mx[i].clone()
        o.cnt = cnt
        scale, zp = o.calculate_qparams(o.min_val, o.max_val)
        s_out, z_out = q._per_tensor_qparam_targets()
        s_out.copy_(scale.reshape(s_out.shape))
        z_out.copy_(zp.reshape(z_out.shape).to(z_out.dtype))
        q.qparam_epoch += 1
