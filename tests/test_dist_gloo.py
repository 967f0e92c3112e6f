"""world_size-2 gloo test (CPU) of the rank-sharded calibration protocol (dist.py): batches dealt
round-robin, one packed all-reduce, recurrence replayed in batch order => bit-identical to the
sequential CPU oracle."""
import os
import socket

import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from oracle import osq_oracle as O

N_OBS, N_BATCH = 5, 7


def _batch(b, o):
    g = torch.Generator().manual_seed(1000 * b + o)
    x = torch.randn(3, 16, 24, generator=g) * (1 + o)
    lens = [16, 5 + b % 4, 9]
    return x, lens


def _sequential():
    out = []
    for o in range(N_OBS):
        st = O.ObserverState()
        for b in range(N_BATCH):
            x, lens = _batch(b, o)
            O.observe_avg_prune_minmax(st, x, 0.9, "x", lens, 1)
        out.append((st.min_val, st.max_val))
    return torch.tensor([[float(a), float(c)] for a, c in out])


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    from outlier_suppression_b200.dist import SlotTable, my_batches
    table = SlotTable(N_OBS, N_BATCH)
    mine = my_batches(N_BATCH, rank, world)
    for b in mine:
        for o in range(N_OBS):
            x, lens = _batch(b, o)
            lo, hi = O.prune_minmax(O.token_matrix(x, lens, 1), 0.9)  # stands in for the per-batch kernel result
            table.slot(o, b).copy_(torch.stack([lo, hi]))
    table.all_reduce()
    mn, mx, cnt = table.replay()
    ret[rank] = (torch.stack([mn, mx], 1), cnt, mine)
    tdist.destroy_process_group()


def test_sharded_calibration_is_bit_identical_to_sequential():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    ref = _sequential()
    assert sorted(ret[0][2] + ret[1][2]) == list(range(N_BATCH))
    for r in (0, 1):
        got, cnt, _ = ret[r]
        assert cnt == N_BATCH
        assert torch.equal(got, ref), (got, ref)


def test_replay_continues_previous_state():
    from outlier_suppression_b200.dist import SlotTable
    vals = torch.tensor([[[-1.0, 2.0], [-3.0, 5.0], [-0.5, 0.25]]])
    t = SlotTable(1, 3); t.buf.copy_(vals)
    mn, mx, cnt = t.replay()
    st = O.ObserverState()
    for b in range(3):
        O.running_average(st, vals[0, b, 0], vals[0, b, 1])
    assert float(mn) == float(st.min_val) and float(mx) == float(st.max_val) and cnt == 3
    t2 = SlotTable(1, 1); t2.buf.copy_(torch.tensor([[[-7.0, 7.0]]]))
    mn2, mx2, cnt2 = t2.replay(cnt, mn, mx)
    O.running_average(st, torch.tensor(-7.0), torch.tensor(7.0))
    assert float(mn2) == float(st.min_val) and float(mx2) == float(st.max_val) and cnt2 == 4
