"""Device time of ONE AvgPruneMinMax observer call at the BERT-base seq512 shape [32, 512, 768] (BASELINE config 2; the call
token-wise clipping makes ~94 k times), new two-launch path vs the earlier select paths, plus the kernels' shares."""
import json, os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from outlier_suppression_b200 import ops

def timeit(fn, reps=50, chain=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(chain):
            fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / chain * 1e3)
    return statistics.median(ts)

out = {}
for shape in ((32, 512, 768), (32, 512, 3072), (32, 128, 768), (32, 2048, 4096)):
    B, S, F = shape
    xs = [torch.randn(B, S, F, device="cuda") for _ in range(max(2, int(300e6 // (B * S * F * 4)) + 1))]
    lens = torch.randint(S // 4, S + 1, (B,), device="cuda"); lens[0] = S
    i = [0]
    def nxt():
        i[0] += 1
        return xs[i[0] % len(xs)]
    rec = {}
    rec["prune_observe_us"] = timeit(lambda: ops.observe_prune_minmax(nxt(), lens, 1, 0.99))
    ws = ops.workspace(xs[0].device)
    torch.cuda.synchronize()
    off = ws.numel() - 32 * 8
    st = ws[off:].view(torch.int64).cpu().tolist()
    names = ["tm_start", "tm_end(cta0)", "tail_enter", "tail_wait_done", "a_done", "b_done", "c_done", "c2_done", "end"]
    rec["trace_ns_since_tm_start"] = {n: st[i] - st[0] for i, n in enumerate(names)}
    rec["legacy_select_us"] = timeit(lambda: ops.observe_prune_minmax(nxt(), lens, 1, 0.99, legacy_select=True))
    rec["token_minmax_only_us"] = timeit(lambda: ops.token_minmax(nxt(), lens, 1))
    rec["avg_minmax_us"] = timeit(lambda: ops.observe_minmax(nxt(), lens, 1))
    valid = float(lens.sum()) / (B * S)
    rec["one_read_us_at_6534GBs"] = valid * B * S * F * 4 / 6534.5e9 * 1e6
    out["x".join(map(str, shape))] = rec
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/observer_call.json", "w"), indent=1)
