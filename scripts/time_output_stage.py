"""Per-launch device time of the fused Linear with and without the output stage (next quantizer, optional GELU) against the
unfused pieces, BERT-base FFN-up shape (M = 16384, 768 -> 3072)."""
import json, os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from outlier_suppression_b200 import ops

M, K, N = 16384, 768, 3072
torch.manual_seed(0)
acts = [torch.randn(M, K, device="cuda") for _ in range(3)]
w = torch.randn(N, K, device="cuda") * 0.05
ws = (w.abs().amax(1) / 31.5).clamp_min(1e-8)
codes, rowsum = ops.pack_weight(w, ws, torch.zeros(N, dtype=torch.int32, device="cuda"), -32, 31)
bias = torch.zeros(N, device="cuda")
sc, zp = torch.tensor([0.12], device="cuda"), torch.tensor([31.0], device="cuda")
osc, ozp = torch.tensor([0.05], device="cuda"), torch.tensor([3.0], device="cuda")
ys = [torch.empty(M, N, device="cuda") for _ in range(2)]
g_out = 1.0 / (M * N * 63) ** 0.5

def timeit(fn, chain=6, reps=10):
    fn(0); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(chain):
            fn(i)
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / chain * 1e3)
    return statistics.median(ts)

res = {}
res["linear_us"] = timeit(lambda i: ops.fused_fq_linear(acts[i % 3], sc, zp, 0, 63, codes, ws, rowsum, bias, lsq_grad_factor=1e-4, out=ys[i % 2]))
for act in (None, "gelu"):
    res["linear+fq(%s)_us" % act] = timeit(lambda i: ops.fused_fq_linear(acts[i % 3], sc, zp, 0, 63, codes, ws, rowsum, bias, lsq_grad_factor=1e-4, out=ys[i % 2],
                                           out_q=dict(scale=osc, zp=ozp, qmin=0, qmax=63, g=g_out, act=act, bins=True)))
    res["linear+fq(%s)_nobins_us" % act] = timeit(lambda i: ops.fused_fq_linear(acts[i % 3], sc, zp, 0, 63, codes, ws, rowsum, bias, lsq_grad_factor=1e-4, out=ys[i % 2],
                                           out_q=dict(scale=osc, zp=ozp, qmin=0, qmax=63, g=g_out, act=act, bins=False)))
res["gelu_us"] = timeit(lambda i: torch.nn.functional.gelu(ys[i % 2]))
res["gelu+fq_bins_one_pass_us"] = timeit(lambda i: ops.fq_per_tensor(ys[i % 2], osc, ozp, 0, 63, lsq_grad_factor=g_out, want_bins=True, act="gelu"))
res["gelu+fq_bins_only_us"] = timeit(lambda i: ops.fq_bins_only(ys[i % 2], osc, ozp, 0, 63, lsq_grad_factor=g_out, act="gelu"))
res["fq_bins_only_us"] = timeit(lambda i: ops.fq_bins_only(ys[i % 2], osc, ozp, 0, 63, lsq_grad_factor=g_out))
res["fq_nobins_us"] = timeit(lambda i: ops.fq_per_tensor(ys[i % 2], osc, ozp, 0, 63, lsq_grad_factor=g_out))
res["fq_bins_us"] = timeit(lambda i: ops.fq_per_tensor(ys[i % 2], osc, ozp, 0, 63, lsq_grad_factor=g_out, want_bins=True))
print(json.dumps(res, indent=1))
json.dump(res, open("gpurun_out/output_stage.json", "w"), indent=1)
