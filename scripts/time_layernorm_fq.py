"""K7 (residual + LayerNorm + quantizer + bins in one pass) against the three launches it replaces, [16384, 768] and [16384, 1024]."""
import json, os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from outlier_suppression_b200 import ops

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, chain=12, reps=10):
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


res = {"peak_gbs": PEAK}
for M, H in ((16384, 768), (16384, 1024)):
    nb = 4   # 4 x (2 x 50 MB in, 50 + 12.6 MB out): rotates past the 126 MB L2
    hs = [torch.randn(M, H, device="cuda") for _ in range(nb)]
    rs = [torch.randn(M, H, device="cuda") for _ in range(nb)]
    gm = torch.rand(H, device="cuda") + 0.5
    bias = torch.randn(H, device="cuda") * 0.1
    sc, zp = torch.tensor([0.1], device="cuda"), torch.tensor([31.0], device="cuda")
    g = 1.0 / (M * H * 63) ** 0.5
    fused = timeit(lambda i: ops.residual_layernorm_fq(hs[i % nb], rs[i % nb], gm, None, bias, 1e-12, sc, zp, 0, 63, lsq_grad_factor=g, want_bins=True))

    def pieces(i):
        u = rs[i % nb] * gm + hs[i % nb]
        ln = torch.nn.functional.layer_norm(u, (H,), None, None, 1e-12)
        ln += bias
        return ops.fq_per_tensor(ln, sc, zp, 0, 63, lsq_grad_factor=g, want_bins=True)
    unfused = timeit(pieces)
    by = 13 * M * H
    res["%dx%d" % (M, H)] = {"fused_us": fused, "unfused_us": unfused, "algorithmic_bytes": by, "gbs": by / fused / 1e3,
                             "frac_of_hbm_peak": by / fused / 1e3 / PEAK}
print(json.dumps(res, indent=1))
json.dump(res, open("gpurun_out/layernorm_fq.json", "w"), indent=1)
