"""Per-launch time and HBM-roofline fraction of the fused kernel for arbitrary (M, K, N) sites.

    python scripts/bench_shapes.py                       # BART-large sites of BASELINE.json config 4 (M = 4 x 1024 tokens)
    python scripts/bench_shapes.py 16384x768x768 ...     # any list of MxKxN

24-launch CUDA-graph chains over rotating buffers larger than L2 (or at least 3 buffers); 6-bit LSQ+ activations,
6-bit symmetric per-channel weights.  Prints one JSON object."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from outlier_suppression_b200 import ops

PEAK = 6650.0
DEFAULT = ["4096x1024x3072", "4096x1024x1024", "4096x1024x4096", "4096x4096x1024", "16384x1024x1024", "16384x1024x4096", "16384x4096x1024"]


def main():
    shapes = [tuple(int(v) for v in a.split("x")) for a in (sys.argv[1:] or DEFAULT)]
    torch.manual_seed(0)
    a_scale = torch.tensor([0.1], device="cuda"); a_zp = torch.tensor([31.0], device="cuda")
    out = {}
    for M, K, N in shapes:
        nbuf = max(3, int(140e6 // (M * K * 4)) + 1)
        nbuf_o = max(3, int(140e6 // (M * N * 4)) + 1)
        acts = [torch.randn(M, K, device="cuda") for _ in range(min(nbuf, 12))]
        outs = [torch.empty(M, N, device="cuda") for _ in range(min(nbuf_o, 12))]
        w = torch.randn(N, K, device="cuda") * 0.05
        ws = (w.abs().amax(1) / 31.5).contiguous(); wz = torch.zeros(N, dtype=torch.int32, device="cuda")
        codes, rowsum = ops.pack_weight(w, ws, wz, -32, 31)
        bias = torch.randn(N, device="cuda")
        i = [0]

        def launch():
            j = i[0]; i[0] += 1
            ops.fused_fq_linear(acts[j % len(acts)], a_scale, a_zp, 0, 63, codes, ws, rowsum, bias, lsq_grad_factor=1e-4,
                                out=outs[j % len(outs)])
        launch(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(24):
                launch()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(5):
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 24 * 1e3)
        us = sorted(ts)[2]
        by = 4 * M * K + 4 * M * N + N * K + 12 * N
        out["%dx%dx%d" % (M, K, N)] = {"us": round(us, 2), "GBs": round(by / us / 1e3, 1), "frac_of_%.0f" % PEAK: round(by / us / 1e3 / PEAK, 3),
                                       "tokens_per_s": round(M / us * 1e6)}
        del acts, outs
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
