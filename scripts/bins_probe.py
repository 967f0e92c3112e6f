"""Per-site time of the fused kernel fed with fp32 activations vs with the upstream fake-quant kernel's uint8 bins,
and of the fake-quant kernel with / without the bins side output (24-launch CUDA-graph chains, M = 16384)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from outlier_suppression_b200 import ops

torch.manual_seed(0)
M = 16384
a_scale = torch.tensor([0.1], device="cuda"); a_zp = torch.tensor([31.0], device="cuda")
res = {}


def chain_time(fn, n=24, reps=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / n * 1e3)
    return sorted(ts)[len(ts) // 2]


for K, N in ((768, 2304), (768, 768), (768, 3072), (3072, 768)):
    acts = [torch.randn(M, K, device="cuda") for _ in range(4 if K == 768 else 2)]
    outs = [torch.empty(M, N, device="cuda") for _ in range(3)]
    w = torch.randn(N, K, device="cuda") * 0.05
    ws = (w.abs().amax(1) / 31.5).contiguous(); wz = torch.zeros(N, dtype=torch.int32, device="cuda")
    codes, rowsum = ops.pack_weight(w, ws, wz, -32, 31)
    bias = torch.randn(N, device="cuda")
    bins = [ops.fq_per_tensor(a, a_scale, a_zp, 0, 63, lsq_grad_factor=1e-4, want_bins=True)[1] for a in acts]
    i = [0]

    def f32():
        j = i[0]; i[0] += 1
        ops.fused_fq_linear(acts[j % len(acts)], a_scale, a_zp, 0, 63, codes, ws, rowsum, bias, lsq_grad_factor=1e-4, out=outs[j % 3])

    def bin_in():
        j = i[0]; i[0] += 1
        ops.fused_fq_linear(acts[j % len(acts)], a_scale, a_zp, 0, 63, codes, ws, rowsum, bias, lsq_grad_factor=1e-4, out=outs[j % 3],
                            a_bins=bins[j % len(acts)])

    res["fused_%dx%d_fp32_in_us" % (K, N)] = chain_time(f32)
    res["fused_%dx%d_bins_in_us" % (K, N)] = chain_time(bin_in)
    if N != 2304:
        ys = [torch.empty_like(a) for a in acts]

        def fq_plain():
            j = i[0]; i[0] += 1
            ops.fq_per_tensor(acts[j % len(acts)], a_scale, a_zp, 0, 63, lsq_grad_factor=1e-4)

        def fq_bins():
            j = i[0]; i[0] += 1
            ops.fq_per_tensor(acts[j % len(acts)], a_scale, a_zp, 0, 63, lsq_grad_factor=1e-4, want_bins=True)
        def fq_only_bins():
            j = i[0]; i[0] += 1
            ops.fq_bins_only(acts[j % len(acts)], a_scale, a_zp, 0, 63, lsq_grad_factor=1e-4)
        res["fq_[%d,%d]_bins_only_us" % (M, K)] = chain_time(fq_only_bins)
        res["fq_[%d,%d]_us" % (M, K)] = chain_time(fq_plain)
        res["fq_[%d,%d]_with_bins_us" % (M, K)] = chain_time(fq_bins)
print(json.dumps(res, indent=1))
json.dump(res, open('gpurun_out/bins_probe.json', 'w'), indent=1)
