"""Profiling aid: prints the clock64 timeline of CTA 0 of the fused kernel for the three BERT-base sites."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from outlier_suppression_b200 import ops

torch.manual_seed(0)
M = 16384
import os
SHAPES = eval(os.environ.get('TRACE_SHAPES', '((768, 768), (768, 3072), (3072, 768))'))
for K, N in SHAPES:
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * 0.05
    bias = torch.randn(N, device="cuda")
    a_scale = torch.tensor([0.1], device="cuda"); a_zp = torch.tensor([31.0], device="cuda")
    w_scale = (w.abs().amax(1) / 31.5).contiguous(); w_zp = torch.zeros(N, dtype=torch.int32, device="cuda")
    codes, rowsum = ops.pack_weight(w, w_scale, w_zp, -32, 31)
    tr = torch.zeros(2048, dtype=torch.int64, device="cuda")
    bins = ops.fq_per_tensor(a, a_scale, a_zp, 0, 63, lsq_grad_factor=1e-4, want_bins=True)[1] if os.environ.get('TRACE_BINS') == '1' else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ops.fused_fq_linear(a, a_scale, a_zp, 0, 63, codes, w_scale, rowsum, bias, lsq_grad_factor=1e-4, a_bins=bins)
    tr.zero_()
    e0.record()
    ops.fused_fq_linear(a, a_scale, a_zp, 0, 63, codes, w_scale, rowsum, bias, lsq_grad_factor=1e-4, trace=tr, a_bins=bins)
    e1.record()
    torch.cuda.synchronize()
    t = tr.cpu().tolist()
    t0 = t[1020]
    rel = lambda v: (v - t0) if v else None
    print("=== K=%d N=%d  end=%s cycles; event time %.1f us" % (K, N, rel(t[1021]), e0.elapsed_time(e1) * 1e3))
    st = [t[1024 + 2 * i] for i in range(148) if t[1024 + 2 * i]]; en = [t[1025 + 2 * i] for i in range(148) if t[1025 + 2 * i]]
    g0 = min(st)
    print("ctas=%d start spread %.1f us; end min/median/max %.1f/%.1f/%.1f us after first start; cta0 %.1f..%.1f" % (len(st), (max(st) - g0) / 1e3, (min(en) - g0) / 1e3, (sorted(en)[len(en) // 2] - g0) / 1e3, (max(en) - g0) / 1e3, (st[0] - g0) / 1e3, (en[0] - g0) / 1e3))
    print("conv a_full arrivals:", [rel(v) for v in t[0:256] if v][:40])
    print("wprod issue times, uses 36..75:", [rel(v) for v in t[1600:1640] if v])
    print("mma per k-block uses 36..71 (a ready, w ready, issued):", [tuple(rel(t[1400 + 3 * i + k]) for k in range(3)) for i in range(36) if t[1400 + 3 * i]])
    print("mma  (start, first operands, commit issued):", [(rel(t[256 + 4 * i]), rel(t[257 + 4 * i]), rel(t[258 + 4 * i])) for i in range(60) if t[256 + 4 * i]][:14])
    print("epi groups of chunk 2 (start, tmem ld done, tile free, math+sts done, fence done, store issued):", [[rel(t[900 + 6 * g + i]) for i in range(6)] for g in range(4)])
    print("epi  (acc_full acquired, chunk done, all warps done, consts published):", [tuple(rel(t[512 + 4 * i + j]) for j in range(4)) for i in range(60) if t[512 + 4 * i]][:14])
