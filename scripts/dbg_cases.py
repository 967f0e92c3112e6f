import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import osq_oracle as O
from outlier_suppression_b200 import ops
for (m, k, n) in ((128, 128, 16), (300, 768, 768), (64, 256, 400)):
    g = torch.Generator().manual_seed(m + k + n)
    a = torch.randn(m, k, generator=g); a[:, :3] *= 20
    w, bias = O.synth_linear(n, k, seed=m + k + n + 1)
    mn, mx = O.global_minmax(a)
    a_scale, a_zp = O.qparams_from_minmax(mn * 0.7, mx * 0.7, 0, 63, False)
    w_scale, w_zp, wqmin, wqmax = O.weight_qparams_minmax(w, 6, True)
    y_ref, qa, qw = O.fused_fq_linear(a, a_scale.reshape(1), a_zp.reshape(1).to(torch.int32), 0, 63, False, w, w_scale, w_zp, wqmin, wqmax, bias)
    codes, rowsum = ops.pack_weight(w.cuda(), w_scale.cuda(), w_zp.cuda(), wqmin, wqmax)
    y = ops.fused_fq_linear(a.cuda(), a_scale.reshape(1).cuda(), a_zp.reshape(1).to(torch.int32).cuda(), 0, 63, codes, w_scale.cuda(), rowsum, bias.cuda())
    torch.cuda.synchronize()
    d = (y.cpu() - y_ref).abs()
    tol = 1e-3 * y_ref.abs() + 1e-3 * y_ref.abs().max()
    bad = d > tol
    print("shape", (m, k, n), "bad", int(bad.sum()), "of", bad.numel(), "max", float(d.max()))
    if bad.any():
        rows = bad.any(1).nonzero().flatten().tolist(); cols = bad.any(0).nonzero().flatten().tolist()
        print("  bad rows (first 20):", rows[:20], "... count", len(rows)); print("  bad cols (first 40):", cols[:40], "... count", len(cols))
        r, c = bad.nonzero()[0].tolist()
        print("  sample", (r, c), float(y[r, c]), float(y_ref[r, c]), "ratio", float(y[r, c]) / float(y_ref[r, c]))
        # is it a constants problem? compare against acc*c1 + c0 with other columns' constants
        acc = ((qa.double() - float(a_zp)) @ qw.double().t())
        print("  y/acc at sample:", float(y[r, c]) , float(acc[r, c]), "c1 true", float(a_scale * w_scale[c]))
