"""Small launches of every kernel added or changed in round 2, for compute-sanitizer (memcheck / synccheck / racecheck);
parity is asserted inside (the same checks as the -m gpu tests, at small sizes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import osq_oracle as O
from outlier_suppression_b200 import ops
from outlier_suppression_b200.quantization.observer import AvgMSEFastObserver, AvgQuantileObserver
from tests.test_gpu_fused_linear import run_case

which = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.manual_seed(0)
if which in ("all", "fused"):
    for (m, k, n) in [(300, 768, 768), (300, 3072, 768), (200, 256, 96)]:
        run_case(m, k, n, 6, 6, True, 11)
        print("fused ok", m, k, n, flush=True)
if which in ("all", "observers"):
    g = torch.Generator().manual_seed(3)
    for n_tok in (700, 5000, 40000):
        x = torch.randn(2, n_tok // 2, 64, generator=g) * torch.rand(2, n_tok // 2, 1, generator=g).mul(4).exp()
        lens = torch.tensor([n_tok // 2, n_tok // 5])
        lo, hi = O.prune_minmax(O.token_matrix(x, lens.tolist(), 1), 0.97)
        cur = ops.observe_prune_minmax(x.cuda(), lens.cuda(), 1, 0.97)
        assert torch.equal(cur.cpu(), torch.stack([lo, hi]))
        pd = torch.tensor([0.97], device="cuda")
        cur = ops.observe_prune_minmax(x.cuda(), lens.cuda(), 1, 0.5, percentile_dev=pd)
        assert torch.equal(cur.cpu(), torch.stack([lo, hi]))
        print("prune ok", n_tok, flush=True)
    # batched call (per-token passes overlapped with the previous select tail) and the cached multi-problem select
    from outlier_suppression_b200.quantization.quantized_module import Quantizer as _Q
    from tests.test_host_logic import QC as _QC
    lens3 = torch.tensor([90, 11, 47], device="cuda")
    xs3 = [torch.randn(3, 90, 64, device="cuda") * (1 + b) for b in range(3)]
    qa, qb = _Q(None, _QC("LSQPlusFakeQuantize", "AvgPruneMinMaxObserver", 6, False, -1)).cuda(), _Q(None, _QC("LSQPlusFakeQuantize", "AvgPruneMinMaxObserver", 6, False, -1)).cuda()
    for q_ in (qa, qb):
        q_.observer.set_name("x"); q_.observer.set_percentile(0.9); q_.enable_observer()
    for x_ in xs3:
        qa(x_, lens3, 1)
    qb.observe_many(xs3, lens3, 1)
    assert torch.equal(qa.scale.detach(), qb.scale.detach()) and torch.equal(qa.observer.min_val, qb.observer.min_val)
    vecs = [ops.token_minmax_hist(x_, lens3, 1) for x_ in xs3]
    table = torch.zeros(3, 2, device="cuda")
    ops.prune_select_cached(ops.select_problems([(v, table[i]) for i, v in enumerate(vecs)], "cuda"), 3, 0.9)
    for i, x_ in enumerate(xs3):
        assert torch.equal(table[i], ops.observe_prune_minmax(x_, lens3, 1, 0.9))
    print("observe_many / cached select ok", flush=True)
    o = AvgQuantileObserver(bit=6).cuda()
    o(torch.randn(4, 50, 96, device="cuda"), torch.tensor([50, 3, 20, 1], device="cuda"), 1)
    torch.cuda.synchronize(); print("quantile ok", float(o.min_val), float(o.max_val), flush=True)
    from outlier_suppression_b200.dist import sharded_calibration
    from outlier_suppression_b200.quantization.quantized_module import Quantizer
    from tests.test_host_logic import QC
    net = torch.nn.Module(); net.a_act_fake_quant = Quantizer(None, QC("LSQPlusFakeQuantize", "AvgPruneMinMaxObserver", 6, False, -1)).cuda()
    q = net.a_act_fake_quant; q.observer.set_name("x"); q.observer.set_percentile(0.9); q.enable_observer()
    with sharded_calibration(net, 3) as ctl:
        for i in range(3):
            ctl.set_batch(i); q(torch.randn(2, 40, 64, device="cuda"), torch.tensor([40, 7], device="cuda"), 1)
    torch.cuda.synchronize(); print("replay ok", float(q.scale), flush=True)
if which in ("all", "mse"):
    o = AvgMSEFastObserver(bit=6, symmetric=False, ch_axis=-1).cuda()
    o(torch.randn(2, 24, 64, device="cuda") * 3, torch.tensor([24, 9], device="cuda"), 1)
    torch.cuda.synchronize(); print("mse tensor ok", float(o.min_val), float(o.max_val), o.loss_evals, flush=True)
if which in ("all", "fq"):
    x = torch.randn(3, 77, 130, device="cuda") * 3
    sc, zp = torch.tensor([0.2], device="cuda"), torch.tensor([11.3], device="cuda")
    a, b = ops.fq_per_tensor(torch.nn.functional.gelu(x), sc, zp, 0, 63, lsq_grad_factor=1e-3, want_bins=True)
    c, d = ops.fq_per_tensor(x, sc, zp, 0, 63, lsq_grad_factor=1e-3, want_bins=True, act="gelu")
    assert torch.equal(a, c) and torch.equal(b, d)
    print("k1c ok", flush=True)
if which in ("all", "f3"):
    # K7 / K8 / K9 (layernorm_fq.cu, attention.cu) at small ragged sizes; parity as in tests/test_gpu_layernorm_fq.py / test_gpu_attention.py
    import math
    sc, zp = torch.tensor([0.11], device="cuda"), torch.tensor([30.0], device="cuda")
    for rows, H in ((37, 768), (9, 1024), (21, 200)):
        h, r = torch.randn(rows, H, device="cuda"), torch.randn(rows, H, device="cuda")
        gm, b = torch.rand(H, device="cuda") + 0.5, torch.randn(H, device="cuda")
        y, bins, ln = ops.residual_layernorm_fq(h, r, gm, None, b, 1e-12, sc, zp, 0, 63, lsq_grad_factor=1e-3, want_bins=True, want_ln=True)
        y1, b1 = ops.fq_per_tensor(ln, sc, zp, 0, 63, lsq_grad_factor=1e-3, want_bins=True)
        assert torch.equal(y, y1) and torch.equal(bins, b1)
        print("k7 ok", rows, H, flush=True)
    for B, h_, Sq, Sk, d in ((1, 2, 70, 132, 64), (1, 1, 20, 600, 32), (1, 1, 130, 64, 128)):
        H = h_ * d
        q3, k3, v3 = torch.randn(B, Sq, H, device="cuda"), torch.randn(B, Sk, H, device="cuda"), torch.randn(B, Sk, H, device="cuda")
        hd = lambda t, S: t.view(B, S, h_, d).permute(0, 2, 1, 3)
        qa = dict(scale=sc, zp=zp, qmin=0, qmax=63, g=1e-3)
        pa = dict(scale=torch.tensor([1 / 63], device="cuda"), zp=torch.tensor([0.0], device="cuda"), qmin=0, qmax=63, g=1e-3)
        mask = torch.zeros(B, 1, 1, Sk, device="cuda")
        s_ = ops.attn_scores_fq(hd(q3, Sq), hd(k3, Sk), qa, qa, out_mul=1 / math.sqrt(d), mask=mask)
        fq = lambda x, q: ops.fq_per_tensor(x.contiguous(), q["scale"], q["zp"], 0, 63, lsq_grad_factor=1e-3)
        ref = torch.matmul(fq(hd(q3, Sq), qa).double(), fq(hd(k3, Sk), qa).double().transpose(-1, -2)) / math.sqrt(d)
        assert float((s_.double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
        probs = torch.softmax(s_, -1)
        ctx, cb = ops.attn_context_fq(probs, hd(v3, Sk), pa, qa, oq=qa, want_bins=True)
        plain = ops.attn_context_fq(probs, hd(v3, Sk), pa, qa)
        refc = torch.matmul(fq(probs, pa).double(), fq(hd(v3, Sk), qa).double()).permute(0, 2, 1, 3).reshape(B, Sq, H)
        assert float((plain.double() - refc).abs().max()) <= 1e-5 * float(refc.abs().max())
        y1, b1 = ops.fq_per_tensor(plain, sc, zp, 0, 63, lsq_grad_factor=1e-3, want_bins=True)
        assert torch.equal(ctx, y1) and torch.equal(cb, b1)
        print("k8 k9 ok", B, h_, Sq, Sk, d, flush=True)
