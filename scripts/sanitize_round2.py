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
