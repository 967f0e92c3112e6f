"""LSQ+ backward kernel (fine stage `learn_scale`, token_wise_clipping.py:72-108) on [32, 512, 768] and [32, 512, 3072]."""
import json, os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from outlier_suppression_b200 import ops
res = {}
for shape in ((32, 512, 768), (32, 512, 3072)):
    x = torch.randn(*shape, device="cuda"); dy = torch.randn(*shape, device="cuda")
    sc, zp = torch.tensor([0.1], device="cuda"), torch.tensor([31.3], device="cuda")
    g = 1.0 / (x.numel() * 63) ** 0.5
    for _ in range(3): ops.lsqplus_backward(x, dy, sc, zp, g, 0, 63)
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.lsqplus_backward(x, dy, sc, zp, g, 0, 63); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    us = statistics.median(ts); by = 12 * x.numel()
    res["x".join(map(str, shape))] = {"us": us, "gbs": by / us / 1e3, "frac_of_6534": by / us / 1e3 / 6534.5}
print(json.dumps(res, indent=1)); json.dump(res, open("gpurun_out/lsq_backward.json", "w"), indent=1)
