"""One launch of K7, K8 and K9 at config 2's size (after one warm-up each), for `ncu --set full`."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from outlier_suppression_b200 import ops

B, S, h, d = 32, 512, 12, 64
H = h * d
hs, rs = torch.randn(B * S, H, device="cuda"), torch.randn(B * S, H, device="cuda")
gm, bias = torch.rand(H, device="cuda") + 0.5, torch.randn(H, device="cuda") * 0.1
def qd(scale, zp, numel):
    return dict(scale=torch.tensor([scale], device="cuda"), zp=torch.tensor([float(zp)], device="cuda"), qmin=0, qmax=63, g=1.0 / (numel * 63) ** 0.5)
q = qd(0.1, 31, B * S * H)
q3, k3, v3 = (torch.randn(B, S, H, device="cuda") for _ in range(3))
heads = lambda t: t.view(B, S, h, d).permute(0, 2, 1, 3)
mask = torch.zeros(B, 1, 1, S, device="cuda")
pq = qd(1 / 63, 0, B * h * S * S)
for _ in range(2):
    ops.residual_layernorm_fq(hs, rs, gm, None, bias, 1e-12, q["scale"], q["zp"], 0, 63, lsq_grad_factor=q["g"], want_bins=True)
    scores = ops.attn_scores_fq(heads(q3), heads(k3), q, q, out_mul=1 / math.sqrt(d), mask=mask)
    probs = torch.softmax(scores, -1)
    ops.attn_context_fq(probs, heads(v3), pq, q, oq=q, want_bins=True)
torch.cuda.synchronize()
print("done")
