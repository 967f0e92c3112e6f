"""One launch of every non-GEMM kernel of the path at the BERT-base seq512 shape, for `ncu --set full` (scripts/gpu_profiles.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from outlier_suppression_b200 import ops
from outlier_suppression_b200.quantization.observer import AvgQuantileObserver

torch.manual_seed(0)
x = torch.randn(32, 512, 768, device="cuda"); x[..., :6] *= 30
x2 = torch.randn(32, 512, 3072, device="cuda")
lens = torch.randint(128, 513, (32,), device="cuda"); lens[0] = 512
sc, zp = torch.tensor([0.5], device="cuda"), torch.tensor([31.0], device="cuda")
for _ in range(2):  # second round = warm
    ops.observe_prune_minmax(x, lens, 1, 0.99)            # token_minmax_kernel + prune_select_tail_kernel
    ops.observe_minmax(x, lens, 1)                         # minmax_masked_kernel
    ops.fq_per_tensor(x2, sc, zp, 0, 63, lsq_grad_factor=1e-4, want_bins=True)              # K1b
    ops.fq_per_tensor(x2, sc, zp, 0, 63, lsq_grad_factor=1e-4, want_bins=True, act="gelu")  # K1c
    o = AvgQuantileObserver(bit=6).cuda(); o(x, lens, 1)   # minmax_masked_kernel + abs_hist_kernel
    ops.mse_multi(x, lens, 1, torch.linspace(0.1, 0.8, 8), torch.full((8,), 31.0), 0, 63)   # mse_multi_kernel<8>
torch.cuda.synchronize()
