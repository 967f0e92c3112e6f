"""Measures what this B200 sustains for pure writes, pure reads and copies with stock kernels (torch fill_, sum, copy_)
on buffers far larger than the 126 MB L2: the practical ceilings the fused kernel's two phases are compared with in DESIGN.md."""
import json
import torch

def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3

out = {}
for mb in (256, 1024, 4096):
    n = mb * 1024 * 1024 // 4
    a = torch.empty(n, device="cuda"); b = torch.empty(n, device="cuda")
    a.normal_()
    t = timeit(lambda: b.fill_(1.5)); out["fill_%dMB_GBs" % mb] = n * 4 / t / 1e9
    t = timeit(lambda: b.zero_()); out["memset_%dMB_GBs" % mb] = n * 4 / t / 1e9
    t = timeit(lambda: a.sum()); out["sum_read_%dMB_GBs" % mb] = n * 4 / t / 1e9
    t = timeit(lambda: b.copy_(a)); out["copy_%dMB_GBs_rw" % mb] = 2 * n * 4 / t / 1e9
    del a, b
print(json.dumps(out, indent=1))
