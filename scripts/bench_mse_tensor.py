"""(Avg)MSEFastObserver per-tensor search, BASELINE.md's probe shape [32, 128, 768] asymmetric 6-bit (12.2 s on the survey
container's CPU through the reference): device-side cooperative search vs the round-1 host-driven search."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from outlier_suppression_b200.quantization.observer import AvgMSEFastObserver

torch.manual_seed(0)
x = torch.randn(32, 128, 768, device="cuda")
x[..., :6] *= 12.0
lens = torch.randint(32, 129, (32,), device="cuda"); lens[0] = 128
res = {}
for name, host in (("device_cooperative", False), ("host_scipy_round1", True)):
    o = AvgMSEFastObserver(bit=6, symmetric=False, ch_axis=-1).cuda()
    o.host_search = host
    o(x, lens, 1); torch.cuda.synchronize()          # warm-up (also decides one_side_dist)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        o(x, lens, 1)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    res[name] = {"s_per_call": dt, "loss_evals_per_call": o.loss_evals / (reps + 1), "min_val": float(o.min_val), "max_val": float(o.max_val),
                 "one_side_dist": o.one_side_dist}
res["reference_cpu_s_per_call_survey_container"] = 12.2
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/mse_tensor.json", "w"), indent=1)
