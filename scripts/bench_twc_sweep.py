"""Token-wise clipping, coarse stage (solver/token_wise_clipping.py:50-66), on a 12-layer random-init BERT-base stack at the
shape of BASELINE config 2 (seq 512, batch 32, 8 calibration batches = calibrate 256; 6-bit -> step 0.0025, 120 iterations):

  graphed   outlier_suppression_b200.twc.GraphedFindRatio: all 120 iterations, 2 x 8 model forwards each, replayed from CUDA graphs
  eager     the reference's own loop (its unmodified set_ratio / calibrate / enable_quantization) on this backend, 3 iterations, scaled
  cpu       the reference on its own quantization package on the host cores: ONE calibration forward + ONE quantized forward of one
            batch, scaled to 120 x 8 x 2 forwards (a full sweep would take hours)

    python scripts/bench_twc_sweep.py [--layers 12] [--seq 512] [--batch 32] [--batches 8] [--iters 120]
"""
import argparse, copy, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ref_model as RM, ref_shim   # CPU baseline + the reference's model code (test infrastructure)

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=12); ap.add_argument("--seq", type=int, default=512)
ap.add_argument("--batch", type=int, default=32); ap.add_argument("--batches", type=int, default=8)
ap.add_argument("--iters", type=int, default=120); ap.add_argument("--no-cpu", action="store_true")
a = ap.parse_args()
step = min(float(format(128 * 32 * 0.01 / a.batch / a.seq, ".2g")), 0.01)      # token_wise_clipping.cac_step_iters
qcfg = RM.quant_config()
mcfg = RM.Cfg(model_type="bert")
fp = RM.fp_bert(layers=a.layers, hidden=768, heads=12, inter=3072, vocab=30522, max_pos=512, seed=0)
cpu_batches = RM.synth_batches(a.batches, a.batch, a.seq, 30522, "cpu", seed=5)
res = {"config": {"layers": a.layers, "seq": a.seq, "batch": a.batch, "batches": a.batches, "iters": a.iters, "step": step}}


def prepare(ns, device, batches):
    model = RM.build_model(ns, copy.deepcopy(fp), qcfg, device)
    Q = ns.quantization
    model = ns.gamma_migration.delay_ln(model, qcfg, mcfg)
    Q.disable_all(model)
    with torch.no_grad():
        fp_out = [(lambda o: o[0] if isinstance(o, tuple) else o.logits)(model(**b)).clone() for b in batches]
    Q.enable_calibration_woquantization(model, quantizer_type="weight_fake_quant")
    with torch.no_grad():
        model(**batches[0])
    Q.disable_all(model)
    Q.state.set_observer_name(model)
    return model, fp_out


if not a.no_cpu:
    ref = RM.load_stack("reference")
    with ref_shim.cpu_only():
        torch.set_num_threads(os.cpu_count() or 1)
        m, fo = prepare(ref, "cpu", cpu_batches[:1])
        twc = ref.token_wise_clipping; twc.task_type = "glue"
        twc.set_ratio(m, 0.99)
        t0 = time.perf_counter(); twc.calibrate(m, cpu_batches[:1]); t_cal = time.perf_counter() - t0
        twc.enable_quantization(m)
        t0 = time.perf_counter(); twc.calibrate(m, cpu_batches[:1], fo); t_q = time.perf_counter() - t0
    res["cpu_reference"] = {"s_calibration_forward": t_cal, "s_quantized_forward": t_q, "cores": os.cpu_count(),
                            "s_per_sweep_scaled": (t_cal + t_q) * a.batches * a.iters,
                            "sample": "one batch: one calibration forward + one quantized forward, scaled x %d batches x %d iterations" % (a.batches, a.iters)}
    del m, fo

ns = RM.load_stack("b200")
from outlier_suppression_b200.twc import GraphedFindRatio
batches = [{k: v.cuda() for k, v in b.items()} for b in cpu_batches]
model, fp_out = prepare(ns, "cuda", batches)
twc = ns.token_wise_clipping; twc.task_type = "glue"
# eager loop of the reference on this backend, 3 iterations
torch.cuda.synchronize(); t0 = time.perf_counter()
eager_losses = []
for i in range(3):
    twc.set_ratio(model, 1.0 - step * i); twc.calibrate(model, batches); twc.enable_quantization(model)
    eager_losses.append(float(twc.calibrate(model, batches, fp_out)))
torch.cuda.synchronize(); t_eager = (time.perf_counter() - t0) / 3
res["eager_on_b200"] = {"s_per_iteration": t_eager, "s_per_sweep_scaled": t_eager * a.iters, "losses_first3": eager_losses}
sweep = GraphedFindRatio(model, batches, fp_out)
torch.cuda.synchronize(); t0 = time.perf_counter()
sweep.capture(1.0)
torch.cuda.synchronize(); t_cap = time.perf_counter() - t0
t0 = time.perf_counter()
ratio, losses = sweep.find_ratio(a.iters, step)
torch.cuda.synchronize(); t_sweep = time.perf_counter() - t0
res["graphed_on_b200"] = {"s_capture": t_cap, "s_per_sweep": t_sweep, "s_per_iteration": t_sweep / a.iters, "best_ratio": ratio,
                          "losses_first3": losses[:3], "losses_match_eager": losses[:3] == eager_losses,
                          "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/twc_sweep.json", "w"), indent=1)
