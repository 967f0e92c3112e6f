"""Weight calibration of BASELINE.json config 3 (RoBERTa-base, 4-bit symmetric per-channel weights, MSEFastObserver
ch_axis=0): the per-channel bounded-Brent searches of one encoder layer's six Linear weights.

GPU: `osq_mse_brent_rows_f32` (one CTA per output channel, the row in shared memory, the whole search on chip).
CPU: the oracle port of observer.py:483-517 (SciPy minimize_scalar per channel) on a 64-channel sample, scaled.
Prints one JSON line; parity (ranges within the tolerance of tests/test_gpu_observers.py) is asserted on the sample."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import osq_oracle as O          # checker / CPU baseline only
from outlier_suppression_b200 import ops

SHAPES = [(768, 768)] * 4 + [(3072, 768), (768, 3072)]


def main():
    torch.manual_seed(0)
    qmin, qmax = O.quant_range(4, True)
    ws = [torch.randn(n, k) * 0.05 for n, k in SHAPES]
    wg = [w.cuda() for w in ws]
    for w in wg:                                   # warm-up
        ops.mse_brent_rows(w, qmin, qmax, "no")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        outs = [ops.mse_brent_rows(w, qmin, qmax, "no", want_evals=True) for w in wg]
    e1.record(); torch.cuda.synchronize()
    gpu_ms = e0.elapsed_time(e1) / reps
    channels = sum(n for n, _ in SHAPES)
    evals = float(sum(o[2].sum() for o in outs)) / channels
    # CPU oracle on a sample of channels of the first and the widest weight
    sample = 32
    t0 = time.perf_counter()
    worst = 0.0
    for wi in (0, 5):
        w = ws[wi][:sample]
        st = O.ObserverState()
        O.observe_mse_fast(st, w, qmin, qmax, True, ch_axis=0)
        g_min, g_max = outs[wi][0][:sample].cpu(), outs[wi][1][:sample].cpu()
        worst = max(worst, float(((g_max - st.max_val.float()).abs() / st.max_val.float().abs()).max()))
    cpu_s = time.perf_counter() - t0
    cpu_ms_layer = cpu_s / (2 * sample) * channels * 1e3
    assert worst < 2e-3, worst                      # same bar as tests/test_gpu_observers.py (search tolerance xatol=1e-5 on the range)
    print(json.dumps({"metric": "MSEFast per-channel weight calibration, one RoBERTa-base encoder layer (6 weights, %d channels, 4-bit sym)" % channels,
                      "gpu_ms_per_layer": gpu_ms, "channels_per_s": channels / (gpu_ms * 1e-3), "loss_evaluations_per_channel": evals,
                      "cpu_port_ms_per_layer_scaled": cpu_ms_layer, "cpu_sample": "%d channels of weights 0 and 5, scaled" % (2 * sample),
                      "max_rel_range_diff_on_sample": worst}))


if __name__ == "__main__":
    main()
