"""-m gpu: token-wise clipping's coarse stage (solver/token_wise_clipping.py:50-66) replayed from CUDA graphs
(outlier_suppression_b200/twc.py) against the reference's own eager loop -- its unmodified set_ratio / calibrate /
enable_quantization driving the unmodified quant_bert on this backend: per-ratio losses, the chosen ratio and every
quantizer's final (scale, zero_point) must be bit-identical."""
import copy

import numpy as np
import pytest
import torch

from oracle import make_ref, ref_model as RM

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(make_ref.root() is None, reason="reference tree not staged (oracle/_ref)")]


def _prepare(ns, qcfg, seed):
    fp = RM.fp_bert(layers=2, hidden=128, heads=2, inter=512, vocab=100, max_pos=64, seed=seed)
    model = RM.build_model(ns, copy.deepcopy(fp), qcfg, "cuda")
    batches = [{k: v.cuda() for k, v in b.items()} for b in RM.synth_batches(3, 4, 32, 100, "cpu", seed=5)]
    Q = ns.quantization
    mcfg = RM.Cfg(model_type="bert")
    model = ns.gamma_migration.delay_ln(model, qcfg, mcfg)                              # ptq_glue_quant.py:230-232
    Q.disable_all(model)
    with torch.no_grad():
        fp_out = [(lambda o: o[0] if isinstance(o, tuple) else o.logits)(model(**b)).clone() for b in batches]  # prepare_input_output
    Q.enable_calibration_woquantization(model, quantizer_type="weight_fake_quant")      # :234-235
    with torch.no_grad():
        model(**batches[0])
    Q.disable_all(model)                                                                # :237-238
    Q.state.set_observer_name(model)
    return model, batches, fp_out


@pytest.mark.parametrize("cache_vectors", [True, False])
def test_graphed_find_ratio_matches_the_reference_loop(cache_vectors):
    """cache_vectors=True: calibration at a ratio = the cached select over the recorded per-token vectors + one replay launch;
    False: replayed calibration forwards.  Both must reproduce the reference's eager loop bit for bit."""
    from outlier_suppression_b200.quantization.fake_quant import QuantizeBase
    from outlier_suppression_b200.twc import GraphedFindRatio
    ns = RM.load_stack("b200")
    qcfg = RM.quant_config()
    iters, step = 5, 0.04
    # ---- the reference's loop, eager, on this backend ----
    model, batches, fp_out = _prepare(ns, qcfg, seed=0)
    twc = ns.token_wise_clipping
    twc.task_type = "glue"
    ref_losses, best, best_i = [], 10000000, 0
    for i in range(iters):
        twc.set_ratio(model, 1.0 - step * i)
        twc.calibrate(model, batches)
        twc.enable_quantization(model)
        cur = twc.calibrate(model, batches, fp_out)
        ref_losses.append(float(cur))
        if best > cur:
            best, best_i = cur, i
    twc.set_ratio(model, 1.0 - step * best_i)
    twc.calibrate(model, batches)
    ref_q = {n: (m.scale.detach().clone(), m.zero_point.detach().clone()) for n, m in model.named_modules() if isinstance(m, QuantizeBase)}
    # ---- the same sweep from CUDA graphs on an identically prepared model ----
    model2, batches2, fp_out2 = _prepare(ns, qcfg, seed=0)
    for a, b in zip(fp_out, fp_out2):
        assert torch.equal(a, b)
    sweep = GraphedFindRatio(model2, batches2, fp_out2, cache_vectors=cache_vectors)
    ratio, losses = sweep.find_ratio(iters, step)
    assert (sweep._cache is not None) == cache_vectors and (len(sweep._cal) == 0) == cache_vectors
    assert ratio == 1.0 - step * best_i
    np.testing.assert_array_equal(np.array(losses, dtype=np.float64), np.array(ref_losses, dtype=np.float64))
    assert len(set(losses)) > 1, "the sweep must actually depend on the ratio"
    for n, m in model2.named_modules():
        if isinstance(m, QuantizeBase):
            s, z = ref_q[n]
            assert torch.equal(m.scale.detach(), s) and torch.equal(m.zero_point.detach(), z), n
            if "act" in n:
                assert m.observer_enabled == 1 and m.fake_quant_enabled == 0 and m.observer.cnt == len(batches2)
