"""-m gpu: the registry entries no shipped twc / minmax / mse config reaches against goldens produced
by the UNMODIFIED reference (tests/golden/gen_golden.py:gen_extra): AvgQuantileObserver, MSEObserver / AvgMSEObserver,
LSQPlusObserver, LSQFakeQuantize, per-channel LSQ+, QEmbedding.  Elementwise results and selections are BIT-EXACT;
mean / std reductions and MSE grid choices carry the tolerance written at the assert."""
import numpy as np
import pytest
import torch

from oracle import osq_oracle as O
from tests.test_host_logic import QC

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def _state(o):
    return np.array([float(o.min_val), float(o.max_val)], dtype=np.float32)


@pytest.mark.parametrize("case", ["quant_mask", "quant_nomask"])
@pytest.mark.parametrize("thr,bins", [(0.999, 2048), (0.99, 256)])
def test_avg_quantile_observer_golden(golden, case, thr, bins):
    """observer.py:240-282: |x| histogram over the valid tokens, first bin whose cumulative count reaches the threshold,
    clip, running average.  Counts are integers, so the trace must match the reference bit for bit."""
    from outlier_suppression_b200.quantization.observer import AvgQuantileObserver
    g = golden("extra")
    lens = g[case + "_lens"]
    mask = None if lens.size == 0 else T(lens).cuda()
    seq_pos = int(g[case + "_meta"][0])
    o = AvgQuantileObserver(bit=6, symmetric=False, ch_axis=-1, threshold=thr, bins=bins).cuda()
    for b in range(3):
        o(T(g["%s_x%d" % (case, b)]).cuda(), observation_mask=mask, seq_pos=seq_pos)
        np.testing.assert_array_equal(_state(o), g["%s_trace_%d" % (case, bins)][b])
    assert o.cnt == 3


def _mse_of(x, lens, lo, hi, bit, sym):
    qmin, qmax = O.quant_range(bit, sym)
    tok = O.token_matrix(x, lens, 1) if lens is not None else x
    return float(O.mse_loss(tok, torch.tensor(lo), torch.tensor(hi), qmin, qmax, sym))


@pytest.mark.parametrize("tag,cls,bit,sym", [("mse_sym", "MSEObserver", 6, True), ("avgmse_sym", "AvgMSEObserver", 6, True),
                                             ("mse_asym4", "MSEObserver", 4, False), ("avgmse_asym4", "AvgMSEObserver", 4, False)])
def test_mse_grid_observers_golden(golden, tag, cls, bit, sym):
    """observer.py:285-409.  The grid (100 ranges x 2^bit zero points) is the reference's; the per-candidate losses are
    accumulated in fp64 here and as an fp32 mean there, so two candidates whose losses tie to fp32 rounding may swap:
    the chosen candidate must be the reference's, or a grid neighbour (<= 2 % of the range) whose loss is no worse than
    the reference's choice by more than 1e-4 relative."""
    from outlier_suppression_b200.quantization import observer as obs
    g = golden("extra")
    lens = T(g["mse_lens"])
    xs = [T(g["mse_x0"]), T(g["mse_x1"])]
    o = getattr(obs, cls)(bit=bit, symmetric=sym, ch_axis=-1).cuda()
    ref = g[tag + "_trace"]
    first = None
    for b, x in enumerate(xs):
        o(x.cuda(), observation_mask=lens.cuda(), seq_pos=1)
        got = _state(o)
        if b == 0:
            first = got
            span = float(ref[0][1] - ref[0][0])
            if not np.array_equal(got, ref[0]):
                assert np.abs(got - ref[0]).max() <= 0.02 * span, (got, ref[0])
                l_got = _mse_of(x, lens.tolist(), float(got[0]), float(got[1]), bit, sym)
                l_ref = _mse_of(x, lens.tolist(), float(ref[0][0]), float(ref[0][1]), bit, sym)
                assert l_got <= l_ref * (1 + 1e-4), (l_got, l_ref)
        else:
            np.testing.assert_allclose(got, ref[b], rtol=0, atol=0.02 * float(ref[b][1] - ref[b][0]))
    assert first is not None


def test_avg_mse_observer_one_sided_golden(golden):
    from outlier_suppression_b200.quantization.observer import AvgMSEObserver
    g = golden("extra")
    o = AvgMSEObserver(bit=6, symmetric=False, ch_axis=-1).cuda()
    o(T(g["mse_pos_x"]).cuda())
    assert o.one_side_dist == "pos"
    got, ref = _state(o), g["mse_pos"]
    assert got[0] == ref[0] == 0.0
    assert abs(got[1] - ref[1]) <= 0.011 * ref[1], (got, ref)  # same or neighbouring grid point (1 % steps)


def test_lsqplus_observer_golden(golden):
    """observer.py:148-173: mean +- 3 std.  mean / std are reductions (order differs on the GPU): 1e-5 relative."""
    from outlier_suppression_b200.quantization.observer import LSQPlusObserver
    g = golden("extra")
    w = T(g["lsq_w"]).cuda()
    for tag, ch in (("lsqobs_t", -1), ("lsqobs_c", 0)):
        o = LSQPlusObserver(bit=4, symmetric=True, ch_axis=ch).cuda()
        o(w)
        np.testing.assert_allclose(o.min_val.cpu().numpy(), g[tag + "_min"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(o.max_val.cpu().numpy(), g[tag + "_max"], rtol=1e-5, atol=1e-7)
        s, z = o.calculate_qparams(o.min_val, o.max_val)
        np.testing.assert_allclose(s.cpu().numpy(), g[tag + "_scale"], rtol=1e-5)
        np.testing.assert_array_equal(z.cpu().numpy(), g[tag + "_zp"])


@pytest.mark.parametrize("tag,quantizer,ch,src", [("lsq_t", "LSQFakeQuantize", -1, "lsq_x"), ("lsq_c", "LSQFakeQuantize", 0, "lsq_w"),
                                                  ("lsqplus_c", "LSQPlusFakeQuantize", 0, "lsq_w")])
def test_lsq_variants_forward_golden(golden, tag, quantizer, ch, src):
    """fake_quant.py:129-167 (LSQ) and the per-channel branch of :170-209 (LSQ+): calibrate with MinMax, then forward.
    Elementwise fp32 with tensor divisors (true division on every backend) -> bit-exact."""
    from outlier_suppression_b200.quantization.quantized_module import Quantizer
    g = golden("extra")
    x = T(g[src]).cuda()
    q = Quantizer(None, QC(quantizer, "MinMaxObserver", 6, True, ch)).cuda()
    q.enable_observer(); q(x); q.disable_observer(); q.enable_fake_quant()
    np.testing.assert_array_equal(q.scale.data.cpu().numpy(), g[tag + "_scale"])
    if quantizer == "LSQPlusFakeQuantize":
        q.zero_point.data += 0.3
    np.testing.assert_array_equal(q.zero_point.data.float().cpu().numpy(), g[tag + "_zp"])
    with torch.no_grad():
        y = q(x)
    np.testing.assert_array_equal(y.cpu().numpy(), g[tag + "_y"])


def test_qembedding_cached_table_golden(golden):
    """quantized_module.py:75-100; here the table is fake-quantized once per weight version, not per forward."""
    from outlier_suppression_b200.quantization import quantized_module as qm
    g = golden("extra")
    emb = torch.nn.Embedding(50, 32, padding_idx=0)
    emb.weight.data = T(g["emb_w"]).clone()
    qe = qm.Quantizer(emb, QC("FixedFakeQuantize", "MinMaxObserver", 6, True, 0)).cuda()
    ids = T(g["emb_ids"]).cuda()
    qe.weight_fake_quant.enable_observer(); qe(ids); qe.weight_fake_quant.disable_observer()
    qe.weight_fake_quant.enable_fake_quant()
    np.testing.assert_array_equal(qe.weight_fake_quant.scale.cpu().numpy(), g["emb_scale"])
    with torch.no_grad():
        y1 = qe(ids)
        cached = qe._fq_weight_cache[1]
        y2 = qe(ids)
        assert qe._fq_weight_cache[1] is cached                 # second forward: no new fake-quant launch
        np.testing.assert_array_equal(y1.cpu().numpy(), g["emb_y"])
        np.testing.assert_array_equal(y2.cpu().numpy(), g["emb_y"])
        qe.weight.data[5] *= 3.0                                # edits through .data are invisible to _version ...
        qe.invalidate_packed()                                  # ... the togglers (state.py) call this
        y3 = qe(ids)
    assert qe._fq_weight_cache[1] is not cached
    ref = O.fq_per_channel(qe.weight.detach().cpu(), T(g["emb_scale"]), torch.zeros(50, dtype=torch.int32), 0, -32, 31)[ids.cpu()]
    np.testing.assert_array_equal(y3.cpu().numpy(), ref.numpy())
