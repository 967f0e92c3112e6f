"""-m gpu: the deferred activation fake-quant (osq_fq_per_tensor_bins_only_f32 / osq_dequant_bins_f32, LazyFakeQuant).

A quantizer whose output is consumed by fused QLinears writes only the uint8 bins; the fp32 tensor of util_quant.py:14 is
produced on demand.  Everything observable must be bit-identical to the eager path: the bins, the Linear behind them, and
the fp32 values whenever anybody asks for them."""
import pytest
import torch

from oracle import osq_oracle as O
from outlier_suppression_b200 import ops
from tests.test_host_logic import QC

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("lsq", [False, True])
@pytest.mark.parametrize("bits", [4, 6, 8])
def test_bins_only_and_dequant_match_k1(bits, lsq):
    g = torch.Generator().manual_seed(bits + 10 * lsq)
    x = torch.randn(3, 50, 256, generator=g) * 3
    x[..., :2] *= 20
    qmin, qmax = 0, 2 ** bits - 1
    sc = torch.tensor([float(x.abs().max()) * 2 / qmax * 0.6])
    z = torch.tensor([float(qmax // 2 + 1)])
    gf = 1.0 / (x.numel() * qmax) ** 0.5 if lsq else 0.0
    zp = (z if lsq else z.to(torch.int32)).cuda()
    y, bins = ops.fq_per_tensor(x.cuda(), sc.cuda(), zp, qmin, qmax, lsq_grad_factor=gf, want_bins=True)
    b2, eff = ops.fq_bins_only(x.cuda(), sc.cuda(), zp, qmin, qmax, lsq_grad_factor=gf)
    assert torch.equal(b2, bins)
    assert torch.equal(ops.dequant_bins(b2, eff, qmin, qmax), y)
    # GELU in front (K1c), bins only
    yg, bg = ops.fq_per_tensor(x.cuda(), sc.cuda(), zp, qmin, qmax, lsq_grad_factor=gf, want_bins=True, act="gelu")
    b3, eff3 = ops.fq_bins_only(x.cuda(), sc.cuda(), zp, qmin, qmax, lsq_grad_factor=gf, act="gelu")
    assert torch.equal(b3, bg) and torch.equal(ops.dequant_bins(b3, eff3, qmin, qmax), yg)
    want = O.fq_lsqplus_per_tensor(x, sc.clone(), z.clone(), qmin, qmax) if lsq else O.fq_per_tensor(x, float(sc), int(z), qmin, qmax)
    assert torch.equal(y.cpu(), want)


def _pair(lsq):
    from outlier_suppression_b200.quantization import quantized_module as qm
    g = torch.Generator().manual_seed(3)
    lin = torch.nn.Linear(256, 384)
    lin.weight.data = torch.randn(384, 256, generator=g) * 0.05
    a_cfg = QC("LSQPlusFakeQuantize" if lsq else "FixedFakeQuantize", "AvgMinMaxObserver", 6, False, -1)
    aq = qm.Quantizer(None, a_cfg).cuda()
    ql = qm.Quantizer(lin, QC("FixedFakeQuantize", "MinMaxObserver", 6, True, 0)).cuda()
    x = (torch.randn(4, 40, 256, generator=g) * 2).cuda()
    ql.weight_fake_quant.enable_observer(); ql.weight_fake_quant(ql.weight); ql.weight_fake_quant.disable_observer()
    aq.enable_observer(); aq(x); aq.disable_observer()
    aq.enable_fake_quant(); ql.weight_fake_quant.enable_fake_quant()
    return aq, ql, x


@pytest.mark.parametrize("lsq", [False, True])
def test_quantizer_defers_once_a_fused_linear_consumes_it(lsq, monkeypatch):
    from outlier_suppression_b200.quantization.fake_quant import LazyFakeQuant, _lazy_stats
    aq, ql, x = _pair(lsq)
    with torch.no_grad():
        monkeypatch.setenv("OSQ_DISABLE_LAZY_FQ", "1")
        y0 = aq(x); out0 = ql(y0)                       # call 1: eager, the Linear asks for bins from now on
        y1 = aq(x); out1 = ql(y1)                       # call 2: eager with bins (round-1 behaviour)
        assert not isinstance(y1, LazyFakeQuant) and torch.equal(out0, out1)
        monkeypatch.delenv("OSQ_DISABLE_LAZY_FQ")
        before = dict(_lazy_stats)
        y2 = aq(x)
        assert isinstance(y2, LazyFakeQuant) and y2.shape == x.shape and y2.dtype == torch.float32 and y2.is_cuda
        out2 = ql(y2)                                   # bins-in launch: the fp32 values were never produced
        assert torch.equal(out2, out1)
        assert _lazy_stats["deferred"] == before["deferred"] + 1 and _lazy_stats["materialized"] == before["materialized"]
        assert y2._real is None and aq._lazy_ok
        # any other consumer gets the real values, bit-identical to the eager output
        assert torch.equal(y2 + 0.0, y1) and torch.equal(y2.cpu(), y1.cpu()) and torch.equal(y2.reshape(-1, 256), y1.reshape(-1, 256))
        assert _lazy_stats["materialized"] == before["materialized"] + 1 and not aq._lazy_ok
        assert not isinstance(aq(x), LazyFakeQuant)    # a quantizer with a non-Linear consumer stops deferring


def test_materialisation_after_the_input_changed_uses_the_bins():
    from outlier_suppression_b200.quantization.fake_quant import LazyFakeQuant
    aq, ql, x = _pair(False)
    with torch.no_grad():
        ql(aq(x))
        want = aq(x) + 0.0 if not isinstance(aq(x), LazyFakeQuant) else None
        aq._lazy_ok = True
        y_eager = ops.fq_per_tensor(x, aq.scale, aq.zero_point, aq.quant_min, aq.quant_max)
        xin = x.clone()
        y = aq(xin)
        assert isinstance(y, LazyFakeQuant)
        xin.mul_(3.0)                                   # the quantizer's input is gone: the recompute route is closed
        assert torch.equal(y + 0.0, y_eager)
        # ops entry points that take an activation pointer see the real tensor, too
        aq._lazy_ok = True
        y = aq(x)
        assert isinstance(y, LazyFakeQuant)
        again = ops.fq_per_tensor(y, aq.scale, aq.zero_point, aq.quant_min, aq.quant_max)   # fake-quant is idempotent
        assert torch.equal(again, y_eager)
