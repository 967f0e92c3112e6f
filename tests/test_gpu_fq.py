"""-m gpu: CUDA fake-quant kernels (through the C ABI) vs golden vectors of the reference and vs the
CPU oracle on seeded inputs.  Integer bins and dequantised floats are BIT-EXACT (NaN pattern included)."""
import numpy as np
import pytest
import torch

from oracle import osq_oracle as O

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def same(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    np.testing.assert_array_equal(a, b)


def dev(v, dtype):
    return torch.tensor([v], dtype=dtype, device="cuda")


def test_fq_per_tensor_golden_bit_exact(golden):
    from outlier_suppression_b200 import ops
    g = golden("fq_per_tensor")
    for i in range(int(g["n"])):
        scale, zp, qmin, qmax = g["p%d" % i]
        x = T(g["x%d" % i]).cuda()
        y, codes = ops.fq_per_tensor(x, dev(scale, torch.float32), dev(int(zp), torch.int32), int(qmin), int(qmax),
                                     want_codes=True)
        same(y, g["y%d" % i])
        q = g["q%d" % i]
        fin = np.isfinite(q)
        same(codes.cpu().numpy()[fin], q[fin].astype(np.int16))
        # unaligned views exercise the scalar head / tail
        y2 = ops.fq_per_tensor(x[1:-2], dev(scale, torch.float32), dev(float(zp), torch.float32), int(qmin), int(qmax))
        same(y2, g["y%d" % i][1:-2])


def test_fq_per_channel_golden_bit_exact(golden):
    from outlier_suppression_b200 import ops
    from outlier_suppression_b200.quantization import util_quant as UQ
    g = golden("fq_per_channel")
    for i in range(int(g["n"])):
        bit, sym, qmin, qmax = (int(v) for v in g["p%d" % i])
        scale, zp = T(g["scale%d" % i]).cuda(), T(g["zp%d" % i]).cuda()
        same(ops.fq_per_channel(T(g["w%d" % i]).cuda(), scale, zp, qmin, qmax), g["y%d" % i])
        same(UQ.fake_quantize_per_channel_affine(T(g["w2_%d" % i]).cuda(), scale, zp, 0, qmin, qmax), g["y2_%d" % i])
        # ch_axis = 1 goes through the transposed view
        yt = UQ.fake_quantize_per_channel_affine(T(g["w%d" % i]).cuda().t().contiguous(), scale, zp, 1, qmin, qmax)
        same(yt.t(), g["y%d" % i])


def test_lsqplus_forward_golden_and_inplace_sanitize(golden):
    from outlier_suppression_b200.quantization.quantized_module import Quantizer
    from tests.test_host_logic import QC
    g = golden("lsqplus")
    for i in range(int(g["n"])):
        bit, qmin, qmax = (int(v) for v in g["p%d" % i])
        m = Quantizer(None, QC("LSQPlusFakeQuantize", "AvgPruneMinMaxObserver", bit, False, -1)).cuda()
        m.scale.data.fill_(float(g["scale_in%d" % i]))
        m.zero_point.data.fill_(float(g["zp_in%d" % i]))
        m.enable_fake_quant()
        with torch.no_grad():
            y = m(T(g["x%d" % i]).cuda())
        same(y, g["y%d" % i])
        same(m.scale.data, g["scale_after%d" % i])
        same(m.zero_point.data, g["zp_after%d" % i])


def test_lsqplus_backward_golden(golden):
    from outlier_suppression_b200.quantization.quantized_module import Quantizer
    from tests.test_host_logic import QC
    g = golden("lsqplus")
    m = Quantizer(None, QC("LSQPlusFakeQuantize", "AvgPruneMinMaxObserver", 6, False, -1)).cuda()
    m.scale.data.fill_(0.09)
    m.zero_point.data.fill_(29.6)
    m.enable_fake_quant()
    x = T(g["gx_x"]).cuda().requires_grad_(True)
    y = m(x)
    same(y, g["gx_y"])
    (y * T(g["gx_gy"]).cuda()).sum().backward()
    same(x.grad, g["gx_dx"])  # (dy*s')/s' replicated literally
    np.testing.assert_allclose(m.scale.grad.cpu().numpy(), g["gx_dscale"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(m.zero_point.grad.cpu().numpy(), g["gx_dzp"], rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("bit,sym", [(4, False), (6, False), (8, False), (6, True)])
def test_fq_per_tensor_large_vs_oracle(bit, sym):
    from outlier_suppression_b200 import ops
    a, _ = O.synth_activation(8, 512, 768, seed=bit)
    qmin, qmax = O.quant_range(bit, sym)
    mn, mx = O.global_minmax(a)
    scale, zp = O.qparams_from_minmax(mn * 0.6, mx * 0.6, qmin, qmax, sym)  # clip so both clamp edges are hit
    ref = O.fq_per_tensor(a, scale.item(), int(zp.item()), qmin, qmax)
    y, codes = ops.fq_per_tensor(a.cuda(), scale.reshape(1).cuda(), zp.reshape(1).to(torch.int32).cuda(), qmin, qmax, want_codes=True)
    same(y, ref)
    same(codes, O.fq_bins(a, scale.item(), int(zp.item()), qmin, qmax).to(torch.int16))
    # permuted-but-dense view (the q / k^T call sites): no copy, same values, same strides as the reference's output
    v = a.cuda().view(8, 512, 12, 64).permute(0, 2, 1, 3)
    yv = ops.fq_per_tensor(v, scale.reshape(1).cuda(), zp.reshape(1).to(torch.int32).cuda(), qmin, qmax)
    assert yv.stride() == v.stride()
    same(yv, ref.view(8, 512, 12, 64).permute(0, 2, 1, 3))
    # idempotence: fq(fq(x)) == fq(x) (what lets QLinear re-derive the bins from a tagged activation)
    same(ops.fq_per_tensor(y, scale.reshape(1).cuda(), zp.reshape(1).to(torch.int32).cuda(), qmin, qmax), ref)


def test_fq_empty_and_tiny():
    from outlier_suppression_b200 import ops
    s, z = dev(0.1, torch.float32), dev(3, torch.int32)
    assert ops.fq_per_tensor(torch.empty(0, 7, device="cuda"), s, z, 0, 63).shape == (0, 7)
    x = torch.tensor([0.26], device="cuda")
    same(ops.fq_per_tensor(x, s, z, 0, 63), O.fq_per_tensor(x.cpu(), s.item(), 3, 0, 63))


def test_calc_qparams_kernel_golden(golden):
    from outlier_suppression_b200 import ops
    g = golden("qparams")
    for bit in (4, 6, 8):
        for sym in (False, True):
            qmin, qmax = O.quant_range(bit, sym)
            s, z = ops.calc_qparams(T(g["mins"]).cuda(), T(g["maxs"]).cuda(), qmin, qmax, sym)
            same(s, g["s_%d_%d" % (bit, sym)])
            same(z, g["z_%d_%d" % (bit, sym)])


@pytest.mark.parametrize("lsq", [False, True])
@pytest.mark.parametrize("shape", [(4, 128, 3072), (3, 77, 130), (5,)])
def test_gelu_then_quantizer_as_one_pass(shape, lsq):
    """K1c (osq_act_fq_per_tensor_bins_f32): intermediate_act_fn + its quantizer (quant_bert.py:278-280) in one elementwise launch
    must equal torch's CUDA GELU followed by K1b bit for bit -- tensor and bins."""
    from outlier_suppression_b200 import ops
    g = torch.Generator().manual_seed(len(shape) * 7 + int(lsq))
    x = (torch.randn(*shape, generator=g) * 3).cuda()
    z = torch.nn.functional.gelu(x)
    scale = torch.tensor([float(z.max() - z.min()) / 63 * 0.8], device="cuda")
    if lsq:
        zp, gf = torch.tensor([3.3], device="cuda"), 1.0 / (x.numel() * 63) ** 0.5
    else:
        zp, gf = torch.tensor([3], dtype=torch.int32, device="cuda"), 0.0
    ref_y, ref_b = ops.fq_per_tensor(z, scale.clone(), zp.clone(), 0, 63, lsq_grad_factor=gf, want_bins=True)
    got_y, got_b = ops.fq_per_tensor(x, scale.clone(), zp.clone(), 0, 63, lsq_grad_factor=gf, want_bins=True, act="gelu")
    np.testing.assert_array_equal(got_y.cpu().numpy(), ref_y.cpu().numpy())
    np.testing.assert_array_equal(got_b.cpu().numpy(), ref_b.cpu().numpy())
    only_y = ops.fq_per_tensor(x, scale.clone(), zp.clone(), 0, 63, lsq_grad_factor=gf, act="gelu")
    np.testing.assert_array_equal(only_y.cpu().numpy(), ref_y.cpu().numpy())
