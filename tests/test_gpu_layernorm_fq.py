"""-m gpu: K7, GammaResidual + LayerNorm + the LayerNorm's output quantizer in one pass (osq_residual_layernorm_fq_f32;
model/util_layernorm.py:14-17 / :34-37 / :41-52), through the C ABI.

The quantizer half is bit-exact: (y, bins) must equal K1 applied to the LayerNorm output the kernel itself produced
(`ln_out`), which in turn is bit-exact with the oracle's fake-quant of that tensor.  The LayerNorm half is plain fp32:
within 4e-6 * max|ln| of an fp64 LayerNorm and of torch's own CUDA kernel; against the full CPU oracle chain
(util_layernorm.py on torch CPU, then oracle fake-quant) a bin may flip only where ln / s sits on a rounding tie."""
import pytest
import torch

from oracle import osq_oracle as O
from outlier_suppression_b200 import ops

pytestmark = pytest.mark.gpu


def _inputs(rows, H, seed, gamma, affine, split):
    g = torch.Generator().manual_seed(seed)
    h = torch.randn(rows, H, generator=g)
    res = torch.randn(rows, H, generator=g) * 2.0
    res[:, : min(6, H)] *= 30.0          # outlier channels, like LayerNorm inputs of BERT
    gm = (torch.rand(H, generator=g) * 2.0 + 0.2) if gamma else None
    w = (torch.rand(H, generator=g) + 0.5) if affine else None
    b = torch.randn(H, generator=g) * 0.1 if (affine or split) else None
    return h, res, gm, w, b


def _ref_ln(h, res, gm, w, b, eps, dtype):
    u = (res * gm if gm is not None else res) + h
    u = u.to(dtype)
    y = torch.nn.functional.layer_norm(u, (u.shape[-1],), None, None, eps)
    if w is not None:
        y = y * w.to(dtype)
    if b is not None:
        y = y + b.to(dtype)
    return y


CASES = [
    # rows, H, gamma, affine, split, lsq
    (1000, 768, True, False, True, True),     # config 2 after gamma migration: split LayerNorm + gamma residual, LSQ+
    (1000, 768, False, True, False, False),   # config 1: affine LayerNorm, FixedFakeQuantize
    (513, 1024, True, True, False, True),     # BART-large width
    (77, 256, False, False, False, True),
    (300, 3072, False, True, False, False),   # generic path (row does not fit the registers)
    (65, 100, True, True, False, True),       # generic path, H % 128 != 0
    (1, 128, False, True, False, False),
]


@pytest.mark.parametrize("rows,H,gamma,affine,split,lsq", CASES)
def test_residual_layernorm_fq(rows, H, gamma, affine, split, lsq):
    h, res, gm, w, b = _inputs(rows, H, rows + H, gamma, affine, split)
    eps = 1e-12
    qmin, qmax = 0, 63
    ln64 = _ref_ln(h, res, gm, w, b, eps, torch.float64)
    lo, hi = float(ln64.min()), float(ln64.max())
    scale = torch.tensor([(hi - lo) / 63 * 0.7], dtype=torch.float32)   # 0.7: some values clamp
    zp_f = torch.tensor([round(-lo / float(scale))], dtype=torch.float32).clamp(qmin, qmax)
    zp = zp_f if lsq else zp_f.to(torch.int32)
    g = 1.0 / (h.numel() * qmax) ** 0.5 if lsq else 0.0
    dev = lambda t: None if t is None else t.cuda()
    y, bins, ln = ops.residual_layernorm_fq(dev(h), dev(res), dev(gm), dev(w), dev(b), eps, scale.cuda(), zp.cuda(), qmin, qmax,
                                            lsq_grad_factor=g, want_bins=True, want_ln=True)
    # LayerNorm half: fp32 accuracy against fp64 and against torch's CUDA kernel
    tol = 4e-6 * float(ln64.abs().max())
    assert float((ln.double().cpu() - ln64).abs().max()) <= tol
    ln_torch = _ref_ln(dev(h), dev(res), dev(gm), dev(w), dev(b), eps, torch.float32)
    assert float((ln - ln_torch).abs().max()) <= tol
    # quantizer half: bit-exact with K1 on the same tensor, and with the oracle's fake-quant of it
    y1, b1 = ops.fq_per_tensor(ln, scale.cuda(), zp.cuda(), qmin, qmax, lsq_grad_factor=g, want_bins=True)
    assert torch.equal(y, y1) and torch.equal(bins, b1)
    if lsq:
        y_or = O.fq_lsqplus_per_tensor(ln.cpu(), scale.clone(), zp_f.clone(), qmin, qmax)
    else:
        y_or = O.fq_per_tensor(ln.cpu(), float(scale), int(zp_f), qmin, qmax)
    assert torch.equal(y.cpu(), y_or)
    # no-bins / no-ln call returns the same y
    y2 = ops.residual_layernorm_fq(dev(h), dev(res), dev(gm), dev(w), dev(b), eps, scale.cuda(), zp.cuda(), qmin, qmax, lsq_grad_factor=g)
    assert torch.equal(y2, y)
    # full CPU oracle chain: bins may differ only by one step and only rarely (ties)
    ln_cpu = _ref_ln(h, res, gm, w, b, eps, torch.float32)
    y_cpu = O.fq_lsqplus_per_tensor(ln_cpu, scale.clone(), zp_f.clone(), qmin, qmax) if lsq else O.fq_per_tensor(ln_cpu, float(scale), int(zp_f), qmin, qmax)
    diff = (y.cpu() - y_cpu).abs()
    assert float(diff.max()) <= float(scale) * 1.001 + tol
    assert float((diff > tol).float().mean()) <= 2e-3


def test_without_residual_and_error_paths():
    h = torch.randn(64, 768).cuda()
    sc, zp = torch.tensor([0.05]).cuda(), torch.tensor([31.0]).cuda()
    y, bins, ln = ops.residual_layernorm_fq(h, None, None, None, None, 1e-5, sc, zp, 0, 63, want_bins=True, want_ln=True)
    assert float((ln - torch.nn.functional.layer_norm(h, (768,), None, None, 1e-5)).abs().max()) <= 4e-6 * float(ln.abs().max())
    y1, b1 = ops.fq_per_tensor(ln, sc, zp, 0, 63, want_bins=True)
    assert torch.equal(y, y1) and torch.equal(bins, b1)
    with pytest.raises(Exception):
        ops.residual_layernorm_fq(h[:, :766].contiguous(), None, None, None, None, 1e-5, sc, zp, 0, 63)
    with pytest.raises(Exception):
        ops.residual_layernorm_fq(h.cpu(), None, None, None, None, 1e-5, sc, zp, 0, 63)
