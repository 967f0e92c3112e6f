"""Functional fake-quant API with the reference's names (quantization/util_quant.py:1-71), backed
by the sm_100a kernels.  ``scale`` / ``zero_point`` may be Python numbers (reference call style,
fake_quant.py:123-125) or device tensors (preferred: no host sync)."""
from __future__ import annotations

import torch

from .. import ops


def _dev_scalar(v, device, dtype):
    if torch.is_tensor(v):
        return v.detach().to(device=device, dtype=dtype).reshape(1)
    return torch.tensor([v], dtype=dtype, device=device)


def round_ste(x: torch.Tensor) -> torch.Tensor:
    """util_quant.py:4-8: value rint(x), gradient 1."""
    return (x.round() - x).detach() + x


def grad_scale(t: torch.Tensor, scale: float) -> torch.Tensor:
    """util_quant.py:70-71: value t (up to 1 ulp), gradient scaled by ``scale``."""
    return (t - (t * scale)).detach() + (t * scale)


def fake_quantize_per_tensor_affine(x, scale, zero_point, quant_min, quant_max):
    """util_quant.py:11-15."""
    zp_dtype = torch.float32 if (torch.is_tensor(zero_point) and zero_point.is_floating_point()) or isinstance(zero_point, float) else torch.int32
    return ops.fq_per_tensor(x, _dev_scalar(scale, x.device, torch.float32), _dev_scalar(zero_point, x.device, zp_dtype),
                             quant_min, quant_max)


def fake_quantize_per_channel_affine(x, scale, zero_point, ch_axis, quant_min, quant_max):
    """util_quant.py:18-26 (any ch_axis; the kernel runs on the ch_axis-major view)."""
    if ch_axis < 0:
        ch_axis += x.dim()
    if ch_axis == 0:
        return ops.fq_per_channel(x, scale, zero_point, quant_min, quant_max)
    y = ops.fq_per_channel(x.movedim(ch_axis, 0).contiguous(), scale, zero_point, quant_min, quant_max)
    return y.movedim(0, ch_axis)


class _LsqPlusPerTensor(torch.autograd.Function):
    """Forward: K1 with the LSQ+ effective parameters; backward: osq_lsqplus_backward_f32
    (the gradients autograd derives from util_quant.py:48-55)."""

    @staticmethod
    def forward(ctx, x, scale, zero_point, quant_min, quant_max, grad_factor):
        y = ops.fq_per_tensor(x, scale.detach(), zero_point.detach(), quant_min, quant_max, lsq_grad_factor=grad_factor)
        ctx.save_for_backward(x, scale, zero_point)
        ctx.cfg = (quant_min, quant_max, grad_factor)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, scale, zero_point = ctx.saved_tensors
        quant_min, quant_max, g = ctx.cfg
        dx, dscale, dzp = ops.lsqplus_backward(x.contiguous(), dy.contiguous(), scale.detach(), zero_point.detach(), g,
                                               quant_min, quant_max)
        return dx, dscale.reshape(scale.shape), dzp.reshape(zero_point.shape), None, None, None


def fake_quantize_learnableplus_per_tensor_affine_training(x, scale, zero_point, quant_min, quant_max, grad_factor):
    """util_quant.py:48-55.  scale / zero_point: fp32 tensors of shape [1] (nn.Parameter)."""
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or scale.requires_grad or zero_point.requires_grad)
    if needs_grad:
        return _LsqPlusPerTensor.apply(x, scale, zero_point, quant_min, quant_max, float(grad_factor))
    return ops.fq_per_tensor(x, scale.detach(), zero_point.detach(), quant_min, quant_max, lsq_grad_factor=float(grad_factor))


# ---- variants that no shipped config reaches: reference formulas on torch ops (out of the CUDA scope) ----
def _bshape(x, ch_axis):
    shape = [1] * x.dim()
    shape[ch_axis] = x.shape[ch_axis]
    return shape


def fake_quantize_learnable_per_tensor_affine_training(x, scale, zero_point, quant_min, quant_max, grad_factor):
    """util_quant.py:29-34 (LSQ, symmetric QAT variant)."""
    s = grad_scale(scale, grad_factor)
    q = torch.clamp(round_ste(x / s) + zero_point, quant_min, quant_max)
    return (q - zero_point) * s


def fake_quantize_learnable_per_channel_affine_training(x, scale, zero_point, ch_axis, quant_min, quant_max, grad_factor):
    """util_quant.py:37-45."""
    s = grad_scale(scale, grad_factor).reshape(_bshape(x, ch_axis))
    z = zero_point.reshape(_bshape(x, ch_axis))
    q = torch.clamp(round_ste(x / s) + z, quant_min, quant_max)
    return (q - z) * s


def fake_quantize_learnableplus_per_channel_affine_training(x, scale, zero_point, ch_axis, quant_min, quant_max, grad_factor):
    """util_quant.py:58-67."""
    z = round_ste(zero_point)
    s = grad_scale(scale, grad_factor).reshape(_bshape(x, ch_axis))
    z = grad_scale(z, grad_factor).reshape(_bshape(x, ch_axis))
    q = torch.clamp(round_ste(x / s) + z, quant_min, quant_max)
    return (q - z) * s
