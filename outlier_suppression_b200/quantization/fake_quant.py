"""Quantizer modules with the reference's names, flags, buffers / Parameters and state_dict keys
(quantization/fake_quant.py), running on the sm_100a kernels.

The forward contract is unchanged: ``q(X, observation_mask=None, seq_pos=-1)`` returns the
fake-quantized fp32 tensor (or X itself when both flags are off).  Additionally the returned
tensor is *tagged* with the quantizer that produced it so that a downstream ``QLinear`` can run the
fused fake-quant + Linear tcgen05 kernel (fake-quant is idempotent).
"""
from __future__ import annotations

import os
import weakref

import torch
import torch.nn as nn

from .. import ops
from .observer import MinMaxObserver
from . import util_quant as UQ

_TAG = "_osq_producer"
_lazy_stats = {"deferred": 0, "materialized": 0}   # how often a quantizer output stayed bins-only / had to be materialised


def producer_of(t: torch.Tensor):
    """The live quantizer whose output ``t`` is, or None.  An in-place edit after the quantizer ran voids the tag
    (the values are no longer fixed points of the fake-quant, so re-quantising them would change the result)."""
    tag = getattr(t, _TAG, None)
    if tag is None or tag[1] != t._version:
        return None
    q = tag[0]()
    if q is None or _qparam_stamp(q) != tag[2]:
        return None  # the quantizer's (scale, zero_point) changed since it produced t: t is no longer its fixed point
    return q


def bins_of(t: torch.Tensor):
    """uint8 bins the producing fake-quant launch wrote next to ``t`` (None if absent, stale or ``t`` is strided)."""
    tag = getattr(t, "_osq_bins", None)
    if tag is None or tag[1] != t._version or not t.is_contiguous() or tag[0].numel() != t.numel():
        return None
    if producer_of(t) is None:  # stale qparams void the bins together with the fusion
        return None
    return tag[0]


def _qparam_stamp(q):
    """Everything that identifies the (scale, zero_point) a quantizer used: rewritten by an observer pass or
    load_state_dict (epoch), or edited in place by an optimizer step / the LSQ+ sanitiser (tensor versions)."""
    return (q.qparam_epoch, q.scale.data_ptr(), q.scale._version, q.zero_point.data_ptr(), q.zero_point._version)


class LazyFakeQuant(torch.Tensor):
    """Deferred output of a per-tensor activation quantizer whose consumers are fused QLinears.

    The launch that produced it wrote ONLY the uint8 bins (osq_fq_per_tensor_bins_only_f32: 5 B / element instead of 9); the bins
    ride along as ``_osq_bins`` and a fused QLinear reads them without ever touching fp32 values.  It still IS the quantizer's
    output -- an fp32 tensor of the input's shape: any other consumer (an add, a print, ``.cpu()``, another kernel of this package)
    reaches ``__torch_dispatch__`` / ``ops.plain`` and gets the real values, produced on first use

      * by running the ordinary fake-quant kernel on the quantizer's input, if that tensor and the quantizer's parameters are
        untouched since (bit-identical to the eager path by construction), else
      * from the bins and the effective (scale, zero_point) the launch recorded (osq_dequant_bins_f32: bit-identical whenever
        the effective zero point is integer valued; within 1 ulp under LSQ+'s rare 1-ulp zero-point drift).

    A quantizer whose deferred output was materialised once stops deferring (``_lazy_ok``): the second pass over the input would
    cost more than the eager launch.  ``OSQ_DISABLE_LAZY_FQ=1`` turns the mechanism off."""

    _osq_lazy = True
    __torch_function__ = torch._C._disabled_torch_function_impl

    @staticmethod
    def __new__(cls, bins, eff, qmin, qmax, recompute, on_materialize):
        return torch.Tensor._make_wrapper_subclass(cls, bins.shape, strides=bins.stride(), dtype=torch.float32, device=bins.device,
                                                   requires_grad=False)

    def __init__(self, bins, eff, qmin, qmax, recompute, on_materialize):
        self._lz = (bins, eff, int(qmin), int(qmax), recompute, on_materialize)
        self._real = None

    def _osq_materialize(self) -> torch.Tensor:
        if self._real is None:
            bins, eff, qmin, qmax, recompute, on_materialize = self._lz
            y = recompute() if recompute is not None else None
            if y is None:
                y = ops.dequant_bins(bins, eff, qmin, qmax)
            self._real = y
            if on_materialize is not None:
                on_materialize()
        return self._real

    def __repr__(self):  # pragma: no cover
        return "LazyFakeQuant(%s)" % (repr(self._osq_materialize()),)

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        from torch.utils._pytree import tree_map
        unwrap = lambda t: t._osq_materialize() if isinstance(t, LazyFakeQuant) else t
        return func(*tree_map(unwrap, args), **tree_map(unwrap, kwargs or {}))


class QuantizeBase(nn.Module):
    """fake_quant.py:15-97."""

    def __init__(self, observer=MinMaxObserver, bit=8, symmetric=False, ch_axis=-1):
        super().__init__()
        self.observer = observer(bit=bit, symmetric=symmetric, ch_axis=ch_axis)
        self.bit, self.symmetric, self.ch_axis = bit, symmetric, ch_axis
        self.observer_enabled = 0
        self.fake_quant_enabled = 0
        self.quant_min, self.quant_max = self.observer.quant_min, self.observer.quant_max
        self.qparam_epoch = 0  # bumped whenever (scale, zero_point) are rewritten; keys derived caches

    def set_name(self, name):
        self.name = name

    @torch.jit.export
    def calculate_qparams(self):
        return self.observer.calculate_qparams(self.observer.min_val, self.observer.max_val)

    @torch.jit.export
    def disable_observer(self):
        self.observer_enabled = 0

    @torch.jit.export
    def enable_observer(self):
        self.observer_enabled = 1

    @torch.jit.export
    def disable_fake_quant(self):
        self.fake_quant_enabled = 0

    @torch.jit.export
    def enable_fake_quant(self):
        self.fake_quant_enabled = 1

    @torch.jit.export
    def extra_repr(self):
        return ("fake_quant_enabled={}, observer_enabled={}, symmetric={}, bit={}, ch_axis={}, quant_min={}, "
                "quant_max={}").format(self.fake_quant_enabled, self.observer_enabled, self.symmetric, self.bit,
                                       self.ch_axis, self.quant_min, self.quant_max)

    # ---- state_dict glue: scale / zero_point change shape after calibration (fake_quant.py:59-97) ----
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        destination[prefix + "scale"] = self.scale
        destination[prefix + "zero_point"] = self.zero_point

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        for name in ("scale", "zero_point"):
            key = prefix + name
            if key in state_dict:
                val = state_dict[key]
                cur = getattr(self, name)
                if isinstance(cur, nn.Parameter):
                    cur.data = torch.ones_like(val.to(cur.device))
                else:
                    cur.resize_(val.shape)
            elif strict:
                missing_keys.append(key)
        self.qparam_epoch += 1
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def _run_observer(self, X, observation_mask, seq_pos, owner):
        """One observer step; True when the launch also refreshed (scale, zero_point).  Goes through the observer's
        ``__call__`` when hooks are registered on it (the reference always does, fake_quant.py:109,180), else
        straight to the kernel wrapper."""
        obs = self.observer
        if obs._forward_hooks or obs._forward_pre_hooks:
            obs._owner_hint = owner
            try:
                obs(X, observation_mask, seq_pos)
            finally:
                obs._owner_hint = None
            return obs._last_fused
        return obs._observe(X, observation_mask, seq_pos, owner)

    def observe_many(self, Xs, observation_mask=None, seq_pos=-1, batch_indices=None) -> None:
        """Calibration on a LIST of batches: the state afterwards is what ``forward`` on each batch in order leaves (observer enabled;
        the fake-quantised outputs are not produced).  AvgPruneMinMaxObserver runs the list as one batched launch sequence
        (osq_prune_observe_many_f32); every other observer is called once per batch."""
        if self.observer_enabled != 1:
            return
        many = getattr(self.observer, "_observe_many", None)
        if many is not None and not (self.observer._forward_hooks or self.observer._forward_pre_hooks):
            fused = many([X.detach() for X in Xs], observation_mask, seq_pos, self, batch_indices)
            if not fused:
                self._refresh_qparams_from_observer()
            self.qparam_epoch += 1
            return
        shard = getattr(self.observer, "_shard", None)
        for i, X in enumerate(Xs):
            if shard is not None and batch_indices is not None:
                shard[0].set_batch(batch_indices[i])
            fq, self.fake_quant_enabled = self.fake_quant_enabled, 0
            try:
                self(X, observation_mask, seq_pos)
            finally:
                self.fake_quant_enabled = fq

    def _refresh_qparams_from_observer(self):
        _scale, _zero_point = self.observer.calculate_qparams(self.observer.min_val, self.observer.max_val)
        _scale, _zero_point = _scale.to(self.scale.device), _zero_point.to(self.zero_point.device)
        if self.scale.shape != _scale.shape:
            self.scale.data.resize_(_scale.shape)
            self.zero_point.data.resize_(_zero_point.shape)
        self.scale.data.copy_(_scale)
        self.zero_point.data.copy_(_zero_point.to(self.zero_point.dtype))

    # ---- where the observer kernels may write the refreshed qparams ----
    def _per_tensor_qparam_targets(self):
        return None, None

    def _per_channel_qparam_targets(self, rows):
        return None, None

    # set by a QLinear that consumed this quantizer's output through the fused kernel: from then on the fake-quant
    # launch also writes the uint8 bins (1 extra byte / element) and the Linear reads those instead of the fp32 tensor
    _emit_bins = False

    def _fq_per_tensor_tagged(self, X, scale, zero_point, g=0.0):
        """No-grad per-tensor fake-quant (K1) whose output carries the producer tag and, when a fused QLinear is
        known to consume it, the bins in the fused kernel's operand format."""
        if (self._emit_bins and X.is_cuda and X.dim() >= 2 and X.shape[-1] % 128 == 0 and X.is_contiguous()
                and self.quant_max - self.quant_min <= 255 and X.numel() > 0):
            if (self._lazy_ok and X.dtype == torch.float32 and not getattr(X, "_osq_lazy", False) and X.data_ptr() % 16 == 0
                    and not torch.is_grad_enabled() and os.environ.get("OSQ_DISABLE_LAZY_FQ") != "1"):
                return self._fq_deferred(X, scale, zero_point, g)
            y, bins = ops.fq_per_tensor(X, scale, zero_point, self.quant_min, self.quant_max, lsq_grad_factor=g, want_bins=True)
            self._tag(y)
            try:
                y._osq_bins = (bins, y._version)
            except Exception:  # pragma: no cover
                pass
            return y
        return self._tag(ops.fq_per_tensor(X, scale, zero_point, self.quant_min, self.quant_max, lsq_grad_factor=g))

    # cleared the first time a deferred output of this quantizer had to be materialised (a non-Linear consumer exists)
    _lazy_ok = True

    def _fq_deferred(self, X, scale, zero_point, g, act=None):
        """bins-only launch; the fp32 values exist only if somebody other than a fused QLinear asks for them (LazyFakeQuant).
        ``act="gelu"``: the activation in front of the quantizer is part of the launch (quant_bert.py:278-280)."""
        bins, eff = ops.fq_bins_only(X, scale, zero_point, self.quant_min, self.quant_max, lsq_grad_factor=g, act=act)
        me, x_ver, stamp = weakref.ref(self), X._version, _qparam_stamp(self)

        def recompute():
            q = me()
            if q is None or X._version != x_ver or _qparam_stamp(q) != stamp:
                return None      # input or parameters changed since: rebuild from the bins
            return ops.fq_per_tensor(X, scale, zero_point, q.quant_min, q.quant_max, lsq_grad_factor=g, act=act)

        def on_materialize():
            q = me()
            if q is not None:
                q._lazy_ok = False
            _lazy_stats["materialized"] += 1

        y = LazyFakeQuant(bins, eff, self.quant_min, self.quant_max, recompute, on_materialize)
        _lazy_stats["deferred"] += 1
        self._tag(y)
        y._osq_bins = (bins, y._version)
        return y

    def _tag(self, y: torch.Tensor) -> torch.Tensor:
        try:
            setattr(y, _TAG, (weakref.ref(self), y._version, _qparam_stamp(self)))
        except Exception:  # pragma: no cover  (tensor subclasses that refuse attributes)
            pass
        return y


class FixedFakeQuantize(QuantizeBase):
    """Non-learnable scale / zero_point (fake_quant.py:100-126)."""

    def __init__(self, observer, bit=8, symmetric=False, ch_axis=-1):
        super().__init__(observer, bit=bit, symmetric=symmetric, ch_axis=ch_axis)
        self.register_buffer("scale", torch.tensor([1.0], dtype=torch.float))
        self.register_buffer("zero_point", torch.tensor([0], dtype=torch.int))

    def _per_tensor_qparam_targets(self):
        if self.scale.shape != torch.Size([]):  # the reference resizes to the observer's 0-dim shape (:115-117)
            self.scale.resize_(())
            self.zero_point.resize_(())
        return self.scale, self.zero_point

    def _per_channel_qparam_targets(self, rows):
        if self.scale.numel() != rows or self.scale.dim() != 1:
            self.scale.resize_(rows)
            self.zero_point.resize_(rows)
        return self.scale, self.zero_point

    def forward(self, X, observation_mask=None, seq_pos=-1):
        if self.observer_enabled == 1 and X.numel() > 0:
            fused = self._run_observer(X.detach(), observation_mask, seq_pos, self)
            if not fused:
                _scale, _zero_point = self.observer.calculate_qparams(self.observer.min_val, self.observer.max_val)
                _scale, _zero_point = _scale.to(self.scale.device), _zero_point.to(self.zero_point.device)
                if self.scale.shape != _scale.shape:
                    self.scale.resize_(_scale.shape)
                    self.zero_point.resize_(_zero_point.shape)
                self.scale.copy_(_scale)
                self.zero_point.copy_(_zero_point)
            self.qparam_epoch += 1
        if self.fake_quant_enabled == 1:
            if self.ch_axis != -1:
                X = UQ.fake_quantize_per_channel_affine(X, self.scale, self.zero_point, self.ch_axis, self.quant_min,
                                                        self.quant_max) if not _needs_grad(X) else \
                    _ste_per_channel(X, self.scale, self.zero_point, self.ch_axis, self.quant_min, self.quant_max)
            elif _needs_grad(X):
                X = _SteFixedPerTensor.apply(X, self.scale, self.zero_point, self.quant_min, self.quant_max)
            else:
                X = self._fq_per_tensor_tagged(X, self.scale, self.zero_point)
        return X


def _needs_grad(x):
    return torch.is_grad_enabled() and x.requires_grad


class _SteFixedPerTensor(torch.autograd.Function):
    """straight-through gradient of util_quant.py:11-15: dy where the un-clamped bin is inside [qmin, qmax]."""

    @staticmethod
    def forward(ctx, x, scale, zero_point, qmin, qmax):
        ctx.save_for_backward(x, scale, zero_point)
        ctx.rng = (qmin, qmax)
        return ops.fq_per_tensor(x, scale, zero_point, qmin, qmax)

    @staticmethod
    def backward(ctx, dy):
        x, scale, zero_point = ctx.saved_tensors
        qmin, qmax = ctx.rng
        v = torch.round(x / scale.reshape(())) + zero_point.reshape(()).to(x.dtype)
        return dy * ((v >= qmin) & (v <= qmax)).to(dy.dtype), None, None, None, None


def _ste_per_channel(x, scale, zero_point, ch_axis, qmin, qmax):
    shape = [1] * x.dim()
    shape[ch_axis] = x.shape[ch_axis]
    s, z = scale.reshape(shape), zero_point.reshape(shape)
    q = torch.clamp(UQ.round_ste(x / s) + z, qmin, qmax)
    return (q - z) * s


class LSQPlusFakeQuantize(QuantizeBase):
    """Learnable scale AND zero point (fake_quant.py:170-209); the fine stage of token-wise clipping
    feeds ``scale`` / ``zero_point`` (fp32 Parameters of shape [1]) to Adam."""

    def __init__(self, observer, bit=8, symmetric=False, ch_axis=-1, use_grad_scaling=True):
        super().__init__(observer, bit=bit, symmetric=symmetric, ch_axis=ch_axis)
        self.scale = torch.nn.Parameter(torch.tensor([1.0], dtype=torch.float))
        self.zero_point = torch.nn.Parameter(torch.tensor([0.0], dtype=torch.float))
        self.register_buffer("eps", torch.tensor([torch.finfo(torch.float32).eps]))
        self.use_grad_scaling = use_grad_scaling

    def _per_tensor_qparam_targets(self):
        return self.scale.data, self.zero_point.data  # shape [1] is kept (copy_ broadcasts, :186-187)

    def grad_factor(self, X):
        """fake_quant.py:193-207."""
        if not self.use_grad_scaling:
            return 1.0
        n = X.numel() if self.ch_axis == -1 else X.numel() / X.shape[self.ch_axis]
        return 1.0 / (n * self.quant_max) ** 0.5

    def forward(self, X, observation_mask=None, seq_pos=-1):
        sanitized_by_kernel = False
        if self.observer_enabled == 1 and X.numel() > 0:
            fused = self._run_observer(X.detach(), observation_mask, seq_pos, self)
            if not fused:
                _scale, _zero_point = self.observer.calculate_qparams(self.observer.min_val, self.observer.max_val)
                _scale, _zero_point = _scale.to(self.scale.device), _zero_point.to(self.zero_point.device)
                if self.ch_axis != -1:
                    self.scale.data = torch.ones_like(_scale, dtype=torch.float32)
                    self.zero_point.data = torch.zeros_like(_zero_point.float())
                self.scale.data.copy_(_scale)
                self.zero_point.data.copy_(_zero_point.float())
            self.qparam_epoch += 1
        elif self.fake_quant_enabled == 1 and self.ch_axis == -1 and X.is_cuda:
            sanitized_by_kernel = True  # K1 / K6 apply fake_quant.py:188-191 in place
        else:
            self.scale.data.abs_()
            self.scale.data.clamp_(min=float(torch.finfo(torch.float32).eps))
            self.zero_point.data.clamp_(self.quant_min, self.quant_max)

        if self.fake_quant_enabled == 1:
            g = self.grad_factor(X)
            if self.ch_axis != -1:
                X = UQ.fake_quantize_learnableplus_per_channel_affine_training(X, self.scale, self.zero_point, self.ch_axis,
                                                                              self.quant_min, self.quant_max, g)
            elif X.is_cuda and not (torch.is_grad_enabled() and (X.requires_grad or self.scale.requires_grad
                                                                   or self.zero_point.requires_grad)):
                X = self._fq_per_tensor_tagged(X, self.scale.detach(), self.zero_point.detach(), float(g))
            else:
                X = UQ.fake_quantize_learnableplus_per_tensor_affine_training(X, self.scale, self.zero_point,
                                                                             self.quant_min, self.quant_max, g)
                if not X.requires_grad:
                    self._tag(X)
        del sanitized_by_kernel
        return X


class LSQFakeQuantize(QuantizeBase):
    """Learnable scale, symmetric (fake_quant.py:129-167).  Not used by any shipped config: kept on
    torch ops (reference formulas) so the registry key resolves."""

    def __init__(self, observer, bit=8, symmetric=False, ch_axis=-1, use_grad_scaling=True):
        super().__init__(observer, bit=bit, symmetric=symmetric, ch_axis=ch_axis)
        self.scale = torch.nn.Parameter(torch.tensor([1.0], dtype=torch.float))
        self.register_buffer("zero_point", torch.tensor([0], dtype=torch.int))
        self.register_buffer("eps", torch.tensor([torch.finfo(torch.float32).eps]))
        self.use_grad_scaling = use_grad_scaling

    def forward(self, X, observation_mask=None, seq_pos=-1):
        if self.observer_enabled == 1 and X.numel() > 0:
            self._run_observer(X.detach(), observation_mask, seq_pos, None)
            _scale, _zero_point = self.observer.calculate_qparams(self.observer.min_val, self.observer.max_val)
            if self.ch_axis != -1:
                self.scale.data = torch.ones_like(_scale, dtype=torch.float32)
                self.zero_point.resize_(_zero_point.shape)
            self.scale.data.copy_(_scale)
            self.zero_point.copy_(_zero_point)
            self.qparam_epoch += 1
        else:
            self.scale.data.abs_()
            self.scale.data.clamp_(min=float(torch.finfo(torch.float32).eps))
        if self.fake_quant_enabled == 1:
            n = X.numel() if self.ch_axis == -1 else X.numel() / X.shape[self.ch_axis]
            g = 1.0 / (n * self.quant_max) ** 0.5 if self.use_grad_scaling else 1.0
            if self.ch_axis != -1:
                X = UQ.fake_quantize_learnable_per_channel_affine_training(X, self.scale, self.zero_point.int(), self.ch_axis,
                                                                          self.quant_min, self.quant_max, g)
            else:
                X = UQ.fake_quantize_learnable_per_tensor_affine_training(X, self.scale, self.zero_point.reshape(()).to(X.dtype),
                                                                         self.quant_min, self.quant_max, g)
        return X
