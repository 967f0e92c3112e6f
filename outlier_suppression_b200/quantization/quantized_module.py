"""Operator wrappers, registries and the ``Quantizer`` factory with the reference's names
(quantization/quantized_module.py).  ``QLinear.forward`` is where the fused fake-quant + Linear
tcgen05 kernel sits behind the unchanged per-module call."""
from __future__ import annotations

import os

import torch  # noqa: F401
import torch.nn.functional as F
from torch import nn

from .. import ops
from .fake_quant import FixedFakeQuantize, LSQFakeQuantize, LSQPlusFakeQuantize, bins_of, producer_of
from .observer import (AvgMinMaxObserver, AvgMSEFastObserver, AvgMSEObserver, AvgPruneMinMaxObserver,
                       AvgQuantileObserver, LSQPlusObserver, MinMaxObserver, MSEFastObserver, MSEObserver)

ObserverDict = {
    "MinMaxObserver": MinMaxObserver,
    "AvgMinMaxObserver": AvgMinMaxObserver,
    "MSEObserver": MSEObserver,
    "AvgMSEObserver": AvgMSEObserver,
    "MSEFastObserver": MSEFastObserver,
    "AvgMSEFastObserver": AvgMSEFastObserver,
    "AvgQuantileObserver": AvgQuantileObserver,
    "LSQPlusObserver": LSQPlusObserver,
    "AvgPruneMinMaxObserver": AvgPruneMinMaxObserver,
}

FakeQuantizeDict = {
    "FixedFakeQuantize": FixedFakeQuantize,
    "LSQFakeQuantize": LSQFakeQuantize,
    "LSQPlusFakeQuantize": LSQPlusFakeQuantize,
}

# statistics for tests / benches: which path QLinear.forward took
stats = {"fused": 0, "unfused": 0, "grouped_launch": 0, "grouped_hit": 0, "bins_in": 0, "epilogue_fused": 0}


def _take_bins(input, aq):
    """Bins the producing quantizer left next to ``input`` (and a request for them from now on)."""
    aq._emit_bins = os.environ.get("OSQ_DISABLE_BINS") != "1"
    bins = bins_of(input)
    if bins is not None:
        stats["bins_in"] += 1
    return bins


class QuantizedModule(nn.Module):
    def __init__(self, backend="academic"):
        super().__init__()
        self.backend = backend


class QuantizedOperator:
    """Mixin: derived-state cache for the weight side (never serialised)."""

    def _weight_key(self):
        wq = self.weight_fake_quant
        return (self.weight.data_ptr(), self.weight._version, tuple(self.weight.shape), wq.qparam_epoch,
                wq.scale.data_ptr(), wq.scale._version, wq.zero_point._version, wq.quant_min, wq.quant_max)

    def invalidate_packed(self):
        """Drop every weight-derived cache.  Call after mutating ``weight.data`` outside torch's
        version tracking (gamma_migration.py:46-76 does ``w.weight.data *= gamma``); state.py's
        togglers and the weight observer call it automatically."""
        self._packed = None
        self._fq_weight_cache = None
        grp = getattr(self, "_sibling_group", None)
        if grp is not None:
            grp._packed = None
            grp._pending = None

    def _weight_is_static(self):
        wq = self.weight_fake_quant
        return (wq.fake_quant_enabled == 1 and wq.observer_enabled == 0 and isinstance(wq, FixedFakeQuantize)
                and not (torch.is_grad_enabled() and self.weight.requires_grad))

    def _cached_fq_weight(self):
        """fake-quantized weight, recomputed only when the weight / its qparams changed
        (the reference re-runs it on every forward, quantized_module.py:72,98)."""
        key = self._weight_key()
        c = getattr(self, "_fq_weight_cache", None)
        if c is None or c[0] != key:
            c = (key, self.weight_fake_quant(self.weight).detach())
            self._fq_weight_cache = c
        return c[1]


class QConv2d(QuantizedOperator, nn.Conv2d):
    """quantized_module.py:38-57."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, padding_mode,
                 w_qconfig):
        super().__init__(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size, stride=stride,
                         padding=padding, dilation=dilation, groups=groups, bias=bias, padding_mode=padding_mode)
        self.weight_fake_quant = WeightQuantizer(w_qconfig)

    def forward(self, input):
        w = self._cached_fq_weight() if self._weight_is_static() else self.weight_fake_quant(self.weight)
        return self._conv_forward(input, w, self.bias)


class QLinear(QuantizedOperator, nn.Linear):
    """quantized_module.py:60-72."""

    def __init__(self, in_features, out_features, bias, w_qconfig):
        super().__init__(in_features=in_features, out_features=out_features, bias=bias)
        self.weight_fake_quant = WeightQuantizer(w_qconfig)
        self._sibling_group = None  # set by group_sibling_linears()

    # ---- fused path ----
    def _fusable_producer(self, input):
        if os.environ.get("OSQ_DISABLE_FUSION") == "1":
            return None
        if not (input.is_cuda and input.dtype == torch.float32 and input.dim() >= 2):
            return None
        aq = producer_of(input)
        if aq is None or aq.fake_quant_enabled != 1 or aq.ch_axis != -1 or aq.quant_max - aq.quant_min > 255:
            return None
        wq = self.weight_fake_quant
        if not self._weight_is_static() or wq.ch_axis != 0 or wq.quant_max - wq.quant_min > 255:
            return None
        # the packed operand is s8 (q - zp): always representable up to 7 bits, at 8 bits only for the symmetric
        # range (zp == 0).  Asymmetric 8-bit weights take the unfused path (osq_pack_weight_s8 rejects them too).
        if wq.quant_max - wq.quant_min > 127 and not (wq.symmetric and wq.quant_min >= -128 and wq.quant_max <= 127):
            return None
        if torch.is_grad_enabled() and (input.requires_grad or aq.scale.requires_grad or
                                        (self.bias is not None and self.bias.requires_grad)):
            return None  # training / learn_scale stage: reference semantics through autograd
        if not ops.fused_linear_supported(self.in_features, self.out_features, input):
            return None
        return aq

    def _packed_weight(self):
        key = self._weight_key()
        c = getattr(self, "_packed", None)
        if c is None or c[0] != key:
            wq = self.weight_fake_quant
            codes, rowsum = ops.pack_weight(self.weight, wq.scale, wq.zero_point, wq.quant_min, wq.quant_max)
            c = (key, codes, rowsum, wq.scale.detach().float().contiguous())
            self._packed = c
        return c[1], c[2], c[3]

    def forward(self, input):
        aq = self._fusable_producer(input)
        if aq is not None:
            group = getattr(self, "_sibling_group", None)
            if group is not None:
                out = group.forward_member(self, input, aq)
                if out is not None:
                    return out
            codes, rowsum, w_scale = self._packed_weight()
            g = aq.grad_factor(input) if isinstance(aq, LSQPlusFakeQuantize) else 0.0
            stats["fused"] += 1
            bins = _take_bins(input, aq)
            return ops.fused_fq_linear(input, aq.scale.detach(), aq.zero_point.detach(), aq.quant_min, aq.quant_max,
                                       codes, w_scale, rowsum, self.bias, lsq_grad_factor=g, a_bins=bins)
        stats["unfused"] += 1
        w = self._cached_fq_weight() if self._weight_is_static() else self.weight_fake_quant(self.weight)
        return F.linear(input, w, self.bias)


class QLinearGroup:
    """Sibling QLinears that consume the SAME quantized activation -- BERT / RoBERTa ``query | key | value``
    (quant_bert.py / quant_roberta.py self-attention), BART ``q_proj | k_proj | v_proj`` -- served by ONE fused
    launch over their concatenated packed weights: the activation is read from HBM and quantised once instead of
    once per sibling.  The first sibling called with a tensor launches and hands the others views of the shared
    ``[..., sum(N_i)]`` output when they are called with the very same tensor object; every output element is
    bit-identical to the ungrouped launch (same integer accumulator, same per-column constants).  Not a module on
    purpose: it must not show up in ``state_dict`` / ``named_modules``."""

    def __init__(self, members):
        self.members = list(members)
        self._packed = None
        self._pending = None  # (input tensor, its version, {id(member): output view})
        self.misses = 0       # grouped launches whose sibling outputs were never collected

    def _packed_weight(self):
        key = tuple((m._weight_key(), None if m.bias is None else (m.bias.data_ptr(), m.bias._version)) for m in self.members)
        c = self._packed
        if c is None or c[0] != key:
            parts = [m._packed_weight() for m in self.members]
            dev = parts[0][0].device
            bias = None
            if any(m.bias is not None for m in self.members):
                bias = torch.cat([m.bias.detach().float() if m.bias is not None else
                                  torch.zeros(m.out_features, device=dev) for m in self.members]).contiguous()
            c = (key, torch.cat([p[0] for p in parts]).contiguous(), torch.cat([p[1] for p in parts]).contiguous(),
                 torch.cat([p[2] for p in parts]).contiguous(), bias)
            self._packed = c
        return c[1:]

    def forward_member(self, member, input, aq):
        pend = self._pending
        if pend is not None and pend[0] is input and pend[1] == input._version and id(member) in pend[2]:
            out = pend[2].pop(id(member))
            if not pend[2]:
                self._pending = None
            stats["grouped_hit"] += 1
            return out
        if pend is not None and pend[2]:
            # the previous grouped launch computed siblings nobody asked for with that tensor (different inputs per
            # sibling, or k/v served from a cache): this group's members do not share an input -> stop grouping it
            self.misses += 1
        self._pending = None
        if os.environ.get("OSQ_DISABLE_GROUPING") == "1" or self.misses >= 2:
            return None
        for m in self.members:  # every sibling must be on the fused path with this very producer
            if m is not member and m._fusable_producer(input) is not aq:
                return None
        if not ops.fused_linear_supported(member.in_features, sum(m.out_features for m in self.members)):
            return None
        codes, rowsum, w_scale, bias = self._packed_weight()
        g = aq.grad_factor(input) if isinstance(aq, LSQPlusFakeQuantize) else 0.0
        y = ops.fused_fq_linear(input, aq.scale.detach(), aq.zero_point.detach(), aq.quant_min, aq.quant_max,
                                codes, w_scale, rowsum, bias, lsq_grad_factor=g, a_bins=_take_bins(input, aq))
        stats["grouped_launch"] += 1
        stats["fused"] += 1
        outs, col = {}, 0
        for m in self.members:
            outs[id(m)] = y[..., col:col + m.out_features]
            col += m.out_features
        mine = outs.pop(id(member))
        self._pending = (input, input._version, outs)
        return mine


SIBLING_NAME_SETS = (("query", "key", "value"), ("q_proj", "k_proj", "v_proj"))
# cross-attention blocks feed q_proj the decoder states and k_proj / v_proj the encoder states (quant_bart.py:166-175):
# their projections never share an input, so they are never grouped
CROSS_ATTENTION_MARKERS = ("encoder_attn", "crossattention", "cross_attn")


def group_sibling_linears(model):
    """Finds the sibling projections of every self-attention block (children named like SIBLING_NAME_SETS, all QLinear
    with the same in_features) and ties them into a QLinearGroup.  Idempotent; returns the number of groups.  Called
    by the state togglers, so a reference driver gets it without any change."""
    n = 0
    for pname, parent in model.named_modules():
        leaf = pname.rsplit(".", 1)[-1]
        if any(mk in leaf for mk in CROSS_ATTENTION_MARKERS) or getattr(parent, "is_cross_attention", False):
            for kid in parent.children():
                if isinstance(kid, QLinear):
                    kid._sibling_group = None
            continue
        kids = dict(parent.named_children())
        for names in SIBLING_NAME_SETS:
            members = [kids.get(k) for k in names]
            if any(not isinstance(m, QLinear) for m in members):
                continue
            if len({m.in_features for m in members}) != 1:
                continue
            n += 1
            cur = getattr(members[0], "_sibling_group", None)
            if cur is not None and cur.members == members:
                continue
            grp = QLinearGroup(members)
            for m in members:
                m._sibling_group = grp
    return n


class QEmbedding(QuantizedOperator, nn.Embedding):
    """quantized_module.py:75-100; the 23 M-element table is fake-quantized once per weight version."""

    def __init__(self, num_embeddings, embedding_dim, padding_idx, max_norm, norm_type, scale_grad_by_freq, sparse,
                 _weight, w_qconfig):
        super().__init__(num_embeddings=num_embeddings, embedding_dim=embedding_dim, padding_idx=padding_idx,
                         max_norm=max_norm, norm_type=norm_type, scale_grad_by_freq=scale_grad_by_freq, sparse=sparse,
                         _weight=_weight)
        self.weight_fake_quant = WeightQuantizer(w_qconfig)

    def forward(self, input):
        w = self._cached_fq_weight() if self._weight_is_static() else self.weight_fake_quant(self.weight)
        return F.embedding(input, w, self.padding_idx, self.max_norm, self.norm_type, self.scale_grad_by_freq,
                           self.sparse)


module_type_to_quant_weight = {nn.Linear: QLinear, nn.Conv2d: QConv2d, nn.Embedding: QEmbedding}


def get_module_args(module):
    if isinstance(module, nn.Linear):
        return dict(in_features=module.in_features, out_features=module.out_features, bias=module.bias is not None)
    if isinstance(module, nn.Conv2d):
        return dict(in_channels=module.in_channels, out_channels=module.out_channels, kernel_size=module.kernel_size,
                    stride=module.stride, padding=module.padding, dilation=module.dilation, groups=module.groups,
                    bias=module.bias is not None, padding_mode=module.padding_mode)
    if isinstance(module, nn.Embedding):
        return dict(num_embeddings=module.num_embeddings, embedding_dim=module.embedding_dim,
                    padding_idx=module.padding_idx, max_norm=module.max_norm, norm_type=module.norm_type,
                    scale_grad_by_freq=module.scale_grad_by_freq, sparse=module.sparse, _weight=None)
    raise NotImplementedError


def Quantizer(module, config):
    """quantized_module.py:144-155: ``None`` -> activation quantizer, Linear/Conv2d/Embedding -> Q* operator."""
    if module is None:
        return ActivationQuantizer(a_qconfig=config)
    cls = module_type_to_quant_weight.get(type(module))
    if cls is None:
        return module
    qmodule = cls(**get_module_args(module), w_qconfig=config)
    qmodule.weight.data = module.weight.data.clone()
    if getattr(module, "bias", None) is not None:
        qmodule.bias.data = module.bias.data.clone()
    return qmodule


def _build(qconfig):
    return FakeQuantizeDict[qconfig.quantizer](ObserverDict[qconfig.observer], bit=qconfig.bit,
                                               symmetric=qconfig.symmetric, ch_axis=qconfig.ch_axis)


def ActivationQuantizer(a_qconfig):
    return _build(a_qconfig)


def WeightQuantizer(w_qconfig):
    return _build(w_qconfig)
