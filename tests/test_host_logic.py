"""CPU-only checks: the C-ABI library builds/loads and exports every symbol of include/osq.h, the
boundary keeps the reference's names / flags / state_dict keys, and the product path refuses to run
without CUDA (no CPU fallback)."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class QC:
    def __init__(self, quantizer, observer, bit, symmetric, ch_axis):
        self.quantizer, self.observer, self.bit, self.symmetric, self.ch_axis = quantizer, observer, bit, symmetric, ch_axis


def test_library_exports_every_declared_symbol():
    from outlier_suppression_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "osq.h")).read()
    declared = set(re.findall(r"\b(osq_[a-z0-9_]+)\s*\(", header))
    declared -= {"osq_tokens_t", "osq_stat_epilogue_t", "osq_fused_linear_t", "osq_replay_target_t"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), "libosq_b200.so does not export %s" % name
    assert set(_lib.EXPORTS) == declared
    assert lib.osq_version() == 100
    assert lib.osq_workspace_bytes() > 0


def test_argument_validation_without_gpu():
    from outlier_suppression_b200 import _lib
    lib = _lib.load()
    rc = lib.osq_fq_per_tensor_f32(None, None, None, 16, None, None, 0, 0.0, 0, 63, None)
    assert rc == -1 and b"null pointer" in lib.osq_last_error()
    rc = lib.osq_fused_fq_linear(None, None)
    assert rc == -1
    # round-2 entry points: argument checks come before any CUDA call
    rc = lib.osq_residual_layernorm_fq_f32(None, None, None, None, None, 1e-5, 4, 768, None, None, 0, 0.0, 0, 63, None, None, None, None)
    assert rc == -1 and b"null pointer" in lib.osq_last_error()
    rc = lib.osq_residual_layernorm_fq_f32(None, None, None, None, None, 1e-5, 0, 768, None, None, 0, 0.0, 0, 63, None, None, None, None)
    assert rc == 0                                    # empty input: nothing to do
    rc = lib.osq_attn_scores_fq_f32(None, None, 1, 2, 8, 8, 64, None, None, None, None, 1.0, None, None, None)
    assert rc == -1 and b"null pointer" in lib.osq_last_error()
    rc = lib.osq_attn_scores_fq_f32(None, None, 0, 2, 8, 8, 64, None, None, None, None, 1.0, None, None, None)
    assert rc == 0
    rc = lib.osq_attn_context_fq_f32(None, None, 1, 2, 8, 8, 64, None, None, None, None, None, None, None)
    assert rc == -1


def test_no_cpu_fallback():
    from outlier_suppression_b200 import ops
    x = torch.randn(4, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.fq_per_tensor(x, torch.tensor([0.1]), torch.tensor([0], dtype=torch.int32), 0, 63)
    from outlier_suppression_b200.quantization.quantized_module import Quantizer
    q = Quantizer(None, QC("FixedFakeQuantize", "AvgMinMaxObserver", 6, False, -1))
    q.enable_fake_quant()
    with pytest.raises(RuntimeError):
        q(x)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "outlier_suppression_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_registries_and_factory(golden):
    from outlier_suppression_b200.quantization import quantized_module as qm
    from outlier_suppression_b200.quantization.fake_quant import (FixedFakeQuantize, LSQFakeQuantize, LSQPlusFakeQuantize,
                                                                  QuantizeBase)
    assert set(qm.ObserverDict) == {"MinMaxObserver", "AvgMinMaxObserver", "MSEObserver", "AvgMSEObserver", "MSEFastObserver",
                                    "AvgMSEFastObserver", "AvgQuantileObserver", "LSQPlusObserver", "AvgPruneMinMaxObserver"}
    assert set(qm.FakeQuantizeDict) == {"FixedFakeQuantize", "LSQFakeQuantize", "LSQPlusFakeQuantize"}
    lin = torch.nn.Linear(128, 48)
    ql = qm.Quantizer(lin, QC("FixedFakeQuantize", "MinMaxObserver", 6, True, 0))
    assert isinstance(ql, qm.QLinear) and isinstance(ql.weight_fake_quant, FixedFakeQuantize)
    assert torch.equal(ql.weight, lin.weight) and ql.weight.data_ptr() != lin.weight.data_ptr()
    assert (ql.weight_fake_quant.quant_min, ql.weight_fake_quant.quant_max) == (-32, 31)
    aq = qm.Quantizer(None, QC("FixedFakeQuantize", "AvgMinMaxObserver", 6, False, -1))
    assert (aq.quant_min, aq.quant_max, aq.observer_enabled, aq.fake_quant_enabled) == (0, 63, 0, 0)
    g = golden("qlinear")
    assert sorted(aq.state_dict().keys()) == list(g["keys_fixed"])
    assert sorted(ql.state_dict().keys()) == list(g["keys_qlinear"])
    lq = qm.Quantizer(None, QC("LSQPlusFakeQuantize", "AvgPruneMinMaxObserver", 6, False, -1))
    assert sorted(lq.state_dict().keys()) == list(g["keys_lsqplus"])
    assert isinstance(lq.scale, torch.nn.Parameter) and isinstance(lq.zero_point, torch.nn.Parameter)
    assert isinstance(lq, QuantizeBase) and isinstance(qm.Quantizer(torch.nn.ReLU(), None), torch.nn.ReLU)
    emb = qm.Quantizer(torch.nn.Embedding(10, 8), QC("FixedFakeQuantize", "MinMaxObserver", 6, True, 0))
    assert isinstance(emb, qm.QEmbedding)
    assert isinstance(qm.Quantizer(None, QC("LSQFakeQuantize", "MinMaxObserver", 8, True, -1)), LSQFakeQuantize)
    assert isinstance(lq, LSQPlusFakeQuantize)


def test_state_dict_roundtrip_resizes():
    from outlier_suppression_b200.quantization import quantized_module as qm
    cfg = QC("FixedFakeQuantize", "MinMaxObserver", 6, True, 0)
    a = qm.Quantizer(torch.nn.Linear(16, 4), cfg)
    a.weight_fake_quant.scale.resize_(4).copy_(torch.tensor([1., 2., 3., 4.]))
    a.weight_fake_quant.zero_point.resize_(4).zero_()
    b = qm.Quantizer(torch.nn.Linear(16, 4), cfg)
    epoch = b.weight_fake_quant.qparam_epoch
    b.load_state_dict(a.state_dict())
    assert torch.equal(b.weight_fake_quant.scale, torch.tensor([1., 2., 3., 4.]))
    assert b.weight_fake_quant.qparam_epoch > epoch


def test_state_togglers():
    from outlier_suppression_b200 import quantization as Q
    from outlier_suppression_b200.quantization.state import set_observer_name
    from outlier_suppression_b200.quantization.quantized_module import Quantizer

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.dense = Quantizer(torch.nn.Linear(8, 8), QC("FixedFakeQuantize", "MinMaxObserver", 6, True, 0))
            self.out_act_fake_quant = Quantizer(None, QC("LSQPlusFakeQuantize", "AvgPruneMinMaxObserver", 6, False, -1))
            self.in_act_fake_quant = Quantizer(None, QC("FixedFakeQuantize", "AvgMinMaxObserver", 6, False, -1))

    net = Net()
    flags = lambda m: (m.observer_enabled, m.fake_quant_enabled)  # noqa: E731
    Q.enable_calibration_woquantization(net, quantizer_type="weight_fake_quant")
    assert flags(net.dense.weight_fake_quant) == (1, 0) and flags(net.out_act_fake_quant) == (0, 0)
    Q.enable_calibration_woquantization(net, quantizer_type="act_fake_quant")
    assert flags(net.dense.weight_fake_quant) == (0, 0) and flags(net.in_act_fake_quant) == (1, 0)
    Q.enable_calibration_quantization(net, quantizer_type="act_fake_quant")
    assert flags(net.out_act_fake_quant) == (0, 1) and flags(net.in_act_fake_quant) == (1, 1)
    Q.enable_quantization(net)
    assert flags(net.dense.weight_fake_quant) == (0, 1) and flags(net.in_act_fake_quant) == (0, 1)
    Q.enable_quantization(net, except_quantizer=["in_act_fake_quant"])
    assert flags(net.in_act_fake_quant) == (0, 0)
    Q.disable_all(net)
    assert flags(net.dense.weight_fake_quant) == (0, 0)
    set_observer_name(net)
    assert net.out_act_fake_quant.observer.name == "out_act_fake_quant.observer"


def test_token_geometry():
    from outlier_suppression_b200 import ops
    x = torch.empty(4, 12, 32)
    t = ops.token_geometry(x, 1)
    assert (t.B, t.S, t.F1, t.F2, t.sb, t.ss, t.sf2) == (4, 12, 1, 32, 384, 32, 1)
    q = torch.empty(4, 12, 3, 8).permute(0, 2, 1, 3)          # [B,h,S,d] view of [B,S,h,d]
    t = ops.token_geometry(q, 2)
    assert (t.B, t.S, t.F1, t.F2, t.ss, t.sf2) == (4, 12, 1, 24, 24, 1)
    kT = q.transpose(-1, -2)                                   # [B,h,d,S]
    t = ops.token_geometry(kT, 3)
    assert (t.B, t.S, t.F1, t.F2, t.ss, t.sf2) == (4, 12, 1, 24, 24, 1)
    p = torch.empty(4, 3, 12, 12)                              # probs: F = h * S_k in h segments
    t = ops.token_geometry(p, 2)
    assert (t.B, t.S, t.F1, t.F2, t.sf1, t.ss, t.sf2) == (4, 12, 3, 12, 144, 12, 1)


def test_host_qparams_match_golden(golden):
    from outlier_suppression_b200.quantization.observer import ObserverBase
    g = golden("qparams")
    for bit in (4, 6, 8):
        for sym in (False, True):
            o = ObserverBase(bit=bit, symmetric=sym)
            s, z = o.calculate_qparams(torch.from_numpy(g["mins"]), torch.from_numpy(g["maxs"]))
            np.testing.assert_array_equal(s.numpy(), g["s_%d_%d" % (bit, sym)])
            np.testing.assert_array_equal(z.numpy(), g["z_%d_%d" % (bit, sym)])


def test_sibling_linear_grouping_discovery():
    """query | key | value QLinears of one parent are tied into one QLinearGroup by the state togglers
    (no GPU needed: only the module graph is inspected); unrelated or mismatching children are left alone."""
    from outlier_suppression_b200 import quantization as Q
    from outlier_suppression_b200.quantization import quantized_module as qm
    w_cfg = QC("FixedFakeQuantize", "MinMaxObserver", 6, True, 0)

    class Attn(torch.nn.Module):
        def __init__(self, names, widths):
            super().__init__()
            for n, k in zip(names, widths):
                setattr(self, n, qm.Quantizer(torch.nn.Linear(k, 16), w_cfg))

    net = torch.nn.Module()
    net.a = Attn(("query", "key", "value"), (32, 32, 32))
    net.b = Attn(("q_proj", "k_proj", "v_proj"), (32, 32, 32))
    net.c = Attn(("query", "key", "value"), (32, 32, 64))      # different in_features: not grouped
    net.d = Attn(("query", "key"), (32, 32))                    # incomplete set: not grouped
    assert qm.group_sibling_linears(net) == 2
    ga = net.a.query._sibling_group
    assert ga is not None and ga is net.a.key._sibling_group is net.a.value._sibling_group
    assert ga.members == [net.a.query, net.a.key, net.a.value]
    assert net.b.q_proj._sibling_group is net.b.v_proj._sibling_group and net.b.q_proj._sibling_group is not ga
    assert net.c.query._sibling_group is None and net.d.query._sibling_group is None
    Q.enable_quantization(net)                                   # idempotent through the togglers
    assert net.a.query._sibling_group is ga
    assert all("_sibling_group" not in k for k in net.state_dict())


def test_ctypes_structures_match_the_header():
    """Field names and order of the three C structs in include/osq.h vs their ctypes mirrors in _lib.py (an ABI drift
    here would silently scramble kernel arguments)."""
    from outlier_suppression_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "osq.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)

    def fields(struct_name):
        end = re.search(r"\}\s*%s\s*;" % struct_name, hdr)
        assert end, struct_name
        start = hdr.rfind("typedef struct", 0, end.start())
        body = hdr[hdr.index("{", start) + 1:end.start()]
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.findall(r"[A-Za-z_][A-Za-z_0-9]*", part)[-1])
        return names

    for cname, cls in (("osq_tokens_t", _lib.Tokens), ("osq_stat_epilogue_t", _lib.StatEpilogue),
                       ("osq_fused_linear_t", _lib.FusedLinearArgs), ("osq_replay_target_t", _lib.ReplayTarget),
                       ("osq_quantizer_t", _lib.QuantizerArgs), ("osq_select_problem_t", _lib.SelectProblem)):
        assert fields(cname) == [f[0] for f in cls._fields_], cname


def test_ctypes_argument_counts_match_the_header():
    """Every prototype of include/osq.h has exactly as many parameters as its ctypes argtypes list in _lib.py."""
    import ctypes as C
    from outlier_suppression_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "osq.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)

    class Rec:                                   # records what _declare() assigns, without loading the library
        def __init__(self):
            self.fns = {}

        def __getattr__(self, name):
            return self.fns.setdefault(name, type("F", (), {})())

    rec = Rec()
    _lib._declare(rec)
    checked = 0
    for name in _lib.EXPORTS:
        m = re.search(r"\b%s\s*\((.*?)\)\s*;" % name, hdr, flags=re.S)
        assert m, "%s is not declared in osq.h" % name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        argtypes = getattr(rec.fns.get(name), "argtypes", None)
        if argtypes is not None:
            assert len(argtypes) == len(params), (name, len(argtypes), len(params))
            checked += 1
    assert checked >= 14
    assert C.sizeof(_lib.Tokens) == 64


def test_cross_attention_projections_are_not_grouped():
    """BART encoder_attn feeds q_proj the decoder states and k_proj / v_proj the encoder states
    (quant_bart.py:166-175); only self-attention siblings share an input and may share a launch."""
    from outlier_suppression_b200.quantization import quantized_module as qm

    def attn():
        m = torch.nn.Module()
        for nme in ("q_proj", "k_proj", "v_proj", "out_proj"):
            setattr(m, nme, qm.Quantizer(torch.nn.Linear(128, 128), QC("FixedFakeQuantize", "MinMaxObserver", 6, True, 0)))
        return m

    layer = torch.nn.Module()
    layer.self_attn, layer.encoder_attn = attn(), attn()
    bert = torch.nn.Module()
    bert.self = torch.nn.Module()
    for nme in ("query", "key", "value"):
        setattr(bert.self, nme, qm.Quantizer(torch.nn.Linear(128, 128), QC("FixedFakeQuantize", "MinMaxObserver", 6, True, 0)))
    net = torch.nn.ModuleList([layer, bert])
    assert qm.group_sibling_linears(net) == 2
    assert layer.self_attn.q_proj._sibling_group is layer.self_attn.v_proj._sibling_group is not None
    assert layer.self_attn.out_proj._sibling_group is None
    assert all(getattr(layer.encoder_attn, n)._sibling_group is None for n in ("q_proj", "k_proj", "v_proj"))
    assert bert.self.query._sibling_group.members == [bert.self.query, bert.self.key, bert.self.value]


def test_pack_weight_rejects_ranges_that_do_not_fit_int8():
    from outlier_suppression_b200 import _lib
    lib = _lib.load()
    fake = 4096  # never dereferenced: argument validation precedes any launch
    assert lib.osq_pack_weight_s8(fake, 8, 128, fake, fake, 0, 255, fake, fake, None) == -1
    assert b"int8" in lib.osq_last_error()
    assert lib.osq_pack_weight_s8(fake, 8, 128, fake, fake, -256, 255, fake, fake, None) == -1


def test_fused_linear_supported_mirrors_the_kernel_contract():
    from outlier_suppression_b200 import ops
    assert ops.fused_linear_supported(768, 768) and ops.fused_linear_supported(32768, 16)
    for k, n in ((100, 768), (768, 100), (65536, 768), (768, 8), (768, (1 << 20) + 16)):
        assert not ops.fused_linear_supported(k, n)


def test_model_level_hooks_are_found_by_duck_typing_and_idempotent():
    """fusion.py attaches to the reference's blocks by their attribute names (no import of reference classes): the FFN block
    (dense / intermediate_act_fn / its quantizer), the residual + LayerNorm block, the self-attention block.  Wrapping twice
    changes nothing; unrelated modules are left alone."""
    from outlier_suppression_b200.quantization import fusion
    from outlier_suppression_b200.quantization.quantized_module import Quantizer
    a = QC("LSQPlusFakeQuantize", "AvgPruneMinMaxObserver", 6, False, -1)
    w = QC("FixedFakeQuantize", "MinMaxObserver", 6, True, 0)

    class Res(torch.nn.Module):
        mul_gamma = False
        def forward(self, x, h):
            return x + h

    class LN(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.qoutput = True
            self.layernorm = torch.nn.LayerNorm(128)
            self.layernorm_post_act_fake_quantize = Quantizer(None, a)

    class Out(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.dense = Quantizer(torch.nn.Linear(128, 128), w)
            self.dropout = torch.nn.Dropout(0.1)
            self.before_LayerNorm_residual = Res()
            self.LayerNorm = LN()
        def forward(self, hidden_states, input_tensor, observation_mask=None):
            return hidden_states

    class Attn(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.num_attention_heads, self.attention_head_size, self.qoutput = 2, 64, True
            for n in ("query", "key", "value"):
                setattr(self, n, Quantizer(torch.nn.Linear(128, 128), w))
            for n in fusion._ATTN_Q + ("context_view_post_act_fake_quantize",):
                setattr(self, n, Quantizer(None, a))
            self.dropout = torch.nn.Dropout(0.1)
        def transpose_for_scores(self, x):
            return x.view(x.shape[0], x.shape[1], 2, 64).permute(0, 2, 1, 3)
        def forward(self, hidden_states, attention_mask=None, head_mask=None, output_attentions=False, observation_mask=None):
            return (hidden_states,)

    net = torch.nn.Module()
    net.a, net.o, net.plain = Attn(), Out(), torch.nn.Linear(4, 4)
    assert fusion.fuse_layernorm_output(net) == 1 and fusion.fuse_self_attention(net) == 1 and fusion.fuse_ffn_activation(net) == 0
    f_o, f_a = net.o.forward, net.a.forward
    assert fusion.fuse_layernorm_output(net) == 1 and fusion.fuse_self_attention(net) == 1
    assert net.o.forward == f_o and net.a.forward == f_a and not hasattr(net.plain, "_osq_unfused_forward")
    # outside the quantized inference state (here: autograd on, quantizers off) the original forward runs
    x = torch.randn(2, 8, 128)
    assert net.a(x)[0] is x
    assert not fusion._attn_fusable(net.a, x, None, None, False)
    with torch.no_grad():
        assert not fusion._attn_fusable(net.a, x, None, None, False)          # fake-quant disabled
        for n in fusion._ATTN_Q + ("context_view_post_act_fake_quantize",):
            getattr(net.a, n).enable_fake_quant(); getattr(net.a, n).disable_observer()
        assert not fusion._attn_fusable(net.a, x, None, None, False)          # CPU tensor: no CPU fallback, reference forward
        assert not fusion._attn_fusable(net.a, x, None, torch.ones(2), False)  # head_mask
