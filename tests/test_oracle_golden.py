"""Pins the CPU oracle (oracle/osq_oracle.py) against vectors produced by the unmodified reference
(tests/golden/gen_golden.py).  Bit-exact unless stated."""
import os
import warnings

import numpy as np
import pytest
import torch

from oracle import osq_oracle as O

warnings.filterwarnings("ignore")
T = torch.from_numpy


def eq(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    np.testing.assert_array_equal(a, b)


def test_fq_per_tensor_bit_exact(golden):
    g = golden("fq_per_tensor")
    for i in range(int(g["n"])):
        scale, zp, qmin, qmax = g["p%d" % i]
        x = T(g["x%d" % i])
        eq(O.fq_bins(x, float(scale), int(zp), int(qmin), int(qmax)), g["q%d" % i])
        eq(O.fq_per_tensor(x, float(scale), int(zp), int(qmin), int(qmax)), g["y%d" % i])


def test_fq_per_channel_and_minmax_qparams(golden):
    g = golden("fq_per_channel")
    for i in range(int(g["n"])):
        bit, sym, qmin, qmax = (int(v) for v in g["p%d" % i])
        w = T(g["w%d" % i])
        st = O.ObserverState()
        O.observe_minmax(st, w, ch_axis=0)
        eq(st.min_val, g["min%d" % i])
        eq(st.max_val, g["max%d" % i])
        assert O.quant_range(bit, bool(sym)) == (qmin, qmax)
        scale, zp = O.qparams_from_minmax(st.min_val, st.max_val, qmin, qmax, bool(sym))
        eq(scale, g["scale%d" % i])
        eq(zp.to(torch.int32), g["zp%d" % i])
        eq(O.fq_per_channel(w, scale, zp.to(torch.int32), 0, qmin, qmax), g["y%d" % i])
        eq(O.fq_per_channel(T(g["w2_%d" % i]), scale, zp.to(torch.int32), 0, qmin, qmax), g["y2_%d" % i])


def test_lsqplus_forward_and_sanitize(golden):
    g = golden("lsqplus")
    for i in range(int(g["n"])):
        bit, qmin, qmax = (int(v) for v in g["p%d" % i])
        scale = torch.tensor([float(g["scale_in%d" % i])], dtype=torch.float32)
        zp = torch.tensor([float(g["zp_in%d" % i])], dtype=torch.float32)
        O.lsqplus_sanitize(scale, zp, qmin, qmax)
        eq(scale, g["scale_after%d" % i])
        eq(zp, g["zp_after%d" % i])
        eq(O.fq_lsqplus_per_tensor(T(g["x%d" % i]), scale, zp, qmin, qmax), g["y%d" % i])


def test_qparams(golden):
    g = golden("qparams")
    for bit in (4, 6, 8):
        for sym in (False, True):
            qmin, qmax = O.quant_range(bit, sym)
            s, z = O.qparams_from_minmax(T(g["mins"]), T(g["maxs"]), qmin, qmax, sym)
            eq(s, g["s_%d_%d" % (bit, sym)])
            eq(z, g["z_%d_%d" % (bit, sym)])


OBS_CASES = ["ln3d", "q4d", "kT4d", "probs", "nomask3d", "flat", "bartprobs3d"]
OBS_NAMES = {"probs": "layer.0.attention_probs_post_act_fake_quantize"}


@pytest.mark.parametrize("case", OBS_CASES)
def test_observers(golden, case):
    g = golden("observers")
    lens = g[case + "_lens"]
    lens = None if lens.size == 0 else [int(v) for v in lens]
    seq_pos = int(g[case + "_meta"][0])
    xs = [T(g["%s_x%d" % (case, b)]) for b in range(3)]
    st = O.ObserverState()
    for b, x in enumerate(xs):
        O.observe_avg_minmax(st, x, lens, seq_pos)
        eq(torch.stack([st.min_val, st.max_val]), g[case + "_avgminmax"][b])
    for p in (0.99, 0.9, 0.7):
        st = O.ObserverState()
        for b, x in enumerate(xs):
            O.observe_avg_prune_minmax(st, x, p, OBS_NAMES.get(case, "x"), lens, seq_pos)
            eq(torch.stack([st.min_val, st.max_val]), g["%s_prune_%d" % (case, int(p * 100))][b])
        s, z = O.qparams_from_minmax(st.min_val, st.max_val, 0, 63, False)
        eq(torch.stack([s, z]), g["%s_prune_%d_qp" % (case, int(p * 100))])


def test_prune_equals_bounds(golden):
    """the clip + global min/max of observer.py:69,227 is exactly the (lower, upper) pair."""
    g = golden("observers")
    x = T(g["ln3d_x0"])
    tok = O.token_matrix(x, [12, 7, 1, 9], 1)
    tmin, tmax = O.token_minmax(tok)
    for p in (0.99, 0.9, 0.7, 0.3):
        lo, hi = O.prune_bounds(tmin, tmax, p)
        mn, mx = O.prune_minmax(tok, p)
        assert lo == mn and hi == mx


def test_token_geometry(golden):
    g = golden("observers")
    x = T(g["geom_x"])
    eq(O.token_matrix(x, [5, 2, 0, 3], 2), g["geom_sp2"])
    eq(O.token_matrix(x.transpose(-1, -2), [5, 2, 0, 3], 3), g["geom_sp3"])
    eq(O.token_matrix(x.reshape(4, 15, 2), [15, 2, 0, 3], 1), g["geom_sp1"])
    eq(O.token_matrix(x, None, 2), g["geom_full"])


def test_minmax_per_channel_running(golden):
    g = golden("minmax_per_channel")
    st = O.ObserverState()
    O.observe_minmax(st, T(g["w1"]), ch_axis=0)
    eq(st.min_val, g["min1"]); eq(st.max_val, g["max1"])
    O.observe_minmax(st, T(g["w2"]), ch_axis=0)
    eq(st.min_val, g["min2"]); eq(st.max_val, g["max2"])
    s, z = O.qparams_from_minmax(st.min_val, st.max_val, -32, 31, True)
    eq(s, g["scale"]); eq(z, g["zp"])


def test_mse_fast(golden):
    g = golden("mse")
    cnt = [0]
    st = O.ObserverState()
    O.observe_mse_fast(st, T(g["w"]), -8, 7, True, ch_axis=0, counter=cnt)
    eq(st.min_val, g["w_min"]); eq(st.max_val, g["w_max"])
    assert cnt[0] == int(g["w_evals"])
    s, z = O.qparams_from_minmax(st.min_val, st.max_val, -8, 7, True)
    eq(s, g["w_scale"])
    cnt = [0]
    st = O.ObserverState()
    for b in range(2):
        O.observe_avg_mse_fast(st, T(g["a_x%d" % b]), 0, 63, False, [10, 4], 1, counter=cnt)
        np.testing.assert_array_equal(np.array([float(st.min_val), float(st.max_val)]), g["a_trace"][b])
    assert cnt[0] == int(g["a_evals"])
    cnt = [0]
    st = O.ObserverState()
    O.observe_avg_mse_fast(st, T(g["p_x"]), 0, 63, False, counter=cnt)
    assert float(st.min_val) == float(g["p_min"]) and float(st.max_val) == float(g["p_max"])
    assert cnt[0] == int(g["p_evals"])
    x = T(g["l_x"])
    for (a, b), ref in zip(g["l_cands"], g["l_loss"]):
        assert float(O.mse_loss(x, float(a), float(b), 0, 63, False)) == float(ref)


def test_qlinear_chain(golden):
    g = golden("qlinear")
    for i in range(int(g["n"])):
        a_bit, w_bit, lsq, aqmin, aqmax, wqmin, wqmax = (int(v) for v in g["p%d" % i])
        x, w, b = T(g["x%d" % i]), T(g["w%d" % i]), T(g["b%d" % i])
        w_scale, w_zp, qmin_w, qmax_w = O.weight_qparams_minmax(w, w_bit, True)
        eq(w_scale, g["w_scale%d" % i]); eq(w_zp, g["w_zp%d" % i])
        a_scale = T(np.atleast_1d(g["a_scale%d" % i])).clone()
        a_zp = T(np.atleast_1d(g["a_zp%d" % i])).clone()
        y, qa, qw = O.fused_fq_linear(x, a_scale, a_zp, aqmin, aqmax, bool(lsq), w, w_scale, w_zp, wqmin, wqmax, b)
        eq(y, g["y%d" % i])
        # integer-code factorisation agrees with the float path within the stated tolerance
        s_a = float(a_scale) if not lsq else float(O.lsqplus_effective_qparams(a_scale, a_zp, x.numel(), aqmax)[0])
        z_a = float(torch.round(a_zp))
        acc = (qa.double() - z_a).reshape(-1, x.shape[-1]) @ qw.double().t()
        y2 = (acc * (s_a * w_scale.double())[None, :] + b.double()[None, :]).reshape(y.shape)
        ref = T(g["y%d" % i]).double()
        assert torch.all((y2 - ref).abs() <= 1e-3 * ref.abs() + 1e-3 * ref.abs().max())


def test_oracle_model_blocks_match_the_reference_modules():
    """tests/golden/blocks.npz holds every tensor crossing layer 0's self-attention block and dense -> residual -> LayerNorm ->
    quantizer block of the reference's unmodified quant_bert.py (config 2: LSQ+ 6-bit after gamma migration; config 1: Fixed
    8-bit).  The oracle restatement must reproduce them: quantizer outputs exactly (same inputs, same arithmetic), matmul /
    softmax / LayerNorm results to the last bits torch's CPU kernels may round differently across builds."""
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "blocks.npz"))
    t = lambda k: torch.from_numpy(g[k])
    for tag in ("c2", "c1"):
        heads, d, bit, lsq = (int(v) for v in g[tag + "_meta"])
        qmin, qmax = O.quant_range(bit, False)
        qp = lambda n: (t("%s_%s_scale" % (tag, n)), t("%s_%s_zp" % (tag, n)))
        scores, probs, ctx = O.attention_block(t(tag + "_q3"), t(tag + "_k3"), t(tag + "_v3"), t(tag + "_mask"), heads, qp("query_permute"),
                                               qp("key_transpose"), qp("attention_probs"), qp("value_permute"), qp("context_view"),
                                               qmin, qmax, bool(lsq))
        assert float((probs - t(tag + "_probs")).abs().max()) <= 1e-6
        # the quantizers themselves, on the reference's own inputs: exact
        assert torch.equal(O.act_fq(t(tag + "_probs"), *qp("attention_probs"), qmin, qmax, bool(lsq)), t(tag + "_probs_fq"))
        assert torch.equal(O.act_fq(t(tag + "_so_ln"), t(tag + "_so_scale"), t(tag + "_so_zp"), qmin, qmax, bool(lsq)), t(tag + "_so_y"))
        step = float(t("%s_context_view_scale" % tag))
        diff = (ctx - t(tag + "_ctx_fq")).abs()
        assert float(diff.max()) <= step * 1.001 and float((diff > 0).float().mean()) <= 1e-3   # a tie may flip one bin
        gam = t(tag + "_so_gamma"); w = t(tag + "_so_ln_weight"); b = t(tag + "_so_ln_bias")
        ln, y = O.residual_layernorm_fq(t(tag + "_so_h"), t(tag + "_so_res"), gam if gam.numel() else None, w if w.numel() else None,
                                        b if b.numel() else None, float(g[tag + "_eps"][0]), (t(tag + "_so_scale"), t(tag + "_so_zp")),
                                        qmin, qmax, bool(lsq))
        assert float((ln - t(tag + "_so_ln")).abs().max()) <= 2e-6 * float(t(tag + "_so_ln").abs().max())
        sdiff = (y - t(tag + "_so_y")).abs()
        assert float(sdiff.max()) <= float(t(tag + "_so_scale")) * 1.001 and float((sdiff > 0).float().mean()) <= 1e-3


def test_integer_contraction_identity_behind_the_attention_kernels():
    """K8 / K9 never form the dequantised operands: fq(a) @ fq(b)^T = s_a s_b (sum_k A B - Zb rowsum(A) - Za rowsum(B) + K Za Zb) with
    A, B the uint8 bins (bin - qmin) and Za, Zb the zero points relative to qmin.  Checked here on the CPU, in exact integer
    arithmetic, against the fp64 product of the oracle's fake-quantised tensors, on the reference-generated block inputs."""
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "blocks.npz"))
    t = lambda k: torch.from_numpy(g[k])
    for tag in ("c2", "c1"):
        heads, d, bit, lsq = (int(v) for v in g[tag + "_meta"])
        qmin, qmax = O.quant_range(bit, False)
        B, S, H = g[tag + "_q3"].shape
        hv = lambda x: x.view(B, S, heads, d).permute(0, 2, 1, 3)

        def bins_and_params(x, name):
            sc, zp = t("%s_%s_scale" % (tag, name)), t("%s_%s_zp" % (tag, name))
            if lsq:
                s, z, _ = O.lsqplus_effective_qparams(sc.clone(), zp.clone(), x.numel(), qmax)
                q = torch.clamp(O.round_ste_value(x / s) + z, qmin, qmax)
                return torch.round(q).to(torch.int64) - qmin, float(s), int(torch.round(z)) - qmin, (q - z) * s
            q = O.fq_bins(x, float(sc), int(zp), qmin, qmax)
            return q.to(torch.int64) - qmin, float(sc), int(zp) - qmin, (q - int(zp)) * float(sc)
        qa, sq, zq, q_fq = bins_and_params(hv(t(tag + "_q3")), "query_permute")
        ka, sk, zk, k_fq = bins_and_params(hv(t(tag + "_k3")), "key_transpose")
        acc = torch.matmul(qa, ka.transpose(-1, -2))                                  # exact: int64
        corr = acc - zk * qa.sum(-1, keepdim=True) - zq * ka.sum(-1, keepdim=True).transpose(-1, -2) + d * zq * zk
        got = corr.double() * (np.float64(np.float32(sq) * np.float32(sk)))           # the kernel's single fp32 scale product
        want = torch.matmul(q_fq.double(), k_fq.double().transpose(-1, -2))
        assert float((got - want).abs().max()) <= 2e-6 * float(want.abs().max())
        assert int(corr.abs().max()) < 2 ** 24                                        # every term the kernel adds in fp32 is exact
