"""Generates tests/golden/*.npz by executing the UNMODIFIED reference (from /root/reference,
CPU, via oracle/ref_shim.py).  Run in the build container only:

    python tests/golden/gen_golden.py

The outputs are committed; the GPU box never needs /root/reference.  Every array is produced by the
reference's own modules / functions -- the oracle and the CUDA path are both checked against them.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402

ref = ref_shim.load()
from quant_transformer.quantization import fake_quant as R_fq  # noqa: E402
from quant_transformer.quantization import observer as R_obs  # noqa: E402
from quant_transformer.quantization import quantized_module as R_qm  # noqa: E402
from quant_transformer.quantization import util_quant as R_uq  # noqa: E402

QC = ref_shim.QConfig


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: (v.shape, str(v.dtype)) for k, v in out.items()})


def edge_tensor(scale, zp, qmin, qmax, g, n_rand=2000):
    """ties at (k+.5)*s, clamp edges, +-0, huge / tiny values, plus random fill."""
    ks = torch.arange(qmin - zp - 3, qmax - zp + 4, dtype=torch.float32)
    s = torch.tensor(scale, dtype=torch.float32)
    ties = (ks + 0.5) * s
    near = torch.cat([ties * (1 + 2 ** -23), ties * (1 - 2 ** -23), ks * s])
    special = torch.tensor([0.0, -0.0, 1e-30, -1e-30, 1e30, -1e30, 3.4e38, -3.4e38, float("inf"), float("-inf")])
    rnd = torch.randn(n_rand, generator=g) * scale * (qmax - qmin) / 3
    return torch.cat([ties, near, special, rnd])


def gen_fq_per_tensor():
    g = torch.Generator().manual_seed(100)
    out = {}
    i = 0
    for bit in (4, 6, 8):
        for sym in (False, True):
            qmin, qmax = (-(2 ** (bit - 1)), 2 ** (bit - 1) - 1) if sym else (0, 2 ** bit - 1)
            for scale in (0.0371, 1.0 / 3.0, 1e-8, 7.25):
                zp = 0 if sym else int(torch.randint(qmin, qmax + 1, (1,), generator=g))
                x = edge_tensor(scale, zp, qmin, qmax, g)
                y = R_uq.fake_quantize_per_tensor_affine(x, float(np.float32(scale)), zp, qmin, qmax)
                q = torch.clamp(R_uq.round_ste(x / float(np.float32(scale))) + zp, qmin, qmax)
                out["x%d" % i], out["y%d" % i], out["q%d" % i] = x, y, q
                out["p%d" % i] = np.array([np.float32(scale), zp, qmin, qmax], dtype=np.float64)
                i += 1
    out["n"] = i
    save("fq_per_tensor", **out)


def gen_fq_per_channel():
    g = torch.Generator().manual_seed(101)
    out = {}
    i = 0
    for bit in (4, 6, 8):
        for sym in (True, False):
            w = torch.randn(48, 40, generator=g) * 0.05
            w[3] *= 20
            w[7] = 0.0
            w[9] = w[9].abs()
            fqm = R_fq.FixedFakeQuantize(R_obs.MinMaxObserver, bit=bit, symmetric=sym, ch_axis=0)
            fqm.enable_observer()
            fqm.enable_fake_quant()
            y = fqm(w)
            scale, zp = fqm.scale.clone(), fqm.zero_point.clone()
            # put exact ties into row 5 to pin rounding under a per-row tensor scale
            w2 = w.clone()
            w2[5, :20] = (torch.arange(20) - 10 + 0.5) * scale[5]
            y2 = R_uq.fake_quantize_per_channel_affine(w2, scale, zp.int(), 0, fqm.quant_min, fqm.quant_max)
            out.update({"w%d" % i: w, "y%d" % i: y, "w2_%d" % i: w2, "y2_%d" % i: y2,
                        "scale%d" % i: scale, "zp%d" % i: zp,
                        "min%d" % i: fqm.observer.min_val, "max%d" % i: fqm.observer.max_val,
                        "p%d" % i: np.array([bit, int(sym), fqm.quant_min, fqm.quant_max])})
            i += 1
    out["n"] = i
    save("fq_per_channel", **out)


def gen_lsqplus():
    g = torch.Generator().manual_seed(102)
    out = {}
    i = 0
    for bit, scale, zp, shape in ((6, 0.0713, 30.99992, (2, 5, 16)), (6, 0.21, 17.3, (3, 7, 24)),
                                  (8, 0.013, 127.5, (2, 4, 8)), (4, 0.5, 7.00001, (1, 9, 8)),
                                  (6, 0.0713, 70.2, (2, 5, 16)), (6, -0.05, 12.0, (2, 3, 8))):
        m = R_fq.LSQPlusFakeQuantize(R_obs.AvgPruneMinMaxObserver, bit=bit, symmetric=False, ch_axis=-1)
        m.scale.data.fill_(scale)
        m.zero_point.data.fill_(zp)
        m.enable_fake_quant()
        x = torch.randn(*shape, generator=g) * abs(scale) * 25
        x.view(-1)[:8] = (torch.arange(8) - 4 + 0.5) * abs(scale)
        with torch.no_grad():
            y = m(x)
        out.update({"x%d" % i: x, "y%d" % i: y, "scale_in%d" % i: np.float32(scale), "zp_in%d" % i: np.float32(zp),
                    "scale_after%d" % i: m.scale.data.clone(), "zp_after%d" % i: m.zero_point.data.clone(),
                    "p%d" % i: np.array([bit, m.quant_min, m.quant_max])})
        i += 1
    # gradients (fine-stage learn_scale path, token_wise_clipping.py:72-108)
    m = R_fq.LSQPlusFakeQuantize(R_obs.AvgPruneMinMaxObserver, bit=6, symmetric=False, ch_axis=-1)
    m.scale.data.fill_(0.09)
    m.zero_point.data.fill_(29.6)
    m.enable_fake_quant()
    x = (torch.randn(4, 6, 32, generator=g) * 2.5).requires_grad_(True)
    y = m(x)
    gy = torch.randn(y.shape, generator=g)
    (y * gy).sum().backward()
    out.update({"gx_x": x.detach(), "gx_gy": gy, "gx_y": y.detach(), "gx_dx": x.grad,
                "gx_dscale": m.scale.grad, "gx_dzp": m.zero_point.grad})
    out["n"] = i
    save("lsqplus", **out)


def gen_qparams():
    mins = torch.tensor([-3.0, -0.5, 0.0, 0.2, -1e-12, -7.3, 0.0, -2.0, -1000.0, -0.3333])
    maxs = torch.tensor([2.0, 4.0, 5.0, 0.9, 1e-12, -1.0, 0.0, 2.0, 3000.0, 0.6667])
    out = {"mins": mins, "maxs": maxs}
    for bit in (4, 6, 8):
        for sym in (False, True):
            o = R_obs.ObserverBase(bit=bit, symmetric=sym, ch_axis=-1)
            s, z = o.calculate_qparams(mins, maxs)
            out["s_%d_%d" % (bit, sym)], out["z_%d_%d" % (bit, sym)] = s, z
    save("qparams", **out)


def _act(g, shape, outlier_dim):
    x = torch.randn(*shape, generator=g)
    idx = [slice(None)] * len(shape)
    idx[outlier_dim] = slice(0, 2)
    x[tuple(idx)] *= 12.0
    return x


def gen_observers():
    g = torch.Generator().manual_seed(103)
    out = {}
    cases = [  # name, shape, seq_pos, lens per batch item (or None), observer name
        ("ln3d", (4, 12, 32), 1, [12, 7, 1, 9], "layer.0.output.layernorm"),
        ("q4d", (4, 3, 12, 8), 2, [12, 7, 1, 9], "layer.0.query_permute_post_act_fake_quantize"),
        ("kT4d", (4, 3, 8, 12), 3, [12, 7, 1, 9], "layer.0.key_transpose_post_act_fake_quantize"),
        ("probs", (4, 3, 12, 12), 2, [12, 7, 1, 9], "layer.0.attention_probs_post_act_fake_quantize"),
        ("nomask3d", (4, 12, 32), 1, None, "decoder.x"),
        ("flat", (4, 32), -1, None, "pooler"),
        ("bartprobs3d", (6, 5, 7), 1, [5, 3], "decoder.attn_probs_3d"),
    ]
    for cname, shape, seq_pos, lens, oname in cases:
        batches = []
        for b in range(3):
            x = _act(g, shape, len(shape) - 1)
            if cname == "probs":
                x = torch.softmax(x, -1)
            batches.append(x)
            out["%s_x%d" % (cname, b)] = x
        mask = None if lens is None else torch.tensor(lens)
        out["%s_lens" % cname] = np.array(lens if lens is not None else [], dtype=np.int64)
        out["%s_meta" % cname] = np.array([seq_pos])
        # AvgMinMax
        o = R_obs.AvgMinMaxObserver(bit=6, symmetric=False, ch_axis=-1)
        tr = []
        for x in batches:
            o(x, observation_mask=mask, seq_pos=seq_pos)
            tr.append([float(o.min_val), float(o.max_val)])
        out["%s_avgminmax" % cname] = np.array(tr, dtype=np.float32)
        # AvgPruneMinMax at two percentiles
        for p in (0.99, 0.9, 0.7):
            o = R_obs.AvgPruneMinMaxObserver(bit=6, symmetric=False, ch_axis=-1)
            o.set_name(oname)
            o.set_percentile(p)
            tr = []
            for x in batches:
                o(x, observation_mask=mask, seq_pos=seq_pos)
                tr.append([float(o.min_val), float(o.max_val)])
            out["%s_prune_%d" % (cname, int(p * 100))] = np.array(tr, dtype=np.float32)
            s, z = o.calculate_qparams(o.min_val, o.max_val)
            out["%s_prune_%d_qp" % (cname, int(p * 100))] = np.array([float(s), float(z)], dtype=np.float32)
    # token matrix for geometry pinning
    x = torch.arange(4 * 3 * 5 * 2, dtype=torch.float32).reshape(4, 3, 5, 2)
    o = R_obs.ObserverBase()
    out["geom_x"] = x
    out["geom_sp2"] = o.remove_padding(x, torch.tensor([5, 2, 0, 3]), 2)
    out["geom_sp3"] = o.remove_padding(x.transpose(2, 3).contiguous().transpose(2, 3).transpose(-1, -2), torch.tensor([5, 2, 0, 3]), 3)
    out["geom_sp1"] = o.remove_padding(x.reshape(4, 15, 2), torch.tensor([15, 2, 0, 3]), 1)
    out["geom_full"] = o.reshape_batch_embedding(x, 2)
    save("observers", **out)


def gen_minmax_per_channel():
    g = torch.Generator().manual_seed(104)
    w1 = torch.randn(24, 40, generator=g)
    w2 = torch.randn(24, 40, generator=g) * 2
    o = R_obs.MinMaxObserver(bit=6, symmetric=True, ch_axis=0)
    o(w1)
    a = (o.min_val.clone(), o.max_val.clone())
    o(w2)
    s, z = o.calculate_qparams(o.min_val, o.max_val)
    save("minmax_per_channel", w1=w1, w2=w2, min1=a[0], max1=a[1], min2=o.min_val, max2=o.max_val, scale=s, zp=z)


class CountingMSE(R_obs.MSEFastObserver):
    evals = 0

    def loss_fx(self, x, new_min, new_max):
        CountingMSE.evals += 1
        return super().loss_fx(x, new_min, new_max)


class CountingAvgMSE(R_obs.AvgMSEFastObserver):
    evals = 0

    def loss_fx(self, x, new_min, new_max):
        CountingAvgMSE.evals += 1
        return super().loss_fx(x, new_min, new_max)


def gen_mse():
    import warnings
    warnings.filterwarnings("ignore")
    g = torch.Generator().manual_seed(105)
    out = {}
    # per-channel symmetric 4-bit weights (config 3)
    w = torch.randn(12, 96, generator=g) * 0.05
    w[2] *= 8
    w[5] = w[5].abs()  # still symmetric search
    o = CountingMSE(bit=4, symmetric=True, ch_axis=0)
    CountingMSE.evals = 0
    o(w)
    out.update({"w": w, "w_min": o.min_val, "w_max": o.max_val, "w_evals": CountingMSE.evals})
    s, z = o.calculate_qparams(o.min_val, o.max_val)
    out.update({"w_scale": s, "w_zp": z})
    # per-tensor asymmetric 6-bit activations, 2-D search, two batches with mask
    lens = torch.tensor([10, 4])
    xs = [_act(g, (2, 10, 24), 2) for _ in range(2)]
    o = CountingAvgMSE(bit=6, symmetric=False, ch_axis=-1)
    CountingAvgMSE.evals = 0
    tr = []
    for x in xs:
        o(x, observation_mask=lens, seq_pos=1)
        tr.append([float(o.min_val), float(o.max_val)])
    out.update({"a_x0": xs[0], "a_x1": xs[1], "a_lens": lens, "a_trace": np.array(tr, dtype=np.float64),
                "a_evals": CountingAvgMSE.evals})
    # one-sided positive activations (post-GELU-like) -> 1-D search, asymmetric
    xp = torch.rand(2, 10, 24, generator=g) * 3
    o = CountingAvgMSE(bit=6, symmetric=False, ch_axis=-1)
    CountingAvgMSE.evals = 0
    o(xp)
    out.update({"p_x": xp, "p_min": float(o.min_val), "p_max": float(o.max_val), "p_evals": CountingAvgMSE.evals})
    # raw loss function values for fixed candidates (pins the per-candidate MSE kernel)
    o = R_obs.MSEFastObserver(bit=6, symmetric=False, ch_axis=-1)
    x = xs[0]
    cands = [(-1.0, 2.0), (-3.5, 3.0), (0.0, 4.0), (-12.0, 14.0), (-0.01, 0.02)]
    out["l_x"] = x
    out["l_cands"] = np.array(cands, dtype=np.float32)
    out["l_loss"] = np.array([float(o.loss_fx(x, np.float32(a).item(), np.float32(b).item())) for a, b in cands], dtype=np.float32)
    save("mse", **out)


def gen_qlinear():
    g = torch.Generator().manual_seed(106)
    out = {}
    i = 0
    for a_bit, w_bit, quantizer, (m_b, m_s, k, n) in ((6, 6, "FixedFakeQuantize", (2, 32, 128, 48)),
                                                      (8, 8, "FixedFakeQuantize", (2, 32, 128, 48)),
                                                      (6, 6, "LSQPlusFakeQuantize", (2, 32, 256, 64)),
                                                      (6, 4, "FixedFakeQuantize", (1, 64, 128, 32))):
        a_cfg = QC(quantizer, "AvgMinMaxObserver", a_bit, False, -1)
        w_cfg = QC("FixedFakeQuantize", "MinMaxObserver", w_bit, True, 0)
        lin = torch.nn.Linear(k, n)
        lin.weight.data = torch.randn(n, k, generator=g) * 0.05
        lin.bias.data = torch.randn(n, generator=g) * 0.02
        ql = R_qm.Quantizer(lin, w_cfg)
        aq = R_qm.Quantizer(None, a_cfg)
        x = _act(g, (m_b, m_s, k), 2)
        lens = torch.tensor([m_s, m_s // 2][:m_b])
        # calibration (ptq_glue_quant.py:234-246): weights then activations
        ql.weight_fake_quant.enable_observer()
        ql(x)
        ql.weight_fake_quant.disable_observer()
        aq.enable_observer()
        aq(x, lens, 1)
        aq.disable_observer()
        if quantizer == "LSQPlusFakeQuantize":
            aq.zero_point.data += 0.37  # a learned, fractional zero point
        aq.enable_fake_quant()
        ql.weight_fake_quant.enable_fake_quant()
        with torch.no_grad():
            x_fq = aq(x, lens, 1)
            y = ql(x_fq)
        out.update({"x%d" % i: x, "lens%d" % i: lens, "w%d" % i: lin.weight.data, "b%d" % i: lin.bias.data,
                    "a_scale%d" % i: aq.scale.data.clone(), "a_zp%d" % i: aq.zero_point.data.clone(),
                    "w_scale%d" % i: ql.weight_fake_quant.scale.clone(), "w_zp%d" % i: ql.weight_fake_quant.zero_point.clone(),
                    "x_fq%d" % i: x_fq, "y%d" % i: y,
                    "p%d" % i: np.array([a_bit, w_bit, int(quantizer == "LSQPlusFakeQuantize"), aq.quant_min, aq.quant_max,
                                         ql.weight_fake_quant.quant_min, ql.weight_fake_quant.quant_max])})
        if i == 0:
            out["keys_fixed"] = np.array(sorted(aq.state_dict().keys()))
            out["keys_qlinear"] = np.array(sorted(ql.state_dict().keys()))
        if quantizer == "LSQPlusFakeQuantize":
            out["keys_lsqplus"] = np.array(sorted(aq.state_dict().keys()))
        i += 1
    out["n"] = i
    save("qlinear", **out)


def gen_extra():
    """Goldens for the restatements no shipped twc/minmax config reaches but the registries expose:
    AvgQuantileObserver, MSEObserver / AvgMSEObserver, LSQPlusObserver, LSQFakeQuantize, per-channel LSQ+, QEmbedding."""
    g = torch.Generator().manual_seed(107)
    out = {}
    # ---- AvgQuantileObserver (observer.py:240-282): three batches, with and without pad mask ----
    for cname, shape, seq_pos, lens in (("quant_mask", (4, 24, 96), 1, [24, 9, 1, 17]), ("quant_nomask", (3, 40, 64), -1, None)):
        xs = [_act(g, shape, len(shape) - 1) for _ in range(3)]
        mask = None if lens is None else torch.tensor(lens)
        for thr, bins in ((0.999, 2048), (0.99, 256)):
            o = R_obs.AvgQuantileObserver(bit=6, symmetric=False, ch_axis=-1, threshold=thr, bins=bins)
            tr = []
            for x in xs:
                o(x, observation_mask=mask, seq_pos=seq_pos)
                tr.append([float(o.min_val), float(o.max_val)])
            out["%s_trace_%d" % (cname, bins)] = np.array(tr, dtype=np.float32)
        for b, x in enumerate(xs):
            out["%s_x%d" % (cname, b)] = x
        out["%s_lens" % cname] = np.array(lens if lens is not None else [], dtype=np.int64)
        out["%s_meta" % cname] = np.array([seq_pos])
    # ---- MSEObserver / AvgMSEObserver (observer.py:285-409): 1-D (symmetric), 1-D one-sided, 2-D (asymmetric, 4-bit) ----
    lens = torch.tensor([10, 4])
    xs = [_act(g, (2, 10, 24), 2) for _ in range(2)]
    out.update({"mse_x0": xs[0], "mse_x1": xs[1], "mse_lens": lens})
    for tag, cls, bit, sym in (("mse_sym", R_obs.MSEObserver, 6, True), ("avgmse_sym", R_obs.AvgMSEObserver, 6, True),
                               ("mse_asym4", R_obs.MSEObserver, 4, False), ("avgmse_asym4", R_obs.AvgMSEObserver, 4, False)):
        o = cls(bit=bit, symmetric=sym, ch_axis=-1)
        tr = []
        for x in xs:
            o(x, observation_mask=lens, seq_pos=1)
            tr.append([float(o.min_val), float(o.max_val)])
        out[tag + "_trace"] = np.array(tr, dtype=np.float32)
    xp = torch.rand(2, 10, 24, generator=g) * 3
    o = R_obs.AvgMSEObserver(bit=6, symmetric=False, ch_axis=-1)
    o(xp)
    out.update({"mse_pos_x": xp, "mse_pos": np.array([float(o.min_val), float(o.max_val)], dtype=np.float32)})
    # ---- LSQPlusObserver (observer.py:148-173): per-tensor and per-channel ----
    w = torch.randn(16, 48, generator=g) * 0.05
    w[3] *= 6
    for tag, ch in (("lsqobs_t", -1), ("lsqobs_c", 0)):
        o = R_obs.LSQPlusObserver(bit=4, symmetric=True, ch_axis=ch)
        o(w)
        s, z = o.calculate_qparams(o.min_val, o.max_val)
        out.update({tag + "_min": o.min_val, tag + "_max": o.max_val, tag + "_scale": s, tag + "_zp": z})
    out["lsq_w"] = w
    # ---- LSQFakeQuantize (fake_quant.py:129-167) and per-channel LSQ+ (fake_quant.py:170-209) forward ----
    x = _act(g, (2, 16, 48), 2)
    out["lsq_x"] = x
    for tag, quantizer, obs, ch, inp in (("lsq_t", "LSQFakeQuantize", "MinMaxObserver", -1, x),
                                         ("lsq_c", "LSQFakeQuantize", "MinMaxObserver", 0, w),
                                         ("lsqplus_c", "LSQPlusFakeQuantize", "MinMaxObserver", 0, w)):
        q = R_qm.Quantizer(None, QC(quantizer, obs, 6, True, ch))
        q.enable_observer()
        q(inp)
        q.disable_observer()
        q.enable_fake_quant()
        if quantizer == "LSQPlusFakeQuantize":
            q.zero_point.data += 0.3
        with torch.no_grad():
            y = q(inp)
        out.update({tag + "_y": y, tag + "_scale": q.scale.data.clone(), tag + "_zp": q.zero_point.data.clone().float()})
    # ---- QEmbedding (quantized_module.py:75-100) ----
    emb = torch.nn.Embedding(50, 32, padding_idx=0)
    emb.weight.data = torch.randn(50, 32, generator=g) * 0.1
    qe = R_qm.Quantizer(emb, QC("FixedFakeQuantize", "MinMaxObserver", 6, True, 0))
    ids = torch.randint(0, 50, (3, 11), generator=g)
    qe.weight_fake_quant.enable_observer()
    qe(ids)
    qe.weight_fake_quant.disable_observer()
    qe.weight_fake_quant.enable_fake_quant()
    with torch.no_grad():
        y = qe(ids)
    out.update({"emb_w": emb.weight.data, "emb_ids": ids, "emb_y": y, "emb_scale": qe.weight_fake_quant.scale.clone()})
    save("extra", **out)


def gen_blocks():
    """Goldens for the model-level blocks of SURVEY section 8 f3, taken from the reference's UNMODIFIED quant_bert.py running its own
    quantization package on the CPU, in the quantized state of config 2 (LSQ+ / AvgPruneMinMax 6-bit, gamma migration) and of
    config 1 (Fixed / AvgMinMax 8-bit): every tensor that crosses the self-attention block (quant_bert.py:134-195) and the
    dense -> residual -> LayerNorm -> quantizer block (quant_bert.py:197-217, util_layernorm.py:14-52) of layer 0."""
    import copy
    from oracle import ref_model as RM
    out = {}
    for tag, kw in (("c2", dict(a_bit=6, w_bit=6, a_quantizer="LSQPlusFakeQuantize", a_observer="AvgPruneMinMaxObserver", delay=True)),
                    ("c1", dict(a_bit=8, w_bit=8, a_quantizer="FixedFakeQuantize", a_observer="AvgMinMaxObserver", delay=False))):
        ns = RM.load_stack("reference")
        qcfg = RM.quant_config(**kw)
        fp = RM.fp_bert(layers=2, hidden=128, heads=2, inter=256)
        model = RM.build_model(ns, copy.deepcopy(fp), qcfg, "cpu")
        batches = RM.synth_batches(3, 4, 32, 100, "cpu", seed=21)
        model, _ = RM.run_schedule(ns, model, qcfg, batches)
        layer = model.bert.encoder.layer[0]
        att, so = layer.attention.self, layer.attention.output
        cap = {}

        def grab(name, what="out"):
            def hook(mod, args, output):
                cap[name] = (args[0] if what == "in" else (output[0] if isinstance(output, tuple) else output)).detach().clone()
            return hook
        hs = [att.query.register_forward_hook(grab("q3")), att.key.register_forward_hook(grab("k3")), att.value.register_forward_hook(grab("v3")),
              att.attention_probs_post_act_fake_quantize.register_forward_hook(grab("probs", "in")),
              att.attention_probs_post_act_fake_quantize.register_forward_hook(grab("probs_fq")),
              att.register_forward_hook(grab("ctx_fq")), att.register_forward_pre_hook(lambda m, a: cap.__setitem__("mask", a[1].detach().clone())),
              so.dense.register_forward_hook(grab("so_h")), so.register_forward_pre_hook(lambda m, a: cap.__setitem__("so_res", a[1].detach().clone())),
              so.LayerNorm.layernorm_post_act_fake_quantize.register_forward_hook(grab("so_ln", "in")),
              so.register_forward_hook(grab("so_y"))]
        with torch.no_grad():
            model(**batches[1])
        for h in hs:
            h.remove()
        for k, v in cap.items():
            out["%s_%s" % (tag, k)] = v
        for qname in ("query_permute", "key_transpose", "value_permute", "attention_probs", "context_view"):
            q = getattr(att, qname + "_post_act_fake_quantize")
            out["%s_%s_scale" % (tag, qname)] = q.scale.detach().reshape(-1).float().clone()
            out["%s_%s_zp" % (tag, qname)] = q.zero_point.detach().reshape(-1).float().clone()
        q = so.LayerNorm.layernorm_post_act_fake_quantize
        out["%s_so_scale" % tag], out["%s_so_zp" % tag] = q.scale.detach().reshape(-1).float().clone(), q.zero_point.detach().reshape(-1).float().clone()
        res = so.before_LayerNorm_residual
        out["%s_so_gamma" % tag] = res.gamma.detach().clone() if getattr(res, "mul_gamma", False) else torch.zeros(0)
        ln = so.LayerNorm
        inner = ln.layernorm
        out["%s_so_ln_weight" % tag] = inner.weight.detach().clone() if inner.weight is not None else torch.zeros(0)
        out["%s_so_ln_bias" % tag] = (inner.bias.detach().clone() if inner.bias is not None else
                                     (ln.bias.detach().clone() if getattr(ln, "bias", None) is not None else torch.zeros(0)))
        out["%s_meta" % tag] = np.array([att.num_attention_heads, att.attention_head_size, kw["a_bit"], int(kw["a_quantizer"] == "LSQPlusFakeQuantize")])
        out["%s_eps" % tag] = np.array([inner.eps], dtype=np.float64)
    save("blocks", **out)


if __name__ == "__main__":
    torch.set_num_threads(1)  # reduction order independent of the thread count
    gen_fq_per_tensor()
    gen_fq_per_channel()
    gen_lsqplus()
    gen_qparams()
    gen_observers()
    gen_minmax_per_channel()
    gen_mse()
    gen_qlinear()
    gen_extra()
    gen_blocks()
