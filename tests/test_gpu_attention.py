"""-m gpu: K8 / K9, the attention-side contractions with their activation quantizers as prologues (osq_attn_scores_fq_f32,
osq_attn_context_fq_f32; model/quant_bert.py:148-150, :169-172, :185-193), through the C ABI.

Oracle: the reference's own chain restated on the CPU -- oracle fake-quant of each operand (bit-exact bins), then the matmul of
the dequantised tensors in fp64 (the exact value of what the reference's fp32 torch.matmul approximates).  The kernels perform
the contraction on the integer bins, so they must agree with the fp64 product to fp32 rounding of the final scale
(<= 1e-6 relative to the largest output), far inside the 1e-3 bar for dequantised floats; the context quantizer in K9's epilogue
must be bit-exact with K1 on the un-quantised context the same kernel produces without it."""
import math

import pytest
import torch

from oracle import osq_oracle as O
from outlier_suppression_b200 import ops

pytestmark = pytest.mark.gpu


def _q(scale, zp, qmin, qmax, numel, lsq):
    sc = torch.tensor([scale], dtype=torch.float32)
    z = torch.tensor([float(zp)], dtype=torch.float32)
    g = 1.0 / (numel * qmax) ** 0.5 if lsq else 0.0
    dev = dict(scale=sc.cuda(), zp=(z if lsq else z.to(torch.int32)).cuda(), qmin=qmin, qmax=qmax, g=g)
    return sc, z, dev


def _oracle_fq(x, sc, z, qmin, qmax, lsq):
    return O.fq_lsqplus_per_tensor(x, sc.clone(), z.clone(), qmin, qmax) if lsq else O.fq_per_tensor(x, float(sc), int(z), qmin, qmax)


CASES = [
    # B, h, Sq, Sk, d, bits, lsq
    (2, 12, 128, 128, 64, 6, True),      # config 2 geometry at a size the CPU oracle finishes quickly
    (1, 4, 200, 264, 64, 6, True),       # ragged tiles on both sides
    (2, 3, 77, 52, 32, 8, False),        # 8-bit FixedFakeQuantize (config 1), d = 32
    (1, 2, 130, 516, 128, 4, False),     # d = 128, > 4 key tiles
    (1, 16, 62, 1024, 64, 6, True),      # BART-like cross attention: few queries, long keys
]


@pytest.mark.parametrize("B,h,Sq,Sk,d,bits,lsq", CASES)
def test_scores_and_context(B, h, Sq, Sk, d, bits, lsq):
    g = torch.Generator().manual_seed(B * 1000 + Sq + Sk + d)
    H = h * d
    q3 = torch.randn(B, Sq, H, generator=g)
    k3 = torch.randn(B, Sk, H, generator=g) * 1.5 + 0.2
    v3 = torch.randn(B, Sk, H, generator=g)
    mask = torch.zeros(B, 1, 1, Sk)
    mask[:, :, :, Sk - Sk // 5:] = -10000.0
    qmin, qmax = 0, 2 ** bits - 1
    heads = lambda t, S: t.view(B, S, h, d).permute(0, 2, 1, 3)            # transpose_for_scores
    qsc, qz, qdev = _q(float(q3.abs().max()) * 2 / qmax * 0.8, qmax // 2, qmin, qmax, q3.numel(), lsq)
    ksc, kz, kdev = _q(float(k3.abs().max()) * 2 / qmax * 0.8, qmax // 2 - 3, qmin, qmax, k3.numel(), lsq)
    out_mul = float(torch.tensor(1.0) / torch.tensor(math.sqrt(d), dtype=torch.float32))
    # ---- K8
    got = ops.attn_scores_fq(heads(q3.cuda(), Sq), heads(k3.cuda(), Sk), qdev, kdev, out_mul=out_mul, mask=mask.cuda().contiguous())
    qf = _oracle_fq(heads(q3, Sq), qsc, qz, qmin, qmax, lsq).double()
    kf = _oracle_fq(heads(k3, Sk), ksc, kz, qmin, qmax, lsq).double()
    want = torch.matmul(qf, kf.transpose(-1, -2)) * out_mul + mask.double()
    err = float((got.double().cpu() - want).abs().max())
    assert err <= 2e-6 * float(want.abs().max()), err
    plain = ops.attn_scores_fq(heads(q3.cuda(), Sq), heads(k3.cuda(), Sk), qdev, kdev)
    want_plain = torch.matmul(qf, kf.transpose(-1, -2))
    assert float((plain.double().cpu() - want_plain).abs().max()) <= 2e-6 * float(want_plain.abs().max())
    # the reference's own fp32 chain on the same device is within the documented 1e-3
    ref32 = torch.matmul(qf.float().cuda(), kf.float().cuda().transpose(-1, -2))
    assert float((plain - ref32).abs().max()) <= 1e-3 * float(ref32.abs().max())
    # ---- K9 on the softmax of those scores
    probs = torch.softmax(got, dim=-1)
    psc, pz, pdev = _q(1.0 / qmax, 0, qmin, qmax, probs.numel(), lsq)
    vsc, vz, vdev = _q(float(v3.abs().max()) * 2 / qmax * 0.9, qmax // 2 + 1, qmin, qmax, v3.numel(), lsq)
    ctx = ops.attn_context_fq(probs, heads(v3.cuda(), Sk), pdev, vdev)
    pf = _oracle_fq(probs.cpu(), psc, pz, qmin, qmax, lsq).double()
    vf = _oracle_fq(heads(v3, Sk), vsc, vz, qmin, qmax, lsq).double()
    want_ctx = torch.matmul(pf, vf).permute(0, 2, 1, 3).contiguous().view(B, Sq, H)
    assert ctx.shape == (B, Sq, H)
    assert float((ctx.double().cpu() - want_ctx).abs().max()) <= 2e-6 * float(want_ctx.abs().max())
    # context quantizer in the epilogue: bit-exact with K1 on the un-quantised context
    osc, oz, odev = _q(float(ctx.abs().max()) * 2 / qmax * 0.7, qmax // 2, qmin, qmax, ctx.numel(), lsq)
    y, bins = ops.attn_context_fq(probs, heads(v3.cuda(), Sk), pdev, vdev, oq=odev, want_bins=True)
    y1, b1 = ops.fq_per_tensor(ctx, odev["scale"], odev["zp"], qmin, qmax, lsq_grad_factor=odev["g"], want_bins=True)
    assert torch.equal(y, y1) and torch.equal(bins, b1)
    assert torch.equal(y.cpu(), _oracle_fq(ctx.cpu(), osc, oz, qmin, qmax, lsq))


def test_error_paths():
    q = torch.randn(1, 2, 8, 64).cuda()
    sc, zp = torch.tensor([0.1]).cuda(), torch.tensor([31.0]).cuda()
    qa = dict(scale=sc, zp=zp, qmin=0, qmax=63, g=0.0)
    with pytest.raises(Exception):
        ops.attn_scores_fq(q[..., :48].contiguous(), q[..., :48].contiguous(), qa, qa)      # head size 48
    with pytest.raises(Exception):
        ops.attn_scores_fq(q, q[:, :, :6], qa, qa)                                            # sk % 4 != 0
    with pytest.raises(Exception):
        ops.attn_scores_fq(q.cpu(), q.cpu(), qa, qa)
    with pytest.raises(Exception):
        ops.attn_scores_fq(q, q, dict(scale=sc, zp=zp, qmin=0, qmax=1023, g=0.0), qa)         # 10 bits


@pytest.mark.parametrize("tag", ["c2", "c1"])
def test_blocks_against_the_reference_goldens(tag):
    """tests/golden/blocks.npz: layer 0 of the reference's unmodified quant_bert.py on the CPU (config 2 after gamma migration,
    config 1).  K8 -> softmax -> K9 must land on the reference's quantized context, K7 on its quantized LayerNorm output: equal
    except where a value sits on a rounding tie (the reference's fp32 GEMM / LayerNorm round differently from an exact
    contraction / another summation order), and then by one quantization step."""
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "blocks.npz"))
    t = lambda k: torch.from_numpy(g[k])
    heads, d, bit, lsq = (int(v) for v in g[tag + "_meta"])
    qmin, qmax = 0, 2 ** bit - 1
    q3, k3, v3 = t(tag + "_q3").cuda(), t(tag + "_k3").cuda(), t(tag + "_v3").cuda()
    B, S, H = q3.shape
    hv = lambda x: x.view(B, S, heads, d).permute(0, 2, 1, 3)

    def qd(name, numel):
        sc, z = t("%s_%s_scale" % (tag, name)), t("%s_%s_zp" % (tag, name))
        return dict(scale=sc.cuda(), zp=(z if lsq else z.to(torch.int32)).cuda(), qmin=qmin, qmax=qmax,
                    g=(1.0 / (numel * qmax) ** 0.5 if lsq else 0.0))
    inv = float(torch.tensor(1.0) / torch.tensor(math.sqrt(d), dtype=torch.float32))
    scores = ops.attn_scores_fq(hv(q3), hv(k3), qd("query_permute", q3.numel()), qd("key_transpose", k3.numel()), out_mul=inv,
                                mask=t(tag + "_mask").cuda().contiguous())
    probs = torch.softmax(scores, -1)
    assert float((probs.cpu() - t(tag + "_probs")).abs().max()) <= 2e-6
    ctx = ops.attn_context_fq(probs, hv(v3), qd("attention_probs", probs.numel()), qd("value_permute", v3.numel()),
                              oq=qd("context_view", q3.numel()))
    want = t(tag + "_ctx_fq")
    diff = (ctx.cpu() - want).abs()
    assert float(diff.max()) <= float(t("%s_context_view_scale" % tag)) * 1.001
    assert float((diff > 0).float().mean()) <= 2e-3
    # K7 on the block behind it
    gam, w, b = t(tag + "_so_gamma"), t(tag + "_so_ln_weight"), t(tag + "_so_ln_bias")
    dev = lambda x: x.cuda() if x.numel() else None
    sc, z = t(tag + "_so_scale"), t(tag + "_so_zp")
    y, _, ln = ops.residual_layernorm_fq(t(tag + "_so_h").cuda(), t(tag + "_so_res").cuda(), dev(gam), dev(w), dev(b), float(g[tag + "_eps"][0]),
                                         sc.cuda(), (z if lsq else z.to(torch.int32)).cuda(), qmin, qmax,
                                         lsq_grad_factor=(1.0 / (q3.numel() * qmax) ** 0.5 if lsq else 0.0), want_ln=True)
    assert float((ln.cpu() - t(tag + "_so_ln")).abs().max()) <= 4e-6 * float(t(tag + "_so_ln").abs().max())
    sdiff = (y.cpu() - t(tag + "_so_y")).abs()
    assert float(sdiff.max()) <= float(sc) * 1.001 and float((sdiff > 0).float().mean()) <= 2e-3
