"""-m gpu: a miniature encoder block wired exactly like quant_bert.py:134-349 (LN-output quantizer -> q/k/v
QLinear, permuted 4-D query quantizer, context quantizer -> output QLinear, GELU quantizer -> FFN QLinear) run through
the schedule of ptq_glue_quant.py:230-253 (gamma fold -> weight calibration -> token-wise-clipping style activation
calibration over two masked batches -> quantized forward) on the drop-in modules, against the CPU oracle doing the
same steps.  Observer states / qparams / fake-quantized activations: bit-exact.  Linear outputs: 1e-3 tolerance."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import osq_oracle as O
from tests.test_host_logic import QC

pytestmark = pytest.mark.gpu
B, S, H, HEADS, FF = 4, 48, 256, 4, 512
P = 0.9


def close(y, ref, rel=1e-3):
    y, ref = y.detach().double().cpu(), ref.detach().double().cpu()
    bad = (y - ref).abs() > rel * ref.abs() + rel * ref.abs().max()
    assert not bool(bad.any()), "%d / %d outside tolerance, max abs diff %g" % (int(bad.sum()), bad.numel(), float((y - ref).abs().max()))


def same(a, b):
    np.testing.assert_array_equal(a.detach().cpu().numpy(), b.detach().cpu().numpy())


def test_mini_encoder_block_matches_oracle():
    from outlier_suppression_b200 import quantization as Q
    from outlier_suppression_b200.quantization import quantized_module as qm
    from outlier_suppression_b200.quantization.state import set_observer_name
    g = torch.Generator().manual_seed(7)
    a_cfg = QC("LSQPlusFakeQuantize", "AvgPruneMinMaxObserver", 6, False, -1)
    w_cfg = QC("FixedFakeQuantize", "MinMaxObserver", 6, True, 0)

    def lin(k, n):
        m = torch.nn.Linear(k, n)
        m.weight.data = torch.randn(n, k, generator=g) * 0.05
        m.bias.data = torch.randn(n, generator=g) * 0.02
        return m

    net = torch.nn.Module()
    net.query, net.value, net.dense, net.ffn = (qm.Quantizer(lin(H, H), w_cfg), qm.Quantizer(lin(H, H), w_cfg),
                                                qm.Quantizer(lin(H, H), w_cfg), qm.Quantizer(lin(H, FF), w_cfg))
    for nme in ("ln_post_act_fake_quantize", "query_permute_post_act_fake_quantize", "context_view_post_act_fake_quantize",
                "gelu_post_act_fake_quantize"):
        setattr(net, nme, qm.Quantizer(None, a_cfg))
    net.cuda()
    set_observer_name(net)
    gamma = torch.rand(H, generator=g) * 2 + 0.2
    xs = [torch.randn(B, S, H, generator=g) for _ in range(2)]
    for x in xs:
        x[..., :3] *= 15
    lens = torch.tensor([S, S // 2, 5, S - 1])

    # ---- schedule on the GPU modules ----
    net.query.weight.data *= gamma.cuda()                 # gamma_migration.py:46-76 (rewrites weight.data)
    net.ffn.weight.data *= gamma.cuda()
    Q.enable_calibration_woquantization(net, quantizer_type="weight_fake_quant")

    def forward(x, mask):
        h = net.ln_post_act_fake_quantize(x, mask, 1)
        q = net.query(h)
        qp = net.query_permute_post_act_fake_quantize(q.view(B, S, HEADS, H // HEADS).permute(0, 2, 1, 3), mask, 2)
        v = net.value(h)
        ctx = net.context_view_post_act_fake_quantize(v, mask, 1)
        o = net.dense(ctx)
        gl = net.gelu_post_act_fake_quantize(F.gelu(o), mask, 1)
        return net.ffn(gl), h, qp, ctx, gl

    with torch.no_grad():
        forward(xs[0].cuda(), lens.cuda())                 # weight observers see one batch (ptq_glue_quant.py:234-235)
        Q.disable_all(net)
        for m in net.modules():                            # token_wise_clipping.set_ratio (token_wise_clipping.py:12-19)
            if isinstance(m, Q.quantized_module.LSQPlusFakeQuantize):
                m.observer.set_percentile(P); m.observer.cnt = 0; m.enable_observer(); m.disable_fake_quant()
        for x in xs:
            forward(x.cuda(), lens.cuda())
        Q.enable_quantization(net)
        before = dict(qm.stats)
        y, h, qp, ctx, gl = forward(xs[0].cuda(), lens.cuda())
    assert qm.stats["fused"] - before["fused"] == 4, "all four QLinear calls must take the fused kernel"

    # ---- the same schedule on the CPU oracle ----
    W = {n: (getattr(net, n).weight.detach().cpu(), getattr(net, n).bias.detach().cpu()) for n in ("query", "value", "dense", "ffn")}
    wq = {n: O.weight_qparams_minmax(W[n][0], 6, True) for n in W}
    for n in W:
        same(getattr(net, n).weight_fake_quant.scale, wq[n][0])
    st = {n: O.ObserverState() for n in ("ln", "qp", "ctx", "gl")}
    ll = lens.tolist()

    def o_lin(x, n):
        s, z, qmin, qmax = wq[n]
        return O.qlinear(x, W[n][0], s, z, qmin, qmax, W[n][1])

    def o_forward(x, observe):
        def aq(name, t, seq_pos):
            if observe:
                O.observe_avg_prune_minmax(st[name], t, P, name, ll, seq_pos)
                return t
            s, z = O.qparams_from_minmax(st[name].min_val, st[name].max_val, 0, 63, False)
            return O.fq_lsqplus_per_tensor(t, s.reshape(1), z.reshape(1), 0, 63)
        h = aq("ln", x, 1)
        # calibration runs with weight fake-quant DISABLED (disable_all), the final forward with it enabled
        lin_ = (lambda t, n: F.linear(t, W[n][0], W[n][1])) if observe else o_lin
        q = lin_(h, "query")
        qp = aq("qp", q.view(B, S, HEADS, H // HEADS).permute(0, 2, 1, 3), 2)
        v = lin_(h, "value")
        ctx = aq("ctx", v, 1)
        o = lin_(ctx, "dense")
        gl = aq("gl", F.gelu(o), 1)
        return lin_(gl, "ffn"), h, qp, ctx, gl

    for x in xs:
        o_forward(x, True)
    # calibration inputs of downstream observers come from cuBLAS fp32 GEMMs on the GPU vs MKL on the CPU (FP passes,
    # outside the hot path): their min/max agree to fp32 GEMM rounding, not bit-for-bit -- the first observer does.
    same(torch.stack([net.ln_post_act_fake_quantize.observer.min_val, net.ln_post_act_fake_quantize.observer.max_val]),
         torch.stack([st["ln"].min_val, st["ln"].max_val]))
    for name, mod in (("qp", net.query_permute_post_act_fake_quantize), ("ctx", net.context_view_post_act_fake_quantize),
                      ("gl", net.gelu_post_act_fake_quantize)):
        np.testing.assert_allclose(float(mod.observer.max_val), float(st[name].max_val), rtol=1e-4)
        np.testing.assert_allclose(float(mod.observer.min_val), float(st[name].min_val), rtol=1e-4, atol=1e-5)
        st[name].min_val, st[name].max_val = mod.observer.min_val.cpu(), mod.observer.max_val.cpu()  # continue from identical state
    y_ref, h_ref, qp_ref, ctx_ref, gl_ref = o_forward(xs[0], False)
    same(h, h_ref)                      # LN-output fake-quant: bit-exact
    close(y, y_ref)
    close(ctx, ctx_ref); close(gl, gl_ref); close(qp, qp_ref)


def test_query_key_value_share_one_launch_bit_identical(monkeypatch):
    """BERT's query | key | value QLinears read the same quantized tensor: grouped they run as ONE fused launch
    (quantized_module.QLinearGroup); every output must be bit-identical to the three separate launches."""
    from outlier_suppression_b200 import quantization as Q
    from outlier_suppression_b200.quantization import quantized_module as qm
    g = torch.Generator().manual_seed(3)
    a_cfg = QC("LSQPlusFakeQuantize", "AvgPruneMinMaxObserver", 6, False, -1)
    w_cfg = QC("FixedFakeQuantize", "MinMaxObserver", 6, True, 0)
    net = torch.nn.Module()
    for name, has_bias in (("query", True), ("key", False), ("value", True)):
        m = torch.nn.Linear(H, H if name != "value" else 2 * H, bias=has_bias)
        m.weight.data = torch.randn(m.out_features, H, generator=g) * 0.05
        setattr(net, name, qm.Quantizer(m, w_cfg))
    net.in_post_act_fake_quantize = qm.Quantizer(None, a_cfg)
    net.cuda()
    x = (torch.randn(B, S, H, generator=g) * 2).cuda()
    Q.enable_calibration_woquantization(net, quantizer_type="fake_quant")
    net.in_post_act_fake_quantize(x)
    for m in (net.query, net.key, net.value):
        m.weight_fake_quant(m.weight)
    Q.enable_quantization(net)
    assert net.query._sibling_group is not None

    @torch.no_grad()                                           # eval path (HF Trainer.evaluate); with grad the LSQ+ scale
    def run():                                                 # is a learnable Parameter and the modules stay unfused
        xq = net.in_post_act_fake_quantize(x)
        return [net.query(xq), net.key(xq), net.value(xq)]

    before = dict(qm.stats)
    grouped = run()
    assert qm.stats["grouped_launch"] == before["grouped_launch"] + 1 and qm.stats["grouped_hit"] == before["grouped_hit"] + 2
    again = run()                                              # a fresh input tensor object: launches again
    assert qm.stats["grouped_launch"] == before["grouped_launch"] + 2
    with torch.no_grad():
        only_key = net.key(net.in_post_act_fake_quantize(x))   # a sibling called first / alone still works
    monkeypatch.setenv("OSQ_DISABLE_GROUPING", "1")
    single = run()
    assert qm.stats["grouped_launch"] == before["grouped_launch"] + 3
    for a, b, c in zip(grouped, again, single):
        assert a.shape == c.shape
        same(a, c)
        same(b, c)
    same(only_key, single[1])
    # the views survive the usual head split of quant_bert.py:transpose_for_scores
    q4 = grouped[0].view(B, S, HEADS, H // HEADS).permute(0, 2, 1, 3)
    same(q4, single[0].view(B, S, HEADS, H // HEADS).permute(0, 2, 1, 3))
    # a weight rewrite through .data + toggler drops the concatenated pack
    net.key.weight.data *= 0.5
    Q.enable_quantization(net)
    net.key.weight_fake_quant.enable_observer(); net.key.weight_fake_quant(net.key.weight); net.key.weight_fake_quant.disable_observer()
    monkeypatch.delenv("OSQ_DISABLE_GROUPING")
    g2 = run()
    monkeypatch.setenv("OSQ_DISABLE_GROUPING", "1")
    s2 = run()
    for a, c in zip(g2, s2):
        same(a, c)
    assert not torch.equal(g2[1], grouped[1])


def test_host_step_runner_matches_direct_calls():
    """hostio.HostStepRunner: pinned upload -> fn -> pinned download on side streams, double-buffered; results must
    equal the plain synchronous calls for every step (buffer reuse across more steps than the pipeline depth)."""
    from outlier_suppression_b200.hostio import HostStepRunner
    g = torch.Generator().manual_seed(5)
    w = torch.randn(64, 64, generator=g).cuda()
    fn = lambda x: torch.tanh(x @ w) * 3.0  # noqa: E731
    batches = [torch.randn(128, 64, generator=g).pin_memory() for _ in range(7)]
    for graph in (False, True):
        runner = HostStepRunner(fn, (128, 64), (128, 64), "cuda", depth=2, graph=graph)
        got = []
        for b in batches:
            i = runner.submit(b)
            if i >= 1:
                got.append(runner.result(i - 1).clone())
        got.append(runner.result(len(batches) - 1).clone())
        runner.drain()
        for b, y in zip(batches, got):
            same(y, fn(b.cuda()))
    with pytest.raises(IndexError):
        runner.result(0)
    assert runner.h2d_bytes == 128 * 64 * 4 and runner.d2h_bytes == 128 * 64 * 4


def test_quantizer_hands_bins_to_qlinear():
    """Second forward on: the activation quantizer's launch also writes uint8 bins and the QLinear (single and
    grouped) consumes them (stats['bins_in']); outputs stay bit-identical, and an in-place edit of the quantized
    tensor invalidates the bins (falls back to re-quantising the fp32 values)."""
    from outlier_suppression_b200 import quantization as Q
    from outlier_suppression_b200.quantization import quantized_module as qm
    g = torch.Generator().manual_seed(9)
    a_cfg = QC("LSQPlusFakeQuantize", "AvgPruneMinMaxObserver", 6, False, -1)
    w_cfg = QC("FixedFakeQuantize", "MinMaxObserver", 6, True, 0)
    net = torch.nn.Module()
    for name in ("query", "key", "value", "dense"):
        m = torch.nn.Linear(H, H)
        m.weight.data = torch.randn(H, H, generator=g) * 0.05
        setattr(net, name, qm.Quantizer(m, w_cfg))
    net.in_post_act_fake_quantize = qm.Quantizer(None, a_cfg)
    net.ctx_post_act_fake_quantize = qm.Quantizer(None, QC("FixedFakeQuantize", "AvgMinMaxObserver", 8, False, -1))
    net.cuda()
    x = (torch.randn(B, S, H, generator=g) * 2).cuda()
    Q.enable_calibration_woquantization(net, quantizer_type="fake_quant")
    net.in_post_act_fake_quantize(x); net.ctx_post_act_fake_quantize(x)
    for m in (net.query, net.key, net.value, net.dense):
        m.weight_fake_quant(m.weight)
    Q.enable_quantization(net)

    @torch.no_grad()
    def run():
        xq = net.in_post_act_fake_quantize(x)
        outs = [net.query(xq), net.key(xq), net.value(xq)]
        cq = net.ctx_post_act_fake_quantize(x)
        outs.append(net.dense(cq))
        return outs

    before = qm.stats["bins_in"]
    first = run()                                   # nobody asked for bins yet
    assert qm.stats["bins_in"] == before
    second = run()                                  # grouped q|k|v launch + the single dense launch take bins
    assert qm.stats["bins_in"] == before + 2
    for a, b in zip(first, second):
        same(a, b)
    with torch.no_grad():
        xq = net.in_post_act_fake_quantize(x)
        assert xq._osq_bins[0].dtype == torch.uint8 and xq._osq_bins[0].shape == xq.shape
        xq.mul_(0.5)                                # in-place edit: the bins no longer describe the tensor
        n_before = qm.stats["bins_in"]
        edited = net.query(xq)
        assert qm.stats["bins_in"] == n_before
        ref = net.query(net.in_post_act_fake_quantize(x * 1.0).mul(0.5).clone())  # untagged clone: unfused path on the same values
    close(edited, ref)
