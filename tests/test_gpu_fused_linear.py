"""-m gpu: the fused fake-quant + Linear tcgen05 kernel.

Parity bar (BASELINE.json north_star): activation / weight bins BIT-EXACT (checked through the kernel's
debug side output and the packed weight codes); Y within 1e-3 relative of the reference's fp32
F.linear path, written as |dY| <= 1e-3*|Y_ref| + 1e-3*max|Y_ref| (the pure elementwise
relative test is ill-posed at cancellation zeros).  Additionally Y must match the EXACT integer
contraction of the bins to fp32 rounding (1e-5), which is size independent."""
import numpy as np
import pytest
import torch

from oracle import osq_oracle as O
from tests.test_host_logic import QC

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def close(y, ref, rel=1e-3):
    y, ref = y.detach().double().cpu(), ref.detach().double().cpu()
    bad = (y - ref).abs() > rel * ref.abs() + rel * ref.abs().max()
    assert not bool(bad.any()), "%d / %d outside tolerance, max abs diff %g" % (int(bad.sum()), bad.numel(), float((y - ref).abs().max()))


def run_case(m, k, n, a_bit, w_bit, lsq, seed, gamma=False, use_code_cache=True):
    from outlier_suppression_b200 import ops
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(m, k, generator=g)
    a[:, :3] *= 20
    w, bias = O.synth_linear(n, k, seed=seed + 1, gamma=gamma)
    a_qmin, a_qmax = O.quant_range(a_bit, False)
    mn, mx = O.global_minmax(a)
    a_scale, a_zp = O.qparams_from_minmax(mn * 0.7, mx * 0.7, a_qmin, a_qmax, False)
    w_scale, w_zp, w_qmin, w_qmax = O.weight_qparams_minmax(w, w_bit, True)
    if lsq:
        a_scale_t = a_scale.reshape(1).clone()
        a_zp_t = a_zp.reshape(1).float() + 0.37
        a_zp_t = a_zp_t.clamp(a_qmin, a_qmax)
    else:
        a_scale_t, a_zp_t = a_scale.reshape(1), a_zp.reshape(1).to(torch.int32)
    y_ref, qa_ref, qw_ref = O.fused_fq_linear(a, a_scale_t.clone(), a_zp_t.clone(), a_qmin, a_qmax, lsq, w, w_scale, w_zp,
                                              w_qmin, w_qmax, bias)
    codes, rowsum = ops.pack_weight(w.cuda(), w_scale.cuda(), w_zp.cuda(), w_qmin, w_qmax)
    np.testing.assert_array_equal(codes.cpu().numpy(), qw_ref.numpy().astype(np.int8))          # weight bins bit-exact
    np.testing.assert_array_equal(rowsum.cpu().numpy(), qw_ref.sum(1).numpy().astype(np.int32))
    gfac = 1.0 / (a.numel() * a_qmax) ** 0.5 if lsq else 0.0
    y, dbg = ops.fused_fq_linear(a.cuda(), a_scale_t.cuda(), a_zp_t.cuda(), a_qmin, a_qmax, codes, w_scale.cuda(), rowsum,
                                 bias.cuda(), lsq_grad_factor=gfac, want_codes=True, use_code_cache=use_code_cache)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dbg.cpu().numpy(), (qa_ref - a_qmin).numpy().astype(np.uint8))  # activation bins bit-exact
    close(y, y_ref)
    # exact contraction of the bins (size-independent property)
    s_eff = a_scale_t if not lsq else O.lsqplus_effective_qparams(a_scale_t, a_zp_t, a.numel(), a_qmax)[0]
    z_int = torch.round(a_zp_t.float())
    acc = (qa_ref.double() - z_int.double()) @ qw_ref.double().t()
    exact = acc * (s_eff.double() * w_scale.double())[None, :] + bias.double()[None, :]
    np.testing.assert_allclose(y.double().cpu().numpy(), exact.numpy(), rtol=2e-6, atol=2e-6 * float(exact.abs().max()))
    return y


@pytest.mark.parametrize("m,k,n", [(128, 128, 16), (300, 768, 768), (257, 768, 3072), (515, 3072, 768), (129, 1024, 1024),
                                   (64, 256, 400), (1, 128, 48)])
def test_fused_shapes_6bit(m, k, n):
    run_case(m, k, n, 6, 6, False, seed=m + k + n)


@pytest.mark.parametrize("m,k,n", [(24, 1024, 1024), (24, 1024, 3072), (24, 4096, 1024), (24, 1024, 4096), (6, 768, 3072), (64, 1024, 1024),
                                   (40, 3072, 768), (17, 256, 640)])
def test_fused_decode_sized_launches(m, k, n):
    """BART generation (config 4) runs every decoder Linear at batch x beams rows per step (4 x 6 = 24): one row tile, N split over
    the CTAs in 128-column chunks; with and without the code cache for K > 1024."""
    y = run_case(m, k, n, 6, 6, True, seed=m + k + n)
    if k > 1024:
        assert torch.equal(y, run_case(m, k, n, 6, 6, True, seed=m + k + n, use_code_cache=False))


@pytest.mark.parametrize("a_bit,w_bit,lsq", [(8, 8, False), (4, 4, False), (6, 4, False), (6, 6, True), (8, 8, True)])
def test_fused_bits_and_lsqplus(a_bit, w_bit, lsq):
    run_case(384, 768, 768, a_bit, w_bit, lsq, seed=a_bit * 10 + w_bit, gamma=True)


def test_fused_streaming_k_with_and_without_code_cache():
    """K = 3072 does not fit in shared memory: pass 0 spills the bins to the L2 code cache and later N chunks
    TMA-load them; without a cache every chunk re-converts.  Both must give identical results.  The
    multi-block case makes every CTA run the cached protocol several times (barrier phases wrap)."""
    y1 = run_case(515, 3072, 768, 6, 6, False, seed=11, use_code_cache=True)
    y2 = run_case(515, 3072, 768, 6, 6, False, seed=11, use_code_cache=False)
    assert torch.equal(y1, y2)
    run_case(148 * 128 + 300, 2048, 512, 8, 8, False, seed=12, use_code_cache=True)
    run_case(100, 4096, 1024, 6, 6, True, seed=13, use_code_cache=True)


def test_fused_multi_block_persistent():
    """more 128-row blocks than SMs: every CTA loops over several blocks (ring phases wrap)."""
    run_case(148 * 128 * 2 + 77, 256, 64, 6, 6, False, seed=5)


def test_fused_full_size_bert_base_site():
    """BASELINE config 2 site 768->768 at M = 32*512 = 16384 against the CPU oracle."""
    run_case(16384, 768, 768, 6, 6, True, seed=2)


def test_qlinear_module_golden(golden):
    """module-level drop-in: Quantizer(None, a_qconfig) -> Quantizer(linear, w_qconfig) exactly as
    quant_bert.py wires them; must take the fused path and reproduce the reference's Y."""
    from outlier_suppression_b200.quantization import quantized_module as qm
    g = golden("qlinear")
    for i in range(int(g["n"])):
        a_bit, w_bit, lsq, aqmin, aqmax, wqmin, wqmax = (int(v) for v in g["p%d" % i])
        w, b, x = T(g["w%d" % i]), T(g["b%d" % i]), T(g["x%d" % i])
        lin = torch.nn.Linear(w.shape[1], w.shape[0])
        lin.weight.data, lin.bias.data = w.clone(), b.clone()
        ql = qm.Quantizer(lin, QC("FixedFakeQuantize", "MinMaxObserver", w_bit, True, 0)).cuda()
        aq = qm.Quantizer(None, QC("LSQPlusFakeQuantize" if lsq else "FixedFakeQuantize", "AvgMinMaxObserver", a_bit, False, -1)).cuda()
        lens = T(g["lens%d" % i]).cuda()
        xg = x.cuda()
        ql.weight_fake_quant.enable_observer(); ql(xg); ql.weight_fake_quant.disable_observer()
        aq.enable_observer(); aq(xg, lens, 1); aq.disable_observer()
        np.testing.assert_array_equal(ql.weight_fake_quant.scale.cpu().numpy(), g["w_scale%d" % i])
        if lsq:
            aq.zero_point.data += 0.37
        np.testing.assert_array_equal(aq.scale.data.reshape(()).cpu().numpy(), g["a_scale%d" % i])
        aq.enable_fake_quant(); ql.weight_fake_quant.enable_fake_quant()
        before = dict(qm.stats)
        with torch.no_grad():
            x_fq = aq(xg, lens, 1)
            y = ql(x_fq)
        np.testing.assert_array_equal(x_fq.cpu().numpy(), g["x_fq%d" % i])
        assert qm.stats["fused"] == before["fused"] + 1, "QLinear did not take the fused kernel"
        close(y, T(g["y%d" % i]))
        # un-tagged input -> reference semantics (separate kernels), still correct
        with torch.no_grad():
            y2 = ql(x_fq.clone())
        assert qm.stats["unfused"] == before["unfused"] + 1
        close(y2, T(g["y%d" % i]), rel=1e-5)


def test_weight_cache_invalidation_after_gamma_fold():
    """gamma_migration.py:46-76 rewrites weight.data in place; the togglers drop the packed cache."""
    from outlier_suppression_b200 import quantization as Q
    from outlier_suppression_b200.quantization import quantized_module as qm
    torch.manual_seed(0)
    lin = torch.nn.Linear(128, 32)
    net = torch.nn.Module()
    net.dense = qm.Quantizer(lin, QC("FixedFakeQuantize", "MinMaxObserver", 6, True, 0))
    net.in_act_fake_quant = qm.Quantizer(None, QC("FixedFakeQuantize", "AvgMinMaxObserver", 6, False, -1))
    net.cuda()
    x = torch.randn(4, 16, 128, device="cuda")

    def calibrate_and_run():
        obs = net.dense.weight_fake_quant.observer  # fresh running extrema, as in a real (single) weight calibration
        obs.min_val = torch.tensor(float("inf"), device="cuda"); obs.max_val = torch.tensor(float("-inf"), device="cuda")
        Q.enable_calibration_woquantization(net, "weight_fake_quant"); net.dense(x)
        Q.enable_calibration_woquantization(net, "act_fake_quant"); net.in_act_fake_quant.observer.cnt = 0; net.in_act_fake_quant(x)
        Q.enable_quantization(net)
        with torch.no_grad():
            return net.dense(net.in_act_fake_quant(x))

    def oracle_run():
        w, b = net.dense.weight.detach().cpu(), net.dense.bias.detach().cpu()
        ws, wz, wqmin, wqmax = O.weight_qparams_minmax(w, 6, True)
        mn, mx = O.global_minmax(x.cpu())
        s, z = O.qparams_from_minmax(mn, mx, 0, 63, False)
        return O.qlinear(O.fq_per_tensor(x.cpu(), s.item(), int(z.item()), 0, 63), w, ws, wz, wqmin, wqmax, b)

    close(calibrate_and_run(), oracle_run())
    gamma = torch.rand(128, device="cuda") * 2 + 0.2
    net.dense.weight.data *= gamma  # not tracked by weight._version
    close(calibrate_and_run(), oracle_run())


PAIR_CASES = "[(256, 768, 768), (1000, 768, 2304), (16384, 768, 768), (300, 1024, 512), (4096, 256, 3072)]"


def test_cta_pair_mode_and_register_path_in_subprocess():
    """The launch plan is read from the environment once per process, so every kernel variant gets a child process of
    its own whatever the default is: OSQ_FUSED_CLUSTER=2 (tcgen05 cta_group::2: one W tile copy per CTA pair),
    OSQ_FUSED_CLUSTER=1 (one CTA per tile), OSQ_FUSED_XTMA=0 (fp32 A through 128-bit register loads instead of TMA
    landing slots) and OSQ_FUSED_STORE3D=0 (per-warp 32 x 128 B tile stores instead of the pair-shared 256-byte-row
    3-D stores).  Same parity bar as every other case."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import torch; from tests.test_gpu_fused_linear import run_case\n"
            "for i, (m, k, n) in enumerate(%s):\n"
            "    run_case(m, k, n, 6, 6, True, 100 + i, gamma=bool(i & 1))\n"
            "    run_case(m, k, n, 8, 8, False, 200 + i)\n"
            "print('variant ok')\n" % PAIR_CASES)
    for env in ({"OSQ_FUSED_CLUSTER": "2"}, {"OSQ_FUSED_CLUSTER": "1"}, {"OSQ_FUSED_XTMA": "0"}, {"OSQ_FUSED_STORE3D": "0"},
                {"OSQ_FUSED_STORE3D": "0", "OSQ_FUSED_CLUSTER": "1"}):
        r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, **env), capture_output=True, text=True,
                           timeout=240)
        assert r.returncode == 0 and "variant ok" in r.stdout, "%s failed:\n%s\n%s" % (env, r.stdout[-2000:], r.stderr[-3000:])


@pytest.mark.parametrize("m,k,n", [(300, 768, 768), (16384, 768, 2304), (1000, 3072, 768), (130, 128, 16), (4096, 1024, 1024)])
@pytest.mark.parametrize("lsq", [False, True])
def test_bins_in_launch_is_bit_identical(m, k, n, lsq):
    """The fake-quant kernel's uint8 side output (bin - qmin) equals the bins the fused kernel derives itself, and a
    bins-in launch (A = NULL, a_codes = bins) reproduces the fp32-in result bit for bit."""
    from outlier_suppression_b200 import ops
    g = torch.Generator().manual_seed(m + k + n)
    a = torch.randn(m, k, generator=g)
    a[:, :3] *= 20
    w, bias = O.synth_linear(n, k, seed=5)
    a_qmin, a_qmax = O.quant_range(6, False)
    mn, mx = O.global_minmax(a)
    a_scale, a_zp = O.qparams_from_minmax(mn * 0.7, mx * 0.7, a_qmin, a_qmax, False)
    if lsq:
        a_scale_t, a_zp_t = a_scale.reshape(1).cuda(), (a_zp.reshape(1).float() + 0.37).clamp(a_qmin, a_qmax).cuda()
        gfac = 1.0 / (a.numel() * a_qmax) ** 0.5
    else:
        a_scale_t, a_zp_t, gfac = a_scale.reshape(1).cuda(), a_zp.reshape(1).to(torch.int32).cuda(), 0.0
    w_scale, w_zp, w_qmin, w_qmax = O.weight_qparams_minmax(w, 6, True)
    codes, rowsum = ops.pack_weight(w.cuda(), w_scale.cuda(), w_zp.cuda(), w_qmin, w_qmax)
    ag = a.cuda()
    y_fq, bins = ops.fq_per_tensor(ag, a_scale_t, a_zp_t, a_qmin, a_qmax, lsq_grad_factor=gfac, want_bins=True)
    y_fq2, q16 = ops.fq_per_tensor(ag, a_scale_t, a_zp_t, a_qmin, a_qmax, lsq_grad_factor=gfac, want_codes=True)
    np.testing.assert_array_equal(y_fq.cpu().numpy(), y_fq2.cpu().numpy())
    np.testing.assert_array_equal(bins.cpu().numpy().astype(np.int16), q16.cpu().numpy() - a_qmin)
    y_ref, own = ops.fused_fq_linear(ag, a_scale_t, a_zp_t, a_qmin, a_qmax, codes, w_scale.cuda(), rowsum, bias.cuda(),
                                     lsq_grad_factor=gfac, want_codes=True)
    np.testing.assert_array_equal(own.cpu().numpy(), bins.cpu().numpy())
    for src in (ag, y_fq):       # raw or already fake-quantised activations: the fp32 tensor is not read in a bins-in launch
        y = ops.fused_fq_linear(src, a_scale_t, a_zp_t, a_qmin, a_qmax, codes, w_scale.cuda(), rowsum, bias.cuda(),
                                lsq_grad_factor=gfac, a_bins=bins)
        np.testing.assert_array_equal(y.cpu().numpy(), y_ref.cpu().numpy())


# ---------------------------------------------------------------------------------------------------------------
# Every shape bench.py times, at the size it times it (M = 32 x 512), and the BART-large sites of BASELINE config 4
# at the per-GPU eval batch (M = 4 x 1024) and at M = 16384.  Same bar as everywhere: bins bit-exact, Y tolerance.
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,n,cache", [(768, 3072, True), (3072, 768, True), (3072, 768, False), (768, 2304, True)])
def test_fused_benchmarked_shapes_at_full_size(k, n, cache):
    run_case(16384, k, n, 6, 6, True, seed=k + n, gamma=(k == 768), use_code_cache=cache)


@pytest.mark.parametrize("m", [4096, 16384])
@pytest.mark.parametrize("k,n", [(1024, 1024), (1024, 4096), (4096, 1024)])
def test_fused_bart_large_sites(m, k, n):
    run_case(m, k, n, 6, 6, True, seed=m + k + n, gamma=(k == 1024))


def test_asymmetric_8bit_weights_take_the_unfused_path():
    """q - zp of an asymmetric 8-bit weight spans [-255, 255] and does not fit the s8 operand.  The
    module must fall back to reference semantics (separate launches) and the C ABI must refuse to pack."""
    from outlier_suppression_b200 import _lib, ops
    from outlier_suppression_b200.quantization import quantized_module as qm
    torch.manual_seed(3)
    lin = torch.nn.Linear(256, 64)
    lin.weight.data += 0.3  # lopsided rows: zero points far from the middle
    ql = qm.Quantizer(lin, QC("FixedFakeQuantize", "MinMaxObserver", 8, False, 0)).cuda()
    aq = qm.Quantizer(None, QC("FixedFakeQuantize", "AvgMinMaxObserver", 8, False, -1)).cuda()
    x = torch.randn(4, 32, 256, device="cuda")
    ql.weight_fake_quant.enable_observer(); ql(x); ql.weight_fake_quant.disable_observer()
    aq.enable_observer(); aq(x); aq.disable_observer()
    aq.enable_fake_quant(); ql.weight_fake_quant.enable_fake_quant()
    before = dict(qm.stats)
    with torch.no_grad():
        y = ql(aq(x))
    assert qm.stats["unfused"] == before["unfused"] + 1 and qm.stats["fused"] == before["fused"]
    w, b = lin.weight.detach(), lin.bias.detach()
    mn, mx = w.min(1).values, w.max(1).values
    ws, wz = O.qparams_from_minmax(mn, mx, 0, 255, False)
    amn, amx = O.global_minmax(x.cpu())
    s, z = O.qparams_from_minmax(amn, amx, 0, 255, False)
    ref = O.qlinear(O.fq_per_tensor(x.cpu(), s.item(), int(z.item()), 0, 255), w, ws, wz.to(torch.int32), 0, 255, b)
    close(y, ref, rel=1e-4)
    wq = ql.weight_fake_quant
    with pytest.raises(_lib.OsqError, match="int8"):
        ops.pack_weight(ql.weight, wq.scale, wq.zero_point, 0, 255)
    # 7-bit asymmetric weights DO fit (|q - zp| <= 127) and stay on the fused path, bit-exact bins
    ql7 = qm.Quantizer(lin, QC("FixedFakeQuantize", "MinMaxObserver", 7, False, 0)).cuda()
    ql7.weight_fake_quant.enable_observer(); ql7(x); ql7.weight_fake_quant.disable_observer()
    ql7.weight_fake_quant.enable_fake_quant()
    before = dict(qm.stats)
    with torch.no_grad():
        y7 = ql7(aq(x))
    assert qm.stats["fused"] == before["fused"] + 1
    ws7, wz7 = O.qparams_from_minmax(mn, mx, 0, 127, False)
    ref7 = O.qlinear(O.fq_per_tensor(x.cpu(), s.item(), int(z.item()), 0, 255), w, ws7, wz7.to(torch.int32), 0, 127, b)
    close(y7, ref7)


def test_stale_qparams_void_the_producer_tag():
    """a tensor produced under OLD (scale, zero_point) must not be re-quantised by the fused kernel with the
    NEW ones (optimizer step, another observer pass, load_state_dict)."""
    from outlier_suppression_b200.quantization import quantized_module as qm
    from outlier_suppression_b200.quantization.fake_quant import bins_of, producer_of
    torch.manual_seed(0)
    lin = torch.nn.Linear(128, 32)
    ql = qm.Quantizer(lin, QC("FixedFakeQuantize", "MinMaxObserver", 6, True, 0)).cuda()
    aq = qm.Quantizer(None, QC("LSQPlusFakeQuantize", "AvgMinMaxObserver", 6, False, -1)).cuda()
    x = torch.randn(2, 64, 128, device="cuda")
    ql.weight_fake_quant.enable_observer(); ql(x); ql.weight_fake_quant.disable_observer()
    aq.enable_observer(); aq(x); aq.disable_observer()
    aq.enable_fake_quant(); ql.weight_fake_quant.enable_fake_quant()
    with torch.no_grad():
        ql(aq(x))            # first fused call: from now on the quantizer emits bins
        x_fq = aq(x)
        assert producer_of(x_fq) is aq and bins_of(x_fq) is not None
        y_good = ql(x_fq)
        aq.scale.mul_(1.5)   # what an optimizer step does
        assert producer_of(x_fq) is None and bins_of(x_fq) is None
        before = dict(qm.stats)
        y = ql(x_fq)
        assert qm.stats["unfused"] == before["unfused"] + 1
    close(y, y_good, rel=1e-5)  # reference semantics: Linear over the tensor as it is, NOT re-quantised with the new scale


def test_multi_site_entry_point_and_persistent_variant():
    """The list entry point in-process (default: sites issued one by one) and, in a child process with OSQ_FUSED_MULTI=1,
    as ONE persistent launch per run of compatible sites."""
    import os
    import subprocess
    import sys
    multi_site_case(300)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("from tests.test_gpu_fused_linear import multi_site_case\nmulti_site_case(300); multi_site_case(16384); print('multi ok')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, OSQ_FUSED_MULTI="1"), capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0 and "multi ok" in r.stdout, "persistent multi-site launch failed:\n%s\n%s" % (r.stdout[-2000:], r.stderr[-3000:])


def multi_site_case(m):
    """osq_fused_fq_linear_multi: one persistent launch walking a list of independent sites must reproduce the individual
    launches bit for bit -- barriers are invalidated and re-initialised per site, TMEM is re-allocated, the code cache of the
    K = 3072 site and the resident plans of the K = 768 sites alternate inside one grid; a site list longer than a run
    (4) and an incompatible site (K % 4 != ... n/a; different M -> different grid) split into several launches."""
    from outlier_suppression_b200 import ops
    shapes = [(768, 2304), (768, 768), (768, 3072), (3072, 768)] * 3   # 12 sites: runs of up to 4
    g = torch.Generator().manual_seed(7)
    sites, refs = [], []
    for i, (k, n) in enumerate(shapes):
        a = torch.randn(m if i != 5 else max(64, m // 2), k, generator=g)   # site 5 has another M: breaks the run
        a[:, :3] *= 20
        w, bias = O.synth_linear(n, k, seed=50 + i, gamma=bool(i & 1))
        mn, mx = O.global_minmax(a)
        a_scale, a_zp = O.qparams_from_minmax(mn * 0.7, mx * 0.7, 0, 63, False)
        a_scale_t, a_zp_t = a_scale.reshape(1).cuda(), (a_zp.reshape(1).float() + 0.37).clamp(0, 63).cuda()
        w_scale, w_zp, w_qmin, w_qmax = O.weight_qparams_minmax(w, 6, True)
        codes, rowsum = ops.pack_weight(w.cuda(), w_scale.cuda(), w_zp.cuda(), w_qmin, w_qmax)
        st = dict(a=a.cuda(), a_scale=a_scale_t, a_zp=a_zp_t, a_qmin=0, a_qmax=63, w_codes=codes, w_scale=w_scale.cuda(),
                  w_rowsum=rowsum, bias=bias.cuda(), lsq_grad_factor=1.0 / (a.numel() * 63) ** 0.5)
        sites.append(st)
        refs.append(ops.fused_fq_linear(st["a"], a_scale_t, a_zp_t, 0, 63, codes, st["w_scale"], rowsum, st["bias"],
                                        lsq_grad_factor=st["lsq_grad_factor"]))
    torch.cuda.synchronize()
    for rep in range(2):   # second pass: same workspace-free path again (barrier phases start from scratch)
        outs = ops.fused_fq_linear_multi(sites)
        torch.cuda.synchronize()
        for i, (y, r) in enumerate(zip(outs, refs)):
            assert torch.equal(y, r), "site %d differs (max |d| %g)" % (i, float((y - r).abs().max()))


@pytest.mark.parametrize("m,k,n", [(300, 768, 3072), (16384, 768, 3072), (515, 3072, 768), (129, 256, 48), (1000, 1024, 4096)])
@pytest.mark.parametrize("act,lsq_out", [("gelu", True), ("gelu", False), (None, True)])
def test_output_stage_matches_the_unfused_chain_bit_for_bit(m, k, n, act, lsq_out):
    """f3: the NEXT activation quantizer (and the GELU in front of it, quant_bert.py:277-280) fused into the Linear's
    epilogue.  Reference chain on the same device: fused Linear -> torch gelu (erf) -> K1 with bins.  The fused launch
    must reproduce the chain's dequantised tensor AND its uint8 bins bit for bit, and a bins-in launch of the next Linear
    fed with those bins must equal the fp32-in launch on the chain's tensor."""
    from outlier_suppression_b200 import ops
    g = torch.Generator().manual_seed(m + k + n)
    a = torch.randn(m, k, generator=g)
    a[:, :3] *= 20
    w, bias = O.synth_linear(n, k, seed=9, gamma=True)
    mn, mx = O.global_minmax(a)
    a_scale, a_zp = O.qparams_from_minmax(mn * 0.7, mx * 0.7, 0, 63, False)
    a_scale_t, a_zp_t = a_scale.reshape(1).cuda(), (a_zp.reshape(1).float() + 0.37).clamp(0, 63).cuda()
    g_in = 1.0 / (a.numel() * 63) ** 0.5
    w_scale, w_zp, w_qmin, w_qmax = O.weight_qparams_minmax(w, 6, True)
    codes, rowsum = ops.pack_weight(w.cuda(), w_scale.cuda(), w_zp.cuda(), w_qmin, w_qmax)
    ag = a.cuda()
    y = ops.fused_fq_linear(ag, a_scale_t, a_zp_t, 0, 63, codes, w_scale.cuda(), rowsum, bias.cuda(), lsq_grad_factor=g_in)
    z = torch.nn.functional.gelu(y) if act == "gelu" else y
    zmin, zmax = float(z.min()), float(z.max())
    o_scale, o_zp = O.qparams_from_minmax(torch.tensor(zmin * 0.8), torch.tensor(zmax * 0.8), 0, 63, False)
    if lsq_out:
        os_t, oz_t, g_out = o_scale.reshape(1).cuda(), (o_zp.reshape(1).float() + 0.21).clamp(0, 63).cuda(), 1.0 / (z.numel() * 63) ** 0.5
    else:
        os_t, oz_t, g_out = o_scale.reshape(1).cuda(), o_zp.reshape(1).to(torch.int32).cuda(), 0.0
    ref_fq, ref_bins = ops.fq_per_tensor(z.contiguous(), os_t, oz_t, 0, 63, lsq_grad_factor=g_out, want_bins=True)
    got_fq, got_bins = ops.fused_fq_linear(ag, a_scale_t, a_zp_t, 0, 63, codes, w_scale.cuda(), rowsum, bias.cuda(), lsq_grad_factor=g_in,
                                           out_q=dict(scale=os_t, zp=oz_t, qmin=0, qmax=63, g=g_out, act=act, bins=True))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(got_bins.cpu().numpy(), ref_bins.cpu().numpy())
    np.testing.assert_array_equal(got_fq.cpu().numpy(), ref_fq.cpu().numpy())
    assert len(torch.unique(got_bins)) > 8   # the quantizer is really exercised (not saturated)
    if n % 128 == 0 and n >= 128:
        # hand-off: the next Linear (n -> 64) bins-in on the fused bins == fp32-in on the chain's tensor
        w2, b2 = O.synth_linear(64, n, seed=10)
        ws2, wz2, q2min, q2max = O.weight_qparams_minmax(w2, 6, True)
        c2, r2 = ops.pack_weight(w2.cuda(), ws2.cuda(), wz2.cuda(), q2min, q2max)
        y_a = ops.fused_fq_linear(ref_fq, os_t, oz_t, 0, 63, c2, ws2.cuda(), r2, b2.cuda(), lsq_grad_factor=g_out)
        y_b = ops.fused_fq_linear(got_fq, os_t, oz_t, 0, 63, c2, ws2.cuda(), r2, b2.cuda(), lsq_grad_factor=g_out, a_bins=got_bins)
        np.testing.assert_array_equal(y_a.cpu().numpy(), y_b.cpu().numpy())


def test_output_stage_against_the_cpu_oracle():
    """The same output stage against the CPU oracle chain (F.linear -> erf GELU -> fake-quant).  CPU and GPU differ in the
    last ulp of the Linear (fp32 F.linear vs exact integer contraction) and of erf, so a value within an ulp of a rounding
    tie may land in the neighbouring bin: bins equal for >= 99.9 % of the elements and never more than one bin apart;
    dequantised values within one quantisation step."""
    from outlier_suppression_b200 import ops
    m, k, n = 512, 768, 3072
    g = torch.Generator().manual_seed(3)
    a = torch.randn(m, k, generator=g)
    a[:, :3] *= 20
    w, bias = O.synth_linear(n, k, seed=4, gamma=True)
    mn, mx = O.global_minmax(a)
    a_scale, a_zp = O.qparams_from_minmax(mn * 0.7, mx * 0.7, 0, 63, False)
    w_scale, w_zp, w_qmin, w_qmax = O.weight_qparams_minmax(w, 6, True)
    y_ref, _, _ = O.fused_fq_linear(a, a_scale.reshape(1), a_zp.reshape(1).to(torch.int32), 0, 63, False, w, w_scale, w_zp, w_qmin, w_qmax, bias)
    z = torch.nn.functional.gelu(y_ref)
    o_scale, o_zp = O.qparams_from_minmax(z.min() * 0.8, z.max() * 0.8, 0, 63, False)
    bins_ref = O.fq_bins(z, o_scale.item(), int(o_zp.item()), 0, 63)
    fq_ref = O.fq_per_tensor(z, o_scale.item(), int(o_zp.item()), 0, 63)
    codes, rowsum = ops.pack_weight(w.cuda(), w_scale.cuda(), w_zp.cuda(), w_qmin, w_qmax)
    got_fq, got_bins = ops.fused_fq_linear(a.cuda(), a_scale.reshape(1).cuda(), a_zp.reshape(1).to(torch.int32).cuda(), 0, 63, codes,
                                           w_scale.cuda(), rowsum, bias.cuda(),
                                           out_q=dict(scale=o_scale.reshape(1).cuda(), zp=o_zp.reshape(1).to(torch.int32).cuda(), qmin=0,
                                                      qmax=63, g=0.0, act="gelu", bins=True))
    d = (got_bins.cpu().to(torch.int32) - bins_ref.to(torch.int32)).abs()
    assert int(d.max()) <= 1 and float((d == 0).float().mean()) >= 0.999, (int(d.max()), float((d == 0).float().mean()))
    assert float((got_fq.cpu() - fq_ref).abs().max()) <= float(o_scale) * 1.0001
