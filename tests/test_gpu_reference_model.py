"""-m gpu: the reference's UNMODIFIED model code (quant_transformer/model/quant_bert.py, util_layernorm.py,
solver/gamma_migration.py, solver/token_wise_clipping.py -- imported from the staged oracle/_ref, see oracle/make_ref.py)
runs on this backend through ``install_as_reference_backend()`` on CUDA, and is compared with the reference running the
same schedule on its own quantization package on CPU, in the same process (tests/lockstep.py):

  (1) module level, teacher forced, BIT-EXACT: every activation-quantizer call our model makes is replayed -- same input
      tensor, same mask, same seq_pos, same flag schedule -- into the reference's own quantizer object of the same
      name; outputs, (scale, zero_point) and observer (min_val, max_val) must be identical after every forward.  Every
      QLinear / QEmbedding call is replayed the same way (QLinear: the documented 1e-3 tolerance, QEmbedding: exact).
  (2) end to end: the reference runs the whole schedule independently on CPU; logits and per-quantizer scales agree to
      the tolerance written at the assert (CPU and GPU GEMM / LayerNorm / softmax round differently, so activations
      drift by ~1e-6 relative and an occasional bin flips; bit-exactness across devices is not defined there).

Schedule = solver/ptq_glue_quant.py:228-253: delay_ln (gamma migration) -> weight calibration -> set_ratio + activation
calibration (token-wise clipping at one ratio) -> enable_quantization -> forward."""
import pytest

from oracle import make_ref, ref_model as RM
from tests import lockstep

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(make_ref.root() is None, reason="reference tree not staged (oracle/_ref)")]

CONFIGS = {
    # BASELINE config 2 (exp/bert_ptq/twc_fine_gamma): LSQ+ / AvgPruneMinMax 6-bit, gamma migration on
    "twc_fine_gamma_6bit": dict(a_bit=6, w_bit=6, a_quantizer="LSQPlusFakeQuantize", a_observer="AvgPruneMinMaxObserver", delay=True),
    # BASELINE config 1 (exp/bert_ptq/minmax/cola at 8 bit): Fixed / AvgMinMax, no migration
    "minmax_8bit": dict(a_bit=8, w_bit=8, a_quantizer="FixedFakeQuantize", a_observer="AvgMinMaxObserver", delay=False),
}


@pytest.mark.parametrize("cfg_name", list(CONFIGS))
def test_unmodified_quant_bert_runs_the_ptq_schedule_on_cuda(cfg_name, monkeypatch):
    from outlier_suppression_b200.quantization import quantized_module as qm
    # module-level teacher forcing needs every quantizer / QLinear call to happen: the fused output stage (which never
    # materialises the tensors in between) is checked separately below
    monkeypatch.setenv("OSQ_DISABLE_EPILOGUE_FUSION", "1")
    monkeypatch.setenv("OSQ_DISABLE_LN_FUSION", "1")
    monkeypatch.setenv("OSQ_DISABLE_ATTN_FUSION", "1")
    stats0 = dict(qm.stats)
    r = lockstep.run("b200", "cuda", RM.quant_config(**CONFIGS[cfg_name]))
    layers = r["layers"]
    # (1) happened inside lockstep.run(): make sure it actually covered the model
    assert r["n_act"] == 1 + 8 * layers - 1 + 2
    assert r["checked"]["q"] == r["n_act"] * r["n_forwards"]
    assert r["checked"]["op"] == (6 * layers + 2 + 3) * r["n_forwards"]
    # the Linear sites ran through the fused tcgen05 kernel in the quantized phase (3 batches x layers x {qkv, attn-out, up, down})
    assert qm.stats["fused"] - stats0["fused"] >= 3 * layers * 4, qm.stats
    assert qm.stats["grouped_launch"] - stats0["grouped_launch"] >= 3 * layers, qm.stats
    # (2) end to end against the independent CPU run of the reference
    assert r["scale_drift"] <= 2e-3, "per-quantizer scales drifted %.3g relative from the reference's own CPU run" % r["scale_drift"]
    worst = 0.0
    for got, want in zip(r["logits"], r["ref_logits"]):
        d = float((got - want).abs().max())
        worst = max(worst, d)
        assert d <= 2e-2 * float(want.abs().max()) + 1e-3, (d, got, want)
    print("%s: %d quantizer and %d operator calls replayed into reference modules (bit-exact / 1e-3); scale drift vs the "
          "independent CPU run %.2e; max |dlogit| %.2e" % (cfg_name, r["checked"]["q"], r["checked"]["op"], r["scale_drift"], worst))


@pytest.mark.parametrize("mode", ["two_launches", "epilogue_stage"])
@pytest.mark.parametrize("cfg_name", list(CONFIGS))
def test_fused_ffn_output_stage_inside_the_unmodified_model(cfg_name, mode, monkeypatch):
    """The state togglers wrap every dense -> GELU -> quantizer block of the reference's model (quant_bert.py:277-280) with the
    fused output stage.  Same model, same calibration, fusion on vs off on the same device: the logits must be bit-identical
    (the fused epilogue reproduces torch's CUDA GELU and K1 exactly), and the fused path must really have been taken."""
    import torch
    from outlier_suppression_b200.quantization import quantized_module as qm
    monkeypatch.setenv("OSQ_DISABLE_EPILOGUE_FUSION", "1")
    monkeypatch.setenv("OSQ_DISABLE_LN_FUSION", "1")   # the fused LayerNorm rounds differently from torch's: checked in its own test
    monkeypatch.setenv("OSQ_DISABLE_ATTN_FUSION", "1")  # likewise the integer-exact attention contractions
    r = lockstep.run("b200", "cuda", RM.quant_config(**CONFIGS[cfg_name]))
    model = r["model"]
    batches = [{k: v.cuda() for k, v in b.items()} for b in RM.synth_batches(3, 4, 32, 100, "cpu", seed=5)]

    def logits():
        with torch.no_grad():
            return [(lambda o: o[0] if isinstance(o, tuple) else o.logits)(model(**b)).clone() for b in batches]

    off = logits()
    for a, b in zip(off, r["logits"]):
        assert torch.equal(a.cpu(), b)
    monkeypatch.delenv("OSQ_DISABLE_EPILOGUE_FUSION")
    if mode == "epilogue_stage":   # GELU + quantizer inside the Linear's epilogue instead of one elementwise pass behind it
        monkeypatch.setenv("OSQ_EPILOGUE_STAGE", "1")
    before = qm.stats["epilogue_fused"]
    on = logits()
    assert qm.stats["epilogue_fused"] - before == 3 * r["layers"], qm.stats
    for a, b in zip(on, off):
        assert torch.equal(a, b), float((a - b).abs().max())


@pytest.mark.parametrize("cfg_name", list(CONFIGS))
def test_fused_residual_layernorm_quantizer_inside_the_unmodified_model(cfg_name, monkeypatch):
    """The togglers also wrap every dense -> dropout -> GammaResidual -> LayerNorm -> quantizer block (quant_bert.py:211-217,
    :296-303) with osq_residual_layernorm_fq_f32.  Its LayerNorm is plain fp32 with another summation order than torch's kernel,
    so activations differ by ~1e-6 relative and an occasional bin flips: the logits agree to the tolerance written below (the
    same one the CPU-vs-GPU comparison of the reference itself uses), and the fused path must really have been taken -- after
    gamma migration (config 2) through the split LayerNorm + gamma-scaled residual, without it through the affine LayerNorm."""
    import torch
    from outlier_suppression_b200.quantization import quantized_module as qm
    monkeypatch.setenv("OSQ_DISABLE_LN_FUSION", "1")
    monkeypatch.setenv("OSQ_DISABLE_ATTN_FUSION", "1")
    r = lockstep.run("b200", "cuda", RM.quant_config(**CONFIGS[cfg_name]))
    model = r["model"]
    batches = [{k: v.cuda() for k, v in b.items()} for b in RM.synth_batches(3, 4, 32, 100, "cpu", seed=5)]

    def logits():
        with torch.no_grad():
            return [(lambda o: o[0] if isinstance(o, tuple) else o.logits)(model(**b)).clone() for b in batches]

    off = logits()
    monkeypatch.delenv("OSQ_DISABLE_LN_FUSION")
    before = qm.stats.get("ln_fused", 0)
    on = logits()
    # two blocks per layer; the last layer's output LayerNorm has no quantizer (qoutput=False, quant_bert.py) and stays unfused
    assert qm.stats.get("ln_fused", 0) - before == 3 * (2 * r["layers"] - 1), qm.stats
    for a, b in zip(on, off):
        d = float((a - b).abs().max())
        assert d <= 2e-2 * float(b.abs().max()) + 1e-3, (d, a, b)


@pytest.mark.parametrize("cfg_name", list(CONFIGS))
def test_fused_attention_contractions_inside_the_unmodified_model(cfg_name, monkeypatch):
    """The togglers wrap every self-attention module (quant_bert.py:134-195): q @ k^T and probs @ v run as integer contractions
    with their quantizers as prologues (K8 / K9).  The integer contraction is exact where the reference's fp32 GEMM over the
    dequantised operands rounds, so scores differ by ~1e-6 relative and an occasional downstream bin flips: same logits
    tolerance as the CPU-vs-GPU comparison of the reference itself; the fused path must really have been taken, once per layer
    and forward."""
    import torch
    from outlier_suppression_b200.quantization import quantized_module as qm
    monkeypatch.setenv("OSQ_DISABLE_LN_FUSION", "1")
    monkeypatch.setenv("OSQ_DISABLE_ATTN_FUSION", "1")
    r = lockstep.run("b200", "cuda", RM.quant_config(**CONFIGS[cfg_name]))
    model = r["model"]
    batches = [{k: v.cuda() for k, v in b.items()} for b in RM.synth_batches(3, 4, 32, 100, "cpu", seed=5)]

    def logits():
        with torch.no_grad():
            return [(lambda o: o[0] if isinstance(o, tuple) else o.logits)(model(**b)).clone() for b in batches]

    off = logits()
    monkeypatch.delenv("OSQ_DISABLE_ATTN_FUSION")
    before = qm.stats.get("attn_fused", 0)
    on = logits()
    assert qm.stats.get("attn_fused", 0) - before == 3 * r["layers"], qm.stats
    for a, b in zip(on, off):
        d = float((a - b).abs().max())
        assert d <= 2e-2 * float(b.abs().max()) + 1e-3, (d, a, b)
