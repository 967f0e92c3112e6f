"""Shared harness of the model-level drop-in tests: runs the PTQ schedule of solver/ptq_glue_quant.py:228-253 on the
reference's UNMODIFIED model code bound to backend ``ours`` and, in lockstep, replays every activation-quantizer /
QLinear / QEmbedding call it makes into the reference's own modules (CPU) with the very same inputs."""
import copy

import numpy as np
import torch

from oracle import ref_model as RM, ref_shim


def named(model, cls):
    return {n: m for n, m in model.named_modules() if isinstance(m, cls)}


def _cpu(v):
    return v.detach().cpu() if torch.is_tensor(v) else v


def run(ours_backend: str, device: str, qcfg, layers=2, hidden=128, heads=2, inter=512, batch=4, seq=32, n_batches=3):
    """Returns a dict of what was checked + both sets of logits.  ``ours_backend`` is "b200" (CUDA) or "reference"
    (harness self-test on CPU)."""
    fp = RM.fp_bert(layers=layers, hidden=hidden, heads=heads, inter=inter, vocab=100, max_pos=max(64, seq))
    cpu_batches = RM.synth_batches(n_batches, batch, seq, 100, "cpu", seed=5)
    dev_batches = [{k: v.to(device) for k, v in b.items()} for b in cpu_batches]

    # ---------------- the reference, on its own quantization package, CPU ----------------
    ref = RM.load_stack("reference")
    R = ref.quantization
    RQ = R.fake_quant.QuantizeBase
    with ref_shim.cpu_only():
        ref_model, ref_logits = RM.run_schedule(ref, RM.build_model(ref, copy.deepcopy(fp), qcfg, "cpu"), qcfg, cpu_batches)
        replay_model = RM.build_model(ref, copy.deepcopy(fp), qcfg, "cpu")          # teacher-forced twin

    # ---------------- the same model file bound to `ours_backend` ----------------
    ours = RM.load_stack(ours_backend) if ours_backend != "reference" else ref
    Q = ours.quantization
    OQ = Q.fake_quant.QuantizeBase
    qm = Q.quantized_module
    model = RM.build_model(ours, copy.deepcopy(fp), qcfg, device)
    calls = []

    def rec_q(name):
        def hook(mod, args, kwargs, out):
            a = list(args) + [None] * (3 - len(args))
            calls.append(("q", name, _cpu(a[0]).clone(), _cpu(kwargs.get("observation_mask", a[1])),
                          kwargs.get("seq_pos", a[2] if a[2] is not None else -1), _cpu(out).clone()))
        return hook

    def rec_op(name):
        def hook(mod, args, out):
            calls.append(("op", name, _cpu(args[0]).clone(), None, None, _cpu(out).clone()))
        return hook

    checked = {"q": 0, "op": 0, "state": 0}

    def replay():
        rq = named(replay_model, RQ)
        rops = dict(replay_model.named_modules())
        oq = named(model, OQ)
        with ref_shim.cpu_only(), torch.no_grad():
            for kind, name, x, mask, seq_pos, out in calls:
                if kind == "q":
                    y = rq[name](x, mask, seq_pos)
                    np.testing.assert_array_equal(y.numpy(), out.numpy(), err_msg=name)
                    checked["q"] += 1
                else:
                    y = rops[name](x)
                    if isinstance(rops[name], torch.nn.Embedding):
                        np.testing.assert_array_equal(y.numpy(), out.numpy(), err_msg=name)
                    else:  # QLinear: |dY| <= 1e-3 |Y| + 1e-3 max|Y|  (BASELINE north_star: dequantized floats within 1e-3)
                        d = (y.double() - out.double()).abs()
                        tol = 1e-3 * y.double().abs() + 1e-3 * float(y.abs().max())
                        assert bool((d <= tol).all()), (name, float(d.max()))
                    checked["op"] += 1
            for name, q in oq.items():
                r = rq[name]
                for a, b in ((q.scale, r.scale), (q.zero_point, r.zero_point), (q.observer.min_val, r.observer.min_val),
                             (q.observer.max_val, r.observer.max_val)):
                    np.testing.assert_array_equal(a.detach().cpu().float().numpy().reshape(-1),
                                                  b.detach().cpu().float().numpy().reshape(-1), err_msg=name)
                checked["state"] += 1
        calls.clear()

    mcfg = RM.Cfg(model_type="bert")
    if qcfg.ln.delay:
        if device == "cpu":
            with ref_shim.cpu_only():
                model = ours.gamma_migration.delay_ln(model, qcfg, mcfg)
        else:
            model = ours.gamma_migration.delay_ln(model, qcfg, mcfg)
        with ref_shim.cpu_only():
            replay_model = ref.gamma_migration.delay_ln(replay_model, qcfg, mcfg)
        assert any(type(m).__name__ == "QuantizedSplitLayerNorm" for m in model.modules())
    hooks = []
    for n, q in named(model, OQ).items():
        if "act_fake_quant" in n:
            hooks.append(q.register_forward_hook(rec_q(n), with_kwargs=True))
    for n, op in model.named_modules():
        if isinstance(op, (qm.QLinear, qm.QEmbedding)):
            hooks.append(op.register_forward_hook(rec_op(n)))

    def both(fn_ours, fn_ref):
        if device == "cpu":
            with ref_shim.cpu_only():
                fn_ours()
        else:
            fn_ours()
        with ref_shim.cpu_only():
            fn_ref()

    def forward_all(batches):
        outs = []
        for b in batches:
            with torch.no_grad():
                if device == "cpu":
                    with ref_shim.cpu_only():
                        o = model(**b)
                else:
                    o = model(**b)
            outs.append((o[0] if isinstance(o, tuple) else o.logits).detach().cpu())
            replay()
        return outs

    both(lambda: Q.enable_calibration_woquantization(model, quantizer_type="weight_fake_quant"),
         lambda: R.enable_calibration_woquantization(replay_model, quantizer_type="weight_fake_quant"))
    forward_all(dev_batches[:1])
    if "PruneMinMaxObserver" in qcfg.a_qconfig.observer:
        both(lambda: (Q.disable_all(model), Q.state.set_observer_name(model), ours.token_wise_clipping.set_ratio(model, 0.99)),
             lambda: (R.disable_all(replay_model), R.state.set_observer_name(replay_model),
                      ref.token_wise_clipping.set_ratio(replay_model, 0.99)))
    else:
        both(lambda: Q.enable_calibration_woquantization(model, quantizer_type="act_fake_quant"),
             lambda: R.enable_calibration_woquantization(replay_model, quantizer_type="act_fake_quant"))
    forward_all(dev_batches)
    both(lambda: Q.enable_quantization(model), lambda: R.enable_quantization(replay_model))
    logits = forward_all(dev_batches)
    for h in hooks:
        h.remove()

    n_act = sum("act_fake_quant" in n for n in named(model, OQ))
    worst = 0.0
    ref_q = named(ref_model, RQ)
    for name, q in named(model, OQ).items():
        s_ref = ref_q[name].scale.detach().reshape(-1).float()
        s = q.scale.detach().cpu().reshape(-1).float()
        worst = max(worst, float(((s - s_ref).abs() / s_ref.abs()).max()))
    return {"checked": checked, "n_act": n_act, "n_forwards": 1 + 2 * n_batches, "logits": logits, "ref_logits": ref_logits,
            "scale_drift": worst, "model": model, "layers": layers}
