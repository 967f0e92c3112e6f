"""Turns the raw ncu artefacts of a gpurun call (gpurun_out/) into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py r01

  profiles/<round>_launches.md     per-kernel launch count / total device time / share (ncu --metrics gpu__time_duration.sum)
  profiles/<round>_fused_full.md   key metrics of the `ncu --set full` capture of the fused kernel, per launch
  profiles/<round>_traffic.json    DRAM bytes per launch per site shape (read by bench.py for roofline.traffic)
  profiles/<round>_bench.json      the bench line of the same call
"""
import collections, csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(PROF, exist_ok=True)

# ---- launch list
rows = [r for r in csv.reader(open(os.path.join(OUT, "launches.csv"))) if len(r) > 10]
hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
first_fused = next(i for i, r in enumerate(rows[1:]) if r[ci["Kernel Name"]].startswith("fused_fq_linear"))
agg = {"setup": collections.OrderedDict(), "step": collections.OrderedDict()}
for i, r in enumerate(rows[1:]):
    try:
        v = float(r[ci["Metric Value"]])
    except ValueError:
        continue
    name = r[ci["Kernel Name"]].split("(")[0]
    key = (name, r[ci["Grid Size"]], r[ci["Block Size"]])
    a = agg["setup" if i < first_fused else "step"].setdefault(key, [0, 0.0, 1e30, 0.0])
    a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
with open(os.path.join(PROF, tag + "_launches.md"), "w") as f:
    f.write("# %s: ncu launch list of `bench.py --steps 2 --warmup 3 --only-value --no-graph`\n\n" % tag)
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fused_fq|pack_weight|minmax|prune_select|fq_per`.\n"
            "Per-launch times are cold-cache and serialised by ncu: compare SHARES, not absolutes.\n")
    for part, title in (("step", "Warm-up + timed steps (5 steps x 48 launches): the region `value` is measured on"),
                        ("setup", "Setup before the first step (untimed): per-channel weight qparams + s8 packing, one calibration batch per activation quantizer")):
        tot = sum(a[1] for a in agg[part].values()) or 1.0
        f.write("\n## %s\n\n| kernel | grid | block | launches | total us | min us | max us | share |\n|---|---|---|---|---|---|---|---|\n" % title)
        for (name, grid, block), a in sorted(agg[part].items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %s | %s | %d | %.1f | %.1f | %.1f | %.1f %% |\n" % (name, grid, block, a[0], a[1] / 1e3, a[2] / 1e3, a[3] / 1e3, 100 * a[1] / tot))
print(open(os.path.join(PROF, tag + "_launches.md")).read())

# ---- full capture
rep = os.path.join(OUT, "prof_fused.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]; units = rows[1]; body = rows[2:]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]
idx = {k: hdr.index(k) for k in keys if k in hdr}
traffic = {}
with open(os.path.join(PROF, tag + "_fused_full.md"), "w") as f:
    f.write("# %s: `ncu --set full --clock-control none --import-source on -k regex:fused_fq_linear` (the 4 launches of one encoder layer)\n\n" % tag)
    f.write("| metric | unit | " + " | ".join("launch %d" % i for i in range(len(body))) + " |\n|---|---|" + "---|" * len(body) + "\n")
    for k, i in idx.items():
        f.write("| %s | %s | %s |\n" % (k, units[i], " | ".join(r[i] for r in body)))
    f.write("\nLaunch order inside a layer: qkv (768->2304, one launch for the three projections), attn_out (768->768), ffn_up (768->3072), ffn_down (3072->768); the capture starts at a layer boundary.\n")
names = ["qkv", "attn_out", "ffn_up", "ffn_down"]
for j, r in enumerate(body[:4]):
    rd, wr = float(r[idx["dram__bytes_read.sum"]]), float(r[idx["dram__bytes_write.sum"]])
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    traffic[names[j]] = rd * scale[units[idx["dram__bytes_read.sum"]]] + wr * scale[units[idx["dram__bytes_write.sum"]]]
json.dump(traffic, open(os.path.join(PROF, tag + "_traffic.json"), "w"), indent=1)
print(open(os.path.join(PROF, tag + "_fused_full.md")).read())
if os.path.exists(os.path.join(OUT, "bench.json")):
    open(os.path.join(PROF, tag + "_bench.json"), "w").write(open(os.path.join(OUT, "bench.json")).read())

# ---- K7 / K8 / K9 full capture (scripts/profile_f3_kernels.py)
rep3 = os.path.join(OUT, "prof_f3.ncu-rep")
if os.path.exists(rep3):
    raw = subprocess.run(["ncu", "-i", rep3, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]; units = rows[1]; body = rows[2:]
    keys3 = keys + ["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
                    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
                    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    idx3 = {k: hdr.index(k) for k in keys3 if k in hdr}
    ni = hdr.index("Kernel Name")
    # the second launch of each kernel (the first is the warm-up)
    seen, pick = {}, []
    for r in body:
        n = r[ni].split("(")[0]
        seen[n] = seen.get(n, 0) + 1
        if seen[n] == 2:
            pick.append(r)
    with open(os.path.join(PROF, tag + "_f3_full.md"), "w") as f:
        f.write("# %s: `ncu --set full --clock-control none --import-source on` of K7 / K8 / K9 at config 2's size (batch 32, seq 512, 12 heads x 64)\n\n" % tag)
        f.write("Second launch of each kernel (`scripts/profile_f3_kernels.py`).  Algorithmic bytes: K7 163.6 MB, K8 503.3 MB, K9 515.9 MB.\n\n")
        f.write("| metric | unit | " + " | ".join(r[ni].split("(")[0].replace("void ", "").replace("osq::", "") for r in pick) + " |\n|---|---|" + "---|" * len(pick) + "\n")
        for k, i in idx3.items():
            f.write("| %s | %s | %s |\n" % (k, units[i], " | ".join(r[i] for r in pick)))
    print(open(os.path.join(PROF, tag + "_f3_full.md")).read())
