#!/usr/bin/env python
"""Summarise one gpu_round.sh output set (gpurun_out/<tag>_*) into profiles/<tag>_*.{md,csv}.

  python scripts/summarize_profile.py <tag>

* <tag>_launches.csv  (ncu --metrics gpu__time_duration.sum launch list)  -> per-kernel totals / shares
* <tag>_prof.ncu-rep  (ncu --set full of the selected kernels)            -> key counters per captured launch
* <tag>_bench.json                                                        -> copied verbatim
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
NSTEPS_IN_RUN = int(sys.argv[2]) if len(sys.argv) > 2 else 4     # steps the ncu'd command ran (3 warm-up + timed)
src = os.path.join(ROOT, "gpurun_out")
dst = os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)
out = [f"# Profile summary `{tag}`", ""]

bench = os.path.join(src, f"{tag}_bench.json")
if os.path.exists(bench) and os.path.getsize(bench):
    shutil.copy(bench, os.path.join(dst, f"{tag}_bench.json"))
    d = json.loads(open(bench).read().strip().splitlines()[-1])
    out += ["## bench line (python bench.py --steps 5 --warmup 3, no profiler)", "",
            f"* value {d['value']:.2f} {d['unit']}  ({d['ms_per_step']:.1f} ms/step), e2e {d['e2e']['value']:.2f}",
            f"* roofline: {d['roofline']['achieved']:.1f} {d['roofline']['unit']} = {d['roofline']['frac']:.3f} of "
            f"{d['roofline']['peak']} ({d['roofline'].get('peak_source')}), conv share of step "
            f"{d['roofline'].get('share_of_step', 0):.3f}",
            f"* clocks {d.get('clocks')}, gpu_launches {d.get('gpu_launches')}", ""]

launches = os.path.join(src, f"{tag}_launches.csv")
if os.path.exists(launches):
    rows = list(csv.reader(open(launches)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    out += ["## ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache serialised: compare SHARES)",
            "", f"total {tot:.1f} ms over {sum(v[0] for v in agg.values())} launches "
            f"(`python bench.py --steps 1 --warmup 3 --value-only`: {NSTEPS_IN_RUN} steps in the run)", "",
            "| kernel | launches | ms | share |", "|---|---:|---:|---:|"]
    with open(os.path.join(dst, f"{tag}_launch_shares.csv"), "w") as f:
        f.write("kernel,launches,ms,share\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{v[0]},{v[1]:.3f},{v[1] / tot:.4f}\n")
            if v[1] / tot >= 0.002:
                out.append(f"| `{k}` | {v[0]} | {v[1]:.2f} | {100 * v[1] / tot:.1f}% |")
    out.append("")

rep = os.path.join(src, f"{tag}_prof.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    out += ["## ncu --set full (per captured launch)", "",
            "| " + " | ".join(f"{w} [{units[i]}]" if units[i] else w for w, i in idx) + " |",
            "|" + "---|" * len(idx)]
    with open(os.path.join(dst, f"{tag}_ncu_full_raw.csv"), "w") as f:
        f.write(raw)
    for r in rows[2:]:
        out.append("| " + " | ".join(r[i][:70] for _, i in idx) + " |")
    out.append("")

traffic = os.path.join(src, f"{tag}_traffic.csv")
if os.path.exists(traffic):
    rows = list(csv.reader(open(traffic)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ii, ki, mi, vi, ui = (hdr.index(c) for c in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    per = collections.OrderedDict()              # launch id -> [kernel, bytes]
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1.0)
        e = per.setdefault(r[ii], [r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", ""), 0.0])
        e[1] += v
    ids = list(per)
    nstep = len(ids) // NSTEPS_IN_RUN             # identical launch sequences per step; the last one is summarised
    last = [per[i] for i in ids[len(ids) - nstep:]]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, b in last:
        agg[k][0] += 1
        agg[k][1] += b
    tot_b = sum(v[1] for v in agg.values())
    out += ["## DRAM traffic of the conv family, one training step (ncu dram__bytes_read.sum + dram__bytes_write.sum)", "",
            f"{nstep} launches, {tot_b / 1e9:.2f} GB per step = {tot_b / max(nstep, 1) / 1e6:.1f} MB per launch", "",
            "| kernel | launches | GB | MB / launch |", "|---|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {v[0]} | {v[1] / 1e9:.2f} | {v[1] / v[0] / 1e6:.1f} |")
    out.append("")
    workload = None
    if os.path.exists(bench) and os.path.getsize(bench):
        workload = json.loads(open(bench).read().strip().splitlines()[-1])["config"]["workload"]
    json.dump({"dram_bytes_per_launch": tot_b / max(nstep, 1), "dram_bytes_per_step": tot_b, "launches_per_step": nstep,
               "workload": workload,
               "source": f"profiles/{tag}_summary.md (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, "
                         "conv-family kernels of one training step)"},
              open(os.path.join(dst, "conv_traffic.json"), "w"), indent=1)

open(os.path.join(dst, f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
