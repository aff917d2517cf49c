#!/usr/bin/env python
"""Summarise gpurun_out/block_<tag>.ncu-rep (scripts/gpu_prof.sh: ncu --set full of every kernel of one fwd+bwd step of the headline block)
into the tracked files  profiles/<tag>_kernels.csv,  profiles/<tag>_summary.md  and  profiles/traffic.json  (DRAM bytes per launch, per workload).

    python scripts/summarize_block_profile.py r02 [workload]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
workload = sys.argv[2] if len(sys.argv) > 2 else "sfno_block_721x1440x73"
rep = os.path.join(ROOT, "gpurun_out", f"block_{tag}.ncu-rep")
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
           "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
STALLS = ["long_scoreboard", "wait", "not_selected", "math_pipe_throttle", "barrier", "short_scoreboard", "mio_throttle", "sleeping", "lg_throttle", "dispatch_stall"]


def short(name):
    name = name.replace("void ", "").replace("b200sht::", "")
    return name.split("(")[0]


r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
rows = list(csv.reader(io.StringIO(r.stdout)))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
out = []
for row in rows[2:]:
    d = {"kernel": short(row[ix["Kernel Name"]])}
    for m in METRICS:
        if m in ix:
            try:
                d[m] = float(row[ix[m]].replace(",", ""))
            except ValueError:
                d[m] = row[ix[m]]
    for s in STALLS:
        k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
        if k in ix:
            d["stall_" + s] = float(row[ix[k]])
    out.append(d)
cols = ["kernel"] + [m for m in METRICS if m in ix] + ["stall_" + s for s in STALLS]
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", f"{tag}_kernels.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(cols)
    for d in out:
        w.writerow([d.get(c, "") for c in cols])

peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
tot = sum(d["gpu__time_duration.sum"] for d in out)
md = [f"# ncu summary {tag}: one forward + backward step of `{workload}` (bf16 activations, TF32 contractions)", "",
      f"Source: `scripts/gpu_prof.sh` -> `gpurun_out/block_{tag}.ncu-rep` (`ncu --set full --import-source on --clock-control none`, third step of "
      "`scripts/prof_block.py`, one B200 under gpurun).  Times under ncu are cold-cache and serialised: compare SHARES with bench.py's CUDA-event numbers.", "",
      "| kernel | us | share | DRAM MB (r+w) | GB/s | frac of measured HBM | warp inst (M) | issue active % | tensor pipe % | fma pipe % | regs | top stalls (cycles per issued instruction) |",
      "|---|---|---|---|---|---|---|---|---|---|---|---|"]
traffic = {}
for d in out:
    us = d["gpu__time_duration.sum"]
    mb = d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
    gbs = mb * 1e6 / (us * 1e-6) / 1e9 if us else 0.0
    st = sorted(((s, d.get("stall_" + s, 0.0)) for s in STALLS), key=lambda kv: -kv[1])[:3]
    md.append(f"| `{d['kernel']}` | {us:.1f} | {100 * us / tot:.1f}% | {mb:.1f} | {gbs:.0f} | {gbs / peaks['hbm_gbs']:.2f} | {d['smsp__inst_executed.sum'] / 1e6:.2f} | "
              f"{d['smsp__issue_active.avg.pct_of_peak_sustained_active']:.1f} | {d.get('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 0):.1f} | "
              f"{d.get('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 0):.1f} | {int(d['launch__registers_per_thread'])} | "
              + ", ".join(f"{s} {v:.2f}" for s, v in st) + " |")
    rec = traffic.setdefault(d["kernel"], {"dram_bytes_per_launch": 0.0, "launches_captured": 0, "capture": tag, "workload": workload})
    rec["dram_bytes_per_launch"] += mb * 1e6
    rec["launches_captured"] += 1
for rec in traffic.values():
    rec["dram_bytes_per_launch"] = round(rec["dram_bytes_per_launch"] / rec["launches_captured"])
md += ["", f"Sum of kernel times under ncu: {tot:.1f} us."]
with open(os.path.join(ROOT, "profiles", f"{tag}_summary.md"), "w") as f:
    f.write("\n".join(md) + "\n")
with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1, sort_keys=True)
print("\n".join(md))
