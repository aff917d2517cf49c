#!/usr/bin/env python
"""Turn the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py r01        # writes profiles/r01_launches.csv, profiles/r01_kernels.csv, profiles/r01_summary.md
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
]


def raw_rows(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    return rows if len(rows) > 2 else []


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    md = [f"# ncu summary {tag}", "", "Source: `scripts/gpu_profile.sh` (ncu under gpurun on one B200; `--clock-control none`).",
          "Per-launch times below are cold-cache and serialised: compare SHARES with bench.py's CUDA-event numbers, not absolutes.", ""]
    # ---- launch list
    lp = os.path.join(OUT, "launches.csv")
    if os.path.exists(lp):
        rows = [r for r in csv.reader(open(lp)) if len(r) > 5]
        hdr = rows[0]
        ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        t, n = defaultdict(float), defaultdict(int)
        for r in rows[1:]:
            try:
                v = float(r[vi].replace(",", ""))
            except ValueError:
                continue
            t[r[ki]] += v
            n[r[ki]] += 1
        tot = sum(t.values())
        with open(os.path.join(PROF, f"{tag}_launches.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["kernel", "launches", "total_us", "share_pct", "avg_us"])
            for k, v in sorted(t.items(), key=lambda kv: -kv[1]):
                w.writerow([k, n[k], round(v / 1e3, 2), round(100 * v / tot, 2), round(v / 1e3 / n[k], 2)])
        md += ["## Launch list (`ncu --metrics gpu__time_duration.sum`, `python bench.py --steps 2 --warmup 1 --no-cpu --no-stages`)", "",
               "| kernel | launches | total us | share | avg us |", "|---|---|---|---|---|"]
        for k, v in sorted(t.items(), key=lambda kv: -kv[1])[:14]:
            md.append(f"| `{k[:90]}` | {n[k]} | {v / 1e3:.1f} | {100 * v / tot:.1f}% | {v / 1e3 / n[k]:.1f} |")
        md.append("")
    # ---- full captures
    allrows = []
    for rep in sorted(f for f in os.listdir(OUT) if f.endswith(".ncu-rep")):
        rows = raw_rows(os.path.join(OUT, rep))
        if not rows:
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = {"capture": rep, "kernel": r[hdr.index("Kernel Name")]}
            for m in METRICS:
                if m in hdr:
                    d[m] = r[hdr.index(m)] + " " + units[hdr.index(m)]
            st = [(hdr[i].replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[i].replace(",", "") or 0))
                  for i in range(len(hdr)) if "smsp__average_warps_issue_stalled" in hdr[i] and "per_issue_active" in hdr[i]]
            d["top_stalls"] = ", ".join(f"{k}={v:.2f}" for k, v in sorted(st, key=lambda kv: -kv[1])[:4])
            allrows.append(d)
    if allrows:
        keys = ["capture", "kernel"] + METRICS + ["top_stalls"]
        with open(os.path.join(PROF, f"{tag}_kernels.csv"), "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=keys)
            w.writeheader()
            for d in allrows:
                w.writerow(d)
        md += ["## `ncu --set full` captures (one row per captured launch)", "",
               "| kernel | time | DRAM read | DRAM write | DRAM % | tensor % | warps active % | issue % | regs | top stalls |", "|---|---|---|---|---|---|---|---|---|---|"]
        for d in allrows:
            g = lambda m: d.get(m, "-").split(" ")[0][:10]
            md.append(f"| `{d['kernel'][:70]}` | {g('gpu__time_duration.sum')} us | {g('dram__bytes_read.sum')} MB | {g('dram__bytes_write.sum')} MB | "
                      f"{g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} | {g('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active')} | "
                      f"{g('sm__warps_active.avg.pct_of_peak_sustained_active')} | {g('smsp__issue_active.avg.pct_of_peak_sustained_active')} | {g('launch__registers_per_thread')} | {d['top_stalls']} |")
        md.append("")
    # DRAM traffic per launch of each captured kernel family -> profiles/traffic.json (bench.py fills roofline.traffic from it)
    if allrows:
        traffic = {}
        for d in allrows:
            try:
                rd = float(d.get("dram__bytes_read.sum", "0 x").split(" ")[0].replace(",", ""))
                wr = float(d.get("dram__bytes_write.sum", "0 x").split(" ")[0].replace(",", ""))
                unit = d.get("dram__bytes_read.sum", "0 Mbyte").split(" ")[1]
            except (ValueError, IndexError):
                continue
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1e6)
            name = d["kernel"].split("(")[0].replace("void ", "").strip()
            traffic.setdefault(name, []).append((rd + wr) * mult)
        out = {k: {"dram_bytes_per_launch": sum(v) / len(v), "launches_captured": len(v), "capture": tag} for k, v in traffic.items()}
        with open(os.path.join(PROF, "traffic.json"), "w") as f:
            json.dump(out, f, indent=1)

    bj = os.path.join(OUT, "bench.json")
    if os.path.exists(bj):
        try:
            b = json.load(open(bj))
            md += ["## bench.py line of the same build (CUDA events, not under ncu)", "", "```json", json.dumps(b, indent=1)[:6000], "```", ""]
        except Exception:
            pass
    with open(os.path.join(PROF, f"{tag}_summary.md"), "w") as f:
        f.write("\n".join(md))
    print("wrote", os.path.join(PROF, f"{tag}_summary.md"))


if __name__ == "__main__":
    main()
