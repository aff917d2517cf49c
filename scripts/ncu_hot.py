#!/usr/bin/env python
"""Rank the source lines of a kernel by warp-stall samples from an `ncu --set full --import-source on` report.

usage: python scripts/ncu_hot.py gpurun_out/prof_fft_analysis.ncu-rep [top_n] [kernel-name substring]
Prints, per source line: samples, share, instructions executed, dominant stall reasons, shared-memory excess wavefronts.
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    want = sys.argv[3] if len(sys.argv) > 3 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, cur_file, lines, cur_fn, seen = None, None, [], "", set()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            cur_fn = r[1]
            if cur_fn not in seen and (want is None or want in cur_fn):
                seen.add(cur_fn)
                print("kernel:", cur_fn[:160])
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[2] != "-":   # keep the per-source-line aggregate rows only (Address == "-")
            continue
        if want is not None and want not in cur_fn:
            continue
        d = dict(zip(hdr, r))
        # the header has two "Source" columns; positional access for the first two
        lines.append((cur_file, r[0], r[1], d))
    def num(d, k):
        try:
            return float(d.get(k, "0").replace(",", ""))
        except ValueError:
            return 0.0
    total = sum(num(d, "# Samples") for _, _, _, d in lines) or 1.0
    tot_inst = sum(num(d, "Instructions Executed") for _, _, _, d in lines) or 1.0
    stall_keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    agg = {k: sum(num(d, k) for _, _, _, d in lines) for k in stall_keys}
    print(f"samples {total:.0f}  warp-instructions {tot_inst:.0f}")
    print("stall mix:", ", ".join(f"{k[6:]} {100 * v / total:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * total))
    exc = sum(num(d, "L1 Wavefronts Shared Excessive") for _, _, _, d in lines)
    wf = sum(num(d, "L1 Wavefronts Shared") for _, _, _, d in lines)
    print(f"shared wavefronts {wf:.0f} (excess {exc:.0f} = {100 * exc / max(wf, 1):.1f}%)")
    lines.sort(key=lambda x: -num(x[3], "# Samples"))
    print(f"{'file:line':<26}{'smp%':>6}{'inst%':>7}  {'stalls':<46} source")
    for f, ln, src, d in lines[:top]:
        s = num(d, "# Samples")
        st = sorted(((num(d, k), k[6:]) for k in stall_keys), reverse=True)[:3]
        sts = " ".join(f"{n}:{100 * v / max(s, 1):.0f}" for v, n in st if v > 0)
        print(f"{f + ':' + ln:<26}{100 * s / total:>6.1f}{100 * num(d, 'Instructions Executed') / tot_inst:>7.1f}  {sts:<46} {src.strip()[:90]}")


if __name__ == "__main__":
    main()
