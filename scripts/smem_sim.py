#!/usr/bin/env python
"""Shared-memory bank-conflict model of the compile-time FFT plans (8-byte accesses, half-warp phases of 16 lanes, 16 8-byte bank
pairs): prints wavefronts / ideal for every access pattern of a plan.  usage: python scripts/smem_sim.py H R0 R1 R2 TPG"""
import sys
from collections import Counter


def skew(i, R0):
    return i + i // R0


def wavefronts(addrs):
    """addrs: per-lane 8-byte-element indices (None = inactive) of one warp instruction"""
    wf = ideal = 0
    for h in range(0, 32, 16):
        lanes = [a for a in addrs[h:h + 16] if a is not None]
        if not lanes:
            continue
        uniq = set(lanes)
        c = Counter(a % 16 for a in uniq)
        wf += max(c.values())
        ideal += (len(uniq) + 15) // 16
    return wf, ideal


def stage(H, R, Ns, R0, TPG, name):
    NB = H // R
    tot = {"ld": [0, 0], "st": [0, 0]}
    for warp0 in range(0, TPG, 32):
        for r in range(R):
            ld, st = [], []
            for lane in range(32):
                j = warp0 + lane
                if j >= NB or j >= TPG:
                    ld.append(None); st.append(None); continue
                k = j % Ns
                j0 = (j - k) * R + k
                ld.append(skew(j + r * NB, R0))
                st.append(skew(j0 + r * Ns, R0))
            for key, a in (("ld", ld), ("st", st)):
                w, i = wavefronts(a)
                tot[key][0] += w; tot[key][1] += i
    print(f"{name:<28} loads {tot['ld'][0]}/{tot['ld'][1]}  stores {tot['st'][0]}/{tot['st'][1]}")


def main():
    H, R0, R1, R2, TPG = (int(x) for x in sys.argv[1:6])
    print(f"H={H} plan {R0}x{R1}x{R2} TPG={TPG}")
    stage(H, R0, 1, R0, TPG, "stage0 (store only matters)")
    stage(H, R1, R0, R0, TPG, "stage1")
    if R2 > 1:
        stage(H, R2, R0 * R1, R0, TPG, "stage2")
    # split pass of the analysis: quarter warp = 8 consecutive m of one quad; reads index skew(m) and skew(H-m) of pair rows
    for QUADS, BS in ((2, (skew(H, R0) + 2) | 1),):
        tot = [0, 0]
        for m0 in range(0, 241, 16):
            for which in (0, 1):
                a = []
                for lane in range(32):
                    qd = (lane // 8) % QUADS
                    m = m0 + (lane % 8) + 8 * (lane // (8 * QUADS))
                    idx = m if which == 0 else (H - m if m else 0)
                    a.append((qd * 2) * BS + skew(idx, R0))
                w, i = wavefronts(a)
                tot[0] += w; tot[1] += i
        print(f"{'split reads':<28} {tot[0]}/{tot[1]}")
    # spectrum build of the synthesis: stores at skew(q) and skew(H-q); quarter warp = 8 consecutive q of one quad
    tot = [0, 0]
    BS = (skew(H, R0) + 2) | 1
    for q0 in range(0, H // 2 + 1, 16):
        for which in (0, 1):
            a = []
            for lane in range(32):
                qd = (lane // 8) % 2
                q = q0 + (lane % 8) + 8 * (lane // 16)
                if q > H // 2:
                    a.append(None); continue
                idx = q if which == 0 else H - q
                a.append((qd * 2) * BS + skew(idx % H if which else idx, R0))
            w, i = wavefronts(a)
            tot[0] += w; tot[1] += i
    print(f"{'build stores':<28} {tot[0]}/{tot[1]}")


if __name__ == "__main__":
    main()
