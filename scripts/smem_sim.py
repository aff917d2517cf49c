#!/usr/bin/env python
"""Shared-memory wavefront model of the compile-time FFT plans in makani_b200/csrc/fft.cu.

All exchange-buffer accesses are 8-byte accesses issued by half warps: 16 lanes are conflict-free iff their slots are distinct
modulo 16 (32 banks x 4 B).  The script replays every access of the analysis kernel of a plan (stage stores/loads, split loads) and
of the synthesis kernel (spectrum-build stores, stage loads/stores, fused last-stage loads) under the two layouts

    LaySkew : slot = i + i/16                      (buffer written by the first stage)
    LayBlock: slot = i + PAD * (i / (R0*R1))       (buffer written by the second stage)
    LayId   : slot = i                             (last analysis stage -> split pass)

and prints wavefronts / ideal per access class.   usage: python scripts/smem_sim.py H R0 R1 R2 TPG [mmax]
"""
import sys
from collections import Counter


def block_pad(R0, R1):
    if R0 >= 16:
        return 0
    p = 0
    while (R0 * R1 + p) % 16 != R0 % 16:
        p += 1
    return p


def wavefronts(slots):
    """slots: per-lane slot (None = inactive lane) of one warp instruction -> (wavefronts, ideal)"""
    wf = ideal = 0
    for h in range(0, 32, 16):
        lanes = [a for a in slots[h:h + 16] if a is not None]
        if not lanes:
            continue
        uniq = set(lanes)
        c = Counter(a % 16 for a in uniq)
        wf += max(c.values())
        ideal += (len(uniq) + 15) // 16
    return wf, ideal


class Tally:
    def __init__(self):
        self.rows = []

    def add(self, name, instrs):
        w = i = 0
        for slots in instrs:
            a, b = wavefronts(slots)
            w += a
            i += b
        self.rows.append((name, w, i))

    def total(self):
        return sum(w for _, w, _ in self.rows), sum(i for _, _, i in self.rows)

    def report(self, title):
        print(title)
        tw = ti = 0
        for name, w, i in self.rows:
            print(f"  {name:<34}{w:>6} / {i:<6} {w / max(i, 1):.2f}x")
            tw += w
            ti += i
        print(f"  {'total':<34}{tw:>6} / {ti:<6} {tw / max(ti, 1):.2f}x")
        return tw, ti


def stage_accesses(H, R, Ns, TPG, lay_in, lay_out):
    NB = H // R
    loads, stores = [], []
    for j0w in range(0, NB, TPG):          # iterations of the j loop
        for w0 in range(0, TPG, 32):       # warps of the group
            for r in range(R):
                ld, st = [], []
                for lane in range(32):
                    t = w0 + lane
                    j = j0w + t
                    if t >= TPG or j >= NB:
                        ld.append(None)
                        st.append(None)
                        continue
                    k = j % Ns
                    j0 = (j - k) * R + k
                    ld.append(lay_in(j + r * NB) if lay_in else None)
                    st.append(lay_out(j0 + r * Ns) if lay_out else None)
                loads.append(ld)
                stores.append(st)
    return loads, stores


def simulate(H, R0, R1, R2, TPG, mmax, verbose=True):
    """-> {"analysis": (wavefronts, ideal), "synthesis": (wavefronts, ideal)}"""
    pad = block_pad(R0, R1)
    blk = R0 * R1
    skew = lambda i: i + (i >> 4)
    block = lambda i: i + pad * (i // blk)
    if verbose:
        print(f"H={H} plan {R0}x{R1}x{R2} TPG={TPG} mmax={mmax}  LayBlock pad={pad} per {blk}")

    # ---- analysis: stage0 regs->S, stage1 S->B, stage2 B->S, split from S (3 stages) or B (2 stages)
    t = Tally()
    _, st = stage_accesses(H, R0, 1, TPG, None, skew)
    t.add("stage 0 stores (LaySkew)", st)
    ld, st = stage_accesses(H, R1, R0, TPG, skew, block)
    t.add("stage 1 loads  (LaySkew)", ld)
    t.add("stage 1 stores (LayBlock)", st)
    res = block
    if R2 > 1:
        ident = lambda i: i
        ld, st = stage_accesses(H, R2, R0 * R1, TPG, block, ident)
        t.add("stage 2 loads  (LayBlock)", ld)
        t.add("stage 2 stores (LayId)", st)
        res = ident
    instrs = []
    for m0 in range(0, mmax, 16):          # a half warp = 16 consecutive m of one quad; model both halves as the same m run
        for which in (0, 1):
            slots = []
            for lane in range(32):
                m = m0 + lane % 16
                if m >= mmax:
                    slots.append(None)
                    continue
                idx = (0 if m == H else m) if which == 0 else (0 if m in (0, H) else H - m)
                slots.append(res(idx) + (lane // 16) * 100003 * 16)   # second half warp: another quad (different rows, own phase)
            instrs.append(slots)
    t.add("split loads", instrs)
    out = {"analysis": t.report("analysis") if verbose else t.total()}

    # ---- synthesis: build -> B, stage0 B->S, stage1 S->B, fused last stage loads from B (3 stages) or S (2 stages)
    t = Tally()
    instrs = []
    for q0 in range(0, H // 2 + 1, 16):
        for which in (0, 1):
            slots = []
            for lane in range(32):
                q = q0 + lane % 16
                if q > H // 2 or (which == 1 and (q == 0 or H - q == q)):
                    slots.append(None)
                    continue
                idx = q if which == 0 else H - q
                slots.append(block(idx) + (lane // 16) * 100003 * 16)
            instrs.append(slots)
    t.add("spectrum build stores (LayBlock)", instrs)
    ld, st = stage_accesses(H, R0, 1, TPG, block, skew)
    t.add("stage 0 loads  (LayBlock)", ld)
    t.add("stage 0 stores (LaySkew)", st)
    if R2 > 1:
        ld, st = stage_accesses(H, R1, R0, TPG, skew, block)
        t.add("stage 1 loads  (LaySkew)", ld)
        t.add("stage 1 stores (LayBlock)", st)
        ld, _ = stage_accesses(H, R2, R0 * R1, TPG, block, None)
        t.add("last stage loads (LayBlock)", ld)
    else:
        ld, _ = stage_accesses(H, R1, R0, TPG, skew, None)
        t.add("last stage loads (LaySkew)", ld)
    out["synthesis"] = t.report("synthesis") if verbose else t.total()
    return out


def main():
    H, R0, R1, R2, TPG = (int(x) for x in sys.argv[1:6])
    mmax = int(sys.argv[6]) if len(sys.argv) > 6 else H // 3 + 1
    simulate(H, R0, R1, R2, TPG, mmax)


if __name__ == "__main__":
    main()
