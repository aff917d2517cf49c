#!/usr/bin/env python
"""Structural diagnosis of the tensor-core DFT synthesis kernel: feed unit impulses Z[m, p, r=0, k] = amp(k) and report which
(row k', order m', re/im) each one lands on.  python scripts/dft_diag.py [nlon nlat mmax]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import makani_b200 as mb
from makani_b200 import _lib
from makani_b200.sht import _ptr, _stream, _VP

nlon, nlat, mmax = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (64, 8, 17)
dev = torch.device("cuda", 0)
plan = mb.get_plan(nlat, nlon, min(nlat, 8), mmax, "legendre-gauss", True, dev)
print("dft_ok", plan.dft_ok, "kp", plan.kp)
B, C = 1, 1
st = _stream(dev)
amp = 1.0 + torch.arange(plan.kp, device=dev) / 16.0
ampc = amp[:nlat].cpu().double()
bad = 0
for m in list(range(min(mmax, 20))) + [mmax - 1]:
    for p in (0, 1):
        Z = torch.zeros(mmax, 2, 1, plan.kp, device=dev)
        Z[m, p, 0] = amp
        M2 = (mmax + 7) // 8
        Zp = torch.zeros(8 * M2, 2, 1, plan.kp, device=dev)
        Zp[:mmax] = Z
        lat = Zp.view(M2, 8, 2, 1, plan.kp // 8, 8).permute(3, 4, 2, 0, 1, 5).contiguous().reshape(-1)   # tiled layout
        y = torch.full((1, 1, nlat, nlon), float("nan"), device=dev)
        _lib.call("b200sht_fft_synthesis", plan.handle, _ptr(lat), _ptr(y), 0, B, C, _VP(0), 1 | 2, st)   # mode 1: y = rowscale * sum_m Re(Z e^{i m phi})
        torch.cuda.synchronize()
        Y = torch.fft.rfft(y[0, 0].double().cpu(), dim=-1) / nlon * 2          # [nlat][nlon/2+1]; a cos(m phi) row -> 1 at m
        Y[:, 0] /= 2
        rs = (Y.abs().max(dim=1).values)
        # expected: row k has magnitude rowscale[k] * amp[k] at order m, phase 0 (p=0) or +90 deg (p=1: -sin -> +i ... )
        sig = (Y.abs() > 1e-3 * Y.abs().max().clamp_min(1e-30)).nonzero().tolist()
        ok = all(mm == m for _, mm in sig) and len(sig) == nlat and bool(torch.isfinite(y).all())
        if not ok:
            bad += 1
            top = sorted(sig, key=lambda km: -Y[km[0], km[1]].abs().item())[:12]
            print(f"m={m} p={p}: nonfinite={int((~torch.isfinite(y)).sum())} hits(k', m', value):", [(k, mm, complex(round(Y[k, mm].real.item(), 3), round(Y[k, mm].imag.item(), 3))) for k, mm in top])
        elif m < 3:
            print(f"m={m} p={p}: ok, row factors", [round(v, 4) for v in (Y[:, m] / ampc).real.tolist()[:4]], [round(v, 4) for v in (Y[:, m] / ampc).imag.tolist()[:4]])
print("bad impulses:", bad)
