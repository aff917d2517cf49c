"""Host-side model of the shared-memory exchange layouts of the compile-time FFT plans (makani_b200/csrc/fft.cu: LaySkew / LayBlock /
LayId).  scripts/smem_sim.py replays every shared-memory access of a plan through the 32-bank / half-warp conflict rule; the plans
listed in CT_PLANS must stay (nearly) conflict-free -- the headline 1440-point plan exactly."""
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import smem_sim  # noqa: E402


def ct_plans():
    src = open(os.path.join(ROOT, "makani_b200", "csrc", "fft.cu")).read()
    block = src[src.index("#define CT_PLANS(X)"):]
    block = block[:block.index("template <typename T>")]
    return [tuple(int(v) for v in m) for m in re.findall(r"X\((\d+), (\d+), (\d+), (\d+), (\d+), (\d+), (\d+)\)", block)]


def test_plan_table_is_consistent():
    plans = ct_plans()
    assert len(plans) >= 10
    seen = set()
    for rows, groups, tpg, r0, r1, r2, minb in plans:
        H = r0 * r1 * r2
        assert H not in seen, f"two plans for nlon={2 * H}"
        seen.add(H)
        assert rows % groups == 0 and (rows // groups) % 2 == 0 and rows % 4 == 0      # row pairs per thread, quads in the split pass
        assert r0 in (2, 4, 8, 16)                                                    # LaySkew assumes a power-of-two first radix
        assert H // r0 <= tpg                                                          # one stage-0 butterfly per thread (register prefetch)
        assert (groups * tpg) % (16 * (rows // 4)) == 0                                # split pass: half warp = 16 orders of one quad
        assert 1 <= minb <= 4
    assert 720 in seen  # nlon 1440, the benchmark grid


def test_headline_plan_is_conflict_free():
    plan = next(p for p in ct_plans() if p[3] * p[4] * p[5] == 720)
    _, _, tpg, r0, r1, r2, _ = plan
    res = smem_sim.simulate(720, r0, r1, r2, tpg, 241, verbose=False)
    assert res["analysis"][0] == res["analysis"][1]
    assert res["synthesis"][0] == res["synthesis"][1]


@pytest.mark.parametrize("plan", ct_plans(), ids=lambda p: f"nlon{2 * p[3] * p[4] * p[5]}")
def test_every_plan_stays_below_1p5x_ideal(plan):
    _, _, tpg, r0, r1, r2, _ = plan
    H = r0 * r1 * r2
    res = smem_sim.simulate(H, r0, r1, r2, tpg, H // 3 + 1, verbose=False)
    for kernel, (wf, ideal) in res.items():
        assert wf <= 1.5 * ideal, (kernel, wf, ideal)
