"""Host-side bookkeeping of bench.py (no GPU): workload table, algorithmic bytes / flops, JSON contract of the
reference arm's line builder."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_nnz_matches_survey():
    assert bench.nnz_modes(240, 241) == 28920
    assert bench.nnz_modes(360, 361) == 64980


@pytest.mark.parametrize("wl", sorted(bench.WORKLOADS))
def test_stage_flops_sum_to_the_block_total(wl):
    assert sum(bench.stage_flops(wl).values()) == bench.flops_fwd_bwd(wl)


def test_headline_workload_numbers():
    wl = "sfno_block_721x1440x73"
    assert bench.WORKLOADS[wl][:2] == (721, 1440) and bench.WORKLOADS[wl][-1] == 73
    # SURVEY 8(d), cfg 2c: SHT 6.09 + mix 1.23 + iSHT 6.09 GFLOP forward
    fl = bench.stage_flops(wl)
    assert abs(fl["legendre_analysis_in"] / 1e9 - 6.09) < 0.01 and abs(fl["mix_forward"] / 1e9 - 1.23) < 0.01
    sb = bench.stage_bytes(wl, 2)
    assert sb["fft_analysis_in"] == 73 * 721 * 1440 * 2 + 73 * 721 * 241 * 8      # bf16 rows in, truncated complex spectrum out
    assert set(sb) == {"fft_analysis_in", "legendre_analysis_in", "mix_forward", "legendre_synthesis_out", "fft_synthesis_out",
                       "fft_analysis_out", "legendre_analysis_out", "mix_backward", "legendre_synthesis_in", "fft_synthesis_in"}


def test_stage_annotation_never_touches_fft_stages():
    st = {"legendre_analysis_in": {"ms": 0.05}, "fft_analysis_in": {"ms": 0.13}, "mix_backward": {"ms": 0.0}}
    out = bench.add_stage_tflops(st, "sfno_block_721x1440x73")
    assert "TFLOPs" in out["legendre_analysis_in"] and "TFLOPs" not in out["fft_analysis_in"] and "TFLOPs" not in out["mix_backward"]


def test_interior_block_flops_match_survey():
    # cfg 2a: 55.4 GFLOP forward, 145.0 GFLOP fwd+bwd
    assert abs(bench.flops_fwd_bwd("sfno_block_240x480x384") / 1e9 - 145.0) < 0.1
