"""precision="fp32x3" (B200SHT_PREC_FP32X3): fp32 operands, Legendre stages as 3 x TF32 on the tensor cores.  Against the fp64 oracle it must
stay inside the fp32 tolerance of BASELINE.json's north_star (rtol 1e-5 element bound); beside it the CUDA-core "fp32" mode on the same inputs."""
import pytest
import torch

import makani_b200 as mb
from oracle import makani_oracle as O
from test_gpu_parity import _run_conv_case
from test_gpu_bench_configs import CFG_2C

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _rel(a, b):
    return float((a.cpu().to(b.dtype) - b).abs().pow(2).sum().sqrt() / b.abs().pow(2).sum().sqrt())


@pytest.mark.parametrize("grid,nlat,nlon,lmax,mmax,B,C", [("equiangular", 181, 360, 90, 91, 2, 5), ("legendre-gauss", 64, 128, 40, 33, 1, 8),
                                                          ("equiangular", 721, 1440, 240, 241, 1, 3)])
def test_sht_fp32x3_vs_fp32(grid, nlat, nlon, lmax, mmax, B, C):
    torch.manual_seed(7)
    x = torch.randn(B, C, nlat, nlon, dtype=torch.float64)
    c64 = O.RealSHT(nlat, nlon, lmax, mmax, grid, dtype=torch.float64)(x)
    y64 = O.InverseRealSHT(nlat, nlon, lmax, mmax, grid, dtype=torch.float64)(c64)
    err = {}
    for prec in ("fp32", "fp32x3"):
        f = mb.RealSHT(nlat, nlon, lmax, mmax, grid, precision=prec)
        i = mb.InverseRealSHT(nlat, nlon, lmax, mmax, grid, precision=prec)
        xd = x.float().to(DEV).requires_grad_()
        c = f(xd)
        y = i(c64.to(torch.complex64).to(DEV))
        (gx,) = torch.autograd.grad(c, xd, c64.to(torch.complex64).to(DEV))   # adjoint of the analysis applied to the spectrum
        err[prec] = {"sht": _rel(c, c64), "isht": _rel(y, y64), "finite_grad": bool(torch.isfinite(gx).all())}
        assert err[prec]["finite_grad"]
    print(f"[parity] {grid} {nlat}x{nlon} rel-L2 vs fp64 oracle: {err}")
    for what in ("sht", "isht"):
        assert err["fp32"][what] < 1e-6, err
        assert err["fp32x3"][what] < 6e-6, err        # truncating fp32 accumulation of the tensor core: grows with nlat (2.5e-6 at 721)


def test_benched_block_fp32x3():
    """bench.py --precision fp32x3 object at the headline size: element bound rtol 1e-5 (as for "fp32"), relative L2 below 1.2e-5"""
    rel = _run_conv_case(CFG_2C, "fp32x3", 1e-5, act_dtype=torch.float32)
    print(f"[benched] sfno_block_721x1440x73 fp32x3 rel_l2: {rel}")
    for k, v in rel.items():
        assert v < 1.2e-5, (k, rel)
