"""Every entry point of include/b200sht.h that a reference-side binding would call is EXECUTED here
through ctypes with plain pointers -- no makani_b200 Python modules in between -- and checked against the CPU oracle:
b200sht_sht_forward / _inverse / _forward_adjoint / _inverse_adjoint (replacing torch_harmonics.RealSHT / InverseRealSHT forward and
autograd backward at makani/models/common/spectral_convolution.py:239,241,253) and b200sht_spectral_conv_forward_host (host numpy
buffers in, host buffers out: the `e2e` path of bench.py)."""
import ctypes

import numpy as np
import pytest
import torch

import makani_b200 as mb
from makani_b200 import _lib
from oracle import makani_oracle as O
from test_gpu_parity import close

pytestmark = pytest.mark.gpu
DEV = "cuda"
VP = ctypes.c_void_p


def _p(t):
    return VP(t.data_ptr())


def _stream():
    return VP(torch.cuda.current_stream(torch.device(DEV)).cuda_stream)


CASES = [("equiangular", 64, 128, 64, 65, 1, 8), ("legendre-gauss", 48, 96, 32, 33, 2, 5), ("equiangular", 91, 180, 91, 91, 1, 3)]


@pytest.mark.parametrize("grid,nlat,nlon,lmax,mmax,B,C", CASES)
@pytest.mark.parametrize("precision,rtol", [(_lib.PREC_FP32, 1e-5), (_lib.PREC_TF32, 1e-3)])
def test_sht_boundary_calls(grid, nlat, nlon, lmax, mmax, B, C, precision, rtol):
    torch.manual_seed(333)
    lib = _lib.load()
    plan = mb.get_plan(nlat, nlon, lmax, mmax, grid, True, torch.device(DEV))
    ws = torch.empty(int(lib.b200sht_sht_workspace_bytes(plan.handle, B, C)), dtype=torch.uint8, device=DEV)
    osht = O.RealSHT(nlat, nlon, lmax, mmax, grid, dtype=torch.float64)
    oisht = O.InverseRealSHT(nlat, nlon, lmax, mmax, grid, dtype=torch.float64)
    keep = torch.tril(torch.ones(lmax, mmax)).bool()
    tag = f"{grid} {nlat}x{nlon} prec={precision}"

    # forward + its adjoint
    x = torch.randn(B, C, nlat, nlon)
    xd = x.to(DEV)
    coeffs = torch.full((B * C, lmax, mmax), float("nan"), dtype=torch.complex64, device=DEV)
    _lib.check(lib.b200sht_sht_forward(plan.handle, _p(xd), _lib.F32, B, C, _p(coeffs), _p(ws), precision, _stream()), "sht_forward")
    xr = x.double().requires_grad_(True)
    cref = osht(xr)
    close(coeffs.view(B, C, lmax, mmax), cref, rtol, f"b200sht_sht_forward {tag}")
    gc = torch.randn(B, C, lmax, mmax, dtype=torch.complex64) * keep
    cref.backward(gc.to(torch.complex128))
    gx = torch.full((B, C, nlat, nlon), float("nan"), device=DEV)
    _lib.check(lib.b200sht_sht_forward_adjoint(plan.handle, _p(gc.to(DEV).contiguous()), _p(gx), _lib.F32, B, C, _p(ws), precision, _stream()), "sht_forward_adjoint")
    close(gx, xr.grad, rtol, f"b200sht_sht_forward_adjoint {tag}")

    # inverse + its adjoint
    cin = torch.randn(B, C, lmax, mmax, dtype=torch.complex64) * keep
    y = torch.full((B, C, nlat, nlon), float("nan"), device=DEV)
    _lib.check(lib.b200sht_sht_inverse(plan.handle, _p(cin.to(DEV).contiguous()), _p(y), _lib.F32, B, C, _p(ws), precision, _stream()), "sht_inverse")
    cr = cin.to(torch.complex128).requires_grad_(True)
    yref = oisht(cr)
    close(y, yref, rtol, f"b200sht_sht_inverse {tag}")
    gy = torch.randn(B, C, nlat, nlon)
    yref.backward(gy.double())
    gcoef = torch.full((B * C, lmax, mmax), float("nan"), dtype=torch.complex64, device=DEV)
    _lib.check(lib.b200sht_sht_inverse_adjoint(plan.handle, _p(gy.to(DEV)), _lib.F32, B, C, _p(gcoef), _p(ws), precision, _stream()), "sht_inverse_adjoint")
    close(gcoef.view(B, C, lmax, mmax).cpu() * keep, cr.grad * keep, rtol, f"b200sht_sht_inverse_adjoint {tag}")


@pytest.mark.parametrize("dtype,np_dtype", [(_lib.F32, np.float32), (_lib.BF16, None)])
@pytest.mark.parametrize("precision,rtol", [(_lib.PREC_FP32, 1e-5), (_lib.PREC_TF32, 1e-3)])
def test_spectral_conv_forward_host(dtype, np_dtype, precision, rtol):
    """host buffer in -> host buffer out, the call an FFI user makes (pageable host memory, the library stages the copies)"""
    torch.manual_seed(333)
    lib = _lib.load()
    nlat, nlon, L, M, B, Ci, Co = 65, 128, 32, 33, 2, 6, 10
    dev = torch.device(DEV)
    pf = mb.get_plan(nlat, nlon, L, M, "equiangular", True, dev)
    pv = mb.get_plan(48, 96, L, M, "legendre-gauss", True, dev)
    w = torch.randn(1, Ci, Co, L, dtype=torch.complex64)
    wd = w.to(DEV)
    wp = torch.empty(int(lib.b200sht_mix_weight_elems(_lib.OP_DHCONV, L, M, 1, Ci, Co)), device=DEV)
    _lib.check(lib.b200sht_mix_weight_pack(_lib.OP_DHCONV, _p(wd), _p(wp), L, 1, Ci, Co, precision, _stream()), "mix_weight_pack")
    bias = torch.randn(Co)
    bd = bias.to(DEV)
    x = torch.randn(B, Ci, nlat, nlon)
    if dtype == _lib.BF16:
        xh = x.bfloat16().contiguous()
        yh = torch.zeros(B, Co, 48, 96, dtype=torch.bfloat16)
        x = xh.float()
        xptr, yptr = VP(xh.data_ptr()), VP(yh.data_ptr())
    else:
        xh = np.ascontiguousarray(x.numpy())
        yh_np = np.zeros((B, Co, 48, 96), dtype=np_dtype)
        xptr, yptr = xh.ctypes.data_as(VP), yh_np.ctypes.data_as(VP)
    desc = _lib.ConvDesc(B, Ci, Co, 1, _lib.OP_DHCONV, dtype, precision)
    _lib.check(lib.b200sht_spectral_conv_forward_host(pf.handle, pv.handle, ctypes.byref(desc), xptr, _p(wp), _p(bd), yptr, _stream()), "spectral_conv_forward_host")
    y = yh.float() if dtype == _lib.BF16 else torch.from_numpy(yh_np)
    of = O.RealSHT(nlat, nlon, L, M, "equiangular", dtype=torch.float64)
    oi = O.InverseRealSHT(48, 96, L, M, "legendre-gauss", dtype=torch.float64)
    yref, _ = O.spectral_conv_forward(x.double(), w.to(torch.complex128), of, oi, bias=bias.double().reshape(1, Co, 1, 1))
    close(y, yref, max(rtol, 4e-3) if dtype == _lib.BF16 else rtol, f"b200sht_spectral_conv_forward_host dtype={dtype} prec={precision}")
