"""CPU tests of the tensor-core DFT's factorisation (csrc/dft.cu, dft_math.cuh): the radix-8 stage, twiddles and index maps are the
same __host__ __device__ code as the kernels (b200sht_debug_dft_host sums the GEMM in double on the host); the reference is numpy's
rfft / irfft, i.e. the semantics torch_harmonics.RealSHT / InverseRealSHT apply along longitude."""
import ctypes

import numpy as np
import pytest

from makani_b200 import _lib

CASES = [(1440, 241), (480, 241), (128, 65), (64, 17), (360, 100), (72, 37), (24, 13), (720, 241), (512, 256), (1440, 121), (16, 9), (1520, 256)]


def _call(N, mmax, direction, mode, rs, inp):
    lib = _lib.load()
    inp = np.ascontiguousarray(inp, dtype=np.float32)
    out = np.zeros(2 * mmax if direction == 0 else N, dtype=np.float32)
    rc = lib.b200sht_debug_dft_host(N, mmax, direction, mode, ctypes.c_float(rs), inp.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, lib.b200sht_last_error()
    return out


@pytest.mark.parametrize("N,mmax", CASES)
@pytest.mark.parametrize("mode", [0, 1])
def test_dft_analysis_host(N, mmax, mode):
    rng = np.random.default_rng(N + mmax + mode)
    x = rng.standard_normal(N).astype(np.float32)
    rs = 0.37
    got = _call(N, mmax, 0, mode, rs, x)
    X = np.fft.rfft(x.astype(np.float64))[:mmax]
    if mode == 0:
        X = X * rs                                     # quadrature weight * 2 pi / nlon folded into the row scale
    else:
        sc = np.full(mmax, 2.0)
        sc[0] = 1.0
        if mmax == N // 2 + 1:
            sc[-1] = 1.0
        X = X * sc                                     # adjoint of irfft(norm="forward")
    ref = np.stack([X.real, X.imag], -1).reshape(-1)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < 2e-6, err


@pytest.mark.parametrize("N,mmax", CASES)
@pytest.mark.parametrize("mode", [0, 1])
def test_dft_synthesis_host(N, mmax, mode):
    rng = np.random.default_rng(7 * N + mmax + mode)
    Z = (rng.standard_normal(mmax) + 1j * rng.standard_normal(mmax)).astype(np.complex64)
    rs = 0.81
    got = _call(N, mmax, 1, mode, rs, np.stack([Z.real, Z.imag], -1).reshape(-1))
    Zd = Z.astype(np.complex128)
    if mode == 0:
        full = np.zeros(N // 2 + 1, dtype=np.complex128)
        full[:mmax] = Zd
        ref = np.fft.irfft(full, n=N) * N             # irfft(norm="forward")
    else:
        j = np.arange(N)
        ref = np.zeros(N)
        for m in range(mmax):
            ref += Zd[m].real * np.cos(2 * np.pi * m * j / N) - Zd[m].imag * np.sin(2 * np.pi * m * j / N)
        ref *= rs
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < 2e-6, err


def test_dft_rejects_unsupported():
    lib = _lib.load()
    buf = np.zeros(4096, dtype=np.float32)
    p = buf.ctypes.data_as(ctypes.c_void_p)
    assert lib.b200sht_debug_dft_host(1442, 100, 0, 0, ctypes.c_float(1.0), p, p) != 0      # nlon not a multiple of 8
    assert lib.b200sht_debug_dft_host(2880, 300, 0, 0, ctypes.c_float(1.0), p, p) != 0      # beyond the kernel's range
