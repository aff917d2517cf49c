"""CPU tests that pin the ORACLE (oracle/makani_oracle.py): invariants the reference's own tests encode, an independent
Y_l^m implementation (scipy), analytic harmonics, and the golden vectors generated from the reference's files."""
import math
import os
import sys

import numpy as np
import pytest
import torch

from oracle import makani_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "contractions_golden.npz")


@pytest.mark.parametrize("n", [2, 3, 32, 33, 64, 181, 721])
def test_clenshaw_curtis_weights(n):
    # /root/reference/tests/test_grids.py:153-186: weights sum to the measure, are non-negative
    x, w = O.clenshaw_curtiss_weights(n)
    assert abs(w.sum() - 2.0) < 1e-12
    assert (w > 0).all()
    assert np.allclose(x, np.cos(np.linspace(np.pi, 0, n)))
    # exact for polynomials of degree < n
    for deg in sorted({0, 1, min(n - 1, 2), min(n - 1, 6)}):
        exact = 0.0 if deg % 2 else 2.0 / (deg + 1)
        assert abs((w * x ** deg).sum() - exact) < 1e-10


def test_clenshaw_curtis_matches_fft_construction():
    # the FFT (Waldvogel) construction torch-harmonics uses, restated inline as an independent check
    def cc_fft(n):
        n1 = n - 1
        N = np.arange(1, n1, 2)
        l = len(N)
        m = n1 - l
        v = np.concatenate([2 / N / (N - 2), 1 / N[-1:], np.zeros(m)])
        v = 0 - v[:-1] - v[-1:0:-1]
        g0 = -np.ones(n1)
        g0[l] += n1
        g0[m] += n1
        g = g0 / (n1 ** 2 - 1 + (n1 % 2))
        w = np.fft.ifft(v + g).real
        return np.concatenate((w, w[:1]))

    for n in (32, 33, 64, 721):
        assert np.abs(O.clenshaw_curtiss_weights(n)[1] - cc_fft(n)).max() < 1e-14


@pytest.mark.parametrize("n", [4, 64, 240])
def test_gauss_weights(n):
    x, w = O.legendre_gauss_weights(n)
    assert abs(w.sum() - 2.0) < 1e-12 and (np.diff(x) > 0).all()


def test_latitudes_north_to_south():
    th, w = O.precompute_latitudes(33, "equiangular")
    assert th[0] == pytest.approx(0.0, abs=1e-7) and th[-1] == pytest.approx(math.pi) and (np.diff(th) > 0).all()
    th, _ = O.precompute_latitudes(32, "legendre-gauss")
    assert (np.diff(th) > 0).all() and 0 < th[0] < th[-1] < math.pi


def test_table_against_scipy_sph_harm():
    from scipy.special import sph_harm_y

    th, _ = O.precompute_latitudes(41, "equiangular")
    P = O.legpoly(24, 28, np.cos(th))
    err = 0.0
    for m in range(24):
        for l in range(m, 28):
            err = max(err, np.abs(P[m, l] - sph_harm_y(l, m, th, 0.0).real).max())
        assert np.all(P[m, :m] == 0)
    assert err < 1e-12


def test_analytic_harmonics():
    # f = Y_1^0 = sqrt(3/4pi) cos(theta) -> c[1,0] = 1;  f = 2 Re Y_2^1 -> c[2,1] = 1 (irfft semantics double m>0)
    nlat, nlon = 32, 64
    sht = O.RealSHT(nlat, nlon, grid="legendre-gauss", dtype=torch.float64)
    th, _ = O.precompute_latitudes(nlat, "legendre-gauss")
    phi = np.linspace(0, 2 * np.pi, nlon, endpoint=False)
    T, Ph = np.meshgrid(th, phi, indexing="ij")
    c = sht(torch.from_numpy(math.sqrt(3 / (4 * math.pi)) * np.cos(T)))
    ref = torch.zeros_like(c)
    ref[1, 0] = 1
    assert (c - ref).abs().max() < 1e-12
    y21 = -math.sqrt(15 / (8 * math.pi)) * np.sin(T) * np.cos(T) * np.exp(1j * Ph)  # Condon-Shortley phase
    c = sht(torch.from_numpy(2 * y21.real))
    ref = torch.zeros_like(c)
    ref[2, 1] = 1
    assert (c - ref).abs().max() < 1e-12


@pytest.mark.parametrize("grid,lmax", [("equiangular", 32), ("legendre-gauss", 64)])
def test_roundtrip_cfg1(grid, lmax):
    # BASELINE.json configs[0]: 64x128 grid, 8 channels.  Clenshaw-Curtis with 64 nodes is exact to degree 63, so the
    # equiangular round trip is exact for lmax <= 32; Gauss-Legendre (degree 127) for the full lmax = 64.
    torch.manual_seed(333)
    sht = O.RealSHT(64, 128, lmax=lmax, mmax=lmax, grid=grid, dtype=torch.float64)
    isht = O.InverseRealSHT(64, 128, lmax=lmax, mmax=lmax, grid=grid, dtype=torch.float64)
    c = torch.tril(torch.randn(1, 8, lmax, lmax, dtype=torch.complex128))  # l >= m
    c[..., 0] = c[..., 0].real.to(torch.complex128)
    x = isht(c)
    assert (sht(x) - c).abs().max() < 1e-10
    assert (isht(sht(x)) - x).abs().max() < 1e-10


def test_parseval_and_h1():
    # /root/reference/tests/test_losses.py:470-499 and utils/losses/base_loss.py:381-387
    nlat, nlon, l0 = 32, 64, 4
    isht = O.InverseRealSHT(nlat, nlon, lmax=nlat, mmax=nlat, grid="equiangular", dtype=torch.float64)
    sht = O.RealSHT(nlat, nlon, lmax=nlat, mmax=nlat, grid="equiangular", dtype=torch.float64)
    e = torch.zeros(nlat, nlat, dtype=torch.complex128)
    e[l0, 0] = 1
    x = isht(e)
    th, w = O.precompute_latitudes(nlat, "equiangular")
    q = torch.from_numpy(w)[:, None] * 2 * math.pi / nlon
    assert ((x ** 2) * q).sum().item() == pytest.approx(1.0, abs=1e-10)
    c = sht(x)
    l = torch.arange(nlat, dtype=torch.float64)[:, None]
    mw = torch.ones(nlat, dtype=torch.float64)
    mw[1:] = 2
    l2 = ((c.abs() ** 2) * mw).sum() / (4 * math.pi)
    h1 = ((c.abs() ** 2) * mw * l * (l + 1)).sum() / (4 * math.pi)
    assert (h1 / l2).item() == pytest.approx(l0 * (l0 + 1), rel=1e-3)  # reference tolerance: 5 % (quadrature is inexact at lmax = nlat)
    e[:] = 0
    e[5, 3] = 1
    assert ((isht(e) ** 2) * q).sum().item() == pytest.approx(2.0, abs=1e-10)
    # constant field: only l = 0 populated (test_losses.py:500-510)
    c = sht(torch.ones(nlat, nlon, dtype=torch.float64))
    c0 = c.clone()
    c0[0, 0] = 0
    assert c0.abs().max() < 1e-4 and c[0, 0].real.item() == pytest.approx(math.sqrt(4 * math.pi))  # reference tol 1e-4


def test_split_shapes():
    # reference partitioning incl. uneven splits
    assert O.compute_split_shapes(721, 4) == [181, 181, 181, 178]
    assert O.compute_split_shapes(241, 2) == [121, 120]
    assert O.compute_split_shapes(240, 4) == [60, 60, 60, 60]
    assert O.compute_split_shapes(5, 4) == [2, 2, 1, 0] or sum(O.compute_split_shapes(5, 4)) == 5
    assert O.compute_split_shapes(7, 1) == [7]
    parts = O.split_tensor_along_dim(torch.arange(10), 0, 3)
    assert [p.numel() for p in parts] == [4, 4, 2]


def test_contractions_against_reference_golden():
    g = np.load(GOLD)
    x = torch.from_numpy(g["x"])
    for name, sep, op in (("dhconv", False, "dhconv"), ("diagonal", False, "diagonal"), ("sep_dhconv", True, "dhconv"), ("sep_diagonal", True, "diagonal")):
        y = O.contract_dense(x, torch.from_numpy(g[f"w_{name}"]), separable=sep, operator_type=op)
        assert torch.allclose(y, torch.from_numpy(g[f"y_{name}"]), atol=1e-5, rtol=1e-4), name  # tol of tests/test_contractions.py
    xa = torch.from_numpy(g["xa"])
    y = O.spectral_attention_mlp(xa, [], torch.from_numpy(g["w_shared"]), operator_type="diagonal")
    assert torch.allclose(y, torch.from_numpy(g["y_shared"]), atol=1e-5, rtol=1e-4)
    y = O.spectral_attention_mlp(xa, [], torch.from_numpy(g["w_ldep"]), operator_type="l-dependant")
    assert torch.allclose(y, torch.from_numpy(g["y_ldep"]), atol=1e-5, rtol=1e-4)


def test_complex_relu_against_reference_golden():
    g = np.load(GOLD)
    z = torch.from_numpy(g["z"])
    for mode in ("real", "cartesian", "modulus", "halfplane"):
        bias = torch.from_numpy(g[f"relu_bias_{mode}"]) if f"relu_bias_{mode}" in g else 0.0
        y = O.complex_relu(z, mode=mode, bias=bias, negative_slope=0.1)
        assert torch.allclose(y, torch.from_numpy(g[f"relu_{mode}"]), atol=1e-6, rtol=1e-5), mode


CONV_GOLD = os.path.join(os.path.dirname(__file__), "golden", "spectral_conv_golden.npz")
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden import CONV_GOLDEN_CASES  # noqa: E402  (case table only; the generator itself needs /root/reference)


@pytest.mark.parametrize("name", sorted(CONV_GOLDEN_CASES))
def test_spectral_conv_against_reference_class_golden(name):
    """oracle.spectral_conv_forward == the REFERENCE's SpectralConv class (spectral_convolution.py:116-264, run by
    tests/golden/make_golden.py on the oracle transforms): output, resampled residual and all gradients."""
    g = np.load(CONV_GOLD)
    nlat_i, nlon_i, grid_i, nlat_o, nlon_o, grid_o, lmax, mmax, B, Cin, Cout, G, op, sep, bias = CONV_GOLDEN_CASES[name]
    sht = O.RealSHT(nlat_i, nlon_i, lmax, mmax, grid_i)
    isht = O.InverseRealSHT(nlat_o, nlon_o, lmax, mmax, grid_o)
    x = torch.from_numpy(g[f"{name}/x"]).requires_grad_(True)
    w = torch.from_numpy(g[f"{name}/weight"]).requires_grad_(True)
    assert list(w.shape) == list(g[f"{name}/weight_shape"])
    b = torch.from_numpy(g[f"{name}/bias"]).requires_grad_(True) if bias else None
    y, res = O.spectral_conv_forward(x, w, sht, isht, num_groups=G, operator_type=op, separable=sep, bias=b)
    tol = dict(atol=2e-5, rtol=1e-4)
    assert torch.allclose(y, torch.from_numpy(g[f"{name}/y"]), **tol)
    loss = (y * torch.from_numpy(g[f"{name}/gy"])).sum()
    if f"{name}/residual" in g:
        assert torch.allclose(res, torch.from_numpy(g[f"{name}/residual"]), **tol)
        loss = loss + (res * torch.from_numpy(g[f"{name}/gres"])).sum()
    else:
        assert res is x
    loss.backward()
    assert torch.allclose(x.grad, torch.from_numpy(g[f"{name}/dx"]), **tol)
    assert torch.allclose(w.grad, torch.from_numpy(g[f"{name}/dweight"]), atol=1e-4, rtol=1e-4)
    if bias:
        assert torch.allclose(b.grad, torch.from_numpy(g[f"{name}/dbias"]), atol=1e-3, rtol=1e-4)


def test_spectral_conv_oracle_shapes_and_residual():
    torch.manual_seed(333)
    sht = O.RealSHT(33, 64, lmax=16, mmax=17, grid="equiangular")
    isht = O.InverseRealSHT(24, 48, lmax=16, mmax=17, grid="legendre-gauss")
    x = torch.randn(2, 6, 33, 64)
    w = torch.randn(2, 3, 2, 16, dtype=torch.complex64)
    y, res = O.spectral_conv_forward(x, w, sht, isht, num_groups=2, operator_type="dhconv", bias=torch.ones(1, 4, 1, 1))
    assert y.shape == (2, 4, 24, 48) and res.shape == (2, 6, 24, 48)
    # identity weight on a same-grid pair reproduces the band-limited projection
    sht2 = O.RealSHT(24, 48, lmax=16, mmax=17, grid="legendre-gauss")
    wid = torch.zeros(1, 6, 6, 16, dtype=torch.complex64)
    wid[0, torch.arange(6), torch.arange(6), :] = 1
    xb = isht(sht2(torch.randn(1, 6, 24, 48)))
    y, res = O.spectral_conv_forward(xb, wid, sht2, isht)
    assert res is xb and torch.allclose(y, xb, atol=1e-4)
