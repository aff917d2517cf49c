"""
CPU ORACLE for the spherical-harmonic hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this module.  Nothing under `makani_b200/` imports it; the product path fails loudly when the CUDA library
is missing (see makani_b200/_lib.py).

What is restated here (pure PyTorch on CPU, fp32 or fp64 selectable):

  * torch-harmonics  (third-party dependency of the reference, NOT vendored in /root/reference and not
    installable here; pinned by the reference at commit 887006c640f1d61c3f80590ecc2b207bbb647072,
    /root/reference/docker/Dockerfile:88-90, pyproject.toml:67 `torch-harmonics>=0.9.0`):
      - quadrature.{legendre_gauss_weights, clenshaw_curtiss_weights, precompute_latitudes}
      - legendre precompute (orthonormal associated Legendre functions, Condon-Shortley phase)
      - RealSHT / InverseRealSHT forward passes
      - distributed.{compute_split_shapes, split_tensor_along_dim}
    Published algorithm: X = 2*pi*rfft(x, norm="forward")[..., :mmax]; c[l,m] = sum_k X[k,m] * P[m,l,k]*w_k ;
    inverse: Z[k,m] = sum_l c[l,m] P[m,l,k]; x = irfft(Z, n=nlon, norm="forward").
    Reference call sites that fix the calling convention:
      /root/reference/makani/models/networks/sfnonet.py:792-805   (ctor arguments)
      /root/reference/makani/models/common/spectral_convolution.py:239-253 (forward/inverse use)
      /root/reference/makani/utils/grids.py:67-68,120-129         (quadrature: tensors, dlambda*w)
      /root/reference/makani/utils/losses/base_loss.py:381-387    (Parseval: m>0 counted twice, 1/(4 pi))
      /root/reference/tests/test_losses.py:470-499                (absolute pin: ortho norm, [l, m] order)

  * makani's own spectral layers:
      - SpectralConv.forward        /root/reference/makani/models/common/spectral_convolution.py:213-264
      - _contract_* einsums         /root/reference/makani/models/common/contractions.py:19-54
      - SpectralAttention (INTENDED semantics; the reference raises at HEAD)
                                    /root/reference/makani/models/common/spectral_convolution.py:433-519
      - ComplexReLU                 /root/reference/makani/models/common/activations.py:20-127

PARITY PINNING STATUS
  * contractions + ComplexReLU: pinned element-wise against the reference's own files, imported in the build
    container by tests/golden/make_golden.py  ->  tests/golden/contractions_golden.npz.
  * spectral_conv_forward: pinned element-wise (output, resampled residual, every gradient) against the reference's own
    SpectralConv class, imported by path with its package imports stubbed and run on this file's transforms
    (tests/golden/make_golden.py  ->  tests/golden/spectral_conv_golden.npz; tests/test_oracle.py).
  * the reference's own test classes that go through torch_harmonics (tests/test_losses.py spectral losses, test_grids.py,
    test_noise.py: 193 tests) run unmodified against this file posed as `torch_harmonics` and pass
    (tests/reference_suites/run_reference_tests.py, report.txt).
  * RealSHT / InverseRealSHT: "PARITY UNPINNED" element-wise against torch-harmonics (package absent, no
    network).  Pinned instead by (a) every invariant the reference's tests encode (Parseval, H1 = l(l+1) L2,
    constant field -> only l=0, quadrature sums, GRF variance), (b) an independent implementation of Y_l^m
    (scipy.special.sph_harm_y, Condon-Shortley phase) for the Legendre table, (c) analytic harmonics.
  * SpectralAttention: "PARITY UNPINNED" (reference raises; intended semantics implemented).
"""

import math
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn


# --------------------------------------------------------------------------------------------------
# quadrature  (torch_harmonics.quadrature restated)
# --------------------------------------------------------------------------------------------------
def legendre_gauss_weights(n: int, a: float = -1.0, b: float = 1.0):
    """Gauss-Legendre nodes (ascending in cos(theta)) and weights on [a,b]."""
    xlg, wlg = np.polynomial.legendre.leggauss(n)
    xlg = (b - a) * 0.5 * xlg + (b + a) * 0.5
    wlg = wlg * (b - a) * 0.5
    return xlg, wlg


def clenshaw_curtiss_weights(n: int, a: float = -1.0, b: float = 1.0):
    """Clenshaw-Curtis nodes cos(linspace(pi, 0, n)) (poles included) and weights on [a,b].

    Weights by the closed-form cosine sum (equivalent to the FFT construction used upstream):
        w_k = c_k/(n-1) * (1 - sum_{j=1}^{floor((n-1)/2)} b_j/(4j^2-1) cos(2 j k pi/(n-1)))
    with c_k = 1 at the end points, 2 otherwise; b_j = 1 if 2j == n-1 else 2.
    """
    assert n > 1
    tcc = np.cos(np.linspace(np.pi, 0.0, n))
    if n == 2:
        wcc = np.array([1.0, 1.0])
    else:
        n1 = n - 1
        k = np.arange(n, dtype=np.float64)
        wcc = np.ones(n, dtype=np.float64)
        for j in range(1, n1 // 2 + 1):
            bj = 1.0 if 2 * j == n1 else 2.0
            wcc -= bj / (4.0 * j * j - 1.0) * np.cos(2.0 * j * k * np.pi / n1)
        ck = np.full(n, 2.0)
        ck[0] = ck[-1] = 1.0
        wcc = ck / n1 * wcc
    tcc = (b - a) * 0.5 * tcc + (b + a) * 0.5
    wcc = wcc * (b - a) * 0.5
    return tcc, wcc


def precompute_latitudes(nlat: int, grid: str = "equiangular"):
    """Colatitudes theta (ascending: row 0 = north pole side) and quadrature weights on cos(theta) in [-1,1]."""
    if grid == "legendre-gauss":
        cost, w = legendre_gauss_weights(nlat, -1.0, 1.0)
    elif grid == "equiangular":
        cost, w = clenshaw_curtiss_weights(nlat, -1.0, 1.0)
    else:
        raise ValueError(f"Unknown quadrature mode {grid}")
    theta = np.flip(np.arccos(np.clip(cost, -1.0, 1.0))).copy()
    w = np.flip(w).copy()
    return theta, w


# --------------------------------------------------------------------------------------------------
# Legendre table  (torch_harmonics.legendre restated)
# --------------------------------------------------------------------------------------------------
def legpoly(mmax: int, lmax: int, x: np.ndarray, csphase: bool = True) -> np.ndarray:
    """P[m,l,k]: orthonormal associated Legendre functions at x_k=cos(theta_k), fp64, shape (mmax,lmax,len(x))."""
    nmax = max(mmax, lmax)
    vdm = np.zeros((nmax + 1, nmax + 1, len(x)), dtype=np.float64)
    vdm[0, 0, :] = 1.0 / math.sqrt(4.0 * math.pi)
    for l in range(1, nmax + 1):
        vdm[l - 1, l, :] = math.sqrt(2 * l + 1) * x * vdm[l - 1, l - 1, :]
        vdm[l, l, :] = np.sqrt((2 * l + 1) * (1 + x) * (1 - x) / (2 * l)) * vdm[l - 1, l - 1, :]
    for l in range(2, nmax + 1):
        for m in range(0, l - 1):
            a = math.sqrt((2 * l - 1) / (l - m) * (2 * l + 1) / (l + m))
            b = math.sqrt((l + m - 1) / (l - m) * (2 * l + 1) / (2 * l - 3) * (l - m - 1) / (l + m))
            vdm[m, l, :] = x * a * vdm[m, l - 1, :] - b * vdm[m, l - 2, :]
    vdm = vdm[:mmax, :lmax]
    if csphase:
        vdm[1::2] *= -1.0
    return np.ascontiguousarray(vdm)


# --------------------------------------------------------------------------------------------------
# RealSHT / InverseRealSHT  (torch_harmonics.sht restated)
# --------------------------------------------------------------------------------------------------
class RealSHT(nn.Module):
    def __init__(self, nlat, nlon, lmax=None, mmax=None, grid="equiangular", norm="ortho", csphase=True, dtype=torch.float32):
        super().__init__()
        if norm != "ortho":
            raise NotImplementedError("only norm='ortho' is requested anywhere in makani")
        self.nlat, self.nlon, self.grid, self.norm, self.csphase = nlat, nlon, grid, norm, csphase
        self.lmax = lmax or nlat
        self.mmax = mmax or nlon // 2 + 1
        theta, w = precompute_latitudes(nlat, grid)
        pct = legpoly(self.mmax, self.lmax, np.cos(theta), csphase=csphase)
        weights = torch.from_numpy(pct * w[None, None, :]).to(dtype)
        self.register_buffer("weights", weights, persistent=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        assert x.shape[-2] == self.nlat and x.shape[-1] == self.nlon
        X = 2.0 * math.pi * torch.fft.rfft(x, dim=-1, norm="forward")
        X = torch.view_as_real(X)[..., : self.mmax, :]
        w = self.weights.to(X.dtype)
        re = torch.einsum("...km,mlk->...lm", X[..., 0], w)
        im = torch.einsum("...km,mlk->...lm", X[..., 1], w)
        return torch.view_as_complex(torch.stack([re, im], dim=-1).contiguous())


class InverseRealSHT(nn.Module):
    def __init__(self, nlat, nlon, lmax=None, mmax=None, grid="equiangular", norm="ortho", csphase=True, dtype=torch.float32):
        super().__init__()
        if norm != "ortho":
            raise NotImplementedError("only norm='ortho' is requested anywhere in makani")
        self.nlat, self.nlon, self.grid, self.norm, self.csphase = nlat, nlon, grid, norm, csphase
        self.lmax = lmax or nlat
        self.mmax = mmax or nlon // 2 + 1
        theta, _ = precompute_latitudes(nlat, grid)
        pct = torch.from_numpy(legpoly(self.mmax, self.lmax, np.cos(theta), csphase=csphase)).to(dtype)
        self.register_buffer("pct", pct, persistent=False)

    def forward(self, c: torch.Tensor) -> torch.Tensor:
        assert c.shape[-2] == self.lmax and c.shape[-1] == self.mmax
        cr = torch.view_as_real(c)
        p = self.pct.to(cr.dtype)
        re = torch.einsum("...lm,mlk->...km", cr[..., 0], p)
        im = torch.einsum("...lm,mlk->...km", cr[..., 1], p)
        im = im.clone()
        im[..., 0] = 0.0
        if self.mmax > self.nlon // 2 and self.nlon % 2 == 0:
            im[..., self.nlon // 2] = 0.0
        Z = torch.view_as_complex(torch.stack([re, im], dim=-1).contiguous())
        return torch.fft.irfft(Z, n=self.nlon, dim=-1, norm="forward")


# --------------------------------------------------------------------------------------------------
# distributed helpers (torch_harmonics.distributed.utils restated)
# --------------------------------------------------------------------------------------------------
def compute_split_shapes(size: int, num_chunks: int) -> List[int]:
    if num_chunks == 1:
        return [size]
    chunk = (size + num_chunks - 1) // num_chunks
    last = max(0, size - chunk * (num_chunks - 1))
    if last == 0:
        chunk = size // num_chunks
        last = size - chunk * (num_chunks - 1)
    return [chunk] * (num_chunks - 1) + [last]


def split_tensor_along_dim(tensor, dim, num_chunks):
    assert dim < tensor.dim()
    assert tensor.shape[dim] >= num_chunks
    return torch.split(tensor, compute_split_shapes(tensor.shape[dim], num_chunks), dim=dim)


# --------------------------------------------------------------------------------------------------
# contractions (makani/models/common/contractions.py:19-151 restated)
# --------------------------------------------------------------------------------------------------
def contract_dense(x, weight, separable=False, operator_type="diagonal"):
    """x (B,G,Ci,L,M) complex, weight per reference shapes.  contractions.py:35-54."""
    if separable:
        if operator_type == "diagonal":
            return torch.einsum("bgixy,gixy->bgixy", x, weight)
        if operator_type == "dhconv":
            return torch.einsum("bgixy,gix->bgixy", x, weight)
    else:
        if operator_type == "diagonal":
            return torch.einsum("bgixy,gioxy->bgoxy", x, weight)
        if operator_type == "dhconv":
            return torch.einsum("bgixy,giox->bgoxy", x, weight)
    raise ValueError(f"Unknown operator type {operator_type}")


def complex_relu(z, mode="real", bias=0.0, negative_slope=0.0):
    """activations.py:88-127 restated (functional)."""
    act = lambda t: torch.nn.functional.leaky_relu(t, negative_slope)
    if mode == "cartesian":
        return torch.view_as_complex(act(torch.view_as_real(z)))
    if mode == "modulus":
        zabs = torch.sqrt(z.real ** 2 + z.imag ** 2)
        return torch.where(zabs + bias > 0, (zabs + bias) * z / zabs, torch.zeros_like(z))
    if mode == "halfplane":
        ang = torch.angle(z) - bias
        cond = torch.logical_and(ang >= 0.0, ang < math.pi / 2.0)
        return torch.where(cond, z, negative_slope * z)
    if mode == "real":
        zr = torch.view_as_real(z)
        outr = zr.clone()
        outr[..., 0] = act(zr[..., 0])
        return torch.view_as_complex(outr)
    raise NotImplementedError(mode)


# --------------------------------------------------------------------------------------------------
# SpectralConv / SpectralAttention forward (functional restatement)
# --------------------------------------------------------------------------------------------------
def spectral_conv_forward(x, weight, sht, isht, num_groups=1, operator_type="dhconv", separable=False, bias=None):
    """spectral_convolution.py:213-264.  Returns (y, residual).  Transforms run in fp32 (or sht's dtype)."""
    dtype = x.dtype
    residual = x
    tdtype = sht.weights.dtype
    xs = sht(x.to(tdtype)).contiguous()
    scale_residual = (sht.nlat != isht.nlat) or (sht.nlon != isht.nlon) or (sht.grid != isht.grid)
    if scale_residual:
        residual = isht(xs).to(dtype)
    B, C, H, W = xs.shape
    xg = xs.reshape(B, num_groups, C // num_groups, H, W)
    yp = contract_dense(xg, weight.to(xs.dtype), separable=separable, operator_type=operator_type)
    Cout = yp.shape[1] * yp.shape[2]
    ys = yp.reshape(B, Cout, H, W).contiguous()
    y = isht(ys).to(dtype)
    if bias is not None:
        y = y + bias.to(dtype)
    return y, residual


def spectral_attention_mlp(x, w_list, wout, b_list=None, act_mode="real", act_bias=None, operator_type="diagonal"):
    """INTENDED semantics of SpectralAttention.forward_mlp (spectral_convolution.py:433-470): complex einsums
    on the complex tensor ("bixy,io->boxy" / "bixy,xio->boxy"), ComplexReLU between layers."""
    eq = "bixy,io->boxy" if operator_type == "diagonal" else "bixy,xio->boxy"
    h = x
    for i, w in enumerate(w_list):
        h = torch.einsum(eq, h, w)
        if b_list is not None:
            h = h + b_list[i]
        ab = 0.0 if act_bias is None else act_bias[i]
        h = complex_relu(h, mode=act_mode, bias=ab)
    return torch.einsum(eq, h, wout)


def spectral_attention_forward(x, w_list, wout, sht, isht, b_list=None, act_mode="real", act_bias=None, operator_type="diagonal"):
    dtype = x.dtype
    residual = x
    tdtype = sht.weights.dtype
    xs = sht(x.to(tdtype))
    if (sht.nlat != isht.nlat) or (sht.nlon != isht.nlon) or (sht.grid != isht.grid):
        residual = isht(xs).to(dtype)
    ys = spectral_attention_mlp(xs, [w.to(xs.dtype) for w in w_list], wout.to(xs.dtype), b_list, act_mode, act_bias, operator_type)
    return isht(ys).to(dtype), residual
