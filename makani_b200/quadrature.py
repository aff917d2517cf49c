"""Quadrature rules of the SHT grids -- host-side mirror of `torch_harmonics.quadrature`
(used by the reference at /root/reference/makani/utils/grids.py:67-68,120-129,225).  Returns torch tensors (float64).
"""
import numpy as np
import torch


def _to_t(*arrs):
    return tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in arrs)


def _legendre_gauss_np(n, a=-1.0, b=1.0):
    x, w = np.polynomial.legendre.leggauss(n)
    return (b - a) * 0.5 * x + (b + a) * 0.5, w * (b - a) * 0.5


def _clenshaw_curtiss_np(n, a=-1.0, b=1.0):
    """Clenshaw-Curtis rule on cos(linspace(pi, 0, n)) via the DCT-I closed form of the weights."""
    if n < 2:
        raise ValueError("clenshaw_curtiss_weights needs n >= 2")
    t = np.cos(np.linspace(np.pi, 0.0, n))
    if n == 2:
        w = np.array([1.0, 1.0])
    else:
        n1 = n - 1
        j = np.arange(1, n1 // 2 + 1, dtype=np.float64)
        coef = np.where(2 * j == n1, 1.0, 2.0) / (4.0 * j * j - 1.0)
        k = np.arange(n, dtype=np.float64)
        w = 1.0 - (coef[None, :] * np.cos(2.0 * np.pi * np.outer(k, j) / n1)).sum(axis=1)
        c = np.full(n, 2.0)
        c[0] = c[-1] = 1.0
        w = c * w / n1
    return (b - a) * 0.5 * t + (b + a) * 0.5, w * (b - a) * 0.5


def legendre_gauss_weights(n, a=-1.0, b=1.0):
    return _to_t(*_legendre_gauss_np(n, a, b))


def clenshaw_curtiss_weights(n, a=-1.0, b=1.0):
    return _to_t(*_clenshaw_curtiss_np(n, a, b))


def _grid_np(nlat, grid):
    """cos(colatitude) in row order (row 0 = north, theta ascending) and the matching weights."""
    if grid == "legendre-gauss":
        cost, w = _legendre_gauss_np(nlat)
    elif grid == "equiangular":
        cost, w = _clenshaw_curtiss_np(nlat)
    else:
        raise ValueError(f"Unknown quadrature mode {grid}")
    return np.ascontiguousarray(cost[::-1]), np.ascontiguousarray(w[::-1])


def precompute_latitudes(nlat, grid="equiangular"):
    """Colatitudes (ascending) and quadrature weights, as torch_harmonics.quadrature.precompute_latitudes."""
    cost, w = _grid_np(nlat, grid)
    return _to_t(np.arccos(np.clip(cost, -1.0, 1.0)), w)


def precompute_longitudes(nlon):
    return torch.linspace(0, 2 * np.pi, nlon + 1, dtype=torch.float64)[:-1]
