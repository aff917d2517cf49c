"""Other callers of the transforms on the CUDA path:

* the spectral losses of makani (`SpectralBaseLoss` weights, /root/reference/makani/utils/losses/base_loss.py:345-409: m > 0 counted twice,
  1 / (4 pi); squared L2 and H1 = l (l + 1) L2 as `h1_loss.py`) restated here in a few lines on top of `makani_b200.RealSHT` at the full
  spherical band limit of the grid (mmax = nlon / 2 + 1, the Nyquist order included) -- value and gradient against the oracle, and
  Parseval against the quadrature of the grid (the identity the reference's tests/test_losses.py:434-510 checks);
* spherical noise as `makani/models/noise.py:537-575` draws it: random coefficients with a power-law spectrum through `InverseRealSHT`
  (a pure synthesis consumer: variance = sum of the spectrum / (4 pi), tests/test_noise.py:406-421);
* `SpectralAttention` with its spectral MLP on the tensor cores (precision tf32: OP_SHARED / OP_LDEP mixes on tcgen05, ComplexReLU on packed
  spectra) against the fp64 oracle of the intended semantics (the reference's own forward raises).
"""
import math

import numpy as np
import pytest
import torch

import makani_b200 as mb
from oracle import makani_oracle as O
from test_gpu_parity import close, oracle_pair

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _lm_weights(lmax, mmax):
    m = 2.0 * torch.ones(mmax, dtype=torch.float64)
    m[0] = 1.0
    return (torch.ones(lmax, dtype=torch.float64)[:, None] * m[None, :]) / (4.0 * math.pi)


def _spectral_losses(sht, prd, tar, lmw):
    """squared spectral L2 and H1 norms of prd - tar per (batch, channel), as base_loss.py:381-387 / h1_loss.py build them"""
    c = sht(prd - tar)
    p = (c.real.square() + c.imag.square()) * lmw.to(c.device, c.real.dtype)
    l = torch.arange(c.shape[-2], device=c.device, dtype=c.real.dtype)
    return p.sum(dim=(-1, -2)), (p * (l * (l + 1.0))[:, None]).sum(dim=(-1, -2))


@pytest.mark.parametrize("grid,nlat,nlon,precision,tol", [("equiangular", 91, 180, "fp32", 2e-5), ("legendre-gauss", 90, 180, "fp32", 2e-5),
                                                          ("equiangular", 181, 360, "tf32", 2e-3)])
def test_spectral_loss_value_and_gradient(grid, nlat, nlon, precision, tol):
    torch.manual_seed(333)
    lmax = mmax = min(nlat, nlon // 2 + 1)   # compute_spherical_bandlimit (utils/grids.py:43-54)
    sht = mb.RealSHT(nlat, nlon, lmax, mmax, grid, precision=precision)
    osht = O.RealSHT(nlat, nlon, lmax, mmax, grid, dtype=torch.float64)
    lmw = _lm_weights(lmax, mmax)
    B, C = 2, 3
    prd, tar = torch.randn(B, C, nlat, nlon), torch.randn(B, C, nlat, nlon)
    pd = prd.to(DEV).requires_grad_(True)
    l2, h1 = _spectral_losses(sht, pd, tar.to(DEV), lmw)
    pr = prd.double().requires_grad_(True)
    l2r, h1r = _spectral_losses(osht, pr, tar.double(), lmw)
    close(l2, l2r, tol, f"spectral L2 loss {grid} {precision}")
    close(h1, h1r, tol, f"spectral H1 loss {grid} {precision}")
    (l2.sum() + 1e-4 * h1.sum()).backward()
    (l2r.sum() + 1e-4 * h1r.sum()).backward()
    close(pd.grad, pr.grad, 5 * tol, f"spectral loss gradient {grid} {precision}")


def test_parseval_against_grid_quadrature():
    """a band-limited field: sum of lm-weighted |c|^2 == quadrature integral of x^2 / (4 pi) (tests/test_losses.py:470-499 on Y_4^0 + others)"""
    torch.manual_seed(333)
    nlat, nlon, L = 65, 128, 40
    isht = mb.InverseRealSHT(nlat, nlon, L, L, "equiangular", precision="fp32")
    sht = mb.RealSHT(nlat, nlon, L, L, "equiangular", precision="fp32")
    c = torch.randn(1, 4, L, L, dtype=torch.complex64).tril()
    c[..., 0] = c[..., 0].real.to(torch.complex64)
    x = isht(c.to(DEV))
    spec = _spectral_losses(sht, x, torch.zeros_like(x), _lm_weights(L, L))[0].double().cpu()
    _, w = O.precompute_latitudes(nlat, "equiangular")
    quad = (x.double().cpu().square().mean(dim=-1) * torch.from_numpy(w)).sum(dim=-1) * (2 * math.pi) / (4 * math.pi)
    assert torch.allclose(spec, quad, rtol=2e-5), (spec, quad)


def test_spherical_noise_synthesis_variance():
    """isotropic Gaussian noise through InverseRealSHT: the area mean of x^2 equals sum_l (2 l + 1) sigma_l^2 / (4 pi) in expectation; against
    the oracle the field itself must agree element-wise (same coefficients)"""
    torch.manual_seed(333)
    nlat, nlon, L = 90, 180, 60
    isht = mb.InverseRealSHT(nlat, nlon, L, L + 1, "legendre-gauss", precision="fp32")
    oi = O.InverseRealSHT(nlat, nlon, L, L + 1, "legendre-gauss", dtype=torch.float64)
    l = torch.arange(L, dtype=torch.float64)
    sigma = (1.0 + l) ** -1.5
    c = (torch.randn(64, 1, L, L + 1, dtype=torch.complex128) * sigma[:, None]).tril()   # E |c_lm|^2 = sigma_l^2
    c[..., 0] = c[..., 0].real * math.sqrt(2.0)
    x = isht(c.to(torch.complex64).to(DEV))
    close(x, oi(c), 2e-5, "noise synthesis")
    _, w = O.precompute_latitudes(nlat, "legendre-gauss")
    var = ((x.double().cpu().square().mean(dim=-1) * torch.from_numpy(w)).sum(dim=-1) * (2 * math.pi) / (4 * math.pi)).mean().item()
    expect = float(((2 * l + 1) * sigma.square()).sum() / (4 * math.pi))
    assert abs(var - expect) / expect < 0.1, (var, expect)   # 64 samples: a few per cent of sampling noise


@pytest.mark.parametrize("op,act", [("diagonal", "real"), ("l-dependant", "cartesian")])
def test_spectral_attention_tf32_tensor_core_mlp(op, act):
    torch.manual_seed(333)
    nlat, nlon, L, M, B, C, Co = 64, 128, 32, 33, 2, 16, 12
    f = mb.RealSHT(nlat, nlon, L, M, "legendre-gauss", precision="tf32")
    i = mb.InverseRealSHT(nlat, nlon, L, M, "legendre-gauss", precision="tf32")
    att = mb.SpectralAttention(f, i, C, Co, operator_type=op, hidden_size_factor=2, complex_activation=act, bias=True, spectral_layers=2, precision="tf32").to(DEV)
    of, oi = oracle_pair(nlat, nlon, nlat, nlon, L, M, "legendre-gauss", "legendre-gauss")
    x = torch.randn(B, C, nlat, nlon)
    xd = x.to(DEV).requires_grad_(True)
    y, _ = att(xd)
    ws = [w.detach().cpu().to(torch.complex128).requires_grad_(True) for w in att.w]
    wo = att.wout.detach().cpu().to(torch.complex128).requires_grad_(True)
    bs = [b.detach().cpu().to(torch.complex128).requires_grad_(True) for b in att.b]
    ab = [a.bias.detach().cpu().double() if isinstance(a.bias, torch.Tensor) else 0.0 for a in att.activations]
    xr = x.double().requires_grad_(True)
    yr, _ = O.spectral_attention_forward(xr, ws, wo, of, oi, b_list=bs, act_mode=act, act_bias=ab, operator_type=op)
    rel = close(y, yr, 4e-3, f"SpectralAttention[{op},{act}] tf32 y")
    assert rel < 2e-3, rel
    gy = torch.randn_like(yr)
    y.backward(gy.float().to(DEV))
    yr.backward(gy)
    # wout sits after the last activation: its gradient is a plain TF32 contraction of the hidden spectrum with gy
    rel = close(att.wout.grad, wo.grad, 6e-3, f"SpectralAttention[{op},{act}] tf32 dwout")
    assert rel < 3e-3, rel
    # everything upstream passes through the ReLU gates, whose derivative is a step function: a pre-activation within the TF32 error
    # (~5e-4 relative) of zero flips its gate and changes that element's gradient by O(1), so the L2 error of these gradients scales like
    # sqrt(fraction of flipped gates) ~ 1e-2, not like the TF32 error itself (measured: 1.6e-3 'real', 1.4e-2 'cartesian', where both the
    # real and the imaginary gate of every mode can flip).  The fp32 test (test_gpu_parity.py) pins the same code path at 2e-5.
    for name, a, b in (("dx", xd.grad, xr.grad), ("dw0", att.w[0].grad, ws[0].grad), ("db1", att.b[1].grad, bs[1].grad)):
        a, b = a.detach().cpu().to(b.dtype), b.detach()
        rel = float((a - b).abs().pow(2).sum().sqrt() / b.abs().pow(2).sum().sqrt())
        print(f"[parity] SpectralAttention[{op},{act}] tf32 {name} (through ReLU gates): rel_l2={rel:.3e}")
        assert rel < 5e-2, (name, rel)
