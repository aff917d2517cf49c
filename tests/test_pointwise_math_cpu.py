"""The arithmetic csrc/norm.cu implements, restated with torch tensor ops and checked against autograd through F.instance_norm / F.gelu (fp64):
pivoted statistics, y = gelu(xhat * gamma + beta), and the two-pass backward dx = rstd * gamma * (g - mean g - xhat * mean(g xhat)) with
g = dy * gelu'(z), dgamma = sum g xhat, dbeta = sum g.  (The kernels themselves are checked on the GPU, tests/test_gpu_pointwise.py.)"""
import math

import pytest
import torch
import torch.nn.functional as F


def _gelu_grad(z):
    return 0.5 * (1 + torch.erf(z / math.sqrt(2))) + z * torch.exp(-0.5 * z * z) / math.sqrt(2 * math.pi)


@pytest.mark.parametrize("gelu", [False, True])
def test_norm_kernel_formulas(gelu):
    torch.manual_seed(333)
    B, C, H, W = 2, 3, 5, 7
    n = H * W
    x = (torch.randn(B, C, H, W, dtype=torch.float64) * 1.7 + 40.0).requires_grad_(True)   # |mean| >> std: the case the pivot is for
    w = torch.randn(C, dtype=torch.float64, requires_grad=True)
    b = torch.randn(C, dtype=torch.float64, requires_grad=True)
    gy = torch.randn(B, C, H, W, dtype=torch.float64)
    y = F.instance_norm(x, weight=w, bias=b, eps=1e-6)
    if gelu:
        y = F.gelu(y)
    y.backward(gy)
    xr, dy = x.detach().reshape(B * C, n), gy.reshape(B * C, n)
    pivot = xr[:, :1]
    d = xr - pivot
    md = d.sum(1, keepdim=True) / n
    var = (d * d).sum(1, keepdim=True) / n - md * md
    mean, rstd = pivot + md, 1 / torch.sqrt(var + 1e-6)
    gam, bet = w.detach().repeat(B).view(-1, 1), b.detach().repeat(B).view(-1, 1)
    xh = (xr - mean) * rstd
    z = xh * gam + bet
    yk = 0.5 * z * (1 + torch.erf(z / math.sqrt(2))) if gelu else z
    g = dy * _gelu_grad(z) if gelu else dy
    s1, s2 = g.sum(1, keepdim=True), (g * xh).sum(1, keepdim=True)
    dx = rstd * gam * (g - s1 / n - xh * s2 / n)
    assert torch.allclose(yk.view_as(y), y.detach(), atol=1e-9)
    assert torch.allclose(dx.view_as(x), x.grad, atol=1e-9)
    assert torch.allclose(s2.view(B, C).sum(0), w.grad, atol=1e-9) and torch.allclose(s1.view(B, C).sum(0), b.grad, atol=1e-9)
    # in fp32 the pivot keeps the variance accurate where E[x^2] - E[x]^2 about zero cancels (x ~ 40 +- 1.7)
    x32 = xr.float()
    d32 = x32 - x32[:, :1]
    var_pivot = (d32 * d32).mean(1) - d32.mean(1) ** 2
    var_naive = (x32 * x32).mean(1) - x32.mean(1) ** 2
    ref = xr.var(dim=1, unbiased=False)
    assert (var_pivot.double() - ref).abs().max() < 1e-5 < (var_naive.double() - ref).abs().max() + 1e-4
