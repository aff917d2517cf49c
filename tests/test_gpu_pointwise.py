"""GPU parity of the pointwise tail of the SFNO block (csrc/norm.cu) against PyTorch's own operators in fp64 on the CPU:
`makani_b200.norm.InstanceNorm2d` (= torch.nn.InstanceNorm2d(eps=1e-6, affine) as built at makani/models/networks/sfnonet.py:618-620), alone and fused
with the GELU that follows it, and `bias_gelu` (bias + GELU of the 1x1-convolution stacks, makani/models/common/layers.py:537-760); forward, input gradient
and parameter gradients.  fp32: rtol 1e-5; bf16: one output rounding (4e-3 element bound)."""
import pytest
import torch
import torch.nn.functional as F

from makani_b200 import norm as mnorm
from makani_b200.sfno import Conv1x1
from test_gpu_parity import close

pytestmark = pytest.mark.gpu
DEV = "cuda"

#          B  C    H    W
SHAPES = [(2, 5, 7, 9),         # odd row length: scalar path, several rows per channel index
          (1, 16, 24, 48),      # vector path, one split
          (3, 8, 64, 130),      # vector path (8320 elements), several splits
          (1, 384, 240, 480)]   # the interior SFNO block


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 8e-3)])
@pytest.mark.parametrize("gelu", [False, True])
def test_instance_norm_matches_torch(shape, dtype, tol, gelu):
    torch.manual_seed(333)
    B, C, H, W = shape
    m = mnorm.InstanceNorm2d(C, eps=1e-6, affine=True, track_running_stats=False).to(DEV)
    with torch.no_grad():
        m.weight.copy_(torch.randn(C) * 0.5 + 1.0)
        m.bias.copy_(torch.randn(C) * 0.3)
    x = (torch.randn(B, C, H, W) * 1.7 + 0.4).to(dtype)   # non-zero mean: exercises the pivoted variance
    gy = torch.randn(B, C, H, W).to(dtype)
    xd = x.to(DEV).requires_grad_(True)
    y = m(xd, gelu=gelu)
    assert y.dtype == dtype and y.shape == x.shape
    y.backward(gy.to(DEV))
    xr = x.double().requires_grad_(True)
    wr = m.weight.detach().cpu().double().requires_grad_(True)
    br = m.bias.detach().cpu().double().requires_grad_(True)
    yr = F.instance_norm(xr, weight=wr, bias=br, eps=1e-6)
    if gelu:
        yr = F.gelu(yr)
    yr.backward(gy.double())
    tag = f"InstanceNorm2d{'+GELU' if gelu else ''} {shape} {dtype}"
    close(y, yr, tol, tag + " y")
    close(xd.grad, xr.grad, tol, tag + " dx")
    close(m.weight.grad, wr.grad, max(tol, 2e-5), tag + " dweight")
    close(m.bias.grad, br.grad, max(tol, 2e-5), tag + " dbias")


def test_instance_norm_without_affine_and_fallbacks():
    torch.manual_seed(333)
    x = torch.randn(2, 6, 16, 24)
    m = mnorm.InstanceNorm2d(6, eps=1e-5, affine=False).to(DEV)
    y = m(x.to(DEV))
    close(y, F.instance_norm(x.double(), eps=1e-5), 1e-5, "InstanceNorm2d(affine=False)")
    # CPU tensors and the switch take torch's operator (same numbers)
    assert torch.allclose(mnorm.InstanceNorm2d(6)(x), F.instance_norm(x), atol=1e-6)
    old = mnorm.set_fused_pointwise(False)
    try:
        y2 = m(x.to(DEV))
    finally:
        mnorm.set_fused_pointwise(old)
    close(y2, y, 1e-5, "fused vs torch operator on the GPU")


@pytest.mark.parametrize("shape", SHAPES[:3] + [(1, 768, 120, 200)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 8e-3)])
def test_bias_gelu_matches_torch(shape, dtype, tol):
    torch.manual_seed(333)
    B, C, H, W = shape
    x = (torch.randn(B, C, H, W) * 1.5).to(dtype)
    b = torch.randn(C)
    gy = torch.randn(B, C, H, W).to(dtype)
    xd = x.to(DEV).requires_grad_(True)
    bd = b.to(DEV).requires_grad_(True)
    y = mnorm.bias_gelu(xd, bd)
    y.backward(gy.to(DEV))
    xr = x.double().requires_grad_(True)
    br = b.double().requires_grad_(True)
    yr = F.gelu(xr + br.view(1, -1, 1, 1))
    yr.backward(gy.double())
    close(y, yr, tol, f"bias_gelu {shape} {dtype} y")
    close(xd.grad, xr.grad, tol, f"bias_gelu {shape} {dtype} dx")
    close(bd.grad, br.grad, max(tol, 2e-5), f"bias_gelu {shape} {dtype} dbias")


def test_conv1x1_is_the_reference_convolution():
    """Conv1x1 = nn.Conv2d(cin, cout, 1) computed as a GEMM on the NCHW tensor: same output and gradients as F.conv2d"""
    torch.manual_seed(333)
    conv = Conv1x1(12, 20, 1, bias=True).to(DEV)
    x = torch.randn(2, 12, 9, 14, device=DEV, requires_grad=True)
    y = conv(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    xr = x.detach().double().cpu().requires_grad_(True)
    wr = conv.weight.detach().double().cpu().requires_grad_(True)
    br = conv.bias.detach().double().cpu().requires_grad_(True)
    yr = F.conv2d(xr, wr, br)
    yr.backward(gy.double().cpu())
    close(y, yr, 1e-5, "Conv1x1 y")
    close(x.grad, xr.grad, 1e-5, "Conv1x1 dx")
    close(conv.weight.grad, wr.grad, 1e-5, "Conv1x1 dweight")
    close(conv.bias.grad, br.grad, 1e-5, "Conv1x1 dbias")
