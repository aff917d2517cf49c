"""Pointwise tail of the SFNO block on the CUDA library: instance norm (+ GELU) and bias + GELU as single-pass kernels.

`InstanceNorm2d` is `torch.nn.InstanceNorm2d` (same constructor, parameter names `weight` / `bias`, state dict) as the reference builds it at
/root/reference/makani/models/networks/sfnonet.py:618-620 (`num_features=embed_dim, eps=1e-6, affine=True, track_running_stats=False`); on CUDA tensors
of dtype float32 / bfloat16 its forward runs `b200sht_instance_norm_forward` (csrc/norm.cu) and can fuse the GELU that follows it in
`NeuralOperatorBlock.forward` (sfnonet.py:387-392).  `bias_gelu(x, bias)` is the `+ bias -> GELU` of the 1x1-convolution stacks
(makani/models/common/layers.py:537-760).  Tensors on the CPU (the oracle-backend reference arm of bench.py, the CPU golden tests) and configurations the
kernels do not cover (running statistics, other dtypes) take torch's own operators -- these layers are outside the spherical-harmonic hot path, whose
no-fallback rule is unchanged.  `B200SHT_FUSED_POINTWISE=0` switches the kernels off.
"""
import ctypes
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

_VP = ctypes.c_void_p
_ENABLED = os.environ.get("B200SHT_FUSED_POINTWISE", "1") != "0"


def set_fused_pointwise(on):
    """switch the fused kernels on / off at run time (returns the previous setting)"""
    global _ENABLED
    old, _ENABLED = _ENABLED, bool(on)
    return old


def fused_pointwise_enabled():
    return _ENABLED


def _ptr(t):
    return _VP(t.data_ptr()) if t is not None else _VP(0)


def _dt(dtype):
    return _lib.BF16 if dtype == torch.bfloat16 else _lib.F32


def _usable(x):
    return _ENABLED and x.is_cuda and x.dim() == 4 and x.dtype in (torch.float32, torch.bfloat16) and x.shape[0] * x.shape[1] <= 65535 and x.numel() > 0


def _workspace(B, C, hw, device):
    n = int(_lib.load().b200sht_pointwise_workspace_floats(B, C, hw))
    return torch.empty(max(n, 2), dtype=torch.float32, device=device)


class _InstanceNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, gelu):
        x = x.contiguous()
        B, C, H, W = x.shape
        hw = H * W
        y = torch.empty_like(x)
        stats = torch.empty(B * C, 2, dtype=torch.float32, device=x.device)
        ws = _workspace(B, C, hw, x.device)
        w32 = weight.detach().to(torch.float32).contiguous() if weight is not None else None
        b32 = bias.detach().to(torch.float32).contiguous() if bias is not None else None
        _lib.call("b200sht_instance_norm_forward", _ptr(x), _ptr(y), _ptr(w32), _ptr(b32), _ptr(stats), _ptr(ws), _dt(x.dtype), B, C, hw, float(eps), int(gelu),
                  _lib.launch_stream(x.device))
        ctx.save_for_backward(x, w32, b32, stats)
        ctx.gelu, ctx.has_affine = int(gelu), (weight is not None, bias is not None)
        ctx.param_dtypes = (weight.dtype if weight is not None else None, bias.dtype if bias is not None else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w32, b32, stats = ctx.saved_tensors
        B, C, H, W = x.shape
        hw = H * W
        dy = dy.contiguous().to(x.dtype)
        dx = torch.empty_like(x)
        sums = torch.empty(B * C, 2, dtype=torch.float32, device=x.device)
        ws = _workspace(B, C, hw, x.device)
        _lib.call("b200sht_instance_norm_backward", _ptr(x), _ptr(dy), _ptr(dx), _ptr(w32), _ptr(b32), _ptr(stats), _ptr(sums), _ptr(ws), _dt(x.dtype), B, C, hw,
                  ctx.gelu, _lib.launch_stream(x.device))
        per_c = sums.view(B, C, 2).sum(dim=0)
        dw = per_c[:, 1].to(ctx.param_dtypes[0]) if (ctx.has_affine[0] and ctx.needs_input_grad[1]) else None
        db = per_c[:, 0].to(ctx.param_dtypes[1]) if (ctx.has_affine[1] and ctx.needs_input_grad[2]) else None
        return (dx if ctx.needs_input_grad[0] else None), dw, db, None, None


class _BiasGeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias):
        x = x.contiguous()
        B, C, H, W = x.shape
        y = torch.empty_like(x)
        b32 = bias.detach().to(torch.float32).contiguous() if bias is not None else None
        _lib.call("b200sht_bias_gelu_forward", _ptr(x), _ptr(b32), _ptr(y), _dt(x.dtype), B, C, H * W, _lib.launch_stream(x.device))
        ctx.save_for_backward(x, b32)
        ctx.bias_dtype = bias.dtype if bias is not None else None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, b32 = ctx.saved_tensors
        B, C, H, W = x.shape
        hw = H * W
        dy = dy.contiguous().to(x.dtype)
        dx = torch.empty_like(x)
        need_b = b32 is not None and ctx.needs_input_grad[1]
        sums = torch.empty(B * C, 2, dtype=torch.float32, device=x.device) if need_b else None
        ws = _workspace(B, C, hw, x.device)
        _lib.call("b200sht_bias_gelu_backward", _ptr(x), _ptr(b32), _ptr(dy), _ptr(dx), _ptr(sums), _ptr(ws), _dt(x.dtype), B, C, hw, _lib.launch_stream(x.device))
        db = sums.view(B, C, 2)[:, :, 0].sum(dim=0).to(ctx.bias_dtype) if need_b else None
        return dx, db


def bias_gelu(x, bias=None):
    """gelu(x + bias[None, :, None, None]) for x (B, C, H, W); exact (erf) GELU"""
    if _usable(x) and (bias is None or bias.is_cuda):
        return _BiasGeluFn.apply(x, bias)
    if bias is not None:
        x = x + bias.to(x.dtype).view(1, -1, 1, 1)
    return F.gelu(x)


class InstanceNorm2d(nn.InstanceNorm2d):
    """torch.nn.InstanceNorm2d whose CUDA forward / backward run on the library's kernels; `forward(x, gelu=True)` returns gelu(norm(x))."""

    def forward(self, x, gelu=False):
        if _usable(x) and not self.track_running_stats and (self.weight is None or self.weight.is_cuda):
            if x.shape[1] != self.num_features:
                raise ValueError(f"expected input with {self.num_features} channels, got {x.shape[1]}")
            return _InstanceNormFn.apply(x, self.weight, self.bias, self.eps, bool(gelu))
        y = super().forward(x)
        return F.gelu(y) if gelu else y
