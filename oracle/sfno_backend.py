"""Oracle-backed transform / filter classes for makani_b200.sfno's `backend` argument -- TEST INFRASTRUCTURE (see makani_oracle.py header).

Builds the SFNO network of makani_b200/sfno.py on the CPU oracle instead of the CUDA kernels: used by the CPU tests (the network
restatement against golden vectors from the reference's own network class) and by bench.py's reference arm for the full-model workload.
"""
import math

import torch
import torch.nn as nn

from . import makani_oracle as O


class OracleSpectralConv(nn.Module):
    """SpectralConv on the oracle: same parameter names / shapes / init as makani/models/common/spectral_convolution.py:116-211."""

    def __init__(self, forward_transform, inverse_transform, in_channels, out_channels, num_groups=1, operator_type="dhconv", separable=False, bias=False,
                 gain=1.0):
        super().__init__()
        self.forward_transform, self.inverse_transform = forward_transform, inverse_transform
        self.num_groups, self.operator_type, self.separable = num_groups, operator_type, separable
        L, M = forward_transform.lmax, forward_transform.mmax
        shape = [num_groups, in_channels // num_groups]
        if not separable:
            shape += [out_channels // num_groups]
        shape += [L] if operator_type == "dhconv" else [L, M]
        scale = math.sqrt(gain / (in_channels // num_groups)) * torch.ones(L, dtype=torch.complex64)
        scale[0] *= math.sqrt(2.0)
        init = scale * torch.randn(*shape, dtype=torch.complex64) if operator_type == "dhconv" else scale.reshape(-1, 1) * torch.randn(*shape, dtype=torch.complex64)
        self.weight = nn.Parameter(init)
        if bias:
            self.bias = nn.Parameter(torch.zeros(1, out_channels, 1, 1))

    def forward(self, x):
        bias = self.bias if hasattr(self, "bias") else None
        with torch.autocast(device_type=x.device.type, enabled=False):     # transforms in fp32, as spectral_convolution.py:237-253
            y, res = self._fwd(x.float(), bias)
        return y.to(x.dtype), res.to(x.dtype)

    def _fwd(self, x, bias):
        return O.spectral_conv_forward(x, self.weight, self.forward_transform, self.inverse_transform, num_groups=self.num_groups,
                                       operator_type=self.operator_type, separable=self.separable, bias=bias)


class OracleBackend:
    def __init__(self, dtype=torch.float32):
        self.dtype = dtype

    def RealSHT(self, nlat, nlon, lmax=None, mmax=None, grid="equiangular"):
        return O.RealSHT(nlat, nlon, lmax, mmax, grid, dtype=self.dtype)

    def InverseRealSHT(self, nlat, nlon, lmax=None, mmax=None, grid="equiangular"):
        return O.InverseRealSHT(nlat, nlon, lmax, mmax, grid, dtype=self.dtype)

    SpectralConv = OracleSpectralConv

    def SpectralAttention(self, *a, **k):
        raise NotImplementedError("the reference's SpectralAttention.forward raises (SURVEY F3): no oracle network for it")
