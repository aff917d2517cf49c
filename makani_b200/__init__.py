"""makani_b200 -- B200-native (sm_100a) implementation of makani's spherical-harmonic hot path.

Public surface (mirrors the reference):
    RealSHT, InverseRealSHT                    <- torch_harmonics.{RealSHT, InverseRealSHT}
    SpectralConv, SpectralAttention, ComplexReLU <- makani.models.common.*
    quadrature                                 <- torch_harmonics.quadrature
    distributed                                <- torch_harmonics.distributed (h x w spatial model parallelism)
    install_torch_harmonics_shim()             <- makes `import torch_harmonics` resolve to this package
    HostFeed                                   <- double-buffered host->device input staging (the data loader's prefetch queue)
    sfno.SphericalFourierNeuralOperatorNet     <- makani.models.networks.sfnonet (same constructor / parameters / state dict)
    norm.InstanceNorm2d, norm.bias_gelu        <- torch.nn.InstanceNorm2d (+ GELU), bias + GELU on the library's kernels (row N2)
"""
from ._lib import B200ShtError, load as load_library  # noqa: F401
from . import quadrature  # noqa: F401
from .sht import RealSHT, InverseRealSHT, get_plan, resolve_precision  # noqa: F401
from .spectral_convolution import SpectralConv, SpectralAttention, ComplexReLU, mix_packed  # noqa: F401
from .host_pipeline import HostFeed  # noqa: F401

__version__ = "0.1.0"
