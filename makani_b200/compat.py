"""`torch_harmonics` shim: lets makani's own Python (networks, trainers, configs) run unchanged on top of makani_b200.

makani imports the package by name at module import time (`import torch_harmonics as th`,
`import torch_harmonics.distributed as thd`, `from torch_harmonics.distributed.primitives import ...`:
/root/reference/makani/models/networks/sfnonet.py:31-32, models/common/spectral_convolution.py:34, mpu/mappings.py:19-25)
and checks class identity (`isinstance(..., thd.DistributedInverseRealSHT)`, spectral_convolution.py:169), so the shim must be
registered in `sys.modules` BEFORE `import makani`:

    import makani_b200.compat as compat
    compat.install_torch_harmonics_shim()      # torch_harmonics -> makani_b200
    compat.patch_makani_spectral_layers()      # makani.models.common.SpectralConv/SpectralAttention -> makani_b200 (optional)
    import makani
"""
import importlib
import sys
import types


def install_torch_harmonics_shim(force=False):
    """Register `torch_harmonics`, `.quadrature`, `.distributed`, `.distributed.primitives` backed by makani_b200."""
    if "torch_harmonics" in sys.modules and not force:
        mod = sys.modules["torch_harmonics"]
        if getattr(mod, "__b200_shim__", False):
            return mod
        raise RuntimeError("a real torch_harmonics is already imported; pass force=True to replace it")
    import makani_b200 as mb
    from makani_b200 import distributed as mbd
    from makani_b200 import quadrature as mbq

    th = types.ModuleType("torch_harmonics")
    th.__b200_shim__ = True
    th.__version__ = "0.9.0+b200"
    th.RealSHT = mb.RealSHT
    th.InverseRealSHT = mb.InverseRealSHT
    th.quadrature = mbq
    th.distributed = mbd
    th.__path__ = []  # mark as package so that submodule imports resolve through sys.modules
    sys.modules["torch_harmonics"] = th
    sys.modules["torch_harmonics.quadrature"] = mbq
    sys.modules["torch_harmonics.distributed"] = mbd
    sys.modules["torch_harmonics.distributed.primitives"] = mbd.primitives
    sys.modules["torch_harmonics.distributed.utils"] = mbd
    return th


def patch_makani_spectral_layers():
    """After `import makani`: point makani.models.common.{SpectralConv, SpectralAttention, ComplexReLU} at the CUDA-backed classes."""
    import makani_b200 as mb

    common = importlib.import_module("makani.models.common")
    for name in ("SpectralConv", "SpectralAttention", "ComplexReLU"):
        setattr(common, name, getattr(mb, name))
    sc = sys.modules.get("makani.models.common.spectral_convolution")
    if sc is not None:
        sc.SpectralConv = mb.SpectralConv
        sc.SpectralAttention = mb.SpectralAttention
    for modname in ("makani.models.networks.sfnonet", "makani.models.networks.fourcastnet3", "makani.models.networks.fourcastnet3_1", "makani.models.networks.snonet"):
        m = sys.modules.get(modname)
        if m is not None:
            for name in ("SpectralConv", "SpectralAttention"):
                if hasattr(m, name):
                    setattr(m, name, getattr(mb, name))
