#!/usr/bin/env python
"""Construct the REFERENCE's SphericalFourierNeuralOperatorNet (/root/reference/makani/models/networks/sfnonet.py, unmodified) twice:

  (a) on the oracle posed as torch_harmonics, with makani's own SpectralConv          -> runs a CPU forward (reference behaviour)
  (b) on makani_b200 installed as torch_harmonics (compat.install_torch_harmonics_shim) with makani's SpectralConv / SpectralAttention
      replaced by makani_b200's (compat.patch_makani_spectral_layers)                 -> construction only on CPU (kernels need a GPU)

and compare what a checkpoint sees: parameter names, shapes, dtypes and the model-parallel tags (`is_shared_mp`, `sharded_dims_mp`)
must be identical, buffers must be absent from the state dict in both (torch-harmonics registers its tables non-persistently).
This is the drop-in claim of SURVEY rows A8/A9: `_init_spectral_transforms` (sfnonet.py:765-838) and `NeuralOperatorBlock`
(:275-286 read `.nlat .nlon .lat_shapes .lon_shapes ...`) work against the makani_b200 classes unchanged.

    python tests/reference_suites/build_reference_sfno.py [a|b]     (each variant needs its own process: both replace sys.modules)
"""
import json
import os
import sys
import types
from dataclasses import dataclass

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import run_reference_tests as R  # noqa: E402

CFG = dict(inp_shape=(33, 64), out_shape=(33, 64), inp_chans=5, out_chans=3, embed_dim=8, num_layers=3, scale_factor=2,
           model_grid_type="equiangular", sht_grid_type="legendre-gauss", bias=True)


# the non-linear spectral filter (SpectralAttention): the reference constructs it (only its forward raises)
CFG_NONLINEAR = dict(CFG, filter_type="non-linear", operator_type="diagonal", num_layers=2)


def stub_physicsnemo():
    pn = types.ModuleType("physicsnemo")

    @dataclass
    class ModelMetaData:
        name: str = ""
        jit: bool = False
        cuda_graphs: bool = False
        amp_cpu: bool = False
        amp_gpu: bool = False

    class Module(torch.nn.Module):
        @classmethod
        def from_torch(cls, torch_model_class, meta=None, name=None, register=False):
            return torch_model_class

    pn.ModelMetaData, pn.Module = ModelMetaData, Module
    core = types.ModuleType("physicsnemo.core")

    class ModelRegistry:
        def register(self, *a, **k):
            pass

    core.ModelRegistry = ModelRegistry
    sys.modules.update({"physicsnemo": pn, "physicsnemo.core": core})


def describe(net):
    out = {}
    for name, p in net.named_parameters():
        out[name] = [list(p.shape), str(p.dtype), getattr(p, "is_shared_mp", None), getattr(p, "sharded_dims_mp", None)]
    return {"params": out, "state_dict_keys": sorted(net.state_dict().keys())}


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "a"
    R.install_environment()          # oracle as torch_harmonics + namespace stubs for the reference tree
    stub_physicsnemo()
    if which == "b":
        import makani_b200.compat as compat

        compat.install_torch_harmonics_shim(force=True)      # torch_harmonics -> makani_b200 (CUDA-backed classes)
        compat.patch_makani_spectral_layers()                # makani.models.common.SpectralConv / SpectralAttention -> makani_b200
    torch.manual_seed(333)
    from makani.models.networks import sfnonet

    cfg = CFG_NONLINEAR if (len(sys.argv) > 2 and sys.argv[2] == "nonlinear") else CFG
    net = sfnonet.SphericalFourierNeuralOperatorNet(**cfg)
    info = describe(net)
    info["spectral_classes"] = sorted({type(m).__module__ + "." + type(m).__name__ for m in net.modules()
                                       if type(m).__name__ in ("SpectralConv", "SpectralAttention", "RealSHT", "InverseRealSHT")})
    if which == "a" and cfg is CFG:
        y = net(torch.randn(1, CFG["inp_chans"], *CFG["inp_shape"]))
        info["forward_shape"] = list(y.shape)
    print(json.dumps(info))


if __name__ == "__main__":
    main()
