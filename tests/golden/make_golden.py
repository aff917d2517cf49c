"""Generate golden vectors from the REFERENCE's own pure-torch files (run in the build container where
/root/reference is mounted; the GPU box never reads /root/reference).

    python tests/golden/make_golden.py   ->  tests/golden/contractions_golden.npz, tests/golden/spectral_conv_golden.npz

Sources imported verbatim (by file path, because `import makani` needs packages absent here):
    /root/reference/makani/models/common/contractions.py          (_contract_*, compl_*mul*2d_fwd)
    /root/reference/makani/models/common/activations.py           (ComplexReLU)
    /root/reference/makani/models/common/spectral_convolution.py  (class SpectralConv: constructor + forward, autograd backward)
The SpectralConv golden runs the REFERENCE class (its own reshape / contraction / dtype / bias / residual code) on the oracle's
RealSHT / InverseRealSHT, because torch-harmonics itself is not installable here: it pins the block wiring element-wise, the
transforms stay pinned by invariants (oracle/makani_oracle.py header).
"""
import importlib.util
import os

import numpy as np
import torch

REF = "/root/reference/makani/models/common"


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    con = _load("ref_contractions", os.path.join(REF, "contractions.py"))
    act = _load("ref_activations", os.path.join(REF, "activations.py"))
    torch.manual_seed(333)  # the reference's test seed (tests/testutils.py:45-52)
    B, G, Ci, Co, L, M = 2, 2, 3, 5, 7, 8
    x = torch.randn(B, G, Ci, L, M, dtype=torch.complex64)
    out = {"x": x.numpy()}
    w_lw = torch.randn(G, Ci, Co, L, dtype=torch.complex64)
    w_lm = torch.randn(G, Ci, Co, L, M, dtype=torch.complex64)
    w_slw = torch.randn(G, Ci, L, dtype=torch.complex64)
    w_slm = torch.randn(G, Ci, L, M, dtype=torch.complex64)
    out["w_dhconv"], out["y_dhconv"] = w_lw.numpy(), con._contract_dense_pytorch(x, w_lw, separable=False, operator_type="dhconv").numpy()
    out["w_diagonal"], out["y_diagonal"] = w_lm.numpy(), con._contract_dense_pytorch(x, w_lm, separable=False, operator_type="diagonal").numpy()
    out["w_sep_dhconv"], out["y_sep_dhconv"] = w_slw.numpy(), con._contract_dense_pytorch(x, w_slw, separable=True, operator_type="dhconv").numpy()
    out["w_sep_diagonal"], out["y_sep_diagonal"] = w_slm.numpy(), con._contract_dense_pytorch(x, w_slm, separable=True, operator_type="diagonal").numpy()
    # ungrouped attention contractions
    xa = torch.randn(B, Ci, L, M, dtype=torch.complex64)
    ws = torch.randn(Ci, Co, dtype=torch.complex64)
    wl = torch.randn(L, Ci, Co, dtype=torch.complex64)
    cb = torch.randn(Co, 1, 1, dtype=torch.complex64)
    out["xa"] = xa.numpy()
    out["w_shared"], out["y_shared"] = ws.numpy(), con.compl_mul2d_fwd(xa, ws).numpy()
    out["w_ldep"], out["y_ldep"] = wl.numpy(), con.compl_exp_mul2d_fwd(xa, wl).numpy()
    out["cbias"] = cb.numpy()
    out["y_shared_bias"] = con.compl_muladd2d_fwd(xa, ws, cb).numpy()
    out["y_ldep_bias"] = con.compl_exp_muladd2d_fwd(xa, wl, cb).numpy()
    # ComplexReLU, all four modes (bias per channel for modulus / halfplane)
    z = torch.randn(B, Co, L, M, dtype=torch.complex64)
    out["z"] = z.numpy()
    for mode in ("real", "cartesian", "modulus", "halfplane"):
        m = act.ComplexReLU(negative_slope=0.1, mode=mode, bias_shape=(Co, 1, 1), scale=0.3)
        with torch.no_grad():
            if isinstance(m.bias, torch.Tensor):
                m.bias.copy_(torch.linspace(-0.5, 0.7, Co).reshape(Co, 1, 1))
                out[f"relu_bias_{mode}"] = m.bias.detach().numpy().copy()
            out[f"relu_{mode}"] = m(z).numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "contractions_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


def _load_reference_spectral_conv():
    """import /root/reference/makani/models/common/spectral_convolution.py with its package imports stubbed"""
    import sys
    import types

    con = _load("makani.models.common.contractions", os.path.join(REF, "contractions.py"))
    act = _load("ref_activations2", os.path.join(REF, "activations.py"))
    comm = types.ModuleType("makani.utils.comm")
    comm.get_rank = lambda name: 0
    comm.get_size = lambda name: 1
    utils = types.ModuleType("makani.utils")
    utils.comm = comm
    common = types.ModuleType("makani.models.common")
    common.ComplexReLU = act.ComplexReLU
    common.contractions = con
    models = types.ModuleType("makani.models")
    models.common = common
    makani = types.ModuleType("makani")
    makani.utils, makani.models = utils, models
    thd = types.ModuleType("torch_harmonics.distributed")

    class DistributedInverseRealSHT:   # only used in an isinstance() test by the reference
        pass

    thd.DistributedInverseRealSHT = DistributedInverseRealSHT
    th = types.ModuleType("torch_harmonics")
    th.distributed = thd
    stubs = {"makani": makani, "makani.utils": utils, "makani.utils.comm": comm, "makani.models": models, "makani.models.common": common,
             "makani.models.common.contractions": con, "torch_harmonics": th, "torch_harmonics.distributed": thd}
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        return _load("ref_spectral_convolution", os.path.join(REF, "spectral_convolution.py"))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


CONV_GOLDEN_CASES = {
    # name: (nlat_i, nlon_i, grid_i, nlat_o, nlon_o, grid_o, lmax, mmax, B, Cin, Cout, G, operator, separable, bias)
    "dhconv_same_grid_bias": (12, 24, "legendre-gauss", 12, 24, "legendre-gauss", 8, 9, 2, 4, 6, 1, "dhconv", False, True),
    "dhconv_regrid_groups": (13, 24, "equiangular", 10, 20, "legendre-gauss", 8, 9, 1, 4, 6, 2, "dhconv", False, True),
    "diagonal": (12, 24, "legendre-gauss", 12, 24, "legendre-gauss", 8, 8, 2, 3, 5, 1, "diagonal", False, False),
    "separable_dhconv": (12, 24, "legendre-gauss", 12, 24, "legendre-gauss", 8, 9, 1, 4, 4, 2, "dhconv", True, False),
}


def conv_golden():
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle import makani_oracle as O

    ref = _load_reference_spectral_conv()
    out = {}
    for name, (nlat_i, nlon_i, grid_i, nlat_o, nlon_o, grid_o, lmax, mmax, B, Cin, Cout, G, op, sep, bias) in CONV_GOLDEN_CASES.items():
        torch.manual_seed(333)
        sht = O.RealSHT(nlat_i, nlon_i, lmax, mmax, grid_i)
        isht = O.InverseRealSHT(nlat_o, nlon_o, lmax, mmax, grid_o)
        conv = ref.SpectralConv(sht, isht, Cin, Cout, num_groups=G, operator_type=op, separable=sep, bias=bias)
        if bias:
            with torch.no_grad():
                conv.bias.copy_(torch.randn_like(conv.bias))
        x = torch.randn(B, Cin, nlat_i, nlon_i, requires_grad=True)
        y, res = conv(x)
        gy = torch.randn_like(y)
        gres = torch.randn_like(res) if conv.scale_residual else None
        loss = (y * gy).sum() + ((res * gres).sum() if gres is not None else 0.0)
        loss.backward()
        out[f"{name}/x"] = x.detach().numpy()
        out[f"{name}/weight"] = conv.weight.detach().numpy()
        out[f"{name}/y"] = y.detach().numpy()
        out[f"{name}/gy"] = gy.numpy()
        out[f"{name}/dx"] = x.grad.numpy()
        out[f"{name}/dweight"] = conv.weight.grad.numpy()
        if bias:
            out[f"{name}/bias"] = conv.bias.detach().numpy()
            out[f"{name}/dbias"] = conv.bias.grad.numpy()
        if conv.scale_residual:
            out[f"{name}/residual"] = res.detach().numpy()
            out[f"{name}/gres"] = gres.numpy()
        # attributes the checkpoint / DDP code reads (spectral_convolution.py:195-203, 210-211)
        out[f"{name}/weight_shape"] = np.array(conv.weight.shape)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "spectral_conv_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")


def check_attention_raises():
    """the reference's SpectralAttention.forward cannot run (its einsums get 5-D operands); recorded here so that the
    "intended semantics, parity unpinned" status of makani_b200.SpectralAttention stays checkable."""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle import makani_oracle as O

    ref = _load_reference_spectral_conv()
    sht, isht = O.RealSHT(12, 24, 8, 8, "legendre-gauss"), O.InverseRealSHT(12, 24, 8, 8, "legendre-gauss")
    for op in ("diagonal", "l-dependant"):
        try:
            att = ref.SpectralAttention(sht, isht, 4, 4, operator_type=op, hidden_size_factor=2, complex_activation="real", spectral_layers=1)
            att(torch.randn(2, 4, 12, 24))
            print(f"reference SpectralAttention({op}) RUNS now: generate golden vectors for it")
        except Exception as e:  # noqa: BLE001
            print(f"reference SpectralAttention({op}) raises {type(e).__name__}: {str(e)[:120]}")


if __name__ == "__main__":
    main()
    conv_golden()
    check_attention_raises()
