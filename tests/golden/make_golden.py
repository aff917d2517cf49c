"""Generate golden vectors from the REFERENCE's own pure-torch files (run in the build container where
/root/reference is mounted; the GPU box never reads /root/reference).

    python tests/golden/make_golden.py   ->  tests/golden/contractions_golden.npz

Sources imported verbatim (by file path, because `import makani` needs packages absent here):
    /root/reference/makani/models/common/contractions.py   (_contract_*, compl_*mul*2d_fwd)
    /root/reference/makani/models/common/activations.py    (ComplexReLU)
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


if __name__ == "__main__":
    main()
