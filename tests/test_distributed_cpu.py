"""World-size 2 / 4 gloo tests (CPU) of the h x w spatial-parallel path: primitives, the all-to-all choreography of
DistributedRealSHT / DistributedInverseRealSHT and its autograd, against the SERIAL oracle.  The local stages are supplied by the
oracle here (the CUDA kernels are covered by the -m gpu tests); pattern of /root/reference/tests/distributed/*."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import makani_b200.distributed as mbd
from oracle import makani_oracle as O


class OracleLocalOps:
    """CPU stand-in for the four local stages (restated torch-harmonics arithmetic on this rank's shard)."""

    def __init__(self, t):
        self.t = t
        theta, w = O.precompute_latitudes(t.nlat, t.grid)
        P = O.legpoly(t.mmax, t.lmax, np.cos(theta), csphase=t.csphase)
        self.P = torch.from_numpy(P[t.m_offset : t.m_offset + t.mmax_local]).double()
        self.w_local = torch.from_numpy(w[t.lat_offset : t.lat_offset + t.nlat_local]).double()

    def fft(self, x):
        X = 2.0 * math.pi * torch.fft.rfft(x.double(), dim=-1, norm="forward")[..., : self.t.mmax]
        return X * self.w_local[:, None]

    def legendre(self, xc):
        return torch.einsum("...km,mlk->...lm", xc, self.P.to(xc.dtype))

    def ilegendre(self, xc):
        return torch.einsum("...lm,mlk->...km", xc.to(torch.complex128), self.P.to(torch.complex128))

    def ifft(self, xc, dtype):
        re, im = xc.real, xc.imag.clone()
        im[..., 0] = 0.0
        return torch.fft.irfft(torch.complex(re, im), n=self.t.nlon, dim=-1, norm="forward")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, h, w, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        # grid: rank = ih * w + iw  (w fastest, like makani's comm tree)
        h_groups = [dist.new_group([ih * w + iw for ih in range(h)]) for iw in range(w)]
        w_groups = [dist.new_group([ih * w + iw for iw in range(w)]) for ih in range(h)]
        ih, iw = rank // w, rank % w
        mbd.init(h_groups[iw] if h > 1 else None, w_groups[ih] if w > 1 else None)
        mbd.set_local_ops(OracleLocalOps)
        torch.manual_seed(333)
        nlat, nlon, lmax, mmax, B, C = 33, 64, 19, 21, 2, 6
        results = {}
        for grid in ("equiangular", "legendre-gauss"):
            dsht = mbd.DistributedRealSHT(nlat, nlon, lmax, mmax, grid)
            disht = mbd.DistributedInverseRealSHT(nlat, nlon, lmax, mmax, grid)
            assert dsht.lat_shapes == O.compute_split_shapes(nlat, h) and dsht.m_shapes == O.compute_split_shapes(mmax, w)
            osht = O.RealSHT(nlat, nlon, lmax, mmax, grid, dtype=torch.float64)
            oisht = O.InverseRealSHT(nlat, nlon, lmax, mmax, grid, dtype=torch.float64)
            x = torch.randn(B, C, nlat, nlon, dtype=torch.float64)
            gc = torch.randn(B, C, lmax, mmax, dtype=torch.complex128)

            def shard(t, hd, wd, hs, ws):
                t = torch.split(t, hs, dim=hd)[ih]
                return torch.split(t, ws, dim=wd)[iw].contiguous()

            xl = shard(x, -2, -1, dsht.lat_shapes, dsht.lon_shapes).requires_grad_(True)
            cl = dsht(xl)
            xs = x.clone().requires_grad_(True)
            cs = osht(xs)
            results[f"{grid}/sht"] = (cl - shard(cs, -2, -1, dsht.l_shapes, dsht.m_shapes)).abs().max().item()
            cl.backward(shard(gc, -2, -1, dsht.l_shapes, dsht.m_shapes))
            cs.backward(gc)
            results[f"{grid}/sht_grad"] = (xl.grad - shard(xs.grad, -2, -1, dsht.lat_shapes, dsht.lon_shapes)).abs().max().item()
            # inverse
            c = torch.randn(B, C, lmax, mmax, dtype=torch.complex128)
            gy = torch.randn(B, C, nlat, nlon, dtype=torch.float64)
            cl = shard(c, -2, -1, disht.l_shapes, disht.m_shapes).to(torch.complex64).to(torch.complex128).requires_grad_(True)
            c = c.to(torch.complex64).to(torch.complex128)
            yl = disht(cl, dtype=torch.float64)
            cs = c.clone().requires_grad_(True)
            ys = oisht(cs)
            results[f"{grid}/isht"] = (yl - shard(ys, -2, -1, disht.lat_shapes, disht.lon_shapes)).abs().max().item()
            yl.backward(shard(gy, -2, -1, disht.lat_shapes, disht.lon_shapes))
            ys.backward(gy)
            results[f"{grid}/isht_grad"] = (cl.grad - shard(cs.grad, -2, -1, disht.l_shapes, disht.m_shapes)).abs().max().item()
        # primitives
        t = torch.arange(world * 3, dtype=torch.float32).reshape(world, 3) + 100 * rank
        g = mbd._gather(t[rank : rank + 1], 0, [1] * world, group=None)
        results["gather"] = float((g[:, 0] - torch.tensor([100.0 * r + 3 * r for r in range(world)])).abs().max())
        s = mbd._split(torch.arange(10.0), 0, group=None)
        results["split"] = float(abs(s.numel() - O.compute_split_shapes(10, world)[rank]))
        r = mbd._reduce(torch.ones(4) * (rank + 1), group=None)
        results["reduce"] = float((r - world * (world + 1) / 2).abs().max())
        q.put((rank, results, None))
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, None, traceback.format_exc()))


@pytest.mark.parametrize("h,w", [(2, 1), (1, 2), (2, 2), (4, 2)])   # (4, 2): the grid of BASELINE configs[3] on 8 ranks, uneven splits in every dimension
def test_distributed_sht_matches_serial_oracle(h, w):
    world = h * w
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, h, w, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, res, err in out:
        assert err is None, f"rank {rank}:\n{err}"
        for k, v in res.items():
            assert v < 1e-9, (rank, k, v)


def test_split_shapes_match_reference_semantics():
    assert mbd.compute_split_shapes(721, 4) == [181, 181, 181, 178]
    assert mbd.compute_split_shapes(241, 2) == [121, 120]
    assert mbd.compute_split_shapes(5, 4) == O.compute_split_shapes(5, 4)
    for n, p in ((10, 3), (7, 7), (240, 4), (1440, 2), (9, 4)):
        assert mbd.compute_split_shapes(n, p) == O.compute_split_shapes(n, p) and sum(mbd.compute_split_shapes(n, p)) == n


def test_torch_harmonics_shim_surface():
    import makani_b200 as mb
    import makani_b200.compat as compat

    compat.install_torch_harmonics_shim()
    import torch_harmonics as th
    import torch_harmonics.distributed as thd
    from torch_harmonics.distributed import compute_split_shapes  # noqa: F401  (makani/mpu/mappings.py:19)
    from torch_harmonics.distributed.primitives import _gather, _reduce, _split, _transpose  # noqa: F401  (mappings.py:20-25)
    from torch_harmonics.quadrature import clenshaw_curtiss_weights, legendre_gauss_weights, precompute_latitudes  # noqa: F401

    assert th.RealSHT is mb.RealSHT and th.InverseRealSHT is mb.InverseRealSHT
    for name in ("init", "is_initialized", "DistributedRealSHT", "DistributedInverseRealSHT", "distributed_transpose_azimuth",
                 "distributed_transpose_polar", "split_tensor_along_dim", "compute_split_shapes"):
        assert hasattr(thd, name), name
    t = thd.DistributedInverseRealSHT(32, 64, 16, 17)
    assert isinstance(t, thd.DistributedInverseRealSHT) and t.l_shapes == [16] and t.lat_shapes == [32]  # spectral_convolution.py:169-173
