#!/usr/bin/env python
"""Run the REFERENCE's distributed SpectralConv test (/root/reference/tests/distributed/tests_distributed_layers.py:69-223,
`TestDistributedLayers.test_distributed_spectral_conv`, six shape cases incl. odd nlat 181 / 91 and up/down-sampling) unmodified on
CPU / gloo against the h x w choreography of makani_b200.distributed.

What is real: the reference's test body, its split / gather helpers, its SpectralConv class, its DDP gradient-reduction hooks
(makani/mpu/mappings.py:init_gradient_reduction_hooks), torch.distributed over gloo, and the all-to-all transposes + autograd of
makani_b200.distributed.  What stands in: the local (per-rank) FFT / Legendre stages are the oracle's arithmetic (the CUDA kernels
need a GPU; they are checked against the same oracle by the -m gpu tests and bit-identical to the single-GPU modules in
scripts/dist_gpu_check.py), `torch_harmonics.RealSHT/InverseRealSHT` are the oracle, `makani.utils.comm` is a small functional
implementation over torch.distributed process groups (the real one needs physicsnemo).

    python tests/reference_suites/run_reference_distributed.py [H W]        (default 2 1; spawns H*W gloo ranks)
"""
import os
import socket
import sys
import types
import unittest

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def make_comm(h, w):
    """functional stand-in for makani.utils.comm on a pure h x w model-parallel grid (rank = ih * w + iw)"""
    comm = types.ModuleType("makani.utils.comm")
    state = {"groups": {}, "sizes": {}, "ranks": {}}

    def init(model_parallel_sizes=None, model_parallel_names=None, data_parallel_sizes=None, data_parallel_names=None, **kw):
        rank, world = dist.get_rank(), dist.get_world_size()
        assert world == h * w
        ih, iw = rank // w, rank % w

        def add(name, members_of_rank):
            # every rank must create every group in the same order
            mine = None
            for members in members_of_rank:
                g = dist.new_group(members)
                if rank in members:
                    mine = (g, members)
            state["groups"][name] = mine[0]
            state["sizes"][name] = len(mine[1])
            state["ranks"][name] = mine[1].index(rank)

        add("h", [[jh * w + jw for jh in range(h)] for jw in range(w)])
        add("w", [[jh * w + jw for jw in range(w)] for jh in range(h)])
        everyone = [list(range(world))]
        singles = [[r] for r in range(world)]
        for name in ("spatial", "model"):
            add(name, everyone)
        for name in ("matmul", "fin", "fout", "ensemble", "batch", "data"):
            add(name, singles)

    comm.init = init
    comm.get_group = lambda name: state["groups"].get(name)
    comm.get_size = lambda name: state["sizes"].get(name, 1)
    comm.get_rank = lambda name: state["ranks"].get(name, 0)
    comm.get_world_rank = lambda: dist.get_rank()
    comm.get_world_size = lambda: dist.get_world_size()
    comm.get_local_rank = lambda: 0
    comm.is_distributed = lambda name: state["sizes"].get(name, 1) > 1
    comm.get_comm_names = lambda: ["h", "w", "spatial", "matmul", "model", "data"]
    comm.get_model_comm_names = lambda: ["h", "w", "matmul"]
    comm.get_names = lambda *a, **k: list(state["groups"])
    return comm


def worker(rank, world, port, h, w, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), GRID_H=str(h), GRID_W=str(w), RANK=str(rank), WORLD_SIZE=str(world))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import run_reference_tests as R

        R.install_environment()
        comm = make_comm(h, w)
        sys.modules["makani.utils.comm"] = comm
        sys.modules["makani.utils"].comm = comm
        import math

        import numpy as np

        import makani_b200.distributed as mbd
        from oracle import makani_oracle as O

        class OracleLocalOps:
            """the four per-rank stages with torch-harmonics' dtype behaviour (fp32 tables, fp32 in -> complex64 out)"""

            def __init__(self, t):
                self.t = t
                theta, wq = O.precompute_latitudes(t.nlat, t.grid)
                P = O.legpoly(t.mmax, t.lmax, np.cos(theta), csphase=t.csphase)
                self.P = torch.from_numpy(P[t.m_offset:t.m_offset + t.mmax_local]).float()
                self.w_local = torch.from_numpy(wq[t.lat_offset:t.lat_offset + t.nlat_local]).float()

            def fft(self, x):
                X = 2.0 * math.pi * torch.fft.rfft(x.float(), dim=-1, norm="forward")[..., :self.t.mmax]
                return X * self.w_local[:, None]

            def legendre(self, xc):
                return torch.einsum("...km,mlk->...lm", xc, self.P.to(xc.dtype))

            def ilegendre(self, xc):
                return torch.einsum("...lm,mlk->...km", xc, self.P.to(xc.dtype))

            def ifft(self, xc, dtype):
                re, im = xc.real, xc.imag.clone()
                im[..., 0] = 0.0
                return torch.fft.irfft(torch.complex(re, im), n=self.t.nlon, dim=-1, norm="forward")

        mbd.set_local_ops(OracleLocalOps)
        ns = types.ModuleType("tests.distributed")
        ns.__path__ = ["/root/reference/tests/distributed"]
        sys.modules["tests.distributed"] = ns
        import importlib

        M = importlib.import_module("tests.distributed.tests_distributed_layers")
        # torch >= 2.8 rejects device_ids=[cpu] (the reference passes [device] unconditionally, mappings.py:441-451; its CI pins
        # torch 2.7): drop the two arguments for CPU modules, everything else of DDP is the real thing
        import makani.mpu.mappings as mappings
        from torch.nn.parallel import DistributedDataParallel as RealDDP

        def ddp_cpu_ok(module, device_ids=None, output_device=None, **kw):
            if device_ids and torch.device(device_ids[0]).type == "cpu":
                device_ids, output_device = None, None
            return RealDDP(module, device_ids=device_ids, output_device=output_device, **kw)

        mappings.DistributedDataParallel = ddp_cpu_ok
        # gloo's all_gather refuses shards of different sizes (NCCL, which the reference runs on, accepts them): the reference's
        # _gather_helper gathers 91 + 90 latitude rows.  Same call, realised as one broadcast per member when the sizes differ.
        real_all_gather = dist.all_gather

        def all_gather_uneven_ok(tensor_list, tensor, group=None, async_op=False):
            if all(t.shape == tensor.shape for t in tensor_list):
                return real_all_gather(tensor_list, tensor, group=group, async_op=async_op)
            grank = dist.get_rank(group=group)
            for i, t in enumerate(tensor_list):
                buf = tensor.contiguous() if i == grank else t
                dist.broadcast(buf, src=dist.get_global_rank(group, i) if group is not None else i, group=group)
                if i == grank and t.data_ptr() != tensor.data_ptr():
                    t.copy_(tensor)
            return None

        dist.all_gather = all_gather_uneven_ok
        # The test draws the full-size input AFTER constructing the sharded module, whose weight has 46 l-modes on one rank and 45 on the
        # other: with CUDA's counter-based generator both ranks still draw the same input, the sequential CPU generator diverges.  Give the
        # CPU generator the same property: every randn / randn_like call uses its own stream (seed, call index), whatever its size.
        rng = {"seed": 0, "calls": 0}
        real_seed, real_randn, real_randn_like = torch.manual_seed, torch.randn, torch.randn_like

        def manual_seed(seed):
            rng["seed"], rng["calls"] = int(seed), 0
            return real_seed(seed)

        def stream():
            rng["calls"] += 1
            return torch.Generator().manual_seed(1000003 * rng["seed"] + rng["calls"])

        def randn(*size, **kw):
            if kw.get("generator") is None and torch.device(kw.get("device") or "cpu").type == "cpu":
                kw["generator"] = stream()
            return real_randn(*size, **kw)

        def randn_like(t, **kw):
            if t.device.type == "cpu":
                return real_randn(t.shape, dtype=kw.get("dtype", t.dtype), generator=stream())
            return real_randn_like(t, **kw)

        torch.manual_seed, torch.randn, torch.randn_like = manual_seed, randn, randn_like
        loader = unittest.defaultTestLoader
        if os.environ.get("REFDIST_SUITE") == "losses":
            # second pin at the SHT boundary: the spectral losses through thd.DistributedRealSHT == through the local transform
            ML = importlib.import_module("tests.distributed.tests_distributed_losses")
            Case = ML.TestDistributedLoss
            keys = ("spectral", "sobolev", "coherence", "quadrature")
            names = [n for n in loader.getTestCaseNames(Case) if any(k in n for k in keys)]
        elif os.environ.get("REFDIST_SUITE") == "fft":
            # the reference's own DistributedRealFFT1/2/3 (makani/mpu/fft.py) call torch_harmonics.distributed's transposes as plain
            # functions (mpu/fft.py:59,169): here those are makani_b200.distributed's
            MF = importlib.import_module("tests.distributed.tests_distributed_fft")
            Case = MF.TestDistributedRealFFT
            names = list(loader.getTestCaseNames(Case))
        elif os.environ.get("REFDIST_SUITE") == "layers_all":
            # every test of the class: the non-spectral ones (distributed MLP, instance / layer norms) reach makani_b200 only through
            # makani/mpu/mappings.py, which imports _gather / _split / _reduce / _transpose from torch_harmonics.distributed.primitives
            Case = M.TestDistributedLayers
            names = list(loader.getTestCaseNames(Case))
        else:
            Case = M.TestDistributedLayers
            names = [n for n in loader.getTestCaseNames(Case) if n.startswith("test_distributed_spectral_conv")]
        suite = unittest.TestSuite(Case(n) for n in names)
        Case.setUpClass()
        if os.environ.get("REFDIST_DEBUG_CASE"):
            a = [float(v) if "e" in v or "." in v else int(v) for v in os.environ["REFDIST_DEBUG_CASE"].split(",")]
            inst = M.TestDistributedLayers("test_distributed_spectral_conv_0")
            inst.setUp()
            try:
                M.TestDistributedLayers._orig_test_distributed_spectral_conv(inst, *a, verbose=(rank == 0))
                q.put((rank, 1, 0, 0, ["debug case passed"]))
            except Exception as e:  # noqa: BLE001
                q.put((rank, 1, 1, 0, [repr(e)[:300]]))
            dist.barrier()
            dist.destroy_process_group()
            return
        r = unittest.TextTestRunner(verbosity=0, stream=open(os.devnull, "w")).run(suite)
        msgs = [t.id().split(".")[-1] + ": " + tb.strip().splitlines()[-1][:600] for t, tb in r.failures + r.errors]
        q.put((rank, r.testsRun, len(r.failures), len(r.errors), msgs))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        import traceback

        q.put((rank, 0, 0, 1, [traceback.format_exc()[-1500:]]))


def run(h, w):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = h * w
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, port, h, w, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=1500) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    return sorted(out)


def main():
    if not os.path.isdir("/root/reference/tests/distributed"):
        print("reference tree not mounted: nothing to run")
        return 0
    h = int(sys.argv[1]) if len(sys.argv) > 2 else 2
    w = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    res = run(h, w)
    bad = 0
    for rank, ran, nf, ne, msgs in res:
        print(f"rank {rank}: ran {ran}  failures {nf}  errors {ne}")
        for m in msgs[:4]:
            print("    " + m)
        bad += nf + ne
    print(f"TOTAL grid {h}x{w}: {'OK' if bad == 0 and all(r[1] > 0 for r in res) else 'FAILED'}")
    return 0 if bad == 0 and all(r[1] > 0 for r in res) else 1


if __name__ == "__main__":
    sys.exit(main())
