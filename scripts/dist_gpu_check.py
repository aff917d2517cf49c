#!/usr/bin/env python
"""Multi-GPU check of the h x w spatial-parallel path over NCCL (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_gpu_check.py --h 4 --w 2

Mirrors /root/reference/tests/distributed/tests_distributed_layers.py:69-223 (test_distributed_spectral_conv): the same
SpectralConv(dhconv, bias) built on Distributed*SHT and on local *SHT must agree in output, input gradient, weight gradient
(gathered along l over h, summed over w) and bias gradient; plus DistributedRealSHT / InverseRealSHT against the local transforms.
Also times the distributed block (CUDA events, max over ranks).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import makani_b200 as mb  # noqa: E402
import makani_b200.distributed as mbd  # noqa: E402


def rel(a, b):
    return ((a - b).abs().double().pow(2).sum().sqrt() / b.abs().double().pow(2).sum().sqrt().clamp_min(1e-30)).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--h", type=int, default=2)
    ap.add_argument("--w", type=int, default=1)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--cases", default="small,odd,sfno")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    h, w = args.h, args.w
    assert h * w == world, (h, w, world)
    h_groups = [dist.new_group([ih * w + iw for ih in range(h)]) for iw in range(w)]
    w_groups = [dist.new_group([ih * w + iw for iw in range(w)]) for ih in range(h)]
    ih, iw = rank // w, rank % w
    hg, wg = (h_groups[iw] if h > 1 else None), (w_groups[ih] if w > 1 else None)
    mbd.init(hg, wg)
    tol = 2e-5 if args.precision == "fp32" else 3e-3
    cases = {
        # nlat_i nlon_i nlat_o nlon_o lmax mmax B C   (shapes of the reference's distributed test + an SFNO-like one)
        "small": (64, 128, 64, 128, 32, 33, 2, 8),
        "odd": (91, 180, 181, 360, 91, 91, 1, 10),
        "sfno": (721, 1440, 240, 480, 240, 241, 1, 16),
        "block73": (721, 1440, 721, 1440, 240, 241, 1, 73),   # the headline block of bench.py, one sample split over h x w ranks
    }
    ok_all = True
    out = {}
    for name in args.cases.split(","):
        nlat_i, nlon_i, nlat_o, nlon_o, lmax, mmax, B, C = cases[name]
        gi, go = "equiangular", ("legendre-gauss" if name == "sfno" else "equiangular")
        fl = mb.RealSHT(nlat_i, nlon_i, lmax, mmax, gi, precision=args.precision)
        il = mb.InverseRealSHT(nlat_o, nlon_o, lmax, mmax, go, precision=args.precision)
        fd = mbd.DistributedRealSHT(nlat_i, nlon_i, lmax, mmax, gi, precision=args.precision)
        idd = mbd.DistributedInverseRealSHT(nlat_o, nlon_o, lmax, mmax, go, precision=args.precision)
        torch.manual_seed(333)
        conv_l = mb.SpectralConv(fl, il, C, C, operator_type="dhconv", bias=True, precision=args.precision).to(dev)
        conv_d = mb.SpectralConv(fd, idd, C, C, operator_type="dhconv", bias=True, precision=args.precision).to(dev)

        def shard(t, hd, wd, hs, ws):
            if hd is not None:
                t = torch.split(t, hs, dim=hd)[ih]
            if wd is not None:
                t = torch.split(t, ws, dim=wd)[iw]
            return t.contiguous()

        with torch.no_grad():
            conv_l.bias.copy_(torch.randn(1, C, 1, 1, device=dev))
            dist.broadcast(conv_l.weight.data, 0)
            dist.broadcast(conv_l.bias.data, 0)
            conv_d.weight.copy_(shard(conv_l.weight, -1, None, fd.l_shapes, None))
            conv_d.bias.copy_(conv_l.bias)
        x = torch.randn(B, C, nlat_i, nlon_i, device=dev)
        dist.broadcast(x, 0)
        xl = x.clone().requires_grad_(True)
        yl, _ = conv_l(xl)
        gy = torch.randn_like(yl)
        dist.broadcast(gy, 0)
        yl.backward(gy)
        xd = shard(x, -2, -1, fd.lat_shapes, fd.lon_shapes).requires_grad_(True)
        yd, _ = conv_d(xd)
        yd.backward(shard(gy, -2, -1, idd.lat_shapes, idd.lon_shapes))
        res = {}
        res["y"] = rel(yd, shard(yl, -2, -1, idd.lat_shapes, idd.lon_shapes))
        res["dx"] = rel(xd.grad, shard(xl.grad, -2, -1, fd.lat_shapes, fd.lon_shapes))
        gw = conv_d.weight.grad.clone()
        if wg is not None:   # dhconv weight is shared over w (spectral_convolution.py:195-198): sum its gradient over the w group
            dist.all_reduce(torch.view_as_real(gw), group=wg)
        res["dw"] = rel(gw, shard(conv_l.weight.grad, -1, None, fd.l_shapes, None))
        gb = conv_d.bias.grad.clone()
        dist.all_reduce(gb)  # bias is shared over the whole model group
        res["db"] = rel(gb, conv_l.bias.grad)
        # bare transforms
        c_l = fl(x)
        c_d = fd(shard(x, -2, -1, fd.lat_shapes, fd.lon_shapes))
        res["sht"] = rel(c_d, shard(c_l, -2, -1, fd.l_shapes, fd.m_shapes))
        cin = torch.randn(B, C, lmax, mmax, dtype=torch.complex64, device=dev)
        dist.broadcast(torch.view_as_real(cin), 0)
        res["isht"] = rel(idd(shard(cin, -2, -1, idd.l_shapes, idd.m_shapes)), shard(il(cin), -2, -1, idd.lat_shapes, idd.lon_shapes))
        # timing of the distributed block fwd+bwd
        def step():
            xd.grad = None
            conv_d.weight.grad = None
            y, _ = conv_d(xd)
            y.backward(shard(gy, -2, -1, idd.lat_shapes, idd.lon_shapes))
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        res["ms_fwd_bwd"] = ms.item()
        worst = torch.tensor([max(v for k, v in res.items() if k != "ms_fwd_bwd")], device=dev)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        ok = worst.item() < tol
        ok_all &= ok
        out[name] = {"ok": ok, "worst_rel": worst.item(), **{k: round(v, 9) for k, v in res.items()}}
    if rank == 0:
        print(json.dumps({"h": h, "w": w, "precision": args.precision, "ok": ok_all, "cases": out}))
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
