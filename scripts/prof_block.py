"""one forward + backward of the headline block (sfno_block_721x1440x73, bf16 + tf32) for ncu captures: python scripts/prof_block.py"""
import sys; sys.path.insert(0, '/root/repo')
import torch, makani_b200 as mb
dev = torch.device("cuda", 0)
f = mb.RealSHT(721, 1440, 240, 241, "equiangular", precision="tf32")
i = mb.InverseRealSHT(721, 1440, 240, 241, "equiangular", precision="tf32")
conv = mb.SpectralConv(f, i, 73, 73, operator_type="dhconv", precision="tf32").to(dev)
conv._wcache.enabled = False
x = torch.randn(1, 73, 721, 1440, device=dev).bfloat16().requires_grad_(True)
gy = torch.randn(1, 73, 721, 1440, device=dev).bfloat16()
for _ in range(3):
    conv.weight.grad = None; x.grad = None
    y, _ = conv(x)
    y.backward(gy)
torch.cuda.synchronize()
