"""Where the full-model step (bench.py --workload sfno_sc3_layers8_edim384) spends its GPU time: torch.profiler over one fwd+bwd, kernels grouped into
the spectral path of this package (b200sht::*) and PyTorch's own operators (1x1 convolutions, instance norm, GELU, casts, adds).  python scripts/prof_model.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from bench import MODEL_WORKLOADS
from makani_b200.sfno import SphericalFourierNeuralOperatorNet

dev = torch.device("cuda", 0)
cfg = MODEL_WORKLOADS["sfno_sc3_layers8_edim384"]
torch.manual_seed(333)
net = SphericalFourierNeuralOperatorNet(**cfg, precision="tf32").to(dev)   # the mode bench.py times (allow_tf32, as makani/train.py:87)
x = torch.randn(1, cfg["inp_chans"], *cfg["inp_shape"], device=dev)


def step():
    for p in net.parameters():
        p.grad = None
    with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
        out = net(x)
    out.float().square().mean().backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count) for e in prof.key_averages()]
rows = [r for r in rows if r[1] > 0]
total = sum(r[1] for r in rows)
ours = sum(r[1] for r in rows if "b200sht" in r[0])
print(f"total device time {total / 1e3:.2f} ms; b200sht kernels {ours / 1e3:.2f} ms ({100 * ours / total:.1f} %); other {100 - 100 * ours / total:.1f} %")
for k, t, n in sorted(rows, key=lambda r: -r[1])[:28]:
    print(f"{t / 1e3:9.3f} ms  x{n:<4d} {k[:150]}")
