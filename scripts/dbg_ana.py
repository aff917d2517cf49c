import sys; sys.path.insert(0, '/root/repo')
import torch, makani_b200 as mb
from makani_b200 import _lib
from makani_b200.sht import _ptr, _stream
dev = torch.device("cuda", 0)
nlat, nlon, mmax = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 128, 65)
plan = mb.get_plan(nlat, nlon, min(nlat, 16), mmax, "equiangular", True, dev)
B, C = 2, 3
x = torch.randn(B, C, nlat, nlon, device=dev)
lat = torch.zeros(plan.latspec_elems(B, C), device=dev)
_lib.call("b200sht_fft_analysis", plan.handle, _ptr(x), 0, B, C, _ptr(lat), 0 | 2, _stream(dev))
torch.cuda.synchronize()
X = lat[: mmax * 2 * B * C * plan.kp].view(mmax, 2, B * C, plan.kp)
got = torch.complex(X[:, 0, :, :nlat], X[:, 1, :, :nlat]).permute(1, 2, 0).reshape(B, C, nlat, mmax)
import math
from oracle import makani_oracle as O
_, w = O.precompute_latitudes(nlat, "equiangular")
ref = torch.fft.rfft(x.double().cpu(), dim=-1)[..., :mmax] * (torch.from_numpy(w) * 2 * math.pi / nlon)[:, None]
print("rel", ((got.cpu() - ref).abs().pow(2).sum().sqrt() / ref.abs().pow(2).sum().sqrt()).item())
