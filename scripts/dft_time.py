"""Time the two DFT kernels alone (CUDA events, 50 launches each, L2 flushed by the 150 MB operands). B200SHT_LIBRARY selects the build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, makani_b200 as mb
from makani_b200 import _lib
from makani_b200.sht import _ptr, _stream, _VP
dev = torch.device("cuda", 0)
plan = mb.get_plan(721, 1440, 240, 241, "equiangular", True, dev)
B, C = 1, 73
x = torch.randn(B, C, 721, 1440, device=dev).bfloat16()
lat = torch.zeros(plan.latspec_elems(B, C), device=dev)
y = torch.empty_like(x)
bias = torch.randn(C, device=dev)
st = _stream(dev)
def timed(fn, n=50):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for rep in range(2):
    ta = timed(lambda: _lib.call("b200sht_fft_analysis", plan.handle, _ptr(x), 1, B, C, _ptr(lat), 0 | 2, st))
    ts = timed(lambda: _lib.call("b200sht_fft_synthesis", plan.handle, _ptr(lat), _ptr(y), 1, B, C, _VP(0), 0 | 2, st))
    tb = timed(lambda: _lib.call("b200sht_fft_synthesis", plan.handle, _ptr(lat), _ptr(y), 1, B, C, _ptr(bias), 1 | 2, st))
    print(os.path.basename(_lib.LIB_PATH), f"analysis {ta:.1f} us  synthesis(mode 0) {ts:.1f} us  synthesis(mode 1, bias) {tb:.1f} us")
