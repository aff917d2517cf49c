#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/dft_diag.py 64 8 17 > gpurun_out/diag_64.log 2>&1
timeout 120 python scripts/dft_diag.py 1440 8 241 > gpurun_out/diag_1440.log 2>&1
timeout 120 python scripts/dft_diag.py 720 8 121 > gpurun_out/diag_720.log 2>&1
timeout 600 python -m pytest tests/test_gpu_dft.py -m gpu -q --timeout=300 -k "analysis" 2>&1 | tail -15 > gpurun_out/dft_pytest.log
cat > /tmp/prof.py <<'PY'
import sys; sys.path.insert(0, '/root/repo')
import torch, makani_b200 as mb
from makani_b200 import _lib
from makani_b200.sht import _ptr, _stream, _VP
dev = torch.device("cuda", 0)
plan = mb.get_plan(721, 1440, 240, 241, "equiangular", True, dev)
B, C = 1, 73
x = torch.randn(B, C, 721, 1440, device=dev).bfloat16()
lat = torch.zeros(plan.latspec_elems(B, C), device=dev)
y = torch.empty_like(x)
st = _stream(dev)
for _ in range(3):
    _lib.call("b200sht_fft_analysis", plan.handle, _ptr(x), 1, B, C, _ptr(lat), 0 | 2, st)
    _lib.call("b200sht_fft_synthesis", plan.handle, _ptr(lat), _ptr(y), 1, B, C, _VP(0), 0 | 2, st)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dft_ -c 4 -o gpurun_out/dft_r02a python /tmp/prof.py > gpurun_out/ncu_dft.log 2>&1
echo "== diag 64"; head -40 gpurun_out/diag_64.log
echo "== diag 1440"; head -30 gpurun_out/diag_1440.log
echo "== diag 720"; head -30 gpurun_out/diag_720.log
echo "== pytest"; cat gpurun_out/dft_pytest.log | cut -c1-300
tail -3 gpurun_out/ncu_dft.log
