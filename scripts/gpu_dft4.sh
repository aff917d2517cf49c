#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/dft_diag.py 64 8 17 > gpurun_out/diag_64.log 2>&1
timeout 120 python scripts/dft_diag.py 1440 8 241 > gpurun_out/diag_1440.log 2>&1
timeout 600 python -m pytest tests/test_gpu_dft.py -m gpu -q --timeout=300 -rA 2>&1 | grep -E "parity\]|passed|failed|Error|FAILED|assert" | tail -70 > gpurun_out/dft_pytest.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_umma.py tests/test_gpu_cabi.py -m gpu -q -x --timeout=600 2>&1 | tail -15 > gpurun_out/dft_parity.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/dft_bench_on.json 2> gpurun_out/dft_bench.err
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
for _ in range(2):
    _lib.call("b200sht_fft_analysis", plan.handle, _ptr(x), 1, B, C, _ptr(lat), 0 | 2, st)
    _lib.call("b200sht_fft_synthesis", plan.handle, _ptr(lat), _ptr(y), 1, B, C, _VP(0), 0 | 2, st)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dft_ --launch-skip 3 -c 2 -o gpurun_out/dft_r02b -f python /tmp/prof.py > gpurun_out/ncu_dft.log 2>&1
echo "== diag 64"; head -12 gpurun_out/diag_64.log | cut -c1-400; tail -1 gpurun_out/diag_64.log
echo "== diag 1440"; head -8 gpurun_out/diag_1440.log | cut -c1-400; tail -1 gpurun_out/diag_1440.log
echo "== dft tests"; cat gpurun_out/dft_pytest.log | cut -c1-250
echo "== parity"; cat gpurun_out/dft_parity.log | cut -c1-300
python - <<'PY'
import json
for f in ["dft_bench_on.json"]:
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "samples/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"], 1), {k: v["ms"] for k, v in d["roofline_stages"].items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -3 gpurun_out/dft_bench.err; tail -2 gpurun_out/ncu_dft.log
