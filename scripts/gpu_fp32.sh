#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fp32_modes.py -m gpu -q --timeout=600 -s 2>&1 | grep -E "\[parity\] (equi|leg)|benched|passed|failed|Error|assert" | cut -c1-500 | tail -12 > gpurun_out/fp32_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-hxw --precision fp32x3 > gpurun_out/bench_fp32x3.json 2> gpurun_out/fp32.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-hxw --precision fp32 > gpurun_out/bench_fp32.json 2>> gpurun_out/fp32.err
cat gpurun_out/fp32_pytest.log
python - <<'PY'
import json
for f in ["bench_fp32x3.json", "bench_fp32.json"]:
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "samples/s", round(d["ms_per_step"], 4), "ms", {k: v["ms"] for k, v in d["roofline_stages"].items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -3 gpurun_out/fp32.err
