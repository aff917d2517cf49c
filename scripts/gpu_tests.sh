#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -q -m gpu --timeout=1200 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --workload sfno_sc3_layers8_edim384 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench.err
echo "=== pytest"; cat gpurun_out/pytest_gpu.log | cut -c1-250
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_cfg3.json").read().strip().splitlines()[-1])
    print("cfg3", round(d["value"], 3), "samples/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"], 3), "launches", d["gpu_launches"], "lib", d.get("gpu_library_baseline"))
except Exception as e:
    print("cfg3 unreadable:", e)
PY
tail -5 gpurun_out/bench.err
