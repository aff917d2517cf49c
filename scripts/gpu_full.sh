#!/bin/bash
# whole GPU test suite + smoke + default bench line + the two other block shapes + the full-model workload (1 GPU)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu --timeout=1200 2>&1 | tail -12 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --workload sfno_block_240x480x384 > gpurun_out/bench_2a.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --workload sfno_block_721to240x384 > gpurun_out/bench_2b.json 2>> gpurun_out/bench.err
timeout 900 python bench.py --steps 5 --warmup 3 --workload sfno_sc3_layers8_edim384 > gpurun_out/bench_cfg3.json 2>> gpurun_out/bench.err
echo "=== pytest"; cat gpurun_out/pytest_gpu.log | cut -c1-250
echo "=== smoke"; tail -2 gpurun_out/smoke.log | cut -c1-300
python - <<'PY'
import json
for f in ["bench.json", "bench_2a.json", "bench_2b.json"]:
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "samples/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"], 1), "lib", (d.get("gpu_library_baseline") or {}).get("value"), {k: v["ms"] for k, v in d["roofline_stages"].items()})
    except Exception as e:
        print(f, "unreadable:", e)
try:
    d = json.loads(open("gpurun_out/bench_cfg3.json").read().strip().splitlines()[-1])
    print("cfg3", round(d["value"], 3), "samples/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"], 3), "launches", d["gpu_launches"], "lib", d.get("gpu_library_baseline"))
except Exception as e:
    print("cfg3 unreadable:", e)
PY
tail -5 gpurun_out/bench.err
