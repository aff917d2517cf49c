#!/bin/bash
# One GPU-box round trip mirroring what the driver runs at round end: the whole GPU test suite, smoke(), the reference arm and the
# default bench line (+ the two other block shapes, + optional ncu profiles).  Everything interesting lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout=900 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench.err
timeout 600 python bench.py > gpurun_out/bench.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --workload sfno_block_240x480x384 > gpurun_out/bench_2a.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --workload sfno_block_721to240x384 > gpurun_out/bench_2b.json 2>> gpurun_out/bench.err
if [ "$1" == "profile" ]; then bash scripts/gpu_profile.sh; fi
echo "=== pytest"; cat gpurun_out/pytest_gpu.log | cut -c1-250 | tail -12
echo "=== smoke"; tail -3 gpurun_out/smoke.log
echo "=== reference arm"; cut -c1-700 gpurun_out/bench_ref.json
echo "=== bench"; cut -c1-3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
for f in ["bench_2a.json", "bench_2b.json"]:
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "samples/s", round(d["ms_per_step"], 4), "ms", {k: v["ms"] for k, v in d["roofline_stages"].items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
