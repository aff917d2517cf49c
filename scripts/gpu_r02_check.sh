#!/bin/bash
# Round-2 state check on one B200: whole GPU suite, smoke(), default bench line, the launch list of the bench command, the
# other block shapes / precisions / the full model, and one --set full capture of a full step.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests/ -x -q -m gpu --timeout=600 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-stages > gpurun_out/ncu_launches.log 2>&1
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --workload sfno_block_240x480x384 > gpurun_out/bench_2a.json 2>> gpurun_out/bench.err
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --workload sfno_block_721to240x384 > gpurun_out/bench_2b.json 2>> gpurun_out/bench.err
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --precision fp32x3 > gpurun_out/bench_fp32x3.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --workload sfno_sc3_layers8_edim384 > gpurun_out/bench_cfg3.json 2>> gpurun_out/bench.err
if [ "$1" == "profile" ]; then bash scripts/gpu_prof.sh; fi
echo "=== pytest"; cut -c1-250 gpurun_out/pytest_gpu.log | tail -8
echo "=== smoke"; tail -2 gpurun_out/smoke.log
echo "=== bench"; cut -c1-2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
for f in ["bench_2a.json", "bench_2b.json", "bench_fp32x3.json", "bench_cfg3.json"]:
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), "samples/s", round(d["ms_per_step"], 4), "ms", {k: v["ms"] for k, v in d.get("roofline_stages", {}).items()}, d.get("gpu_library_baseline"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
