#!/bin/bash
# Final round-2 record on one B200: quick test subset (the whole suite ran green on the PDL build), smoke, reference arm, default bench line, the launch
# list of the bench command and the --set full block capture for profiles/, the other workloads with their library baselines, the model-step profile.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cabi.py -q -m gpu --timeout=600 2>&1 | tail -3 | cut -c1-200 > gpurun_out/final_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final.err
timeout 400 python bench.py > gpurun_out/final_bench.json 2>> gpurun_out/final.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-stages --no-graph > gpurun_out/ncu_launches.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --workload sfno_block_240x480x384 > gpurun_out/final_bench_2a.json 2>> gpurun_out/final.err
timeout 400 python bench.py --steps 5 --warmup 3 --workload sfno_sc3_layers8_edim384 > gpurun_out/final_bench_cfg3.json 2>> gpurun_out/final.err
timeout 200 python scripts/prof_model.py > gpurun_out/final_model_profile.log 2>&1
echo "=== pytest"; cat gpurun_out/final_pytest.log
echo "=== smoke"; tail -2 gpurun_out/final_smoke.log
echo "=== ref"; cut -c1-400 gpurun_out/final_bench_ref.json
echo "=== bench"; cut -c1-1500 gpurun_out/final_bench.json
python - <<'PY'
import json
for f in ["final_bench", "final_bench_2a", "final_bench_cfg3"]:
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), "samples/s", round(d["ms_per_step"], 4), "ms; graph", d.get("cuda_graph_replay"), "host", d.get("host_enqueue_ms_per_step"), "lib", d.get("gpu_library_baseline"), "cpu", (d.get("cpu_baseline") or {}).get("value"),
              {k: v["ms"] for k, v in d.get("roofline_stages", {}).items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
echo "=== model profile"; head -32 gpurun_out/final_model_profile.log | cut -c1-200
tail -4 gpurun_out/final.err
