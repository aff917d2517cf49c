#!/bin/bash
# Programmatic dependent launch: whole GPU suite with it on (default), then the step with / without it, eager and graph-replayed.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -q -m gpu --timeout=600 2>&1 | tail -15 | cut -c1-250 > gpurun_out/pdl_pytest.log
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --no-stages --graph > gpurun_out/bench_pdl_$tag.json 2>> gpurun_out/pdl.err; }
run off B200SHT_PDL=0
run on B200SHT_PDL=1
run off2 B200SHT_PDL=0
run on2 B200SHT_PDL=1
B200SHT_PDL=0 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-stages --graph --workload sfno_block_240x480x384 > gpurun_out/bench_pdl_2a_off.json 2>> gpurun_out/pdl.err
B200SHT_PDL=1 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-stages --graph --workload sfno_block_240x480x384 > gpurun_out/bench_pdl_2a_on.json 2>> gpurun_out/pdl.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_pdl_default.json 2>> gpurun_out/pdl.err
cat gpurun_out/pdl_pytest.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_pdl_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "eager", round(d["ms_per_step"], 4), "ms; host enqueue", round(d.get("host_enqueue_ms_per_step") or -1, 4), "ms; graph", d.get("cuda_graph_replay"), "e2e", round(d["e2e"]["value"], 1), "serial", round(d["e2e"]["serial_value"], 1))
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -5 gpurun_out/pdl.err
