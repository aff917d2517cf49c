#!/bin/bash
# Is the step launch-bound on the host?  host_enqueue_ms_per_step and the CUDA-graph replay next to the eager step, for the unchunked
# and the latitude-chunked analysis / synthesis pairs (graph replay removes the per-launch host cost from the comparison).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_consumers.py -k attention -m gpu -q --timeout=600 -s 2>&1 | grep -E "passed|failed|Error|assert|FAILED|gates" | cut -c1-200 | tail -12 > gpurun_out/graph_pytest.log
B200SHT_LAT_CHUNKS_SYN=3 timeout 600 python -m pytest "tests/test_gpu_bench_configs.py::test_benched_block_fp32_activations_tf32" tests/test_gpu_cabi.py -m gpu -q --timeout=600 -s 2>&1 \
  | grep -E "benched|passed|failed|Error|error|assert|FAILED" | cut -c1-260 | tail -8 >> gpurun_out/graph_pytest.log
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --no-stages --graph > gpurun_out/bench_graph_$tag.json 2>> gpurun_out/graph.err; }
run base B200SHT_LAT_CHUNKS=1
run ana2 B200SHT_LAT_CHUNKS=2
run ana3 B200SHT_LAT_CHUNKS=3
run syn3 B200SHT_LAT_CHUNKS_SYN=3
run both3 B200SHT_LAT_CHUNKS=3 B200SHT_LAT_CHUNKS_SYN=3
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-stages --graph --workload sfno_block_240x480x384 > gpurun_out/bench_graph_2a.json 2>> gpurun_out/graph.err
cat gpurun_out/graph_pytest.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_graph_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "eager", round(d["ms_per_step"], 4), "ms; host enqueue", round(d.get("host_enqueue_ms_per_step") or -1, 4), "ms; graph", d.get("cuda_graph_replay"), "launches", d.get("gpu_launches"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -5 gpurun_out/graph.err
