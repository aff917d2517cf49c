#!/bin/bash
# Smallest useful GPU round trip: FFT / SHT / conv parity subset + the headline bench line (no CPU legs).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fft or sht or SHT or golden or full_size" 2>&1 | tail -3
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --graph > gpurun_out/micro_bench.json 2> gpurun_out/micro_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/micro_bench.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), "samples/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"], 1), {k: v["ms"] for k, v in d["roofline_stages"].items()})
print("cuda graph replay:", d.get("cuda_graph_replay"))
PY
tail -3 gpurun_out/micro_bench.err
