#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/dft_bench_on.json 2> gpurun_out/dft_bench.err
timeout 2400 python -m pytest tests/ -q -x -m gpu --timeout=1200 -rA 2>&1 | grep -v "^PASSED\|parity\]" | tail -70 > gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for f in ["dft_bench_on.json"]:
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "samples/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"], 1), {k: v["ms"] for k, v in d["roofline_stages"].items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -3 gpurun_out/dft_bench.err
echo "=== pytest"; cat gpurun_out/pytest_gpu.log | cut -c1-250
