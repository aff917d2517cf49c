#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dft.py tests/test_gpu_umma.py tests/test_gpu_parity.py -m gpu -q -x --timeout=600 2>&1 | tail -5 > gpurun_out/dft_pytest.log
python scripts/dft_time.py > gpurun_out/dft_time.log 2>&1
python scripts/dft_waitprof.py > gpurun_out/waitprof.log 2>&1
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --no-hxw > gpurun_out/dft_bench_on.json 2> gpurun_out/dft_bench.err
echo "== tests"; cut -c1-300 gpurun_out/dft_pytest.log
grep -v Warn gpurun_out/dft_time.log; tail -14 gpurun_out/waitprof.log
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
