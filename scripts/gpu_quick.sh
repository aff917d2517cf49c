#!/bin/bash
# Short GPU round trip: parity tests + headline bench (A/B of an FFT variant through B200SHT_FFT_VARIANT).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_umma.py -m gpu -q -x --timeout=600 -k "not kernels_agree" 2>&1 | tail -15 > gpurun_out/quick_pytest.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
B200SHT_FFT_VARIANT=1 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/quick_bench_v1.json 2>> gpurun_out/quick_bench.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --workload sfno_block_240x480x384 > gpurun_out/quick_bench_2a.json 2>> gpurun_out/quick_bench.err
echo "=== pytest"; cat gpurun_out/quick_pytest.log | cut -c1-300
python - <<'PY'
import json
for f in ["quick_bench.json", "quick_bench_v1.json", "quick_bench_2a.json"]:
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "samples/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"], 1), {k: v["ms"] for k, v in d["roofline_stages"].items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -5 gpurun_out/quick_bench.err
