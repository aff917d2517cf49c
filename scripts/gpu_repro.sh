#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dft.py tests/test_gpu_parity.py -m gpu -q -x --timeout=600 -k "dft or fft_stages or forward_inverse_fp32" 2>&1 | tail -60 > gpurun_out/repro.log
timeout 600 python -m pytest tests/test_gpu_dft.py -m gpu -q --timeout=300 -rA 2>&1 | grep -E "parity\]|passed|failed|Error|FAILED" | tail -12 > gpurun_out/dft_pytest.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/dft_bench_on.json 2> gpurun_out/dft_bench.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --workload sfno_block_240x480x384 > gpurun_out/dft_bench_2a.json 2>> gpurun_out/dft_bench.err
echo "== repro"; cat gpurun_out/repro.log | cut -c1-220
echo "== dft tests"; cat gpurun_out/dft_pytest.log | cut -c1-200
python - <<'PY'
import json
for f in ["dft_bench_on.json", "dft_bench_2a.json"]:
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "samples/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"], 1), {k: v["ms"] for k, v in d["roofline_stages"].items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -3 gpurun_out/dft_bench.err
