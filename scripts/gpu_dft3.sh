#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/dft_diag.py 64 8 17 > gpurun_out/diag_64.log 2>&1
timeout 120 python scripts/dft_diag.py 1440 8 241 > gpurun_out/diag_1440.log 2>&1
timeout 120 python scripts/dft_diag.py 720 8 121 > gpurun_out/diag_720.log 2>&1
timeout 600 python -m pytest tests/test_gpu_dft.py -m gpu -q --timeout=300 -rA -k "analysis" 2>&1 | grep -E "parity\]|passed|failed|Error" | tail -40 > gpurun_out/dft_pytest.log
B200SHT_DFT=0 timeout 1200 python -m pytest tests/test_gpu_cabi.py tests/test_gpu_bench_configs.py -m gpu -q --timeout=900 -rA 2>&1 | grep -E "parity\]|benched\]|passed|failed|Error|FAILED" | tail -60 > gpurun_out/new_tests_dftoff.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/dft_bench_on.json 2> gpurun_out/dft_bench.err
echo "== diag 64"; head -50 gpurun_out/diag_64.log | cut -c1-600
echo "== diag 1440"; head -30 gpurun_out/diag_1440.log | cut -c1-600
echo "== diag 720"; head -30 gpurun_out/diag_720.log | cut -c1-600
echo "== dft analysis tests"; cat gpurun_out/dft_pytest.log | cut -c1-250
echo "== new tests (DFT off)"; cat gpurun_out/new_tests_dftoff.log | cut -c1-250
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
