#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fp32_modes.py tests/test_gpu_parity.py tests/test_gpu_cabi.py -m gpu -q -x --timeout=600 -s 2>&1 | grep -E "strict fp32|passed|failed|Error|assert" | tail -12 > gpurun_out/fp32_pytest.log
timeout 600 python -m pytest tests/test_gpu_bench_configs.py -m gpu -q -x --timeout=600 2>&1 | tail -3 >> gpurun_out/fp32_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-hxw --precision fp32 > gpurun_out/bench_fp32_tc.json 2> gpurun_out/fp32.err
B200SHT_FP32_SIMT=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-hxw --precision fp32 > gpurun_out/bench_fp32_simt.json 2>> gpurun_out/fp32.err
cut -c1-400 gpurun_out/fp32_pytest.log
python - <<'PY'
import json
for f in ["bench_fp32_tc.json", "bench_fp32_simt.json"]:
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "samples/s", round(d["ms_per_step"], 4), "ms", {k: v["ms"] for k, v in d["roofline_stages"].items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -3 gpurun_out/fp32.err
