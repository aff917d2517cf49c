#!/bin/bash
# GPU round trip for the tensor-core DFT: its own tests, the TF32 parity suites, and the headline bench with the DFT on / off.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dft.py -m gpu -q --timeout=300 2>&1 | tail -40 > gpurun_out/dft_pytest.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_umma.py -m gpu -q -x --timeout=600 2>&1 | tail -15 > gpurun_out/dft_parity.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/dft_bench_on.json 2> gpurun_out/dft_bench.err
B200SHT_DFT=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/dft_bench_off.json 2>> gpurun_out/dft_bench.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --workload sfno_block_240x480x384 > gpurun_out/dft_bench_2a.json 2>> gpurun_out/dft_bench.err
echo "=== dft pytest"; cut -c1-400 gpurun_out/dft_pytest.log
echo "=== parity"; cut -c1-300 gpurun_out/dft_parity.log
python - <<'PY'
import json
for f in ["dft_bench_on.json", "dft_bench_off.json", "dft_bench_2a.json"]:
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "samples/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"], 1), {k: v["ms"] for k, v in d["roofline_stages"].items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -5 gpurun_out/dft_bench.err
