#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dft.py tests/test_gpu_sfno.py -m gpu -q --timeout=300 -rA 2>&1 | grep -E "parity\]|passed|failed|Error|FAILED|assert" | tail -90 > gpurun_out/dft_pytest.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_umma.py -m gpu -q -x --timeout=600 2>&1 | tail -8 > gpurun_out/dft_parity.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/dft_bench_on.json 2> gpurun_out/dft_bench.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --workload sfno_block_240x480x384 > gpurun_out/dft_bench_2a.json 2>> gpurun_out/dft_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:dft_ --launch-skip 3 -c 2 --csv --log-file gpurun_out/ncu_dft_counts.csv python scripts/prof_dft.py > /dev/null 2>&1
echo "== dft+sfno tests"; cat gpurun_out/dft_pytest.log | grep -v "dft_analysis\|dft_synthesis" | cut -c1-250
echo "== parity"; cat gpurun_out/dft_parity.log | cut -c1-300
python - <<'PY'
import json
for f in ["dft_bench_on.json", "dft_bench_2a.json"]:
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "samples/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"], 1), {k: v["ms"] for k, v in d["roofline_stages"].items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -3 gpurun_out/dft_bench.err; grep -v "^==" gpurun_out/ncu_dft_counts.csv | cut -d, -f5,12- | tail -12
