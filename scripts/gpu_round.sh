#!/bin/bash
# One GPU-box round trip: tests, smoke, bench (+ optional ncu profiles); everything interesting lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python scripts/umma_diag.py all small odd tiles wide cfg2c > gpurun_out/umma_diag.log 2>&1
echo "diag exit: $?" >> gpurun_out/umma_diag.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA --timeout=600 2>&1 | grep -E "parity\]|PASS|FAIL|ERROR|passed|failed|Error|assert" | tail -150 > gpurun_out/pytest_gpu.log
timeout 1500 python -m pytest tests/test_gpu_umma.py -m gpu -q -rA --timeout=900 -k "not kernels_agree" 2>&1 | grep -E "parity\]|PASS|FAIL|ERROR|passed|failed|Error|assert" | tail -80 > gpurun_out/pytest_umma.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --workload sfno_block_240x480x384 > gpurun_out/bench_2a.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --workload sfno_block_721to240x384 > gpurun_out/bench_2b.json 2>> gpurun_out/bench.err

[ -x scripts/micro/f32x2 ] && timeout 120 scripts/micro/f32x2 > gpurun_out/micro_f32x2.log 2>&1
if [ "$1" == "profile" ]; then bash scripts/gpu_profile.sh; fi
echo "=== micro"; cat gpurun_out/micro_f32x2.log 2>/dev/null
echo "=== diag"; grep -E "failures|rc=1" gpurun_out/umma_diag.log | cut -c1-260 | tail -20
echo "=== pytest"; grep -E "FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-200; echo "=== pytest umma"; grep -E "FAILED|passed|failed|parity" gpurun_out/pytest_umma.log | cut -c1-220 | tail -40
echo "=== smoke"; tail -3 gpurun_out/smoke.log; echo "=== bench"; cat gpurun_out/bench.json | cut -c1-2500; tail -5 gpurun_out/bench.err
