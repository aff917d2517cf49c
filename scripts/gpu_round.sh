#!/bin/bash
# One GPU-box round trip: tests, smoke, bench; everything interesting lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rA -x --timeout=600 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -60 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -5; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
