#!/bin/bash
# ncu --set full of the three mix kernels on the C=384 interior block (cfg 2a)
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu --no-stages --workload sfno_block_240x480x384"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_kernel -s 7 -c 7 -f -o gpurun_out/prof_mix384 $BENCH > gpurun_out/ncu_mix.log 2>&1
ls -la gpurun_out/prof_mix384.ncu-rep; tail -2 gpurun_out/ncu_mix.log
