#!/bin/bash
# ncu evidence for profiles/: (1) launch list with device time per kernel, (2) --set full capture of the hot kernels.
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu --no-stages"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_analysis_ct -s 4 -c 1 -f -o gpurun_out/prof_fft_analysis $BENCH > gpurun_out/ncu_fft_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_synthesis_ct -s 4 -c 1 -f -o gpurun_out/prof_fft_synthesis $BENCH > gpurun_out/ncu_fft_s.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_kernel -s 10 -c 5 -f -o gpurun_out/prof_umma $BENCH > gpurun_out/ncu_umma.log 2>&1
ls -la gpurun_out/*.ncu-rep 2>/dev/null; tail -3 gpurun_out/ncu_umma.log
