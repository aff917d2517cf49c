#!/bin/bash
# ncu --set full of the two FFT kernels for one B200SHT_FFT_VARIANT (default 0); reports land in gpurun_out/prof_fft_{a,s}_v$V.ncu-rep
V=${1:-0}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu --no-stages"
B200SHT_FFT_VARIANT=$V timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_analysis_ct -s 4 -c 1 -f -o gpurun_out/prof_fft_a_v$V $BENCH > gpurun_out/ncu_fft_a.log 2>&1
B200SHT_FFT_VARIANT=$V timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_synthesis_ct -s 4 -c 1 -f -o gpurun_out/prof_fft_s_v$V $BENCH > gpurun_out/ncu_fft_s.log 2>&1
ls -la gpurun_out/prof_fft_*_v$V.ncu-rep
