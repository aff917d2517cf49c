#!/bin/bash
# ncu --set full of every kernel of the third fwd+bwd step of the headline block (13 kernels) -> gpurun_out/block_r02.ncu-rep
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none --launch-skip 26 -c 13 -f -o gpurun_out/block_r02 python scripts/prof_block.py > gpurun_out/ncu_block.log 2>&1
tail -3 gpurun_out/ncu_block.log
ncu -i gpurun_out/block_r02.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active 2>/dev/null | cut -d, -f5,12- | tail -15
