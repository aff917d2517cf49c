#!/bin/bash
# Multi-GPU round trip (gpurun --gpus N): NCCL checks of the h x w path + the data-parallel bench line.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" == "2" ]; then GRIDS="2x1 1x2"; elif [ "$N" == "4" ]; then GRIDS="2x2 4x1"; else GRIDS="4x2"; fi
P=29600
for g in $GRIDS; do
  h=${g%x*}; w=${g#*x}; P=$((P+1))
  timeout 600 $TR --master-port $P scripts/dist_gpu_check.py --h $h --w $w --precision fp32 > gpurun_out/dist_${g}_fp32.log 2>&1
  echo "dist $g fp32 rc=$?"; grep -E '^\{' gpurun_out/dist_${g}_fp32.log | cut -c1-1500; grep -E "Error|error|Traceback" gpurun_out/dist_${g}_fp32.log | head -5
  P=$((P+1))
  timeout 600 $TR --master-port $P scripts/dist_gpu_check.py --h $h --w $w --precision tf32 > gpurun_out/dist_${g}_tf32.log 2>&1
  echo "dist $g tf32 rc=$?"; grep -E '^\{' gpurun_out/dist_${g}_tf32.log | cut -c1-1500; grep -E "Error|error|Traceback" gpurun_out/dist_${g}_tf32.log | head -5
done
P=$((P+1))
timeout 600 $TR --master-port $P bench.py --gpus $N --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_dp${N}.json 2> gpurun_out/bench_dp${N}.err
echo "bench dp$N rc=$?"; cut -c1-700 gpurun_out/bench_dp${N}.json; tail -3 gpurun_out/bench_dp${N}.err
