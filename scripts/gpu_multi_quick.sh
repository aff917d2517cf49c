#!/bin/bash
# Short multi-GPU round trip: one NCCL h x w check (tf32) + the data-parallel bench line (+ reference arm under torchrun).
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" == "2" ]; then h=2; w=1; elif [ "$N" == "4" ]; then h=2; w=2; else h=4; w=2; fi
timeout 300 $TR --master-port 29611 scripts/dist_gpu_check.py --h $h --w $w --precision tf32 --cases small,odd,block73 > gpurun_out/dist_${h}x${w}_tf32.log 2>&1
echo "dist ${h}x${w} tf32 rc=$?"; grep -E '^\{' gpurun_out/dist_${h}x${w}_tf32.log | cut -c1-900; grep -E "Error|error|Traceback" gpurun_out/dist_${h}x${w}_tf32.log | head -5
timeout 300 $TR --master-port 29612 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_dp${N}.json 2> gpurun_out/bench_dp${N}.err
echo "bench dp$N rc=$?"; cut -c1-900 gpurun_out/bench_dp${N}.json; tail -3 gpurun_out/bench_dp${N}.err
timeout 300 $TR --master-port 29613 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_ref_dp${N}.json 2> gpurun_out/bench_ref_dp${N}.err
echo "bench ref dp$N rc=$?"; cut -c1-600 gpurun_out/bench_ref_dp${N}.json; tail -2 gpurun_out/bench_ref_dp${N}.err
