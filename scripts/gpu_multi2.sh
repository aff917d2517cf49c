#!/bin/bash
# N-GPU validation: h x w correctness vs the local modules over NCCL, then the bench line at N GPUs (DP + hxw)
N=${1:-2}; H=${2:-1}; W=${3:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_gpu_check.py --h $H --w $W --precision tf32 --cases small,odd,sfno > gpurun_out/dist_${N}_tf32.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/dist_gpu_check.py --h $H --w $W --precision fp32 --cases small,odd > gpurun_out/dist_${N}_fp32.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err
echo "== dist tf32"; tail -2 gpurun_out/dist_${N}_tf32.log | cut -c1-900
echo "== dist fp32"; tail -2 gpurun_out/dist_${N}_fp32.log | cut -c1-600
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n${N}.json").read().strip().splitlines()[-1])
    print("N=${N}", round(d["value"], 1), "samples/s", round(d["ms_per_step"], 4), "ms; e2e", round(d["e2e"]["value"], 1), "hxw", d.get("hxw"))
except Exception as e:
    print("bench unreadable:", e)
PY
tail -4 gpurun_out/bench_n${N}.err
