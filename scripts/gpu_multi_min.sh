#!/bin/bash
# Minimal 2-GPU round trip (charged 2x): NCCL h x w check of the CUDA local stages + the data-parallel bench line with its hxw object.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29611 scripts/dist_gpu_check.py --h 2 --w 1 --precision tf32 --cases small,block73 > gpurun_out/dist_2x1_tf32.log 2>&1
echo "dist 2x1 tf32 rc=$?"; grep -E '^\{' gpurun_out/dist_2x1_tf32.log | cut -c1-700; grep -E "Error|error|Traceback" gpurun_out/dist_2x1_tf32.log | head -5
timeout 250 $TR --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-stages > gpurun_out/bench_dp2.json 2> gpurun_out/bench_dp2.err
echo "bench dp2 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_dp2.json").read().strip().splitlines()[-1])
    print("dp2 value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "hxw", d.get("hxw"))
except Exception as e:
    print("unreadable", e)
PY
tail -4 gpurun_out/bench_dp2.err
