#!/bin/bash
# data-parallel step at N GPUs with the SM reservation for the overlapped all-reduce on (8, 16) and off (0); N=1 beside it
N=${1:-2}
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 5 --no-cpu --no-hxw > gpurun_out/dp_n1.json 2> gpurun_out/dp.err
for R in 8 0 16; do
  B200SHT_OVERLAP_SMS=$R timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$((R % 10)) bench.py --gpus $N --steps 50 --warmup 5 --no-cpu --no-hxw > gpurun_out/dp_n${N}_r${R}.json 2>> gpurun_out/dp.err
done
python - <<PY
import json
for f in ["dp_n1", "dp_n${N}_r8", "dp_n${N}_r0", "dp_n${N}_r16"]:
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "samples/s", round(d["ms_per_step"], 4), "ms")
    except Exception as e:
        print(f, "unreadable:", e)
PY
grep -v "^$\|Setting OMP\|\*\*\*\*" gpurun_out/dp.err | tail -5
