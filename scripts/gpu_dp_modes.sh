#!/bin/bash
# 2 GPUs (charged 2x): data-parallel step with the weight-gradient all-reduce overlapped (side stream, reserved SMs) vs trailing on the compute stream.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { tag=$1; shift; env "$@" timeout 150 $TR --master-port $((29700 + RANDOM % 200)) bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu --no-stages --no-hxw > gpurun_out/bench_dpmode_$tag.json 2> gpurun_out/bench_dpmode_$tag.err; }
run overlap B200SHT_DP_MODE=overlap
run trailing B200SHT_DP_MODE=trailing
run overlap16 B200SHT_DP_MODE=overlap B200SHT_OVERLAP_SMS=16
run overlap_nores B200SHT_DP_MODE=overlap B200SHT_OVERLAP_SMS=0 B200SHT_DP_MAXCTAS=16
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_dpmode_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), d["config"].get("dp_allreduce"))
    except Exception as e:
        print(f, "unreadable:", e, open(f.replace(".json", ".err")).read()[-400:])
PY
