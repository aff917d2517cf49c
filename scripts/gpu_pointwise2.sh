#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pointwise.py tests/test_gpu_sfno.py -q -m gpu --timeout=600 -k "conv1x1 or sfno or 16-24" 2>&1 | tail -6 | cut -c1-260 > gpurun_out/pointwise2_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --workload sfno_sc3_layers8_edim384 > gpurun_out/bench_cfg3_fused.json 2> gpurun_out/pointwise.err
B200SHT_FUSED_POINTWISE=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --workload sfno_sc3_layers8_edim384 > gpurun_out/bench_cfg3_unfused.json 2>> gpurun_out/pointwise.err
timeout 200 python scripts/prof_model.py > gpurun_out/model_profile_tf32.log 2>&1
cat gpurun_out/pointwise2_pytest.log
python - <<'PY'
import json
for f in ["bench_cfg3_fused", "bench_cfg3_unfused"]:
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), "samples/s", round(d["ms_per_step"], 3), "ms; e2e", round(d["e2e"]["value"], 2), "launches", d.get("gpu_launches"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
head -30 gpurun_out/model_profile_tf32.log | cut -c1-190
tail -4 gpurun_out/pointwise.err
