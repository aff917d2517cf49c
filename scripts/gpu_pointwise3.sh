#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pointwise.py tests/test_gpu_sfno.py -q -m gpu --timeout=600 2>&1 | tail -4 | cut -c1-260 > gpurun_out/pointwise3_pytest.log
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --workload sfno_sc3_layers8_edim384 > gpurun_out/bench_cfg3_fused.json 2> gpurun_out/pointwise.err
timeout 200 python scripts/prof_model.py > gpurun_out/model_profile_tf32.log 2>&1
cat gpurun_out/pointwise3_pytest.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_cfg3_fused.json").read().strip().splitlines()[-1])
print("cfg3", round(d["value"], 2), "samples/s", round(d["ms_per_step"], 3), "ms; e2e", round(d["e2e"]["value"], 2), d["clocks"])
PY
head -3 gpurun_out/model_profile_tf32.log | tail -1; grep norm_kernel gpurun_out/model_profile_tf32.log | cut -c1-120
tail -3 gpurun_out/pointwise.err
