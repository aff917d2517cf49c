#!/bin/bash
# Latitude-chunked (DFT analysis -> Legendre analysis) experiment + the new consumer / TF32-attention tests.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_consumers.py tests/test_gpu_cabi.py "tests/test_gpu_parity.py::test_distributed_local_stages_cuda_subplans" \
  "tests/test_gpu_parity.py::test_distributed_modules_world1_and_dense_conv" tests/test_gpu_bench_configs.py -k "not 384" -m gpu -q --timeout=600 -s 2>&1 \
  | grep -E "chunked|passed|failed|Error|error|assert|FAILED|\[parity\] (spectral|noise|SpectralAttention)" | cut -c1-260 | tail -40 > gpurun_out/chunks_pytest.log
B200SHT_LAT_CHUNKS_SYN=3 timeout 600 python -m pytest "tests/test_gpu_bench_configs.py::test_benched_block_fp32_activations_tf32" tests/test_gpu_cabi.py -m gpu -q --timeout=600 -s 2>&1 \
  | grep -E "benched|passed|failed|Error|error|assert|FAILED" | cut -c1-260 | tail -8 > gpurun_out/chunks_syn_pytest.log
for n in 2 3; do
  B200SHT_LAT_CHUNKS_SYN=$n timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --no-stages > gpurun_out/bench_chunks_syn_$n.json 2>> gpurun_out/chunks.err
done
B200SHT_LAT_CHUNKS=3 B200SHT_LAT_CHUNKS_SYN=3 timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --no-stages > gpurun_out/bench_chunks_both_3.json 2>> gpurun_out/chunks.err
for n in 1 2 3 4; do
  B200SHT_LAT_CHUNKS=$n timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --no-stages > gpurun_out/bench_chunks_$n.json 2>> gpurun_out/chunks.err
done
for n in 1 3; do
  B200SHT_LAT_CHUNKS=$n timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-stages --workload sfno_block_240x480x384 > gpurun_out/bench2a_chunks_$n.json 2>> gpurun_out/chunks.err
done
for n in 1 8; do
  B200SHT_LAT_CHUNKS=$n timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-stages --workload sfno_block_721to240x384 > gpurun_out/bench2b_chunks_$n.json 2>> gpurun_out/chunks.err
done
cat gpurun_out/chunks_pytest.log
echo '== syn chunks'; cat gpurun_out/chunks_syn_pytest.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench*_chunks_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "samples/s", round(d["ms_per_step"], 4), "ms")
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -5 gpurun_out/chunks.err
