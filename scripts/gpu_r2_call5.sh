#!/bin/bash
# round 2, GPU call 5: staged epilogue A/B, recalibrated plan model (pair kernel on by default), loopback exchange
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/stable-diffusion.cpp_b200/lib/libggml-b200.so
echo "== gemm_bench staged epilogue on / off"
GEMM_BENCH_PAIR=1 GEMM_BENCH_FEW=1 timeout 200 stable-diffusion.cpp_b200/lib/gemm_bench 20 2>&1 | cut -c1-160 > gpurun_out/r2c5_bench_vec1.log; echo rc=$?
GGML_B200_GEMM2_VEC_EPI=0 GEMM_BENCH_PAIR=1 GEMM_BENCH_FEW=1 timeout 200 stable-diffusion.cpp_b200/lib/gemm_bench 20 2>&1 | cut -c1-160 > gpurun_out/r2c5_bench_vec0.log; echo rc=$?
grep -c "mismatches 0 " gpurun_out/r2c5_bench_vec1.log gpurun_out/r2c5_bench_vec0.log; grep "pair bn" gpurun_out/r2c5_bench_vec1.log | grep -v "mismatches 0 " | head
echo "== tests"
timeout 900 python -m pytest tests/test_gpu_cfg_split.py tests/test_gpu_ops.py tests/test_gpu_models.py tests/test_gpu_parity_config.py -q -m gpu -s \
  -k "loopback or ops or wan or sd15_unet_full or flux or mmdit or truth or fixture or batched or bit_identical or vae" 2>&1 | grep -E "passed|failed|rel_l2|truth|rror|assert|FAILED|fused" | tail -30 | tee gpurun_out/r2c5_tests.log
echo "== step A/B"
for v in "GGML_B200_GEMM2_VEC_EPI=1" "GGML_B200_GEMM2_VEC_EPI=0" "GGML_B200_GEMM2=0"; do
  env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); e=d['extra_workloads']; print('[$v]', round(d['value'],2), 'steps/s e2e', round(d['e2e']['value'],2), 'serial', round(d['alt_layout']['value'],2), 'GEMM', round(d['roofline']['achieved'],1), 'TFLOP/s, vae', round(d['vae_decode']['value'],2), 'ms, vae1024', round(d['vae_decode']['at_1024']['value'],2), 'sdxl', round(e['sdxl']['forward_ms'],2), 'flux', round(e['flux']['forward_ms'],2), 'pairs', d['backend']['cta_pair_gemm_launches'])" | tee -a gpurun_out/r2c5_ab.log
done
