#!/bin/bash
# round 2, GPU call 4: validation of the second batch (loopback exchange, Q8_0 activation round trip, T5, ingest, attention changes),
# producer-thread A/B in the pair GEMM (graph-timed gemm_bench), attention tile A/B on Flux, step A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/stable-diffusion.cpp_b200/lib/libggml-b200.so
echo "== tests"
timeout 900 python -m pytest tests/test_gpu_cfg_split.py tests/test_gpu_ops.py tests/test_gpu_models.py tests/test_gpu_parity_config.py -q -m gpu -s \
  -k "loopback or ops or wan or t5 or clip or sd15_unet_full or flux or mmdit or truth or upload or weight or fixture or batched" 2>&1 | grep -E "passed|failed|rel_l2|truth|rror|assert|FAILED" | tail -40 | tee gpurun_out/r2c4_tests.log
echo "== gemm_bench (graph-timed), 2 producers vs 1"
GEMM_BENCH_PAIR=1 GEMM_BENCH_FEW=1 timeout 200 stable-diffusion.cpp_b200/lib/gemm_bench 20 2>&1 | cut -c1-160 > gpurun_out/r2c4_bench_np2.log; echo rc=$?
GGML_B200_GEMM2_NPROD=1 GEMM_BENCH_PAIR=1 GEMM_BENCH_FEW=1 timeout 200 stable-diffusion.cpp_b200/lib/gemm_bench 20 2>&1 | cut -c1-160 > gpurun_out/r2c4_bench_np1.log; echo rc=$?
grep -c "mismatches 0 " gpurun_out/r2c4_bench_np2.log gpurun_out/r2c4_bench_np1.log; grep "pair bn" gpurun_out/r2c4_bench_np2.log | grep -v "mismatches 0 " | head
echo "== flux attention A/B"
for v in "GGML_B200_FA_BN128=0" "GGML_B200_FA_BN128=1"; do env $v GGML_B200_GEMM2=1 timeout 200 python scripts/one_forward.py flux 3 2>&1 | grep "forward 2" | sed "s/^/[$v] /" | cut -c1-170 | tee -a gpurun_out/r2c4_ab.log; done
echo "== step A/B"
for v in "GGML_B200_GEMM2=0" "GGML_B200_GEMM2=1" "GGML_B200_GEMM2=1 GGML_B200_GEMM2_NPROD=1"; do
  env $v timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --extra none 2>&1 | grep "^{" | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$v]', round(d['value'],2), 'steps/s e2e', round(d['e2e']['value'],2), 'serial', round(d['alt_layout']['value'],2), 'GEMM', round(d['roofline']['achieved'],1), 'TFLOP/s, vae', round(d['vae_decode']['value'],2), 'ms, vae1024', round(d['vae_decode']['at_1024']['value'],2), 'host', d['host'])" | tee -a gpurun_out/r2c4_ab.log
done
