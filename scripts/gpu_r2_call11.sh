#!/bin/bash
# round 2, GPU call 11: f16-only attention result at batch 1; cold-weight GEMM measurements and the L2 slab prefetch experiment
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/stable-diffusion.cpp_b200/lib/libggml-b200.so
echo "== targeted: bit identity"
timeout 600 python -m pytest tests/test_gpu_models.py -q -m gpu -k "bit_identical" 2>&1 | tail -12 | tee gpurun_out/r2c11_targeted.log
echo "== gemm_bench: weights from HBM (cold) vs warm, with / without the up-front L2 request"
GEMM_BENCH_COLD=1 GEMM_BENCH_PF=1 GEMM_BENCH_ONLY="conv 16x16,conv 8x8,conv 32x32 1920,conv 32x32 640,linear geglu L256,linear ff-out L256,linear qkv L256,linear geglu L1024,flux qkv" \
  timeout 300 stable-diffusion.cpp_b200/lib/gemm_bench 20 2>&1 | grep -v "single" | tee gpurun_out/r2c11_gemm_cold.log
echo "== A/B (same box)"
ab() { env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" | python -c "
import sys,json; d=json.loads(sys.stdin.read()); e=d['extra_workloads']
print('[$*]', round(d['value'],2), 'steps/s e2e', round(d['e2e']['value'],2), 'serial', round(d['alt_layout']['value'],2), 'GEMM', round(d['roofline']['achieved'],1), 'vae', round(d['vae_decode']['value'],2), 'vae1024', round(d['vae_decode']['at_1024']['value'],2), 'sdxl', round(e['sdxl']['forward_ms'],2), 'flux', round(e['flux']['forward_ms'],2), 'launches', d['gpu_launches']//d['steps'])" | tee -a gpurun_out/r2c11_ab.log; }
ab X=1
ab GGML_B200_WPREFETCH=1
ab X=2
