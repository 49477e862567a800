#!/bin/bash
# round 2, GPU call 10: K/V match fix (ggml_cast self reference), side streams (hoisted context projections, K/V beside Q): targeted tests + same-box A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/stable-diffusion.cpp_b200/lib/libggml-b200.so
echo "== targeted: bit identity, cfg split loopback, sdxl / sd15 parity at config size"
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_cfg_split.py tests/test_gpu_parity_config.py -q -m gpu -k "bit_identical or loopback or sdxl or sd15 or arbiter or batched or unet" 2>&1 | tail -25 | tee gpurun_out/r2c10_targeted.log
echo "== A/B (same box)"
ab() { env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" | python -c "
import sys,json; d=json.loads(sys.stdin.read()); e=d['extra_workloads']
print('[$*]', round(d['value'],2), 'steps/s e2e', round(d['e2e']['value'],2), 'serial', round(d['alt_layout']['value'],2), 'GEMM', round(d['roofline']['achieved'],1), 'vae', round(d['vae_decode']['value'],2), 'vae1024', round(d['vae_decode']['at_1024']['value'],2), 'sdxl', round(e['sdxl']['forward_ms'],2), 'flux', round(e['flux']['forward_ms'],2), 'launches', d['gpu_launches']//d['steps'])" | tee -a gpurun_out/r2c10_ab.log; }
ab X=1
ab GGML_B200_SIDE_STREAMS=0
ab GGML_B200_SIDE_STREAMS=0 GGML_B200_KV_DIRECT=0
ab X=2
