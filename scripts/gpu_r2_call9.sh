#!/bin/bash
# round 2, GPU call 9: attention operands in place (K/V rows from the projection GEMM, split batch, f16-only result), gated-residual epilogue,
# vectorised row norm, multi-row GEMV: targeted tests first (fail fast, per-feature switches), full suite, same-box A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/stable-diffusion.cpp_b200/lib/libggml-b200.so
echo "== targeted: bit identity / flux / sdxl / mmdit"
timeout 900 python -m pytest tests/test_gpu_models.py -q -m gpu -x -k "bit_identical or flux or sdxl or mmdit or wan or t5" 2>&1 | tail -25 | tee gpurun_out/r2c9_targeted.log
echo "== targeted: ops"
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backend_ops.py -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2c9_ops.log
echo "== full gpu suite"
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2c9_pytest.log
echo "== A/B (same box)"
ab() { env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" | python -c "
import sys,json; d=json.loads(sys.stdin.read()); e=d['extra_workloads']
print('[$*]', round(d['value'],2), 'steps/s e2e', round(d['e2e']['value'],2), 'serial', round(d['alt_layout']['value'],2), 'GEMM', round(d['roofline']['achieved'],1), 'vae', round(d['vae_decode']['value'],2), 'vae1024', round(d['vae_decode']['at_1024']['value'],2), 'sdxl', round(e['sdxl']['forward_ms'],2), 'flux', round(e['flux']['forward_ms'],2), 'launches', d['gpu_launches']//d['steps'])" | tee -a gpurun_out/r2c9_ab.log; }
ab X=1
ab GGML_B200_KV_DIRECT=0 GGML_B200_FA_OUT16=0
ab GGML_B200_GATE_FUSION=0
ab GGML_B200_GEMV_RPW=1
ab X=2
