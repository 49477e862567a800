#!/bin/bash
# First GPU call of the next round: everything that was written after the round-1 GPU budget ran out, each stage with its own timeout.
#   1. the gated parity tests (tiled VAE decode, CLIP text encoder, Wan VAE decoder)
#   2. test-backend-ops IM2COL_3D (supports_op says yes since round 1; the op group list of tests/test_gpu_backend_ops.py does not include it yet)
#   3. tools/gemm_bench with the experimental persistent GEMM compared element by element against the regular kernel
#   4. bench A/B of GGML_B200_FOLD_BATCH and GGML_B200_PERSISTENT on the SD1.5 step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/stable-diffusion.cpp_b200/lib/libggml-b200.so
echo "== gated tests"; SDB200_UNVALIDATED=1 timeout 300 python -m pytest tests/test_gpu_models.py -q -m gpu -k "tiled_vae or clip_text or wan_vae" 2>&1 | tail -15 | tee gpurun_out/r2_gated.log
echo "== tbo IM2COL_3D"; timeout 200 oracle/_ref/test-backend-ops test -b B200_0 -o IM2COL_3D 2>&1 | sed 's/\x1b\[[0-9;]*m//g' | grep -E "FAIL|tests passed|Backend" | tail -8 | tee gpurun_out/r2_tbo_im2col3d.log
echo "== gemm_bench persistent"; (cd stable-diffusion.cpp_b200/csrc && make gemm_bench >/dev/null 2>&1); GEMM_BENCH_PERSISTENT=1 timeout 120 stable-diffusion.cpp_b200/lib/gemm_bench 20 2>&1 | cut -c1-230 | tee gpurun_out/r2_gemm_bench.log
echo "== bench A/B"
for v in "" "GGML_B200_FOLD_BATCH=1" "GGML_B200_PERSISTENT=1"; do
  env $v timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-vae --no-alt 2>&1 | grep "^{" | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$v]', round(d['value'],2), 'steps/s, GEMM', round(d['roofline']['achieved'],1), 'TFLOP/s')" | tee -a gpurun_out/r2_ab.log
done
