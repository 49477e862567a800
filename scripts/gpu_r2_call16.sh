#!/bin/bash
# round 2, GPU call 16: GEGLU projection epilogue, batched / prefetched residual reads, fast-SiLU NHWC transform, 32 KB GroupNorm chunks,
# warp-per-row norms: op tests, model bit-identity / parity, same-box A/B, Flux block launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/stable-diffusion.cpp_b200/lib/libggml-b200.so
echo "== op tests"
timeout 400 python -m pytest tests/test_gpu_ops.py -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/r2c16_ops.log
echo "== model tests (bit identity, vae, sd15, sdxl, tiny)"
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_parity_config.py -q -m gpu -k "bit_identical or vae or arbiter or sd15_unet_vs or unet_tiny or sdxl or flux" 2>&1 | tail -25 | tee gpurun_out/r2c16_models.log
echo "== A/B (same box)"
ab() { env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" | tee -a gpurun_out/r2c16_bench_lines.jsonl | python -c "
import sys,json; d=json.loads(sys.stdin.read()); e=d['extra_workloads']
print('[$*]', round(d['value'],2), 'steps/s e2e', round(d['e2e']['value'],2), 'serial', round(d['alt_layout']['value'],2), 'GEMM', round(d['roofline']['achieved'],1), 'vae', round(d['vae_decode']['value'],2), 'vae1024', round(d['vae_decode']['at_1024']['value'],2), 'sdxl', round(e['sdxl']['forward_ms'],2), 'flux', round(e['flux']['forward_ms'],2), 'launches', d['gpu_launches']//d['steps'])" | tee -a gpurun_out/r2c16_ab.log; }
ab X=1
ab GGML_B200_GEGLU_EPI=0
ab GGML_B200_RES_PREFETCH=0 GGML_B200_NORM_WARP=0
echo "== warm launch lists"
GGML_B200_CUDA_GRAPHS=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 600 --csv --log-file gpurun_out/r2c16_launches_flux1_warm.csv \
    python scripts/one_forward.py flux1 3 > gpurun_out/r2c16_ncu_flux1.log 2>&1; echo "exit $?"; wc -l gpurun_out/r2c16_launches_flux1_warm.csv
GGML_B200_CUDA_GRAPHS=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1000 --csv --log-file gpurun_out/r2c16_launches_vae_warm.csv \
    python scripts/one_forward.py vae 3 > gpurun_out/r2c16_ncu_vae.log 2>&1; echo "exit $?"; wc -l gpurun_out/r2c16_launches_vae_warm.csv
