#!/bin/bash
# round 2, GPU call 1: round-1 leftovers (gated tests, persistent GEMM validation, A/B) + launch lists of the VAE decode
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_round2_first.sh 2>&1 | tee gpurun_out/r2_first.log
echo "== VAE launch list"
GGML_B200_CUDA_GRAPHS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_vae.csv python scripts/one_forward.py vae 3 > gpurun_out/r2_vae_ncu.log 2>&1
tail -3 gpurun_out/r2_vae_ncu.log
timeout 120 python scripts/one_forward.py vae 4 2>&1 | tail -4 | tee gpurun_out/r2_vae.log
timeout 200 python scripts/one_forward.py vae128 3 2>&1 | tail -4 | tee gpurun_out/r2_vae128.log
