#!/bin/bash
# round 2, GPU call 17: held-back adaLN modulate, RMS x weight, relaxed result-over-activation check, per-tap plan for HBM-resident residuals;
# ncu --set full of the VAE 512 x 512 convolutions and of the epilogue-bound 320-feature Linear
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/stable-diffusion.cpp_b200/lib/libggml-b200.so
echo "== model tests (DiT families, bit identity, vae)"
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_parity_config.py -q -m gpu -x -k "flux or mmdit or wan_1_3b or bit_identical or vae_decode or unet_tiny or t5 or clip" --durations=6 2>&1 | tail -22 | tee gpurun_out/r2c17_models.log
echo "== norm / mul_mat op tests"
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "norm or mul_mat or rope" 2>&1 | tail -4 | tee gpurun_out/r2c17_ops.log
echo "== A/B (same box)"
ab() { env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" | tee -a gpurun_out/r2c17_bench_lines.jsonl | python -c "
import sys,json; d=json.loads(sys.stdin.read()); e=d['extra_workloads']
print('[$*]', round(d['value'],2), 'steps/s e2e', round(d['e2e']['value'],2), 'serial', round(d['alt_layout']['value'],2), 'GEMM', round(d['roofline']['achieved'],1), 'vae', round(d['vae_decode']['value'],2), 'vae1024', round(d['vae_decode']['at_1024']['value'],2), 'sdxl', round(e['sdxl']['forward_ms'],2), 'flux', round(e['flux']['forward_ms'],2), 'launches', d['gpu_launches']//d['steps'], e['flux']['launches_per_forward'], e['sdxl']['launches_per_forward'])" | tee -a gpurun_out/r2c17_ab.log; }
ab X=1
ab GGML_B200_DEFER_MODULATE=0 GGML_B200_RMS_MUL=0
echo "== ncu full: VAE convolutions 40..42 of the decode (512 x 512 x 128, with / without residual)"
GGML_B200_CUDA_GRAPHS=0 timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc2 -s 39 -c 4 -o gpurun_out/r2c17_full_vae_conv -f \
    python scripts/one_forward.py vae 1 > gpurun_out/r2c17_ncu_full_vae.log 2>&1; echo "exit $?"
echo "== ncu full: SD1.5 batched forward, pair GEMMs 60..63"
GGML_B200_CUDA_GRAPHS=0 timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc2 -s 60 -c 4 -o gpurun_out/r2c17_full_sd15_gemm -f \
    python scripts/one_forward.py sd15x2 1 > gpurun_out/r2c17_ncu_full_sd15.log 2>&1; echo "exit $?"
ls -la gpurun_out/*.ncu-rep
