#!/bin/bash
# round 2, GPU call 15: fitted halo plan model (bench), new tests (left-padded mask, sched fallback), WARM launch lists (ncu, caches kept)
# of the batched SD1.5 forward and the 512x512 VAE decode
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/stable-diffusion.cpp_b200/lib/libggml-b200.so
echo "== new tests"
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -q -m gpu -k "left_padded or sched_fallback or flash_attn" 2>&1 | tail -25 | tee gpurun_out/r2c15_tests.log
echo "== bench (fitted halo model)"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" | tee gpurun_out/r2c15_bench.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); e=d['extra_workloads']
print(round(d['value'],2), 'steps/s e2e', round(d['e2e']['value'],2), 'serial', round(d['alt_layout']['value'],2), 'GEMM', round(d['roofline']['achieved'],1), 'vae', round(d['vae_decode']['value'],2), 'vae1024', round(d['vae_decode']['at_1024']['value'],2), 'sdxl', round(e['sdxl']['forward_ms'],2), 'flux', round(e['flux']['forward_ms'],2), 'launches', d['gpu_launches']//d['steps'])" | tee gpurun_out/r2c15_ab.log
echo "== conv plans picked"
GGML_B200_GEMM_LOG=1 GGML_B200_CUDA_GRAPHS=0 timeout 120 python scripts/one_forward.py sd15x2 1 2>&1 | grep "GEMMLOG" | sed 's/model1.*//' | sort | uniq -c | sort -rn > gpurun_out/r2c15_gemmlog_sd15x2.txt; wc -l gpurun_out/r2c15_gemmlog_sd15x2.txt
GGML_B200_GEMM_LOG=1 GGML_B200_CUDA_GRAPHS=0 timeout 120 python scripts/one_forward.py vae 1 2>&1 | grep "GEMMLOG" | sed 's/model1.*//' | sort | uniq -c | sort -rn > gpurun_out/r2c15_gemmlog_vae.txt
echo "== warm launch lists"
GGML_B200_CUDA_GRAPHS=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 4000 --csv --log-file gpurun_out/r2c15_launches_sd15x2_warm.csv \
    python scripts/one_forward.py sd15x2 3 > gpurun_out/r2c15_ncu_sd15.log 2>&1; echo "exit $?"; wc -l gpurun_out/r2c15_launches_sd15x2_warm.csv
GGML_B200_CUDA_GRAPHS=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1000 --csv --log-file gpurun_out/r2c15_launches_vae_warm.csv \
    python scripts/one_forward.py vae 3 > gpurun_out/r2c15_ncu_vae.log 2>&1; echo "exit $?"; wc -l gpurun_out/r2c15_launches_vae_warm.csv
