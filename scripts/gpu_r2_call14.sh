#!/bin/bash
# round 2, GPU call 14: halo-reuse 3x3 convolution on hardware -- gemm_bench conv sweep (halo vs per-tap, element compare), op tests,
# model parity at config size, same-box A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/stable-diffusion.cpp_b200/lib/libggml-b200.so
echo "== gemm_bench conv sweep"
GEMM_BENCH_CONV=1 timeout 240 stable-diffusion.cpp_b200/lib/gemm_bench 20 > gpurun_out/r2c14_conv_sweep.log 2>&1; echo "exit $?"
grep -c "" gpurun_out/r2c14_conv_sweep.log; grep "dispatcher\|FAILED\|status" gpurun_out/r2c14_conv_sweep.log
awk '$NF+0 > 1e-3 && /\|/ {print "DIFF", $0}' gpurun_out/r2c14_conv_sweep.log | head
echo "== conv op tests"
GGML_B200_GEMM_LOG=1 timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "conv or resblock or upsample" 2> gpurun_out/r2c14_gemmlog.txt | tail -6 | tee gpurun_out/r2c14_ops.log
grep -c "taps 9" gpurun_out/r2c14_gemmlog.txt; grep -c "taps 3" gpurun_out/r2c14_gemmlog.txt
if grep -q "failed\|rror" gpurun_out/r2c14_ops.log; then echo "!! conv tests failed: halo mode off for the rest of this call"; export GGML_B200_CONV_HALO=0; fi
echo "== model parity (vae 512, sd15, bit identity, cfg split loopback)"
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_parity_config.py tests/test_gpu_cfg_split.py -q -m gpu -k "bit_identical or vae or arbiter or loopback or sd15_unet_vs or sdxl" 2>&1 | tail -12 | tee gpurun_out/r2c14_models.log
echo "== A/B (same box)"
ab() { env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" | tee -a gpurun_out/r2c14_bench_lines.jsonl | python -c "
import sys,json; d=json.loads(sys.stdin.read()); e=d['extra_workloads']
print('[$*]', round(d['value'],2), 'steps/s e2e', round(d['e2e']['value'],2), 'serial', round(d['alt_layout']['value'],2), 'GEMM', round(d['roofline']['achieved'],1), 'vae', round(d['vae_decode']['value'],2), 'vae1024', round(d['vae_decode']['at_1024']['value'],2), 'sdxl', round(e['sdxl']['forward_ms'],2), 'flux', round(e['flux']['forward_ms'],2), 'launches', d['gpu_launches']//d['steps'])" | tee -a gpurun_out/r2c14_ab.log; }
ab X=1
ab GGML_B200_CONV_HALO=0
