#!/bin/bash
# round 2, GPU call 8: V consumed in place by the attention kernel, 16-bit-only MLP epilogue, emb-ADD fold, wider NHWC tiles: tests + same-box A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/stable-diffusion.cpp_b200/lib/libggml-b200.so
echo "== attention op test with V in place"
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "flash or attention" 2>&1 | tail -4 | tee gpurun_out/r2c8_fa.log
if grep -q failed gpurun_out/r2c8_fa.log; then echo "!! V-in-place attention disabled for the rest of this call"; export GGML_B200_FA_VMN=0; fi
echo "== full gpu suite"
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12 | tee gpurun_out/r2c8_pytest.log
echo "== A/B (same box)"
ab() { env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" | python -c "
import sys,json; d=json.loads(sys.stdin.read()); e=d['extra_workloads']
print('[$*]', round(d['value'],2), 'steps/s e2e', round(d['e2e']['value'],2), 'serial', round(d['alt_layout']['value'],2), 'GEMM', round(d['roofline']['achieved'],1), 'vae', round(d['vae_decode']['value'],2), 'vae1024', round(d['vae_decode']['at_1024']['value'],2), 'sdxl', round(e['sdxl']['forward_ms'],2), 'flux', round(e['flux']['forward_ms'],2), 'launches', d['gpu_launches']//d['steps'])" | tee -a gpurun_out/r2c8_ab.log; }
ab X=1
ab GGML_B200_FA_VMN=0
ab GGML_B200_D16=0
ab GGML_B200_CHAIN_FUSION=0
ab X=2
