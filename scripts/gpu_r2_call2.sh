#!/bin/bash
# round 2, GPU call 2: CTA-pair GEMM validation (gemm_bench sweep, element-exact vs the one-CTA kernel), the new parity tests, A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/stable-diffusion.cpp_b200/lib/libggml-b200.so
echo "== gemm_bench pair sweep"
GEMM_BENCH_PAIR=1 timeout 240 stable-diffusion.cpp_b200/lib/gemm_bench 20 2>&1 | cut -c1-200 > gpurun_out/r2c2_gemm_bench.log; echo "rc=$?"; grep -c "mismatches 0 " gpurun_out/r2c2_gemm_bench.log; grep -v "mismatches 0 " gpurun_out/r2c2_gemm_bench.log | grep "pair bn" | head -20
echo "== new parity tests (default dispatch: pair kernel off)"
timeout 900 python -m pytest tests/test_gpu_parity_config.py tests/test_gpu_models.py -q -m gpu -x -s 2>&1 | grep -E "passed|failed|rel_l2|truth|Error|error|assert" | tail -40 | tee gpurun_out/r2c2_parity.log
echo "== op + model tests with the pair kernel forced"
GGML_B200_GEMM2=2 timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py tests/test_gpu_parity_config.py -q -m gpu -s 2>&1 | grep -E "passed|failed|rel_l2|truth|Error|error|assert|FAILED" | tail -40 | tee gpurun_out/r2c2_forced.log
echo "== bench A/B"
for v in "GGML_B200_GEMM2=0" "GGML_B200_GEMM2=1" "GGML_B200_GEMM2=2"; do
  env $v timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-alt 2>&1 | grep "^{" | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$v]', round(d['value'],2), 'steps/s, GEMM', round(d['roofline']['achieved'],1), 'TFLOP/s, vae', round(d['vae_decode']['value'],2), 'ms')" | tee -a gpurun_out/r2c2_ab.log
done
for v in "GGML_B200_GEMM2=0" "GGML_B200_GEMM2=1"; do
  env $v timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-alt --no-vae --extra sdxl,flux 2>&1 | grep "^{" | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); e=d['extra_workloads']; print('[$v] sdxl', round(e['sdxl']['forward_ms'],2), 'ms; flux', round(e['flux']['forward_ms'],2), 'ms', round(e['flux']['tensor_tflops'],1), 'TFLOP/s')" | tee -a gpurun_out/r2c2_ab.log
done
