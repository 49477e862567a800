#!/bin/bash
# round 2, GPU call 3: loopback CFG-split exchange, 3xTF32, arbiter; in-situ tuning of the CTA-pair kernel (ncu launch lists joined
# with the dispatcher's shape log); ncu --set full of the pair kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/stable-diffusion.cpp_b200/lib/libggml-b200.so
echo "== tests"
timeout 600 python -m pytest tests/test_gpu_cfg_split.py tests/test_gpu_ops.py tests/test_gpu_parity_config.py -q -m gpu -s -k "loopback or mul_mat or truth or weight or no_cuda_core" 2>&1 | grep -E "passed|failed|rel_l2|truth|Error|error|assert|FAILED" | tail -30 | tee gpurun_out/r2c3_tests.log
echo "== tune sd15x2"
: > gpurun_out/r2c3_tune.txt
run() {   # label, env...
  label=$1; shift
  env "$@" GGML_B200_CUDA_GRAPHS=0 GGML_B200_GEMM_LOG=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/_tune.csv \
      python scripts/one_forward.py $CASE 2 > gpurun_out/_tune.out 2> gpurun_out/_tune.err
  python scripts/tune_join.py gpurun_out/_tune.err gpurun_out/_tune.csv "$CASE:$label" >> gpurun_out/r2c3_tune.txt
  grep "forward 1" gpurun_out/_tune.out | cut -c1-120
}
CASE=sd15x2
run one GGML_B200_GEMM2=0
run model GGML_B200_GEMM2=2
for bn in 256 128; do for sp in 1 2 4; do run "bn${bn}s${sp}" GGML_B200_GEMM2=2 GGML_B200_GEMM2_BN=$bn GGML_B200_GEMM2_SPLITS=$sp; done; done
run bn64s1 GGML_B200_GEMM2=2 GGML_B200_GEMM2_BN=64 GGML_B200_GEMM2_SPLITS=1
run bn64s4 GGML_B200_GEMM2=2 GGML_B200_GEMM2_BN=64 GGML_B200_GEMM2_SPLITS=4
CASE=sd15
run one GGML_B200_GEMM2=0
run model GGML_B200_GEMM2=2
for bn in 256 128; do for sp in 2 4; do run "bn${bn}s${sp}" GGML_B200_GEMM2=2 GGML_B200_GEMM2_BN=$bn GGML_B200_GEMM2_SPLITS=$sp; done; done
CASE=vae
run one GGML_B200_GEMM2=0
run model GGML_B200_GEMM2=2
run bn128s1 GGML_B200_GEMM2=2 GGML_B200_GEMM2_BN=128 GGML_B200_GEMM2_SPLITS=1
CASE=sdxl
run one GGML_B200_GEMM2=0
run model GGML_B200_GEMM2=2
wc -l gpurun_out/r2c3_tune.txt
echo "== ncu full of the pair kernel (3 launches of the sd15x2 forward)"
GGML_B200_GEMM2=2 GGML_B200_CUDA_GRAPHS=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc2 -s 300 -c 4 -o gpurun_out/r2c3_pair python scripts/one_forward.py sd15x2 2 > gpurun_out/r2c3_ncu_full.log 2>&1
ls -la gpurun_out/r2c3_pair.ncu-rep
rm -f gpurun_out/_tune.csv gpurun_out/_tune.out gpurun_out/_tune.err
