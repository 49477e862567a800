#!/bin/bash
# First GPU bring-up: op-level parity through the reference's own test-backend-ops harness, then a
# whole-model CPU-vs-B200 comparison through the synthetic-weight harness.  Writes to gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/stable-diffusion.cpp_b200/lib/libggml-b200.so
TBO=oracle/_ref/test-backend-ops
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Flags" | cut -c1-400 > gpurun_out/cpu.txt
OPS_BASIC="ADD,SUB,MUL,DIV,SCALE,CLAMP,SQR,SQRT,SIN,COS,LOG,LEAKY_RELU,CPY,CONT,DUP,CONCAT,REPEAT,PAD,UPSCALE,TIMESTEP_EMBEDDING,GET_ROWS,ARANGE,FILL,SUM_ROWS,MEAN,GROUP_NORM,NORM,RMS_NORM,L2_NORM,SOFT_MAX,IM2COL,GLU,GELU,SILU,RELU,SIGMOID,TANH,EXP,GELU_QUICK,NEG,ABS,GELU_ERF,HARDSWISH,HARDSIGMOID,STEP,SGN,ELU,FLOOR,CEIL,ROUND,TRUNC,EXPM1,SOFTPLUS,REGLU,GEGLU,SWIGLU,GEGLU_ERF,GEGLU_QUICK"
echo "== basic ops (no tensor cores involved)" | tee gpurun_out/tbo_basic.log
timeout 900 $TBO test -b B200_0 -o "$OPS_BASIC" >> gpurun_out/tbo_basic.log 2>&1
echo "exit $?" >> gpurun_out/tbo_basic.log
grep -E "tests passed|FAIL|Backend B200" gpurun_out/tbo_basic.log | tail -5
echo "== MUL_MAT with the CUDA-core reference GEMM" | tee gpurun_out/tbo_mm_ref.log
GGML_B200_TC_GEMM=0 timeout 900 $TBO test -b B200_0 -o MUL_MAT >> gpurun_out/tbo_mm_ref.log 2>&1
echo "exit $?" >> gpurun_out/tbo_mm_ref.log
grep -E "tests passed|Backend B200" gpurun_out/tbo_mm_ref.log | tail -3
echo "== MUL_MAT with the tcgen05 GEMM" | tee gpurun_out/tbo_mm_tc.log
timeout 600 $TBO test -b B200_0 -o MUL_MAT >> gpurun_out/tbo_mm_tc.log 2>&1
echo "exit $?" >> gpurun_out/tbo_mm_tc.log
grep -E "tests passed|Backend B200" gpurun_out/tbo_mm_tc.log | tail -3
grep -c FAIL gpurun_out/tbo_mm_tc.log
echo "== FLASH_ATTN_EXT" | tee gpurun_out/tbo_fa.log
timeout 900 $TBO test -b B200_0 -o FLASH_ATTN_EXT >> gpurun_out/tbo_fa.log 2>&1
echo "exit $?" >> gpurun_out/tbo_fa.log
grep -E "tests passed|Backend B200" gpurun_out/tbo_fa.log | tail -3
echo "== whole-model parity"
timeout 1500 python scripts/model_parity.py "$@" 2>&1 | grep -v "^load_backend" | tee gpurun_out/model_parity.log | tail -40
