#!/bin/bash
# Short GPU check (every stage has its own tight timeout so a hang costs seconds, not minutes).
# usage: gpu_quick.sh [stage ...]   stages: gemm fa tbo pytest bench ncu   (default: gemm fa pytest bench)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/stable-diffusion.cpp_b200/lib/libggml-b200.so
TBO=oracle/_ref/test-backend-ops
STAGES="${@:-gemm fa pytest bench}"
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
for st in $STAGES; do
  echo "=================== $st"
  case $st in
    gemm)  timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "mul_mat or conv" 2>&1 | tail -15 | tee gpurun_out/q_gemm.log ;;
    fa)    timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "flash or attention" 2>&1 | tail -15 | tee gpurun_out/q_fa.log ;;
    ops)   timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu 2>&1 | tail -25 | tee gpurun_out/q_ops.log ;;
    tbo)   for ops in MUL_MAT FLASH_ATTN_EXT "GROUP_NORM,NORM,RMS_NORM,SOFT_MAX,IM2COL,ADD,MUL,CPY,CONT,CONCAT,UPSCALE,SILU,GELU,SCALE,TIMESTEP_EMBEDDING,PAD,REPEAT,L2_NORM"; do
             timeout 600 $TBO test -b B200_0 -o "$ops" > gpurun_out/tbo_$(echo $ops | cut -c1-12).log 2>&1
             echo "$ops: exit $? $(grep -E 'tests passed' gpurun_out/tbo_$(echo $ops | cut -c1-12).log | tail -1)"
             grep -a FAIL gpurun_out/tbo_$(echo $ops | cut -c1-12).log | sed 's/\x1b\[[0-9;]*m//g' | head -20
           done ;;
    pytest) timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/q_pytest.log ;;
    models) timeout 900 python scripts/model_parity.py 2>&1 | grep -v "^load_backend" | tee gpurun_out/model_parity.log | tail -30 ;;
    bench) timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | grep -v "^load_backend" | tail -5 | tee gpurun_out/bench.log ;;
    benchg) GGML_B200_CUDA_GRAPHS=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep -v "^load_backend" | tail -5 | tee gpurun_out/bench_graphs.log ;;
    ncu)   timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
             python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit $?"; wc -l gpurun_out/launches.csv
           timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 200 -c 3 -o gpurun_out/prof_gemm -f \
             python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?" ;;
  esac
done
