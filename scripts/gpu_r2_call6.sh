#!/bin/bash
# round 2, GPU call 6: ncu --set full of the pair kernel on one output-heavy and one K-heavy shape (stall reasons, store path), 2-GPU tests are separate
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for shape in "geglu L4096" "ff-out L1024" "conv 32x32 640"; do
  tag=$(echo "$shape" | tr ' ' '_' | tr -d '>-')
  GEMM_BENCH_ONLY="$shape" GEMM_BENCH_PAIR=1 GEMM_BENCH_FEW=1 timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc2 -s 4 -c 2 -o gpurun_out/r2c6_$tag stable-diffusion.cpp_b200/lib/gemm_bench 3 > gpurun_out/r2c6_$tag.log 2>&1
  ls -la gpurun_out/r2c6_$tag.ncu-rep
done
