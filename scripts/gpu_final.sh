#!/bin/bash
# Round-end validation on one B200: smoke, bench (both arms), Flux forward, ncu launch list + full capture of the GEMMs, full GPU test suite.
# Every stage has its own timeout so a hang costs seconds.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/stable-diffusion.cpp_b200/lib/libggml-b200.so
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
echo "== smoke";  timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^load_backend" | tail -2 | tee gpurun_out/f_smoke.log
echo "== bench";  timeout 300 python bench.py --steps 20 --warmup 3 2>&1 | grep -v "^load_backend" | tail -2 | tee gpurun_out/f_bench.log
echo "== ref";    timeout 120 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | grep -v "^load_backend" | tail -1 | tee gpurun_out/f_bench_ref.log
echo "== flux";   timeout 200 python scripts/one_forward.py flux 3 2>&1 | grep -v "^load_backend" | tee gpurun_out/f_flux.log
echo "== sdxl";   timeout 100 python scripts/one_forward.py sdxl 3 2>&1 | grep -v "^load_backend" | tee gpurun_out/f_sdxl.log
echo "== ncu list"; GGML_B200_CUDA_GRAPHS=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_final_batched.csv \
    python scripts/one_forward.py sd15x2 3 > gpurun_out/f_ncu_list.log 2>&1; echo "exit $?"; wc -l gpurun_out/launches_final_batched.csv
echo "== ncu full"; GGML_B200_CUDA_GRAPHS=0 timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 400 -c 10 -o gpurun_out/prof_gemm_final -f \
    python scripts/one_forward.py sd15x2 3 > gpurun_out/f_ncu_full.log 2>&1; echo "exit $?"
echo "== pytest"; timeout 900 python -m pytest tests -q -m gpu --durations=8 2>&1 | tail -30 | tee gpurun_out/f_pytest.log
