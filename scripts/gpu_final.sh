#!/bin/bash
# Round-end validation on one B200: smoke, bench (both arms), ncu metric passes over every launch of the SD1.5 batched forward / VAE decode /
# Flux block pair, ncu --set full of the kernels that changed, NHWC transform A/B, full GPU test suite.  Every stage has its own timeout.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/stable-diffusion.cpp_b200/lib/libggml-b200.so
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
echo "== smoke";  timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^load_backend" | tail -2 | tee gpurun_out/f_smoke.log
echo "== bench";  timeout 400 python bench.py --steps 20 --warmup 3 2>&1 | grep "^{" | tee gpurun_out/r02_final_bench.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); e=d['extra_workloads']
print(round(d['value'],2), 'steps/s e2e', round(d['e2e']['value'],2), 'serial', round(d['alt_layout']['value'],2), 'GEMM', round(d['roofline']['achieved'],1), 'vae', round(d['vae_decode']['value'],2), 'vae1024', round(d['vae_decode']['at_1024']['value'],2), 'sdxl', round(e['sdxl']['forward_ms'],2), 'flux', round(e['flux']['forward_ms'],2), 'launches', d['gpu_launches']//d['steps'], 'cpu', d['cpu_baseline'])"
echo "== ref";    timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | grep "^{" | tee gpurun_out/r02_final_bench_ref.json | cut -c1-300
echo "== NHWC transform A/B (NSUB 2)"
GGML_B200_NHWC_NSUB=2 timeout 120 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "conv or resblock or upsample" 2>&1 | tail -2
for v in 4 2; do GGML_B200_NHWC_NSUB=$v timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-alt --extra none 2>&1 | grep "^{" | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('[NSUB=$v]', round(d['value'],2), 'vae', round(d['vae_decode']['value'],3), 'vae1024', round(d['vae_decode']['at_1024']['value'],3))" | tee -a gpurun_out/r02_final_nsub_ab.log; done
echo "== ncu metric passes"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct
for c in sd15x2 vae flux1; do
  GGML_B200_CUDA_GRAPHS=0 timeout 300 ncu --clock-control none --csv --log-file gpurun_out/r02_final_metrics_$c.csv --metrics $M python scripts/one_forward.py $c 2 > gpurun_out/r02_final_metrics_$c.log 2>&1
  tail -1 gpurun_out/r02_final_metrics_$c.log | cut -c1-150
done
echo "== ncu full"
for k in k_gemm_tc2 k_to_nhwc_f16 k_row_norm_warp k_flash_attn; do
  GGML_B200_CUDA_GRAPHS=0 timeout 150 ncu --set full --clock-control none --import-source on -k regex:$k -s 20 -c 2 -o gpurun_out/r02_final_full_$k -f python scripts/one_forward.py sd15x2 1 > gpurun_out/r02_final_full_$k.log 2>&1
  ls -la gpurun_out/r02_final_full_$k.ncu-rep 2>&1 | cut -c1-120
done
echo "== pytest"; timeout 800 python -m pytest tests -q -m gpu --durations=8 2>&1 | tail -16 | tee gpurun_out/f_pytest.log
