#!/bin/bash
# round 2, GPU call 7: full GPU suite on the current tree + step / VAE / SDXL / Flux numbers + a 3-metric ncu pass over ONE SD1.5 forward
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/stable-diffusion.cpp_b200/lib/libggml-b200.so
echo "== gemm_bench: bulk-store epilogue on / off"
GEMM_BENCH_PAIR=1 GEMM_BENCH_FEW=1 timeout 200 stable-diffusion.cpp_b200/lib/gemm_bench 20 2>&1 | cut -c1-160 > gpurun_out/r2c7_bench_tma1.log; echo rc=$?
GGML_B200_GEMM2_TMA_STORE=0 GEMM_BENCH_PAIR=1 GEMM_BENCH_FEW=1 timeout 200 stable-diffusion.cpp_b200/lib/gemm_bench 20 2>&1 | cut -c1-160 > gpurun_out/r2c7_bench_tma0.log; echo rc=$?
ok=$(grep -c "mismatches 0 " gpurun_out/r2c7_bench_tma1.log); all=$(grep -c "pair bn" gpurun_out/r2c7_bench_tma1.log); echo "bulk-store epilogue: $ok / $all exact"
if [ "$ok" != "$all" ] || [ "$all" -lt 50 ]; then echo "!! bulk-store epilogue disabled for the rest of this call"; export GGML_B200_GEMM2_TMA_STORE=0; fi
echo "== full gpu suite"
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/r2c7_pytest.log
echo "== smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2c7_bench.json') if l.startswith('{')][-1])
e=d['extra_workloads']
print(round(d['value'],2),'steps/s e2e',round(d['e2e']['value'],2),'serial',round(d['alt_layout']['value'],2),'roofline',round(d['roofline']['achieved'],1),'TF/s frac',round(d['roofline']['frac'],3),
      'vae',round(d['vae_decode']['value'],2),'ms vae1024',round(d['vae_decode']['at_1024']['value'],2),'sdxl',round(e['sdxl']['forward_ms'],2),'flux',round(e['flux']['forward_ms'],2),'launches/step',d['gpu_launches']/d['steps'])
print(d['host']); print(d['backend'])
PY
echo "== ncu metrics, one sd15x2 forward"
GGML_B200_CUDA_GRAPHS=0 timeout 600 ncu --clock-control none --csv --log-file gpurun_out/r2c7_metrics_sd15x2.csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active python scripts/one_forward.py sd15x2 1 > gpurun_out/r2c7_metrics.log 2>&1
tail -1 gpurun_out/r2c7_metrics.log | cut -c1-150
