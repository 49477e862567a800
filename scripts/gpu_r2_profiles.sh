#!/bin/bash
# round 2: committed ncu evidence.  (1) metrics pass over EVERY launch of one SD1.5 batched-CFG forward, one VAE decode and one full-width
# Flux block pair (duration, DRAM bytes, tensor-pipe %, occupancy, L2 hit rate per kernel); (2) ncu --set full of one launch of each
# non-GEMM kernel the north star names + the pair GEMM
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct
for c in sd15x2 vae flux1; do
  GGML_B200_CUDA_GRAPHS=0 timeout 400 ncu --clock-control none --csv --log-file gpurun_out/r2p_metrics_$c.csv --metrics $M python scripts/one_forward.py $c 2 > gpurun_out/r2p_metrics_$c.log 2>&1
  tail -1 gpurun_out/r2p_metrics_$c.log | cut -c1-150
done
for k in k_flash_attn k_to_nhwc_f16 k_gn_stats k_row_norm k_gemv k_rope_rows k_gemm_tc2 k_geglu; do
  case=sd15x2; skip=6
  [ $k = k_rope_rows ] && case=flux1
  GGML_B200_CUDA_GRAPHS=0 timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 2 -o gpurun_out/r2p_full_$k python scripts/one_forward.py $case 2 > gpurun_out/r2p_full_$k.log 2>&1
  ls -la gpurun_out/r2p_full_$k.ncu-rep 2>&1 | cut -c1-120
done
