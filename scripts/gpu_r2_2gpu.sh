#!/bin/bash
# round 2, 2-GPU call: CFG split with the device-side exchange -- bit-identity tests and the bench line at N = 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | head -4
echo "== pair tests"
timeout 900 python -m pytest tests/test_gpu_cfg_split.py -q -m gpu -s 2>&1 | grep -E "passed|failed|rror|assert|FAILED|fused|\{" | tail -20 | tee gpurun_out/r2g2_tests.log
echo "== pair script (sd15)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/cfg_split_pair.py sd15_unet 10 2>&1 | grep "^{" | tee gpurun_out/r2g2_pair.json
echo "== bench N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2g2_bench.json 2> gpurun_out/r2g2_bench.err; tail -3 gpurun_out/r2g2_bench.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2g2_bench.json') if l.startswith('{')][-1])
    print('N=2', d['config']['layout'], round(d['value'],2),'steps/s e2e',round(d['e2e']['value'],2),'ms/step',round(d['ms_per_step'],3),'alt',d['alt_layout']['layout'],round(d['alt_layout']['value'],2), 'scaling', d['scaling'])
    print(d['backend'])
except Exception as e: print('no bench line', e)
PY
