#!/bin/bash
# 8-GPU e2e check (no mode comparison): default 8 x 1 M reads
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 8 --steps 3 --warmup 3 --no-verify > gpurun_out/bench_n8_quick.json 2> gpurun_out/bench_n8_quick.err
tail -2 gpurun_out/bench_n8_quick.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n8_quick.json'))
print('value', d['value'], d['ms_per_step'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('pinned_h2d_gb_per_s_per_gpu_all_ranks_copying'), d['e2e']['ms_per_step_floor_from_h2d'])
print('e2e_svb16', d['e2e_svb16']['value'], d['e2e_svb16']['ms_per_step'])
PY
