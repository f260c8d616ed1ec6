#!/bin/bash
# 8-GPU call: topology, the default bench (1 M reads per GPU, what the driver's scaling run uses) and
# the literal configs[4] line (8 x 1.25 M = 10 M reads per step)
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
(nproc; lscpu | grep -i numa; free -g | head -2) >> gpurun_out/topo.txt 2>&1
for cfg in "n8:1000000" "n8_10M:1250000"; do
  name=${cfg%%:*}; reads=${cfg##*:}
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 8 --steps 3 --warmup 3 --reads $reads $( [ $name = n8_10M ] && echo --no-verify ) > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  tail -2 gpurun_out/bench_$name.err
  python tools/bench_brief.py gpurun_out/bench_$name.json 2>/dev/null | head -2
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$name.json'))
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('pinned_h2d_gb_per_s_per_gpu_all_ranks_copying'))
print('e2e_svb16', d['e2e_svb16']['value'], d['e2e_svb16']['ms_per_step'])
PY
done
