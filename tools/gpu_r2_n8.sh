#!/bin/bash
# 8-GPU call: topology, then the literal configs[4] line (8 x 1.25 M = 10 M reads per step)
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
(nproc; cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null; lscpu | grep -i numa; free -g | head -2) >> gpurun_out/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 3 --warmup 3 --reads 1250000 > gpurun_out/bench_n8_10M.json 2> gpurun_out/bench_n8_10M.err
tail -3 gpurun_out/bench_n8_10M.err
python tools/bench_brief.py gpurun_out/bench_n8_10M.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n8_10M.json'))
print(d['e2e']); print(d['config']['cpus_bound_per_rank'])
PY
cat gpurun_out/topo.txt | head -40
