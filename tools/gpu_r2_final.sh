#!/bin/bash
# Last measurement call of round 2 (one GPU, about 6 minutes of box time):
#   /usr/local/graft/bin/gpurun --timeout 700 -- 'bash tools/gpu_r2_final.sh [validation batches]'
# GPU tests, smoke, the driver's bench command, BASELINE configs[3], the ncu launch list of the bench
# command and a fast-path validation run on the current synthetic population.
set -u
NB=${1:-30}
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/tests.log
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -2 > gpurun_out/smoke.log
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 200 python bench.py --steps 3 --config full > gpurun_out/bench_full_n1.json 2> gpurun_out/bench_full_n1.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 400 --csv \
    --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-verify > gpurun_out/ncu_launches.log 2>&1
timeout $((NB * 8 + 60)) python tools/validate_fast_path.py --batches $NB --seed0 2020000 \
    > gpurun_out/validate_fast.json 2> gpurun_out/validate_fast.err
cat gpurun_out/tests.log gpurun_out/smoke.log
python tools/bench_brief.py gpurun_out/bench_n1.json 2>&1 | head -12
python tools/bench_brief.py gpurun_out/bench_full_n1.json 2>&1 | head -8
tail -1 gpurun_out/validate_fast.err | cut -c1-400
python - <<'E'
import json
try:
    d = json.load(open('gpurun_out/validate_fast.json'))
    print('validation', d['total_reads'], d['total_mismatches'], d['accepted_by_barcode'], d['exact_rerun_fraction'])
except Exception as e:
    print('validation document missing:', e)
E
