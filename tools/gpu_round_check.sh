#!/bin/bash
# One gpurun call that produces everything a round's profiles/ needs (run from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_round_check.sh'
# Outputs land in gpurun_out/: tests.log, bench_n1.json, launches.csv (ncu launch list of the
# same bench command, kernels only), full_tc.ncu-rep (+ raw csv) of the tensor-core LSTM kernels.
# Each step has its own timeout so that a hang in one cannot eat the whole box time.
set -u
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/tests.log
timeout 200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
    --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-verify > gpurun_out/ncu_launches.log 2>&1
# full-set capture: ncu replays each kernel ~40 times, so a smaller batch and one launch of each
# tensor-core instantiation + the segmentation kernel
timeout 420 ncu --set full --clock-control none --import-source on \
    -k regex:'k_lstm_tc|k_segment' -c 8 -f -o gpurun_out/full_tc \
    python bench.py --reads 148000 --steps 1 --warmup 1 --no-e2e --no-cpu --no-verify > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/full_tc.ncu-rep --page raw --csv > gpurun_out/full_tc_raw.csv 2>/dev/null
python tools/bench_brief.py gpurun_out/bench_n1.json 2>&1 | head -6
cat gpurun_out/tests.log
