#!/bin/bash
# One gpurun call that produces everything a round's profiles/ needs (run from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_round_check.sh'
# Outputs land in gpurun_out/: tests.log, bench_n1.json, bench_full_n1.json, launches.csv (ncu launch
# list of the same bench command), full_raw.csv (ncu --set full, raw page), sweep_lengths.json,
# validate_*.json, host_path_time.log.  tools/profile_digest.py + tools/ncu_summary.py turn the two
# ncu outputs into the files kept under profiles/.
# Each step has its own timeout so that a hang in one cannot eat the whole box time.
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/tests.log
timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 python bench.py --config full > gpurun_out/bench_full_n1.json 2> gpurun_out/bench_full_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 400 --csv \
    --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-verify > gpurun_out/ncu_launches.log 2>&1
# full-set capture: ncu replays each kernel ~40 times, so a smaller batch and one launch of each
# tensor-core kernel + the HBM / fp64 kernels
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'k_lstm_tc|k_segment|k_pool|k_windows' -c 12 -f -o gpurun_out/full \
    python bench.py --reads 148000 --steps 1 --warmup 1 --no-e2e --no-cpu --no-verify > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/full.ncu-rep --page raw --csv > gpurun_out/full_raw.csv 2>/dev/null
rm -f gpurun_out/full.ncu-rep
timeout 300 python tools/host_path_time.py > gpurun_out/host_path_time.log 2>&1
timeout 500 python tools/validate_fast_path.py --batches 30 --seed0 424000 > gpurun_out/validate_fast.json 2> gpurun_out/validate_fast.err
timeout 200 python tools/validate_fast_path.py --batches 4 --length 16000 --reads 500000 --seed0 525000 > gpurun_out/validate_fast_16k.json 2> gpurun_out/validate_fast_16k.err
timeout 200 python tools/validate_fast_path.py --batches 4 --mode strict --seed0 626000 > gpurun_out/validate_strict.json 2> gpurun_out/validate_strict.err
timeout 450 python tools/sweep_lengths.py > gpurun_out/sweep_lengths.json 2> gpurun_out/sweep_lengths.err
python tools/bench_brief.py gpurun_out/bench_n1.json 2>&1 | head -8
python tools/bench_brief.py gpurun_out/bench_full_n1.json 2>&1 | head -8
cat gpurun_out/tests.log
tail -4 gpurun_out/host_path_time.log
tail -1 gpurun_out/validate_fast.err | cut -c1-300
