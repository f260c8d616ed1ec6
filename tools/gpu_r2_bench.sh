#!/bin/bash
# Round-2 measurement call: default bench, config full, validation at scale, ncu launch list
set -u
mkdir -p gpurun_out
timeout 400 python bench.py --steps 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python tools/bench_brief.py gpurun_out/bench_n1.json 2>&1 | head -12
timeout 400 python bench.py --steps 3 --config full > gpurun_out/bench_full_n1.json 2> gpurun_out/bench_full_n1.err
tail -3 gpurun_out/bench_full_n1.err
python tools/bench_brief.py gpurun_out/bench_full_n1.json 2>&1 | head -12
timeout 900 python tools/validate_fast_path.py --batches 40 --seed0 880000 > gpurun_out/validate_fast.json 2> gpurun_out/validate_fast.err
tail -1 gpurun_out/validate_fast.err
timeout 300 python tools/validate_fast_path.py --batches 4 --length 16000 --reads 500000 --seed0 990000 > gpurun_out/validate_fast_16k.json 2> gpurun_out/validate_fast_16k.err
tail -1 gpurun_out/validate_fast_16k.err
timeout 300 python tools/validate_fast_path.py --batches 6 --mode strict --seed0 660000 > gpurun_out/validate_strict.json 2> gpurun_out/validate_strict.err
tail -1 gpurun_out/validate_strict.err
