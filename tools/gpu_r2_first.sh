#!/bin/bash
# Round-2 first GPU call: tests, bench (all three modes side by side), guard study and
# fast-path validation on the barcoded population.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -30 > gpurun_out/tests.log
cat gpurun_out/tests.log | tail -5
timeout 400 python bench.py --steps 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -3 gpurun_out/bench_n1.err
python tools/bench_brief.py gpurun_out/bench_n1.json 2>&1 | head -12
timeout 300 python tools/guard_study.py --reads 400000 --batches 2 > gpurun_out/guard_short.json 2> gpurun_out/guard_short.err
timeout 300 python tools/guard_study.py --reads 100000 --batches 2 --length 16000 > gpurun_out/guard_stock.json 2> gpurun_out/guard_stock.err
timeout 600 python tools/validate_fast_path.py --batches 12 > gpurun_out/validate_fast.json 2> gpurun_out/validate_fast.err
tail -2 gpurun_out/validate_fast.err
