#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/tests_probes.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_probes.json 2> gpurun_out/bench_probes.err
tail -3 gpurun_out/bench_probes.err
python tools/bench_brief.py gpurun_out/bench_probes.json 2>&1 | head -8
