#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --durations=8 2>&1 | tail -16 | tee gpurun_out/tests_host_path.log
timeout 400 python tools/host_path_time.py 2>&1 | tail -12 | tee gpurun_out/host_path_time.log
