#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/tests_host_path.log
timeout 400 python tools/host_path_time.py 2>&1 | tail -14 | tee gpurun_out/host_path_time.log
