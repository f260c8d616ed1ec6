#!/bin/bash
# ncu --set full of the tensor-core LSTM kernels (one launch of each instantiation) on a small batch
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'k_lstm_tc' -c 6 -f -o gpurun_out/full_tc \
    python bench.py --reads 148000 --steps 1 --warmup 1 --no-e2e --no-cpu --no-verify > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/full_tc.ncu-rep --page raw --csv > gpurun_out/full_tc_raw.csv 2>/dev/null
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
