#!/bin/bash
# Large validation of the fast path (ring kernels, split tentative / resolve) against the exact-only kernels
set -u
mkdir -p gpurun_out
timeout 600 python tools/validate_fast_path.py --batches 120 --seed0 1313000 > gpurun_out/validate_fast_120M.json 2> gpurun_out/validate_fast_120M.err
tail -1 gpurun_out/validate_fast_120M.err | cut -c1-200
timeout 200 python tools/validate_fast_path.py --batches 16 --length 8000 --reads 500000 --seed0 1414000 > gpurun_out/validate_fast_8k.json 2> gpurun_out/validate_fast_8k.err
tail -1 gpurun_out/validate_fast_8k.err | cut -c1-200
python - <<PY
import json
for f in ('validate_fast_120M','validate_fast_8k'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, d['total_reads'], d['total_mismatches'], d['exact_rerun_fraction'], d['accepted_by_barcode'], d['reads_over_1e-5_rel'], d['max_rel_error_normalised_signal_at_100pA'], d['max_scaler_z0_error'], d['max_scaler_z1_error'])
    except Exception as e:
        print(f, 'failed', e)
PY
