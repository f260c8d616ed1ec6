"""Summarise an `ncu --set full` capture (exported with `ncu -i X.ncu-rep --page raw --csv`)
into the per-kernel JSON kept under profiles/.   python tools/ncu_summary.py raw.csv out.json"""
import csv
import json
import sys

KEEP = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum']


def main(raw, out):
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')]
    res = []
    for r in rows[2:]:
        ent = {'kernel': r[idx['Kernel Name']]}
        for k in KEEP:
            if k in idx and r[idx[k]] != '':
                ent[k] = {'value': float(r[idx[k]].replace(',', '')), 'unit': units[idx[k]]}
        top = sorted(((float(r[idx[h]] or 0), h) for h in stalls), reverse=True)[:6]
        ent['top_stalls_per_issue_active'] = {
            h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''): v for v, h in top}
        res.append(ent)
    json.dump(res, open(out, 'w'), indent=1)
    print('wrote', out, len(res), 'kernels')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
