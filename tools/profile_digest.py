"""Digest of the two ncu passes of tools/gpu_round_check.sh into the files kept under profiles/:

    python tools/profile_digest.py shares  launches.csv  out.json   # per-kernel share of the launch list
    python tools/profile_digest.py traffic ncu_full.json reads windows out.json
                                                                    # DRAM bytes per read of the --set full capture
`reads` / `windows`: reads of the captured batch and windows its classifier kernels stepped
(bench.py --reads R prints classified_reads)."""
import csv
import json
import re
import sys

ROWS = (('k_pool', r'k_pool<', 'reads'), ('k_windows', r'k_windows\(', 'reads'),
        ('k_segment', r'k_segment<', 'reads'),
        ('k_lstm_tc_scaler', r'k_lstm_tc_scaler2<', 'reads'),
        ('k_lstm_tc_scaler_l1', r'k_lstm_tc<48, 0, 1, 0>', None),     # resolved by grid size below
        ('k_lstm_tc_scaler_l2', r'k_lstm_tc<48, 48, 0, 0>', 'reads'),
        ('k_lstm_tc_demux_l2', r'k_lstm_tc<64, 96, 0, 0>', 'windows'),
        ('k_lstm_tc_demux_l2_probe', r'k_lstm_tc_probes<64, 96>|k_lstm_tc<64, 96, 0, [12]>', 'windows'))


def shares(csv_path, out):
    tot, per = 0.0, {}
    # ncu's --log-file starts with its own ==PROF== / ==WARNING== lines
    for r in csv.DictReader(l for l in open(csv_path) if not l.startswith('==')):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '').replace('pb::', '')
        ms = float(r['Metric Value'].replace(',', '')) / 1e6
        e = per.setdefault(name, {'ms': 0.0, 'launches': 0})
        e['ms'] += ms
        e['launches'] += 1
        tot += ms
    for e in per.values():
        e['share'] = e['ms'] / tot
    json.dump({'command': 'ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 '
                          'python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-verify',
               'note': 'first 400 launches of our kernels (warm-up steps included): cold-cache, serialised '
                       'times; shares, not absolutes, are comparable with bench.py',
               'kernels': dict(sorted(per.items(), key=lambda kv: -kv[1]['ms']))}, open(out, 'w'), indent=1)
    print('wrote', out)


def traffic(summary, reads, windows, out):
    """First captured launch of every kernel: DRAM bytes it moved / the reads (or windows) it stepped."""
    res = {}
    ks = json.load(open(summary))
    fused_scaler = any('k_lstm_tc_scaler2' in k['kernel'] for k in ks)
    for k in ks:
        name = k['kernel']
        byt = (k.get('dram__bytes_read.sum', {}).get('value', 0) + k.get('dram__bytes_write.sum', {}).get('value', 0))
        unit = k.get('dram__bytes_read.sum', {}).get('unit', 'Gbyte')
        byt *= {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}.get(unit, 1e9)
        for row, pat, per in ROWS:
            if per is None or not re.search(pat, name) or row in res:
                continue
            n = reads if per == 'reads' else windows
            res[row] = {'dram_bytes_per_read': byt / n, 'reads_in_capture': n, 'kernel': name[:60]}
        if re.search(r'k_lstm_tc<48, 0, 1, 0>', name):
            # scalar-input layer: the scaler's first layer (two-launch scaler only) or classifier layer 1
            grid = k.get('launch__grid_size', {}).get('value', 0)
            tiles_w = (windows + 127) // 128
            is_demux = fused_scaler or abs(grid - 2 * tiles_w) <= 2
            row, n = ('k_lstm_tc_demux_l1', windows) if is_demux else ('k_lstm_tc_scaler_l1', reads)
            res.setdefault(row, {'dram_bytes_per_read': byt / n, 'reads_in_capture': n, 'kernel': name[:60]})
    json.dump(res, open(out, 'w'), indent=1)
    print('wrote', out, sorted(res))


if __name__ == '__main__':
    if sys.argv[1] == 'shares':
        shares(sys.argv[2], sys.argv[3])
    else:
        traffic(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5])
