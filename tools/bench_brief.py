import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], 'reads/s %.0f  ms/step %.1f  e2e %s' % (d['value'], d['ms_per_step'], d.get('e2e', {}).get('value')))
c = d['config']
print('  reruns', c.get('exact_reruns_per_step'), c.get('exact_rerun_causes'), 'mismatches', c.get('mismatches_vs_exact_only_kernels'))
print('  ' + '  '.join('%s %.1f' % (k['kernel'].replace('k_lstm_tc_', 'tc_'), k['ms_per_step']) for k in d['kernels'][:10]))
if 'modes' in d:
    print('  modes', {k: (round(v['value']), round(v['ms_per_step'], 1)) for k, v in d['modes'].items() if isinstance(v, dict)})
    print('  barcode mix', c.get('barcode_mix_of_classified'), 'guess', c.get('best_guess_mix_of_classified'))
    r = d['roofline']
    print('  roofline', r.get('kernel'), 'frac %.3f' % (r.get('frac') or 0), 'whole step frac %.3f' % r['whole_step']['frac'],
          'cpu', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('outputs_match_gpu'))
