import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], 'reads/s %.0f  ms/step %.1f  e2e %s' % (d['value'], d['ms_per_step'], d.get('e2e', {}).get('value')))
c = d['config']
print('  reruns', c.get('exact_reruns_per_step'), c.get('exact_rerun_causes'), 'mismatches', c.get('mismatches_vs_exact_only_kernels'))
print('  ' + '  '.join('%s %.1f' % (k['kernel'].replace('k_lstm_tc_', 'tc_'), k['ms_per_step']) for k in d['kernels'][:10]))
