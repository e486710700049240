#!/usr/bin/env python
"""Print the figures of a bench.py JSON line that the notes quote: python profiles/show_bench.py <file.json>"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'ms_per_batch', 'gpu_launches', 'serial_ms_per_batch', 'n_gpus') if k in d})
print('per_batch_ms', d.get('per_batch_ms'))
e = d['e2e']
print('e2e', {k: e[k] for k in ('value', 'ms_per_batch', 'batches_per_step', 'timed_region_s') if k in e}, 'link', e.get('link', {}).get('d2h_copy_ms'), e.get('link', {}).get('gbs'))
print('roofline', {k: d['roofline'][k] for k in ('achieved', 'peak', 'frac', 'traffic', 'share_of_step')})
for s in d.get('stages', []):
    print('  ', s['stage'], round(s['ms'] * 1e3, 1), 'us frac', round(s['frac'], 3), 'issue', round(s.get('issue_frac', 0), 3))
print('clocks', d.get('clocks'))
print('per_rank', d.get('per_rank'))
for k, b in d.get('configs', {}).items():
    print('==', k)
    if 'error' in b:
        print(b); continue
    for kk, vv in b.items():
        if kk in ('stages', 'stages_rank0'):
            for s in vv:
                print('     ', s['stage'], round(s['ms'] * 1e3, 1), 'us frac', round(s['frac'], 3))
        elif kk not in ('workload', 'counts', 'counts_rank0'):
            print('  ', kk, vv)
print('cpu_baseline', d.get('cpu_baseline'))
print('cpu_port', d.get('cpu_port'))
