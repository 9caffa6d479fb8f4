"""Aggregate an `ncu --csv` launch list (gpu__time_duration + dram bytes) by kernel name -> markdown table."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i + 1
        break
idx = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[start:]:
    if len(r) < len(hdr):
        continue
    name, metric, unit = r[idx['Kernel Name']], r[idx['Metric Name']], r[idx['Metric Unit']]
    val = float(r[idx['Metric Value']].replace(',', ''))
    if metric == 'gpu__time_duration.sum':
        val *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1}.get(unit, 1e-6)
    else:
        val *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
    key = name.split('(')[0].replace('void ', '').replace('b200::', '')[:64]
    d = agg.setdefault(key, {'ms': 0, 'rd': 0, 'wr': 0, 'ids': set()})
    d[{'gpu__time_duration.sum': 'ms', 'dram__bytes_read.sum': 'rd', 'dram__bytes_write.sum': 'wr'}[metric]] += val
    d['ids'].add(r[idx['ID']])
tot = sum(d['ms'] for d in agg.values())
print('| kernel | launches | total ms | share | DRAM read MB | DRAM write MB | DRAM GB/s |')
print('|---|---|---|---|---|---|---|')
for k, d in sorted(agg.items(), key=lambda kv: -kv[1]['ms']):
    if d['ms'] / tot < 0.002:
        continue
    print('| %s | %d | %.3f | %.1f %% | %.0f | %.0f | %.0f |' % (k, len(d['ids']), d['ms'], 100 * d['ms'] / tot, d['rd'] / 1e6,
                                                              d['wr'] / 1e6, (d['rd'] + d['wr']) / d['ms'] / 1e6 if d['ms'] else 0))
print('\ntotal %.3f ms over %d launches' % (tot, sum(len(d['ids']) for d in agg.values())))
