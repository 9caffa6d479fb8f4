"""Per-launch dump of an `ncu --csv` launch list: id, kernel, grid, block, us, DRAM MB (read/write)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i + 1
        break
idx = {h: i for i, h in enumerate(hdr)}
L = collections.OrderedDict()
for r in rows[start:]:
    if len(r) < len(hdr):
        continue
    unit, metric = r[idx['Metric Unit']], r[idx['Metric Name']]
    val = float(r[idx['Metric Value']].replace(',', ''))
    if metric == 'gpu__time_duration.sum':
        val *= {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(unit, 1e-3)
    else:
        val *= {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}.get(unit, 1e-6)
    d = L.setdefault(r[idx['ID']], {'name': r[idx['Kernel Name']].split('(')[0].replace('void ', '').replace('b200::', '')[:56],
                                    'grid': r[idx['Grid Size']], 'block': r[idx['Block Size']]})
    d[metric] = val
for i, d in L.items():
    print('%4s %-56s %-18s %-14s %8.1f us  rd %7.1f wr %7.1f MB' % (i, d['name'], d['grid'], d['block'], d.get('gpu__time_duration.sum', 0),
          d.get('dram__bytes_read.sum', 0), d.get('dram__bytes_write.sum', 0)))
