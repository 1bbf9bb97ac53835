"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_*` launch list by kernel.
usage: launch_shares.py file.csv [marker_kernel_substring occurrence_from occurrence_to]
With a marker (e.g. yolo_loss) the launches between two occurrences of that kernel = one step."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, mi, vi, ui, idi = (hdr.index(x) for x in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit', 'ID'))
L = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    d = L.setdefault(int(r[idi]), {'name': r[ki], 'mb': 0.0})
    v = float(r[vi].replace(',', ''))
    if r[mi] == 'gpu__time_duration.sum':
        d['us'] = v * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}.get(r[ui], 1)
    else:
        d['mb'] += v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}.get(r[ui], 1)
ids = list(L)
if len(sys.argv) > 2:
    marks = [k for k in ids if sys.argv[2] in L[k]['name']]
    a, b = marks[int(sys.argv[3])], marks[int(sys.argv[4])]
    ids = [k for k in ids if a <= k < b]
    nsteps = int(sys.argv[4]) - int(sys.argv[3])
else:
    nsteps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for k in ids:
    n = L[k]['name']
    short = n.split('(')[0].split('::')[-1]
    if 'conv_tc_kernel' in n:
        short = 'conv_tc_kernel<%s>' % ('32' if '32>' in n.split('(')[0] else '64')
    a_ = agg[short]
    a_[0] += 1; a_[1] += L[k]['us']; a_[2] += L[k]['mb']
tot = sum(v[1] for v in agg.values())
print('%d launches, %.1f us per step (%d step(s))' % (sum(v[0] for v in agg.values()) / nsteps, tot / nsteps, nsteps))
print('| kernel | launches | time (us) | share | DRAM (MB) |')
print('|---|---|---|---|---|')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('| `%s` | %d | %.1f | %.1f %% | %.0f |' % (k, v[0] / nsteps, v[1] / nsteps, 100 * v[1] / tot, v[2] / nsteps))
