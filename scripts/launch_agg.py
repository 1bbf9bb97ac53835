"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name; list the slowest launches."""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1], errors='replace')))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]
ki, mi, vi = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value')
ui = hdr.index('Metric Unit')
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg, order = collections.defaultdict(lambda: [0, 0.0]), []
n = 0
for r in rows[hi + 1:]:
    if len(r) <= vi or r[mi] != 'gpu__time_duration.sum':
        continue
    n += 1
    if n <= skip:
        continue
    t = float(r[vi].replace(',', ''))
    u = r[ui]
    t_us = t / 1e3 if u in ('ns', 'nsecond') else (t * 1e3 if u in ('ms', 'msecond') else t)
    name = re.sub(r'\(.*', '', r[ki]).replace('void ', '').replace('dy::<unnamed>::', '').replace('dy::', '')
    agg[name][0] += 1
    agg[name][1] += t_us
    order.append((t_us, name, n))
tot = sum(v[1] for v in agg.values())
print('launches %d  total %.1f us' % (len(order), tot))
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%8.1f us %5.1f%%  x%-4d %s' % (t, 100 * t / tot, c, name[:90]))
print('--- slowest launches')
for t, name, i in sorted(order, reverse=True)[:25]:
    print('%8.1f us  #%d %s' % (t, i, name[:80]))
