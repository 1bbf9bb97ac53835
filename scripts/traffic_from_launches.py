"""profiles/traffic.json from an ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum
per launch): DRAM bytes per step of the conv_tc_kernel launches.  usage: traffic_from_launches.py launches.csv steps"""
import csv, json, sys, re
rows = list(csv.reader(open(sys.argv[1], errors='replace')))
steps = int(sys.argv[2])
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]
ki, mi, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
tot, n = 0.0, 0
for r in rows[hi + 1:]:
    if len(r) <= vi or 'conv_tc_kernel' not in r[ki]:
        continue
    if r[mi] in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        tot += float(r[vi].replace(',', '')) * scale.get(r[ui], 1.0)
    if r[mi] == 'gpu__time_duration.sum':
        n += 1
out = dict(conv_tc_dram_bytes_per_step=tot / steps, conv_tc_launches_per_step=n // steps,
           source='ncu dram__bytes_read.sum + dram__bytes_write.sum over the conv_tc_kernel launches of %s (%d steps)'
                  % (sys.argv[1].split('/')[-1], steps))
json.dump(out, open('profiles/traffic.json', 'w'), indent=1)
print(out)
