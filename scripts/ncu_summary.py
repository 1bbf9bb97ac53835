"""Summarise an `ncu --page raw --csv` export: one line per kernel launch."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [('dur_us', 'gpu__time_duration.sum'), ('rd_MB', 'dram__bytes_read.sum'), ('wr_MB', 'dram__bytes_write.sum'),
        ('l2_MB', 'lts__t_bytes.sum'), ('tensor%', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
        ('dram%', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
        ('lts%', 'lts__throughput.avg.pct_of_peak_sustained_elapsed'),
        ('l1%', 'l1tex__throughput.avg.pct_of_peak_sustained_active'),
        ('issue%', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed'),
        ('smGHz', 'gpc__cycles_elapsed.avg.per_second')]
scale = {'ms': 1e3, 'us': 1.0, 'ns': 1e-3, 's': 1e6, 'Gbyte': 1e3, 'Mbyte': 1.0, 'Kbyte': 1e-3, 'byte': 1e-6,
         'Ghz': 1.0, 'Mhz': 1e-3, '%': 1.0, 'GHz': 1.0, 'MHz': 1e-3}
ki = hdr.index('Kernel Name')
print('%3s %-8s ' % ('#', 'kernel') + ' '.join('%9s' % w[0] for w in want))
for i, r in enumerate(data):
    name = r[ki]
    short = 'conv1' if 'conv1' in name else ('tc32' if '<32>' in name else ('tc64' if 'conv_tc' in name else name.split('::')[-1][:8]))
    vals = []
    for _, m in want:
        if m not in hdr:
            vals.append(float('nan')); continue
        c = hdr.index(m)
        try:
            vals.append(float(r[c].replace(',', '')) * scale.get(units[c], 1.0))
        except Exception:
            vals.append(float('nan'))
    print('%3d %-8s ' % (i + 1, short) + ' '.join('%9.1f' % v if abs(v) >= 10 else '%9.2f' % v for v in vals))
