"""Top stall sites of an `ncu --page source --csv` export (SASS view): address, samples, instruction."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ci = hdr.index('# Samples'); si = hdr.index('Source'); ei = hdr.index('Instructions Executed')
tot = sum(int(r[ci] or 0) for r in data)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
print('total samples', tot, 'instructions', len(data))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ci] or 0))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
    print('%5d %6.2f%% exec=%-9s %-70s %s' % (i, 100.0 * int(r[ci] or 0) / max(tot, 1), r[ei], r[si][:70],
                                             ' '.join('%s:%d' % (n, v) for v, n in st if v)))
