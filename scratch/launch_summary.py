"""Fold an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals: python scratch/launch_summary.py file.csv [skip_first_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors='replace')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
ki, vi = rows[hdr].index('Kernel Name'), rows[hdr].index('Metric Value')
seq = [(r[ki], float(r[vi].replace(',', ''))) for r in rows[hdr + 2:] if len(r) > vi and r[vi].replace(',', '').replace('.', '').isdigit()]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
seq = seq[skip:]
tot = collections.OrderedDict()
for k, v in seq:
    k = k.split('(')[0][:70]
    c = tot.setdefault(k, [0, 0.0]); c[0] += 1; c[1] += v
total = sum(v for _, v in seq)
print("%d launches, %.3f ms" % (len(seq), total / 1e6))
for k, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %5d  %10.1f us  %5.1f%%" % (k, c, v / 1e3, 100 * v / total))
