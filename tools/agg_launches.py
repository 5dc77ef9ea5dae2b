"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: python tools/agg_launches.py file.csv"""
import csv, sys, collections, re
rows = list(csv.reader(l for l in open(sys.argv[1], errors='replace') if l.startswith('"')))
hdr = rows[0]
ki, vi, mi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
ui = hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0.0, 0])
for r in rows[1:]:
    if 'gpu__time_duration' not in r[mi]:
        continue
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] in ('ns', 'nsecond') else (v * 1e3 if r[ui] in ('ms', 'msecond') else v)      # -> us
    name = re.sub(r'\(.*', '', r[ki])
    agg[name][0] += v
    agg[name][1] += 1
tot = sum(a[0] for a in agg.values())
print(f'total {tot / 1e3:.2f} ms over {sum(a[1] for a in agg.values())} launches')
for k, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f'{us / 1e3:9.3f} ms {100 * us / tot:5.1f}% n={n:5d} {us / n:9.1f} us  {k}')
