"""Aggregate a bench.py --dump-profile file by (kernel family, shape)."""
import collections
import sys

rows = [l.rstrip('\n').split('\t') for l in open(sys.argv[1])]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.OrderedDict()
for name, fam, ms, fl, desc in rows:
    if float(ms) < 0:
        continue
    a = agg.setdefault((fam, desc), [0, 0.0, 0.0])
    a[0] += 1; a[1] += float(ms); a[2] += float(fl)
tot = sum(a[1] for a in agg.values())
print(f'total {tot:.3f} ms')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f'{a[1]:7.3f} ms {a[1] / tot * 100:5.1f}% n={a[0]:3d} {a[1] / a[0] * 1e3:7.1f} us '
          f'{a[2] / a[1] / 1e9 if a[1] else 0:7.1f} TF  {k[0]:16s} {k[1]}')
