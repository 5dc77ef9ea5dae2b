"""Top stall lines of an `ncu --page source --csv` dump.  python tools/ncu_top.py file.csv [N]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
k = ci['Warp Stall Sampling (All Samples)']
data = []
for r in rows[hi + 1:]:
    if len(r) <= k:
        continue
    try:
        v = float(r[k])
    except ValueError:
        continue
    stalls = {h[6:]: r[ci[h]] for h in hdr if h.startswith('stall_') and 'Not' not in h and r[ci[h]] not in ('0', '')}
    data.append((v, r[ci['Source']][:100], stalls))
tot = sum(v for v, _, _ in data)
for v, s, st in sorted(data, key=lambda t: -t[0])[:n]:
    top = sorted(st.items(), key=lambda kv: -float(kv[1]))[:2]
    print(f'{v:7.0f} {100 * v / tot:5.1f}%  {s:100s} {top}')
