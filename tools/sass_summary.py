"""Per-kernel counts of the Blackwell-specific SASS instructions in librsg_b200.so (cuobjdump -sass):
UTCHMMA/UTCQMMA (tcgen05.mma), UTMALDG (TMA tensor load), UBLKCP (cp.async.bulk), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit),
HMMA (mma.sync), LDGSTS (cp.async), ATOM/RED.   python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'rsgnet_b200', 'librsg_b200.so')
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
pats = collections.OrderedDict([('UTCHMMA', r'\bUTC[A-Z]*MMA'), ('UTMALDG', r'\bUTMALDG'), ('UBLKCP', r'\bUBLKCP'), ('LDTM', r'\bLDTM'),
                                ('UTCBAR', r'\bUTCBAR'), ('HMMA', r'\bHMMA'), ('LDGSTS', r'\bLDGSTS'), ('RED/ATOM', r'\b(RED|ATOM)[. ]')])
rows, cur = [], None
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = [m.group(1), collections.Counter(), 0]
        rows.append(cur)
        continue
    if cur is None or '/*' not in line:
        continue
    cur[2] += 1
    for k, p in pats.items():
        if re.search(p, line):
            cur[1][k] += 1
def demangle(n):
    try:
        d = subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
        d = d.replace('(anonymous namespace)::', '').replace('void ', '')
        return d.split('(')[0]
    except Exception:
        return n
print('librsg_b200.so, sm_100a SASS instruction counts per kernel (cuobjdump -sass)')
print('%-72s %7s ' % ('kernel', 'instr') + ' '.join('%8s' % k for k in pats))
tot = collections.Counter()
for name, c, n in sorted(rows, key=lambda r: -r[2]):
    tot.update(c)
    print('%-72s %7d ' % (demangle(name)[-72:], n) + ' '.join('%8d' % c[k] for k in pats))
print('%-72s %7d ' % ('TOTAL', sum(r[2] for r in rows)) + ' '.join('%8d' % tot[k] for k in pats))
