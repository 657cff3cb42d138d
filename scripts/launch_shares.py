"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel launches, total ms, share.
usage: python scripts/launch_shares.py launches.csv [first_launch last_launch]  (ids of the timed step)"""
import csv, sys, collections, re
path = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.DictReader(lines)
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(r['Metric Value'].replace(',', ''))
    unit = r['Metric Unit']
    ms = v * {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'nsecond': 1e-6, 'ms': 1.0, 'msecond': 1.0}[unit]
    rows.append((int(r['ID']), r['Kernel Name'], ms, r.get('Grid Size', '')))
rows = [r for r in rows if lo <= r[0] < hi]
tot = sum(r[2] for r in rows)
agg = collections.OrderedDict()
for _, k, ms, _ in rows:
    k = re.sub(r'\(.*', '', k)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ms
print('launches %d total_ms %.3f small(<10us) %d sum %.3f ms' % (len(rows), tot, sum(1 for r in rows if r[2] < 0.01), sum(r[2] for r in rows if r[2] < 0.01)))
print('kernel,launches,total_ms,share')
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%s,%d,%.3f,%.3f' % (k, n, ms, ms / tot))
