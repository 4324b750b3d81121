"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    v = float(r['Metric Value'].replace(',', ''))
    unit = r.get('Metric Unit', 'ns')
    v = v * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(unit, 1e-3)
    tot[name][0] += 1
    tot[name][1] += v
allt = sum(v[1] for v in tot.values())
print('%-70s %7s %12s %8s %7s' % ('kernel', 'count', 'total_us', 'avg_us', 'share'))
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print('%-70s %7d %12.1f %8.1f %6.1f%%' % (k[:70], n, t, t / n, 100 * t / allt))
print('%-70s %7d %12.1f' % ('TOTAL', sum(v[0] for v in tot.values()), allt))
