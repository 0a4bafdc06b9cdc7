"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): kernels of
ONE accepted ROSW step (between two k_complete_step launches), grouped by name."""
import csv, re, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.DictReader(lines)
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(r['Metric Value'].replace(',', ''))
    unit = r['Metric Unit']
    us = v / 1e3 if unit in ('ns', 'nsecond') else v if unit in ('us', 'usecond') else v * 1e3
    rows.append((r['Kernel Name'], us))
idx = [i for i, (n, _) in enumerate(rows) if 'k_complete_step' in n]
a, b = idx[1] + 1, idx[2] + 1       # the first timed step (warm-up step is idx[0])
step = rows[a:b]
agg = {}
for n, us in step:
    n = re.sub(r'\((?:long long|Geom|MarchArgs|int|GmFin|GmBegin|const).*$', '', n).replace('void ', '')
    c, t = agg.get(n, (0, 0.0))
    agg[n] = (c + 1, t + us)
tot = sum(t for _, t in agg.values())
print('ONE accepted ROSW step (between two k_complete_step launches): %d launches, %.1f us summed kernel time' % (len(step), tot))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('  %3d launches  %8.1f us  %4.1f%%  avg %7.2f us  %s' % (c, t, 100 * t / tot, t / c, n))
