"""Summarise a scripts/tune_tma.cu log: best time per variant ordinal and problem."""
import re
import sys

best, order, sec = {}, [], None
for line in open(sys.argv[1]):
    if line.startswith('=='):
        sec = line.strip()
        continue
    m = re.match(r"#(\d+) (\S+)\s+(REF|TMA) (.*?)\s+([\d.]+) us\s+([\d.]+) Gpts/s\s+frac ([\d.]+)(.*)", line)
    if not m:
        if line.strip():
            print('??', line.strip())
        continue
    o, name, kind, cfg, us, g, frac, rest = m.groups()
    key = (sec, int(o))
    if key not in best:
        order.append(key)
    if key not in best or float(us) < best[key][0]:
        best[key] = (float(us), name, kind, re.sub(r"\s+", " ", cfg), frac, rest.strip())
cur = None
for key in order:
    if key[0] != cur:
        cur = key[0]
        print(cur)
    v = best[key]
    print('  #%02d %-9s %s %-62s %8.1f us frac %s %s' % (key[1], v[1], v[2], v[3], v[0], v[4], v[5][-12:]))
