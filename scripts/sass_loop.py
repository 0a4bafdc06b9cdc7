"""Opcode histogram of the hottest loop (largest backward-branch span) of a SASS dump.
usage: cuobjdump -sass -fun <mangled> <binary> | python scripts/sass_loop.py"""
import re
import sys
from collections import Counter

ins = []
for line in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for a, t in ins:
    m = re.search(r"BRA(?:\.\w+)*\s+(?:`\(\S+\)|0x([0-9a-f]+))", t)
    mm = re.search(r"BRA.*0x([0-9a-f]+)", t)
    if mm:
        tgt = int(mm.group(1), 16)
        if tgt < a:
            loops.append((a - tgt, tgt, a))
loops.sort(reverse=True)
print('backward branches (span, from, to):', [(s // 16, hex(t), hex(a)) for s, t, a in loops[:6]])
if loops:
    s, t, a = loops[0]
    body = [x for x in ins if t <= x[0] <= a]
    c = Counter()
    for _, txt in body:
        op = txt.split()[0]
        if op.startswith('@'):
            op = txt.split()[1]
        c[op.split('.')[0] + ('.' + op.split('.')[1] if op.startswith(('LDS', 'STS', 'LDG', 'STG')) and '.' in op else '')] += 1
    print('loop body: %d instructions' % len(body))
    for k, v in c.most_common(40):
        print('%5d %s' % (v, k))
