"""SASS opcode histograms of the hot kernels of libksfd_b200.so (whole kernel and hottest loop).
usage: python scripts/sass_hist.py [library] > profiles/rNN_sass_histograms.txt"""
import re
import subprocess
import sys
from collections import Counter

LIB = sys.argv[1] if len(sys.argv) > 1 else 'ksfd_b200/libksfd_b200.so'
WANT = [r'k_tma_marchILi2ELi256ELi1E7SweepOpILi2ELi2E', r'k_tma_marchILi3ELi16ELi16E7SweepOpILi3ELi2E',
        r'k_tma_marchILi3ELi32ELi16E7SweepOpILi3ELi2E',
        r'k_tma_marchILi2ELi256ELi1E5JvpOpILi2ELi2ELb1E', r'k_tma_marchILi3ELi16ELi16E5JvpOpILi3ELi2ELb1E',
        r'k_tma_marchILi3ELi32ELi16E5JvpOpILi3ELi2ELb1E',
        r'k_marchILi2ELi124ELi1E10ResidualOpILi2ELi2ELb1E', r'k_marchILi2ELi252ELi1E10ResidualOpILi2ELi2ELb1E',
        r'k_marchILi3ELi16ELi16E10ResidualOpILi3ELi2ELb1E']
funcs = re.findall(r'Function : (\S+)', subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True,
                                                      text=True).stdout)
for pat in WANT:
    for fn in [f for f in funcs if re.search(pat, f)]:
        sass = subprocess.run(['cuobjdump', '-sass', '-fun', fn, LIB], capture_output=True, text=True).stdout
        ins = []
        for line in sass.splitlines():
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))

        def hist(body):
            c = Counter()
            for _, t in body:
                op = t.split()[1] if t.startswith('@') else t.split()[0]
                parts = op.split('.')
                key = parts[0]
                if key in ('LDS', 'STS', 'LDG', 'STG', 'UTMALDG', 'SYNCS', 'LDGSTS', 'UBLKCP') and len(parts) > 1:
                    key = '.'.join(parts[:3])
                c[key] += 1
            return c
        loops = []
        for a, t in ins:
            mm = re.search(r"BRA.*0x([0-9a-f]+)", t)
            if mm and int(mm.group(1), 16) < a:
                loops.append((a - int(mm.group(1), 16), int(mm.group(1), 16), a))
        loops.sort(reverse=True)
        name = subprocess.run(['c++filt', fn], capture_output=True, text=True).stdout.strip()
        print('=' * 100)
        print(re.sub(r'\(MarchArgs.*', '', name))
        whole = hist(ins)
        print('whole kernel: %d instructions;  TMA / mbarrier / cp.async opcodes: %s' % (
            len(ins), {k: v for k, v in whole.items() if k.startswith(('UTMALDG', 'SYNCS', 'LDGSTS', 'UBLKCP', 'UTMAPF'))}))
        if loops:
            s, t, a = loops[0]
            body = [x for x in ins if t <= x[0] <= a]
            print('hottest loop (J.v kernels: five plane iterations, the loop is unrolled five-fold; residual: one): %d instructions' % len(body))
            print('  ' + '  '.join('%s %d' % kv for kv in hist(body).most_common(28)))
