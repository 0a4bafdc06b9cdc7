"""Static resource usage of every kernel in the built library (cuobjdump --dump-resource-usage):
registers, stack (spills), static shared memory — the check that goes with `-Xptxas -v` before GPU
time is spent.  python scripts/resource_usage.py [filter] > profiles/..."""
import os, re, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(R, 'ksfd_b200', 'libksfd_b200.so')
out = subprocess.run(['cuobjdump', '--dump-resource-usage', lib], capture_output=True, text=True).stdout
names = re.findall(r'Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)', out)
dem = subprocess.run(['c++filt'], input='\n'.join(n[0] for n in names), capture_output=True, text=True).stdout.splitlines()
flt = sys.argv[1] if len(sys.argv) > 1 else ''
rows = []
for (m, reg, stack, sh, loc), d in zip(names, dem):
    d = re.sub(r'\(.*', '', d)          # template name without the argument list
    d = d.replace('void ', '')
    if flt and flt not in d:
        continue
    rows.append((d, int(reg), int(stack), int(sh), int(loc)))
rows.sort()
print('%d kernels in %s (sm_100a); REG = registers per thread, STACK = bytes of stack frame (spills / local arrays), '
      'SHARED = static shared memory (the marchers use dynamic shared memory on top)' % (len(rows), os.path.basename(lib)))
for d, reg, stack, sh, loc in rows:
    print('%-118s REG %3d  STACK %4d  SHARED %6d' % (d[:118], reg, stack, sh))
spill = [r for r in rows if r[2] > 0]
print('\nkernels with a stack frame: %d' % len(spill))
