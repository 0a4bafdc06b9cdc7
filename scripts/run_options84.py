"""BASELINE configs[1]: the reference's options84 problem at FULL size (2-D
1536x1536, two ligand groups, TSAdapt basic from dt = 1e-8) through the
ksfdsolver2 entry on one GPU, for a limited number of steps.  Prints wall time
per accepted step (includes monitors; no --save)."""
import os
import sys
import tempfile
import time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
from ksfd_b200.solver import main

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
pc_type = sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] != '-' else None   # pbjacobi | fft | (default: lu -> automatic)
ksp_type = sys.argv[3] if len(sys.argv) > 3 else None    # gmres | richardson | (default: preonly -> the library's choice)
lines = []
for line in open(os.path.join(R, 'tests', 'options', 'options84.args')):
    key = line.split('=', 1)[0].strip()
    if key.startswith('--save') or key.startswith('--check'):
        continue
    if key == 'maxsteps':
        line = 'maxsteps=%d\n' % nsteps
    if pc_type and line.startswith('-pc_type'):
        line = '-pc_type %s\n' % pc_type
    if ksp_type and line.startswith('-ksp_type'):
        line = '-ksp_type %s\n' % ksp_type
    lines.append(line)
with tempfile.NamedTemporaryFile('w', suffix='.args', delete=False) as f:
    f.write(''.join(lines))
t0 = time.perf_counter()
rc = main('ksfdsolver2.py', '@' + f.name)
wall = time.perf_counter() - t0
print('options84 full size (pc %s, ksp %s): rc %d, %d steps in %.2f s wall (%.1f ms/step incl. set-up and '
      'monitors)' % (pc_type or 'auto', ksp_type or 'auto', rc, nsteps, wall, 1e3 * wall / nsteps))
