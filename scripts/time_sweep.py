"""Kernel-only timing of the stencil kernels (bench.kernel_rooflines) for the library named by
KSFD_B200_LIB (A/B of build variants): 2-D 1024^2 and 3-D 256^3."""
import json, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import bench
for dim, n, reps in ((2, (1024, 1024), 20), (3, (256, 256, 256), 10)):
    k, peak, _ = bench.kernel_rooflines(dim, n, reps=reps)
    print(os.environ.get('KSFD_B200_LIB', 'default'), n,
          {a: (round(v['us'], 1), round(v['frac'], 3)) for a, v in k.items()}, flush=True)
