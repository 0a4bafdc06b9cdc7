"""
Fixture generator (test infrastructure): a PATTERNED options84 state for the
preconditioner tests.  Integrates the options84 physics on a 96x96 periodic
tile (h = 1/384) from rho = 9000 + 90 N(0,1) with the oracle's adaptive ROSW
(TSAdapt basic, atol = rtol = 1e-5, SuperLU solves) through the linear-growth
phase into the pattern phase and saves the state after 180 accepted steps
(t ~ 2.7e3: rho between ~5e2 and the density cap 2.6e4) to
tests/golden/host_pattern96.npz.  Takes ~15 minutes on one core; run once:
    python oracle/make_pattern_state.py
"""
import os
import sys

import numpy as np

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, 'tests'))
from helpers import oracle_physics, phys84          # noqa: E402
from oracle import ksfd_oracle as O                 # noqa: E402

if __name__ == '__main__':
    n = (96, 96)
    ph = oracle_physics(phys84(2, n))
    rng = np.random.default_rng(5)
    rho = 9000 + 90 * rng.standard_normal(n)
    u = np.stack([rho, rho, rho]).reshape(-1, order='F')
    out = O.integrate(u, 0.0, 1e-8, 180, ph,
                      adapt=dict(atol=1e-5, rtol=1e-5, clip=(0.1, 5.0), dt_min=1e-20, dt_max=1e4))
    t, state = out[-1]
    np.savez_compressed(os.path.join(R, 'tests', 'golden', 'host_pattern96.npz'),
                        u=state.reshape(-1, order='F'), t=t, n=np.array(n))
    print('t = %.6g, rho in [%.4g, %.4g]' % (t, state[0].min(), state[0].max()))
