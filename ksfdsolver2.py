#! /usr/bin/env python3
"""Drop-in for the reference's ksfdsolver2.py: same command line and option
files, implicit time stepping on B200 GPUs.  See ksfd_b200/solver.py."""
import sys

from ksfd_b200.solver import main

if __name__ == '__main__':
    sys.exit(main())
