"""
Build the oracle's C restatement (oracle/ksfd_oracle_c.c -> oracle/_build/libksfd_oracle.so).

TEST INFRASTRUCTURE: the library is the checker / CPU baseline, never the product.
Run by __graft_entry__.build() and, when the library is missing or stale, by
oracle/ksfd_oracle_c.py at import.  Plain gcc, OpenMP, no -march (the file travels to the GPU
box), -ffp-contract=off so that the arithmetic follows the numpy oracle's association order.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'ksfd_oracle_c.c')
OUT = os.path.join(HERE, '_build', 'libksfd_oracle.so')


def stale():
    return (not os.path.exists(OUT)) or os.path.getmtime(OUT) < os.path.getmtime(SRC)


def build(force=False):
    if not force and not stale():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    tmp = OUT + '.%d.tmp' % os.getpid()
    cmd = ['gcc', '-O3', '-fopenmp', '-fPIC', '-shared', '-ffp-contract=off', '-Wall',
           '-o', tmp, SRC, '-lm']
    subprocess.run(cmd, check=True)
    os.replace(tmp, OUT)          # atomic: several ranks / test workers may build at once
    return OUT


if __name__ == '__main__':
    print('oracle C restatement:', build(force='--force' in sys.argv))
