#!/usr/bin/env python3
"""
tsmerge — gather and merge KSFD time series (same command line as the reference's
tsmerge.py:40-111, which it replaces for series written by this package):

    python tsmerge.py -o merged  runs4@                # gather the 4 per-rank files of `run`
    python tsmerge.py -o all     first second          # merge two sequential series
    python tsmerge.py -o part -s 10 -e 50  runs8@ tail # both, keeping 10 <= t <= 50

The output is the sequential series <outfile>s1r0 (.h5 where h5py exists, else the .npz
stand-in), which `--resume` accepts on any number of ranks.  Runs on the CPU alone.
"""
import sys
from argparse import ArgumentParser


def main(argv=None):
    parser = ArgumentParser(description='Merge time series', allow_abbrev=True)
    parser.add_argument('-o', '--outfile', required=True, help='merged file basename')
    parser.add_argument('-s', '--start', type=float, default=0.0, help='start time')
    parser.add_argument('-e', '--end', type=float, help='end time')
    parser.add_argument('infiles', nargs='+', help='files to merge')
    parser.add_argument('-v', '--verbose', action='count', default=0)
    a = parser.parse_args(argv)
    from ksfd_b200.timeseries import tsmerge
    name = tsmerge(a.outfile, a.infiles, start=a.start, end=a.end, verbose=a.verbose)
    if a.verbose:
        print('wrote', name)
    return 0


if __name__ == '__main__':
    sys.exit(main())
