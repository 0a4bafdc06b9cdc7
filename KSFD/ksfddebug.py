"""`from KSFD.ksfddebug import log` (reference KSFD/ksfddebug.py): prints when the
colon-separated KSFDDEBUG environment variable names the system or ALL."""
import os


def log(*args, system='KSFD', **kwargs):
    wanted = set(os.getenv('KSFDDEBUG', default='').split(':'))
    if system in wanted or 'ALL' in wanted:
        rank = int(os.environ.get('RANK', os.environ.get('OMPI_COMM_WORLD_RANK', 0)))
        print('%s, rank=%d:' % (system, rank), *args, flush=True, **kwargs)
