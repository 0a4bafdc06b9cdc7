"""
ksfd_b200 — B200-native implicit time-stepping hot path of leonavery/KSFD.

The names a KSFD user knows are re-exported here (reference KSFD/__init__.py):
Parser, SolutionParameters, LigandGroups, Grid, Derivatives, SpatialExpression,
implicitTS, ksfdTS, TimeSeries, Generator, random_function, dillnp/dillunp.
Compute happens in libksfd_b200.so (CUDA, sm_100a) through ksfd_b200._lib.
"""
from ._lib import KSFDError  # noqa: F401
from .params import (KSFDException, Ligand, LigandGroup, LigandGroups, Parameter,  # noqa: F401
                     ParameterList, Parser, SolutionParameters, default_parameters, find_duplicates,
                     parse_commandline, petsc_init, safe_sympify)


def __getattr__(name):
    # heavier modules (torch, the CUDA library) load on first use
    import importlib
    table = {
        'Grid': ('.grid', 'Grid'), 'Vec': ('.vec', 'Vec'),
        'Derivatives': ('.derivs', 'Derivatives'),
        'SpatialExpression': ('.derivs', 'SpatialExpression'),
        'implicitTS': ('.ts', 'make_implicitTS'), 'ksfdTS': ('.ts', 'ksfdTS'),
        'KSFDTS': ('.ts', 'KSFDTS'),
        'TimeSeries': ('.timeseries', 'TimeSeries'),
        'Gatherer': ('.timeseries', 'Gatherer'), 'tsmerge': ('.timeseries', 'tsmerge'),
        'dillnp': ('.timeseries', 'dillnp'), 'dillunp': ('.timeseries', 'dillunp'),
        'Generator': ('.random', 'Generator'),
        'random_function': ('.random', 'random_function'),
        'Context': ('.core', 'Context'),
    }
    if name in table:
        mod, attr = table[name]
        return getattr(importlib.import_module(mod, __name__), attr)
    raise AttributeError(name)
