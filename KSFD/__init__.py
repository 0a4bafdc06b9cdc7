"""
`KSFD` import name for existing user scripts (reference ksfdsolver2.py:354-360:
`from KSFD import (KSFDException, Grid, TimeSeries, random_function, LigandGroups,
ParameterList, Parser, default_parameters, SolutionParameters, SpatialExpression,
Derivatives, implicitTS, Generator, dillnp)`, `from KSFD.ksfddebug import log`).
Everything resolves to ksfd_b200; compute runs in libksfd_b200.so on the GPU.
"""
import ksfd_b200 as _impl
from ksfd_b200 import (KSFDError, KSFDException, Ligand, LigandGroup, LigandGroups,  # noqa: F401
                       Parameter, ParameterList, Parser, SolutionParameters,
                       default_parameters, find_duplicates, parse_commandline, petsc_init,
                       safe_sympify)

__all__ = ['Parser', 'KSFDException', 'Generator', 'random_function', 'TimeSeries', 'dillnp',
           'dillunp', 'Parameter', 'ParameterList', 'Ligand', 'LigandGroup', 'LigandGroups',
           'find_duplicates', 'SolutionParameters', 'default_parameters', 'Grid',
           'safe_sympify', 'SpatialExpression', 'Derivatives', 'ksfdTS', 'implicitTS']


def __getattr__(name):
    # Grid, Derivatives, implicitTS, TimeSeries ... load torch / the CUDA library on first use
    return getattr(_impl, name)
