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

__all__ = ['Parser', 'KSFDException', 'Generator', 'random_function', 'TimeSeries', 'Gatherer', 'dillnp',
           'dillunp', 'Parameter', 'ParameterList', 'Ligand', 'LigandGroup', 'LigandGroups',
           'find_duplicates', 'SolutionParameters', 'default_parameters', 'Grid',
           'safe_sympify', 'SpatialExpression', 'Derivatives', 'ksfdTS', 'implicitTS']


# names of the reference's __all__ (KSFD/__init__.py) that belong to the subsystems this
# implementation REPLACES: the sympy -> C code generation (ksfdufunc.py,
# ksfdsym.py:81-110, 1400-1439) and the assembled PETSc matrix (ksfdmat.py, ksfdMat.pyx) — or that
# the reference lists without defining (remap_from_files, makeKSFDSolver's dead module)
_REPLACED = {
    'getMat': 'the assembled Jacobian is replaced by the matrix-free operator Derivatives.Jacobian returns',
    'UFUNC_MAXARGS': 'there are no generated ufuncs: the stencil kernels are parameterised by ksfd_physics',
    'UfuncifyCodeWrapperMultiple': 'there are no generated ufuncs',
    'ufuncify': 'there are no generated ufuncs',
    'StencilUfunc': 'there are no generated ufuncs',
    'cartesian_product': 'row/column index arrays of the assembled Jacobian are not built',
    'spatial_expression': 'use SpatialExpression (sources and initial values are evaluated by it)',
    'Solution': 'the FEniCS-era Solution base class is not on the time-stepping path; use SolutionParameters',
    'makeKSFDSolver': 'dead code in the reference (ksfdmakesolver.py); use ksfdsolver2.main / implicitTS',
    'remap_from_files': 'listed in the reference\'s __all__ but defined nowhere in its tree',
}


def __getattr__(name):
    # Grid, Derivatives, implicitTS, TimeSeries ... load torch / the CUDA library on first use
    if name in _REPLACED:
        raise AttributeError('KSFD.%s is not provided by ksfd_b200: %s' % (name, _REPLACED[name]))
    return getattr(_impl, name)
