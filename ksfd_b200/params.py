"""
Host-side problem description, mirroring the reference's user-facing surface:

  Parser / parse_commandline   option files ('@file', '#' comments, shlex
                               quoting, '--petsc ... --' pass-through block)
                               reference KSFD/ksfdargparse.py:57-128,
                               ksfdsolver2.py:380-422
  default_parameters           reference KSFD/ksfdargparse.py:11-55
  LigandGroups / Ligand        reference KSFD/ksfdligand.py:256-746
  SolutionParameters           reference KSFD/ksfdsoln.py:58-347
  PetscOptions                 the '--petsc' list, read the way PETSc's options
                               database would be (only the keys the hot path
                               understands)

Everything here is plain Python + sympy; it produces the plain-number
`ksfd_physics` block the CUDA kernels consume (SolutionParameters.physics).
"""
import argparse
import collections
import copy
import re
import shlex

import sympy as sy

from ._lib import KSFDError


class KSFDException(Exception):
    pass


# name, default, help — same table the reference ships
default_parameters = [
    ('degree', 3, 'order of finite difference approximations'),
    ('dim', 1, 'spatial dimensions'),
    ('nelements', 8, 'number grid points in each dimension'),
    ('nwidth', 8, 'number grid points in width'),
    ('nheight', 8, 'number grid points in height'),
    ('ndepth', 8, 'number grid points in depth'),
    ('randgridnw', 0, 'random grid width'),
    ('randgridnh', 0, 'random grid height'),
    ('randgridnd', 0, 'random grid depth'),
    ('width', 1.0, 'width of spatial domain'),
    ('height', 1.0, 'height of spatial domain'),
    ('depth', 1.0, 'depth of spatial domain'),
    ('CFL_safety_factor', 0.0, 'CFL upper bound on timestep'),
    ('conserve_worms', False, 'enforce conservation of worms'),
    ('variance_rate', 0.0, 'rate of increase in random rho variance'),
    ('variance_interval', 100.0, 'frequency of increase in random rho variance'),
    ('variance_timing_function', 't/variance_interval', 'when to inject noise'),
    ('Umin', 1e-7, 'minimum allowed value of U'),
    ('rhomin', 1e-7, 'minimum allowed value of rho'),
    ('rhomax', 28000, 'approximate max value of rho'),
    ('cushion', 2000, 'cushion on rho'),
    ('maxscale', 2.0, 'scale of cap potential'),
    ('s2', 5.56e-4, 'random worm movement (sigma)'),
    ('Nworms', 0.0, 'total number of worms'),
    ('srho0', 90.0, 'standard deviation of rho(0)'),
    ('rho0', 9000.0, 'function for rho0, added to random rho0'),
    ('U0_1_1', '', 'function for U0_1_1'),
    ('ngroups', 1, 'number of ligand groups'),
    ('nligands_1', 1, 'number of ligands in group 1'),
    ('alpha_1', 1500.0, 'alpha for ligand group 1'),
    ('beta_1', 5.56e-4, 'beta for ligand group 1'),
    ('s_1_1', 0.01, 's for ligand group 1, ligand 1'),
    ('gamma_1_1', 0.01, 'gamma for ligand group 1, ligand 1'),
    ('D_1_1', 1e-6, 'D for ligand group 1, ligand 1'),
    ('maxsteps', 1000, 'maximum number of time steps'),
    ('t0', 0.0, 'initial time'),
    ('dt', 0.001, 'first time step'),
    ('lastvart', 0.0, 'last variance injection time'),
    ('tmax', 200000, 'time to simulate'),
    ('rtol', 1e-5, 'relative tolerance for step size adaptation'),
    ('atol', 1e-5, 'absolute tolerance for step size adaptation'),
]

# per-group / per-ligand defaults (reference KSFD/ksfdligand.py:578-600)
GROUP_DEFAULTS = collections.OrderedDict(alpha=1.0, beta=1.0, nligands=1)
LIGAND_DEFAULTS = collections.OrderedDict(weight=1.0, s=1.0, gamma=1.0, D=1.0,
                                          series=1, depth=0.4)


def safe_sympify(exp):
    """sympify with '' -> None, 'True'/'False' -> bool, and a readable error
    for Python keywords (reference KSFD/ksfdsym.py:55-79)."""
    import keyword
    if isinstance(exp, str):
        if exp == '':
            return None
        if exp in ('True', 'False'):
            return exp == 'True'
        for word in re.findall(r'\b\w+\b', exp):
            if word in keyword.kwlist:
                raise ValueError('expression contains keyword %s' % word)
    return sy.sympify(exp)


def find_duplicates(items):
    seen, dups = set(), []
    for x in items:
        if x in seen:
            dups.append(x)
        seen.add(x)
    return dups


def decode_value(text):
    """'k=v' right-hand side -> bool / int / float / sympy expression."""
    v = safe_sympify(text)
    if v is None or isinstance(v, bool) or getattr(v, 'is_Boolean', False):
        return bool(v)
    if v.is_Integer:
        return int(v)
    if v.is_Float:
        return float(v)
    return v


# ---------------------------------------------------------------------------
# command line
# ---------------------------------------------------------------------------
class Parser(argparse.ArgumentParser):
    """argparse with '@file' indirection (shlex syntax, '#' comments) that also
    peels off '--petsc arg ... --' blocks into namespace.petsc."""

    subsystems = ['petsc']

    def __init__(self, *args, **kwargs):
        kwargs.setdefault('fromfile_prefix_chars', '@')
        kwargs.setdefault('allow_abbrev', False)
        super().__init__(*args, **kwargs)
        self.add_argument('--petsc', action='append', default=argparse.SUPPRESS,
                          help='PETSc subsystem arguments, terminated by --')

    def convert_arg_line_to_args(self, arg_line):
        return shlex.split(arg_line, comments=True)

    def parse_args(self, args=None, namespace=None):
        import sys
        args = list(sys.argv[1:] if args is None else args)
        args = self._read_args_from_files(args)
        peeled = {s: [] for s in self.subsystems}
        for s in self.subsystems:
            flag = '--' + s
            while flag in args:
                i = args.index(flag)
                try:
                    j = args.index('--', i + 1)
                except ValueError:
                    j = len(args)
                peeled[s] += args[i + 1:j]
                del args[i:j + 1]
        # parameters may be interleaved with --options
        ns = super().parse_intermixed_args(args, namespace)
        for s in self.subsystems:
            setattr(ns, s, peeled[s])
        return ns


LIG_HELP = """
Use --showparams to see ligand and user-defined parameters
"""


def parameter_help(param_list=None, add_help=LIG_HELP):
    """`--help` epilogue: every name=value parameter with its default and meaning
    (reference ksfdsolver2.py:366-376)."""
    text = 'Parameters:\n'
    for t, d, h in (default_parameters if param_list is None else param_list):
        text += t + '=' + str(d) + ' -- ' + h + '\n'
    text += """
You may define additional user parameters for use in rho0 or sources.
These should be of type float (e.g. 'k0=10.0' rather than 'k0=10')\n
"""
    return text + add_help


def parse_commandline(args=None):
    """The ksfdsolver2.py command line (reference ksfdsolver2.py:380-422)."""
    p = Parser(description='Solve Keller-Segel PDEs (B200-native hot path)',
               epilog=parameter_help(), formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument('--cappotential', choices=['tophat', 'witch'], default='tophat')
    p.add_argument('--save', help='filename prefix in which to save results')
    p.add_argument('--check', help='filename prefix for checkpoints')
    p.add_argument('--resume', help='resume from last point of a TimeSeries')
    p.add_argument('--restart', help='restart (t=t0) from last point of a TimeSeries')
    p.add_argument('--series_retries', type=int, default=0)
    p.add_argument('--series_retry_interval', type=int, default=60)
    p.add_argument('--mpiok', action='store_true')
    p.add_argument('--showparams', action='store_true')
    p.add_argument('--noperiodic', action='store_true')
    p.add_argument('--onestep', action='store_true')
    p.add_argument('--solver', default='petsc')
    p.add_argument('--seed', type=int, default=793817931)
    p.add_argument('--source', type=str, action='append', default=[])
    p.add_argument('params', type=str, nargs='*')
    return p.parse_args(args=args, namespace=argparse.Namespace())


class PetscOptions:
    """The '--petsc' list as a key/value database ('-ts_type rosw', ...)."""

    def __init__(self, args=()):
        self.db = collections.OrderedDict()
        args = list(args)
        i = 0
        while i < len(args):
            a = args[i]
            if a.startswith('-') and not _is_number(a):
                key = a.lstrip('-')
                if i + 1 < len(args) and (not args[i + 1].startswith('-')
                                          or _is_number(args[i + 1])):
                    self.db[key] = args[i + 1]
                    i += 2
                else:
                    self.db[key] = ''
                    i += 1
            else:
                i += 1

    def get(self, key, default=None):
        return self.db.get(key, default)

    def getReal(self, key, default=None):
        if key not in self.db:
            if default is None:
                raise KeyError(key)
            return default
        return float(self.db[key])

    def getInt(self, key, default=None):
        if key not in self.db:
            if default is None:
                raise KeyError(key)
            return default
        return int(self.db[key])

    def getRealArray(self, key, default=None):
        if key not in self.db:
            return default
        return [float(x) for x in self.db[key].split(',')]


def _is_number(s):
    try:
        float(s)
        return True
    except ValueError:
        return False


_petsc_options = PetscOptions()


def petsc_init(args=()):
    """Stand-in for petsc4py.init(args): remember the --petsc list."""
    global _petsc_options
    _petsc_options = PetscOptions(args)
    return _petsc_options


def petsc_options():
    return _petsc_options


# ---------------------------------------------------------------------------
# Parameter / ParameterList (names kept for `from KSFD import ParameterList`,
# reference KSFD/ksfdligand.py:14-63, 65-240): an ordered name -> value table
# whose entries may live elsewhere (accessor pairs)
# ---------------------------------------------------------------------------
class Parameter:
    """Accessor pair: p() / p.val / p.get() read, p(v) / p.val = v / p.set(v) write."""

    def __init__(self, getter, setter):
        self.get, self.set = getter, setter

    def __call__(self, val=None):
        if val is not None:
            self.set(val)
        return self.get()

    val = property(lambda self: self.get(), lambda self, v: self.set(v))


class ParameterList:
    """[(key, default[, help]) | (key, Parameter, default, help), ...]"""

    def __init__(self, parameters=()):
        self.values = collections.OrderedDict()
        self.ps = collections.OrderedDict()
        self.defaults = collections.OrderedDict()
        self.helps = collections.OrderedDict()
        self.keys = self.ps.keys
        self.add(parameters)

    def _own(self, key):
        store = self.values
        return Parameter(lambda: store[key], lambda v: store.__setitem__(key, v))

    def add(self, parameters):
        for item in parameters:
            if len(item) in (2, 3):
                key, default = item[0], item[1]
                doc = item[2] if len(item) == 3 else None
                acc = self.ps[key] if key in self.ps else self._own(key)
                acc.set(default)
            elif len(item) == 4:
                key, acc, default, doc = item
            else:
                raise ValueError('parameter element has length %d, 2, 3 or 4 is required'
                                 % len(item))
            self.ps[key], self.defaults[key], self.helps[key] = acc, default, doc

    def update(self, parameters):
        pairs = parameters.items() if callable(getattr(parameters, 'items', None)) else parameters
        for key, value in pairs:
            self[key] = value

    def items(self):
        return ((k, acc()) for k, acc in self.ps.items())

    __iter__ = items

    def __contains__(self, key):
        return key in self.ps

    def __getitem__(self, key):
        return self.ps[key]()

    def __setitem__(self, key, value):
        if key not in self.ps:
            self.ps[key] = self._own(key)
            self.defaults.setdefault(key, value)
            self.helps.setdefault(key, None)
        self.ps[key].set(value)

    def __delitem__(self, key):
        del self.ps[key]
        self.values.pop(key, None)
        self.defaults.pop(key, None)
        self.helps.pop(key, None)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def decode(self, params, allow_new=False):
        """apply 'key=value' strings (values decoded as on the command line)"""
        keys = [a.split('=', 1)[0] for a in params]
        dups = find_duplicates(keys)
        if dups:
            raise KSFDException('duplicated parameters: ' + ', '.join(dups))
        for a in params:
            key, _, text = a.partition('=')
            if key not in self and not allow_new:
                raise KSFDException('unknown parameter ' + key)
            self[key] = decode_value(text)

    def str(self):
        return '\n'.join('%s = %s' % kv for kv in self.items())


# ---------------------------------------------------------------------------
# ligands
# ---------------------------------------------------------------------------
class Ligand(collections.OrderedDict):
    """dict with attribute access; keys weight, s, gamma, D, series, depth,
    groupnum, ligandnum."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(e)

    def __setattr__(self, name, value):
        self[name] = value

    def name(self):
        return 'U_%d_%d' % (self.groupnum, self.ligandnum)


class LigandGroup:
    def __init__(self, groupnum=1, nligands=1):
        self.groupnum = groupnum
        self.alpha = GROUP_DEFAULTS['alpha']
        self.beta = GROUP_DEFAULTS['beta']
        self.ligands = []
        for i in range(1, nligands + 1):
            lig = Ligand(LIGAND_DEFAULTS)
            lig.groupnum, lig.ligandnum = groupnum, i
            self.ligands.append(lig)

    @property
    def nligands(self):
        return len(self.ligands)

    def names(self):
        return [l.name() for l in self.ligands]

    def V(self, Us):
        """-beta*log(alpha + sum w*U)  (reference KSFD/ksfdligand.py:527-547)"""
        if len(Us) != self.nligands:
            raise KSFDException('wrong number of ligands %d, should be %d'
                                % (len(Us), self.nligands))
        if self.nligands == 0:
            return 0.0
        sU = sum(l.weight * U for l, U in zip(self.ligands, Us))
        return -self.beta * sy.log(self.alpha + sU)

    def fourier_series(self):
        """Expand ligands with series > 1 into cosine modes in depth
        (reference KSFD/ksfdligand.py:315-388, 511-518)."""
        out = []
        for lig in self.ligands:
            try:
                n = round(lig.series)
            except (AttributeError, TypeError):
                n = 1
            n = max(int(n), 1)
            parts = []
            for i in range(n):
                li = copy.deepcopy(lig)
                li.fourier_term = i
                li.s = li.s / n
                li.weight = li.weight / n
                li.omega = sy.pi * i / li.depth
                li.gamma = li.gamma + li.D * li.omega ** 2
                parts.append(li)
            single = lig.s / lig.gamma
            series = sum(li.s / li.gamma for li in parts)
            for li in parts:
                li.s = li.s * single / series
            out += parts
        self.ligands = out
        for i, l in enumerate(self.ligands):
            l.ligandnum = i + 1


class LigandGroups:
    def __init__(self, params):
        """params: dict name -> value holding ngroups, nligands_g."""
        ngroups = int(params.get('ngroups', 1) or 1)
        self.groups = [LigandGroup(g, int(params.get('nligands_%d' % g, 1)))
                       for g in range(1, ngroups + 1)]

    def nligands(self):
        return sum(g.nligands for g in self.groups)

    def ligands(self):
        for g in self.groups:
            for l in g.ligands:
                yield l

    def names(self):
        return [l.name() for l in self.ligands()]

    def fourier_series(self):
        for g in self.groups:
            g.fourier_series()

    def V(self, Us):
        if len(Us) != self.nligands():
            raise KSFDException('provided %d ligands, need %d'
                                % (len(Us), self.nligands()))
        V, first = 0, 0
        for g in self.groups:
            V = V + g.V(Us[first:first + g.nligands])
            first += g.nligands
        return V


# ---------------------------------------------------------------------------
# parameters
# ---------------------------------------------------------------------------
NON_SYMBOLIC = [re.compile(p) for p in (
    'degree', 'dim', 'nelements', 'nwidth', 'nheight', 'ndepth', 'width',
    'Nworms', 'ngroups', r'nligands_\d+', 'maxsteps', 'rtol', 'atol',
    r'series_\d+_\d+', 'rho0', r'U0_\d+_\d+')]


class SolutionParameters:
    """
    All parameters of one run.  Members follow the reference class
    (KSFD/ksfdsoln.py:58-161): params0 (values as given, possibly sympy
    expressions), values(t) (everything numeric at time t), values0, constants,
    tdfuncs, funcs, groups/Vgroups, nligands, dim, nwidth/nheight/ndepth,
    width/height/depth, V(Us, rho, params).
    """

    def __init__(self, clargs):
        self.clargs = clargs
        given = collections.OrderedDict()
        keys = [a.split('=', 1)[0] for a in clargs.params]
        dups = find_duplicates(keys)
        if dups:
            raise KSFDException('duplicated parameters: ' + ', '.join(dups))
        for arg in clargs.params:
            if '=' not in arg:
                raise KSFDException('parameter %r is not of the form name=value' % arg)
            k, v = arg.split('=', 1)
            given[k] = decode_value(v)
        self.cparams = given
        p0 = collections.OrderedDict()
        for k, d, _ in default_parameters:
            p0[k] = decode_value(d) if isinstance(d, str) else d
        self.t0 = p0['t0']
        p0['t'] = self.t0
        # ligand structure first (needs ngroups / nligands_g)
        struct = dict(p0)
        struct.update(given)
        self.groups = LigandGroups(struct)
        # group / ligand parameters: group defaults shadow the table's group-1
        # entries unless given on the command line (reference precedence)
        for g in self.groups.groups:
            p0['alpha_%d' % g.groupnum] = g.alpha
            p0['beta_%d' % g.groupnum] = g.beta
            p0['nligands_%d' % g.groupnum] = g.nligands
            for l in g.ligands:
                for name in LIGAND_DEFAULTS:
                    p0['%s_%d_%d' % (name, g.groupnum, l.ligandnum)] = l[name]
        p0.update(given)
        for k in ('nwidth', 'nheight', 'ndepth'):
            if k not in given:
                p0[k] = p0['nelements']
        # push values back into the ligand objects, expand Fourier series
        for g in self.groups.groups:
            g.alpha = p0['alpha_%d' % g.groupnum]
            g.beta = p0['beta_%d' % g.groupnum]
            for l in g.ligands:
                for name in LIGAND_DEFAULTS:
                    l[name] = p0['%s_%d_%d' % (name, g.groupnum, l.ligandnum)]
        self.groups.fourier_series()
        for g in self.groups.groups:
            p0['nligands_%d' % g.groupnum] = g.nligands
            for l in g.ligands:
                for name in LIGAND_DEFAULTS:
                    p0['%s_%d_%d' % (name, g.groupnum, l.ligandnum)] = l[name]
        self.Vgroups = copy.deepcopy(self.groups)
        self.params0 = p0
        self.nwidth, self.nheight, self.ndepth = p0['nwidth'], p0['nheight'], p0['ndepth']
        self.width, self.height, self.depth = p0['width'], p0['height'], p0['depth']
        self.dim = p0['dim']
        self.degree = p0['degree']
        self.nligands = self.groups.nligands()
        self.rhomax, self.cushion, self.maxscale = p0['rhomax'], p0['cushion'], p0['maxscale']
        self.t0 = p0['t0']
        self._resolve()
        self.values0 = self.values()
        self.constants = collections.OrderedDict(
            (k, v) for k, v in self.values0.items() if k not in self.tdfuncs)

    # pickling: the command-line namespace is the whole state
    def __getstate__(self):
        return self.clargs

    def __setstate__(self, clargs):
        self.__init__(clargs)

    def _resolve(self):
        """Substitute parameters into one another until each is a number, a
        function of t, or a function of t and space (reference pfuncs,
        KSFD/ksfdsoln.py:254-347; here by fixed-point substitution)."""
        leaves = set(sy.symbols('t x y z')[:self.dim + 1])
        exprs = collections.OrderedDict()
        for k, v in self.params0.items():
            if k == 't':
                continue
            exprs[k] = v
        names = set(exprs)

        def is_num(v):
            return (v is None or v == '' or isinstance(v, (bool, int, float)))

        for _ in range(len(exprs) + 2):
            changed = False
            for k, v in exprs.items():
                if is_num(v):
                    continue
                free = {str(s) for s in v.free_symbols} & names
                if not free:
                    continue
                sub = {sy.Symbol(n): exprs[n] for n in free
                       if not (exprs[n] is None or exprs[n] == '')}
                if k in free:
                    raise KSFDException('parameter %s depends on itself' % k)
                nv = v.subs(sub)
                if nv != v:
                    exprs[k] = nv
                    changed = True
            if not changed:
                break
        else:
            raise KSFDException('cyclic parameter dependencies')
        funcs, tdfuncs = collections.OrderedDict(), collections.OrderedDict()
        tsym = sy.Symbol('t')
        for k, v in exprs.items():
            if is_num(v):
                funcs[k] = (lambda t, p0=v: p0)
                continue
            free = v.free_symbols
            unknown = free - leaves
            if unknown:
                raise KSFDException('parameter %s uses unknown symbols %s'
                                    % (k, sorted(map(str, unknown))))
            if not free:
                val = v.evalf()
                val = float(val) if val.is_real else val
                funcs[k] = (lambda t, p0=val: p0)
            elif free == {tsym}:
                f = sy.lambdify(tsym, v, 'math')
                funcs[k] = (lambda t, f=f: float(f(t)))
                tdfuncs[k] = funcs[k]
            else:
                funcs[k] = (lambda t, e=v: e.subs({tsym: t}))
                if tsym in free:
                    tdfuncs[k] = funcs[k]
        funcs['t'] = lambda t: t
        tdfuncs['t'] = funcs['t']
        self.funcs, self.tdfuncs = funcs, tdfuncs

    def values(self, t=None):
        t = self.t0 if t is None else t
        return collections.OrderedDict((k, f(t)) for k, f in self.funcs.items())

    def time_dependent_symbols(self):
        tds = collections.OrderedDict(self.values0)
        for k in self.tdfuncs:
            tds[k] = sy.Symbol(k)
        return tds

    def is_time_dependent(self, names):
        return any(n in self.tdfuncs for n in names)

    def V(self, Us, rho, params=None):
        """symbolic potential V(U, rho) (reference KSFD/ksfdsoln.py:147-161)"""
        p = self.values0 if params is None else params
        for g in self.Vgroups.groups:
            g.alpha, g.beta = p['alpha_%d' % g.groupnum], p['beta_%d' % g.groupnum]
            for l in g.ligands:
                l.weight = p['weight_%d_%d' % (g.groupnum, l.ligandnum)]
        th = sy.tanh((rho - p['rhomax']) / p['cushion'])
        cap = p['maxscale'] * p['s2'] * (th + 1)
        if self.clargs.cappotential == 'witch':
            cap = cap * (rho / p['rhomax'])
        return self.Vgroups.V(list(Us)) + cap

    # ---- what the kernels consume ---------------------------------------
    PHYS_KEYS = ('s2', 'rhomax', 'cushion', 'maxscale', 'rhomin', 'Umin')

    def physics_names(self):
        names = list(self.PHYS_KEYS)
        for g in self.groups.groups:
            names += ['alpha_%d' % g.groupnum, 'beta_%d' % g.groupnum]
            for l in g.ligands:
                names += ['%s_%d_%d' % (n, g.groupnum, l.ligandnum)
                          for n in ('weight', 's', 'gamma', 'D')]
        return names

    def physics_is_time_dependent(self):
        return self.is_time_dependent(self.physics_names())

    def physics(self, spacing, t=None):
        """plain-number `ksfd_physics` block at time t for grid spacing."""
        from . import core
        v = self.values(t)
        groups = []
        for g in self.groups.groups:
            ligs = []
            for l in g.ligands:
                gl = (g.groupnum, l.ligandnum)
                ligs.append(tuple(float(v['%s_%d_%d' % ((n,) + gl)])
                                  for n in ('weight', 's', 'gamma', 'D')))
            groups.append((float(v['alpha_%d' % g.groupnum]),
                           float(v['beta_%d' % g.groupnum]), ligs))
        try:
            return core.make_physics(
                self.dim, spacing, groups, float(v['s2']), float(v['rhomax']),
                float(v['cushion']), float(v['maxscale']),
                self.clargs.cappotential, float(v['rhomin']), float(v['Umin']))
        except TypeError as e:
            raise KSFDError('a kernel parameter does not evaluate to a number '
                            'at t=%r: %s' % (t, e))
