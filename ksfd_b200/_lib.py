"""
ctypes binding of libksfd_b200.so (the C ABI in include/ksfd_b200.h).

There is no CPU fallback: if the library is missing this module raises, and
every compute entry point needs a CUDA device.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# KSFD_B200_LIB: an experimental build of the same sources (ksfd_b200/build.py --out)
LIB_PATH = os.environ.get('KSFD_B200_LIB') or os.path.join(HERE, 'libksfd_b200.so')

MAX_LIGANDS = 7
MAX_GROUPS = 7

EXPORTS = [
    'ksfd_abi_version', 'ksfd_last_error', 'ksfd_launch_count',
    'ksfd_ctx_create', 'ksfd_ctx_destroy', 'ksfd_set_physics',
    'ksfd_set_option', 'ksfd_local_size', 'ksfd_profile_fetch',
    'ksfd_to_internal', 'ksfd_from_internal',
    'ksfd_nccl_unique_id', 'ksfd_comm_init', 'ksfd_p2p_export', 'ksfd_p2p_import',
    'ksfd_halo_exchange',
    'ksfd_groom', 'ksfd_residual', 'ksfd_velocity_max', 'ksfd_velocity',
    'ksfd_jvp_setup', 'ksfd_jvp', 'ksfd_jvp_precond', 'ksfd_pc_apply',
    'ksfd_block_diagonal',
    'ksfd_mdot', 'ksfd_maxpy', 'ksfd_norm2', 'ksfd_sum_dof0',
    'ksfd_scale_dof0', 'ksfd_mul_exp_dof0', 'ksfd_gmres', 'ksfd_ksp_solve', 'ksfd_sweep', 'ksfd_ts_step',
    'ksfd_allreduce_max', 'ksfd_allreduce_sum',
]


class Physics(C.Structure):
    """mirror of `ksfd_physics`"""
    _fields_ = [
        ('ngroups', C.c_int32), ('nlig', C.c_int32), ('cap_type', C.c_int32),
        ('reserved', C.c_int32),
        ('s2', C.c_double), ('rhomax', C.c_double), ('cushion', C.c_double),
        ('maxscale', C.c_double), ('rhomin', C.c_double), ('Umin', C.c_double),
        ('alpha', C.c_double * MAX_GROUPS), ('beta', C.c_double * MAX_GROUPS),
        ('lig_group', C.c_int32 * (MAX_LIGANDS + 1)),
        ('weight', C.c_double * MAX_LIGANDS), ('s', C.c_double * MAX_LIGANDS),
        ('gamma', C.c_double * MAX_LIGANDS), ('D', C.c_double * MAX_LIGANDS),
        ('w1', (C.c_double * 5) * 3), ('w2', (C.c_double * 5) * 3),
    ]


class KspOpts(C.Structure):
    _fields_ = [('rtol', C.c_double), ('atol', C.c_double), ('dtol', C.c_double),
                ('max_it', C.c_int32), ('restart', C.c_int32),
                ('reorth', C.c_int32), ('precond', C.c_int32),
                ('ksp_type', C.c_int32), ('reserved', C.c_int32)]


class KspResult(C.Structure):
    _fields_ = [('its', C.c_int32), ('reason', C.c_int32),
                ('rnorm0', C.c_double), ('rnorm', C.c_double)]


class TsOpts(C.Structure):
    _fields_ = [('ts_type', C.c_int32), ('adapt', C.c_int32),
                ('atol', C.c_double), ('rtol', C.c_double),
                ('clip_lo', C.c_double), ('clip_hi', C.c_double),
                ('dt_min', C.c_double), ('dt_max', C.c_double),
                ('safety', C.c_double), ('reject_safety', C.c_double),
                ('max_reject', C.c_int32), ('flags', C.c_int32),
                ('ksp', KspOpts)]


class TsResult(C.Structure):
    _fields_ = [('t_new', C.c_double), ('h_used', C.c_double),
                ('h_next', C.c_double), ('enorm', C.c_double),
                ('accepted', C.c_int32), ('rejections', C.c_int32),
                ('ksp_its', C.c_int32), ('ksp_fail', C.c_int32),
                ('vmax', C.c_double * 3), ('have_vmax', C.c_int32), ('reserved', C.c_int32)]


TIME_CB = C.CFUNCTYPE(None, C.c_double, C.c_void_p)

_lib = None


class KSFDError(RuntimeError):
    pass


def load():
    """dlopen the library (no CUDA call is made) and declare prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KSFDError(
            'ksfd_b200: %s is missing; build it with `python -m ksfd_b200.build` '
            '(there is no CPU fallback)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, dp, i32, i64 = C.c_void_p, C.c_void_p, C.c_int, C.c_int64
    lib.ksfd_abi_version.restype = C.c_int
    lib.ksfd_last_error.restype = C.c_char_p
    lib.ksfd_launch_count.restype = C.c_int64
    lib.ksfd_local_size.restype = C.c_int64
    lib.ksfd_local_size.argtypes = [vp]
    lib.ksfd_ctx_create.argtypes = [C.POINTER(vp), i32, C.POINTER(i64), i64, i64,
                                    i32, i32]
    lib.ksfd_ctx_destroy.argtypes = [vp]
    lib.ksfd_set_physics.argtypes = [vp, C.POINTER(Physics)]
    lib.ksfd_set_option.argtypes = [vp, C.c_char_p, i64]
    lib.ksfd_profile_fetch.argtypes = [vp, C.POINTER(C.c_double), vp]
    lib.ksfd_to_internal.argtypes = [vp, dp, dp, i32, vp]
    lib.ksfd_from_internal.argtypes = [vp, dp, dp, i32, vp]
    lib.ksfd_nccl_unique_id.argtypes = [C.c_char_p, C.c_char_p]
    lib.ksfd_comm_init.argtypes = [vp, C.c_char_p, i32, i32, C.c_char_p]
    lib.ksfd_p2p_export.argtypes = [vp, C.c_char_p]
    lib.ksfd_p2p_import.argtypes = [vp, C.c_char_p, i32]
    lib.ksfd_halo_exchange.argtypes = [vp, dp, i32, vp]
    lib.ksfd_groom.argtypes = [vp, dp, vp]
    lib.ksfd_residual.argtypes = [vp, dp, dp, dp, dp, vp]
    lib.ksfd_velocity_max.argtypes = [vp, dp, dp, vp]
    lib.ksfd_velocity.argtypes = [vp, dp, dp, vp]
    lib.ksfd_jvp_setup.argtypes = [vp, dp, C.c_double, vp]
    lib.ksfd_jvp.argtypes = [vp, dp, dp, vp]
    lib.ksfd_jvp_precond.argtypes = [vp, dp, dp, vp]
    lib.ksfd_pc_apply.argtypes = [vp, dp, dp, vp]
    lib.ksfd_block_diagonal.argtypes = [vp, dp, vp]
    lib.ksfd_mdot.argtypes = [vp, i32, C.POINTER(C.c_void_p), dp, dp, vp]
    lib.ksfd_maxpy.argtypes = [vp, i32, C.POINTER(C.c_double),
                               C.POINTER(C.c_void_p), dp, vp]
    lib.ksfd_norm2.argtypes = [vp, dp, C.POINTER(C.c_double), vp]
    lib.ksfd_sum_dof0.argtypes = [vp, dp, C.POINTER(C.c_double), vp]
    lib.ksfd_scale_dof0.argtypes = [vp, dp, C.c_double, vp]
    lib.ksfd_mul_exp_dof0.argtypes = [vp, dp, dp, C.c_double, vp]
    lib.ksfd_gmres.argtypes = [vp, dp, dp, C.POINTER(KspOpts),
                               C.POINTER(KspResult), vp]
    lib.ksfd_ksp_solve.argtypes = [vp, dp, dp, C.POINTER(KspOpts),
                                   C.POINTER(KspResult), vp]
    lib.ksfd_sweep.argtypes = [vp, dp, dp, dp, i32, C.POINTER(C.c_double), vp]
    lib.ksfd_ts_step.argtypes = [vp, dp, C.c_double, C.c_double,
                                 C.POINTER(TsOpts), dp, TIME_CB, vp,
                                 C.POINTER(TsResult), vp]
    lib.ksfd_allreduce_max.argtypes = [vp, C.POINTER(C.c_double), i32, vp]
    lib.ksfd_allreduce_sum.argtypes = [vp, C.POINTER(C.c_double), i32, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ('ksfd_abi_version', 'ksfd_last_error',
                        'ksfd_launch_count', 'ksfd_local_size'):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().ksfd_last_error()
        raise KSFDError('KSFD Exception %d: %s' % (
            rc, msg.decode() if msg else 'unknown error'))


def launch_count():
    return int(load().ksfd_launch_count())
