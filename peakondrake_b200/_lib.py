"""
ctypes binding of libpeakon_b200.so (C ABI: include/peakon_b200.h).

This is the binding INTEGRATION.md shows a maintainer of the reference adding.  There is NO
fallback: if the shared library is missing or the device is absent, loading/creating raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('PK_LIB_PATH') or os.path.join(_HERE, 'libpeakon_b200.so')    # PK_LIB_PATH: A/B builds (profiles/)

PK_NOISE_SCALAR, PK_NOISE_X, PK_NOISE_FULL = 0, 1, 2

# column indices of pk_diagnostics (keep in sync with include/peakon_b200.h)
D_ENERGY, D_L2U2, D_H1SEMI2, D_MASS_U, D_MAX_DU, D_MIN_DU, D_MAX_JUMP_LOCAL, D_MAX_U, \
    D_L2M2, D_H1SEMI_M2, D_MASS_M, D_M_MAX, D_UMU, D_ERR_SDE2, NDIAG = range(15)
I_ARGMAX_DU, I_ARGMIN_DU, I_ARGMAX_U, I_ARGMAX_JUMP, NIDX = range(5)

EXPORTS = ['pk_create', 'pk_destroy', 'pk_last_error', 'pk_set_tables', 'pk_step', 'pk_step_host',
           'pk_set_step_counter', 'pk_get_step_counter', 'pk_diagnostics', 'pk_get_layout',
           'pk_last_step_kernel_ms', 'pk_ensemble_stats', 'pk_peakon_suite', 'pk_step_upwind', 'pk_diagnostics_dg1', 'pk_get_work_split', 'pk_reset_warm_start', 'pk_mu', 'pk_step_kernel_ms_since', 'pk_ensemble_scalars', 'pk_peakon_sde_step', 'pk_du_smooth_loc', 'pk_launch_count', 'pk_step_dump', 'pk_step_snap', 'pk_diagnostics_rows', 'pk_peakon_suite_rows']


class PkConfig(ctypes.Structure):
    _fields_ = [('ne', ctypes.c_int32), ('R', ctypes.c_int32),
                ('Ld', ctypes.c_double), ('dt', ctypes.c_double), ('alphasq', ctypes.c_double),
                ('noise_mode', ctypes.c_int32), ('num_xis', ctypes.c_int32),
                ('nxi_updates', ctypes.c_int32), ('device', ctypes.c_int32),
                ('seed', ctypes.c_uint64), ('rid_offset', ctypes.c_uint32), ('max_it', ctypes.c_int32),
                ('rtol', ctypes.c_double), ('atol', ctypes.c_double), ('stol', ctypes.c_double),
                ('divtol', ctypes.c_double), ('xi_const', ctypes.c_double), ('scheme', ctypes.c_int32)]


_lib = None


def load():
    """Load the CUDA library; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError('peakondrake_b200: %s not found -- build it with '
                           '`python -c "import __graft_entry__ as g; g.build()"`. '
                           'There is no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, dp, ip = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p
    lib.pk_create.restype = ctypes.c_int
    lib.pk_create.argtypes = [ctypes.POINTER(PkConfig), ctypes.POINTER(vp)]
    lib.pk_destroy.restype = None
    lib.pk_destroy.argtypes = [vp]
    lib.pk_last_error.restype = ctypes.c_char_p
    lib.pk_last_error.argtypes = [vp]
    lib.pk_set_tables.restype = ctypes.c_int
    lib.pk_set_tables.argtypes = [vp, dp, dp, dp]
    lib.pk_step.restype = ctypes.c_int
    lib.pk_step.argtypes = [vp, ctypes.c_int32, dp, dp, dp, dp, dp, dp, ip, ip, ip, vp]
    lib.pk_step_host.restype = ctypes.c_int
    lib.pk_step_host.argtypes = [vp, ctypes.c_int32, dp, dp, dp, dp, dp, dp, ip, ip, ip]
    lib.pk_set_step_counter.restype = ctypes.c_int
    lib.pk_set_step_counter.argtypes = [vp, ctypes.c_int64]
    lib.pk_get_step_counter.restype = ctypes.c_int64
    lib.pk_get_step_counter.argtypes = [vp]
    lib.pk_diagnostics.restype = ctypes.c_int
    lib.pk_diagnostics.argtypes = [vp, dp, dp, ctypes.c_int32, ctypes.c_int32, dp, ip, dp, ip, vp]
    lib.pk_get_layout.restype = ctypes.c_int
    lib.pk_get_layout.argtypes = [vp] + [ctypes.POINTER(ctypes.c_int32)] * 4
    lib.pk_step_kernel_ms_since.restype = ctypes.c_double
    lib.pk_step_kernel_ms_since.argtypes = [vp, ctypes.POINTER(ctypes.c_int32)]
    lib.pk_mu.restype = ctypes.c_int
    lib.pk_mu.argtypes = [vp, dp, dp, ip, dp, ip, vp]
    lib.pk_reset_warm_start.restype = ctypes.c_int
    lib.pk_reset_warm_start.argtypes = [vp]
    lib.pk_get_work_split.restype = ctypes.c_int
    lib.pk_get_work_split.argtypes = [vp] + [ctypes.POINTER(ctypes.c_int32)] * 4
    lib.pk_last_step_kernel_ms.restype = ctypes.c_double
    lib.pk_last_step_kernel_ms.argtypes = [vp]
    lib.pk_ensemble_stats.restype = ctypes.c_int
    lib.pk_ensemble_stats.argtypes = [vp, dp, ip, dp, ip, ctypes.c_int32] + [ctypes.c_double] * 4 + [dp, vp]
    lib.pk_peakon_suite.restype = ctypes.c_int
    lib.pk_peakon_suite.argtypes = [vp, dp, ip, ip, dp, ip, vp]
    lib.pk_step_upwind.restype = ctypes.c_int
    lib.pk_step_upwind.argtypes = [vp, ctypes.c_int32, dp, dp, dp, dp, ip, vp]
    lib.pk_diagnostics_dg1.restype = ctypes.c_int
    lib.pk_diagnostics_dg1.argtypes = [vp, dp, dp, ip, dp, vp]
    lib.pk_ensemble_scalars.restype = ctypes.c_int
    lib.pk_ensemble_scalars.argtypes = [vp, dp, ip, dp, dp, vp]
    lib.pk_peakon_sde_step.restype = ctypes.c_int
    lib.pk_peakon_sde_step.argtypes = [vp, ctypes.c_int32, dp, dp, dp, ctypes.c_int64, dp, dp, dp, ctypes.c_int32,
                                       ctypes.c_double, dp, ctypes.c_int32, vp]
    lib.pk_du_smooth_loc.restype = ctypes.c_int
    lib.pk_du_smooth_loc.argtypes = [vp, dp, ip, dp, dp, ip, dp, vp]
    lib.pk_step_dump.restype = ctypes.c_int
    lib.pk_step_dump.argtypes = [vp, ctypes.c_int32, dp, dp, dp, dp, dp, dp, ip, ip, ip, ctypes.c_int32, dp, ip, dp, ip, vp]
    lib.pk_step_snap.restype = ctypes.c_int
    lib.pk_step_snap.argtypes = [vp, ctypes.c_int32, dp, dp, dp, dp, dp, dp, ip, ip, ip, ctypes.c_int32, dp, ip, vp]
    lib.pk_diagnostics_rows.restype = ctypes.c_int
    lib.pk_diagnostics_rows.argtypes = [vp, dp, ctypes.c_int32, dp, ctypes.c_int32, ctypes.c_int32, dp, ip, dp, ip, vp]
    lib.pk_peakon_suite_rows.restype = ctypes.c_int
    lib.pk_peakon_suite_rows.argtypes = [vp, dp, ctypes.c_int32, ip, ip, dp, ip, vp]
    lib.pk_launch_count.restype = ctypes.c_int64
    lib.pk_launch_count.argtypes = [vp]
    _lib = lib
    return lib


class PkError(RuntimeError):
    pass


def check(lib, ctx, rc):
    if rc != 0:
        msg = lib.pk_last_error(ctx)
        raise PkError(msg.decode() if msg else 'pk error %d' % rc)
