"""ctypes access to the CPU emulation harness (tests/emu/libpk_emu.so) -- test infrastructure."""
import ctypes
import numpy as np

_lib = None


def lib():
    global _lib
    if _lib is None:
        import __graft_entry__ as g
        _lib = ctypes.CDLL(g.build_emu())
        _lib.pk_emu_step.restype = ctypes.c_int
        _lib.pk_emu_philox_normal.restype = ctypes.c_double
        _lib.pk_emu_philox_normal.argtypes = [ctypes.c_ulonglong] + [ctypes.c_uint] * 4
    return _lib


def _p(a, t=ctypes.c_double):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


def emu_step(ne, Ld, dt, mode, u, u1, dF, dG, tabs, dW, nsteps, xi_const=0.0, order=0, T=0, M=0,
             alphasq=1.0, nupd=1, seed=0, rid0=0, export=False, nxi=None, step0=0):
    R = u.shape[0]
    if nxi is None:
        nxi = 0 if dW is None else dW.shape[2]
    fs = -np.ones(R, np.int32); il = np.zeros(R, np.int32); itot = np.zeros(R, np.int32); fr = np.zeros(R, np.int32)
    xi, xix, xixx = tabs if tabs is not None else (None, None, None)
    dWo = np.zeros((nsteps, R, nxi)) if export else None
    rc = lib().pk_emu_step(ctypes.c_int(ne), ctypes.c_double(Ld), ctypes.c_double(dt), ctypes.c_double(alphasq),
                           mode, nxi, nupd, ctypes.c_double(xi_const), R, nsteps, step0, _p(u), _p(u1), _p(dF), _p(dG),
                           _p(xi), _p(xix), _p(xixx), _p(dW), _p(dWo), ctypes.c_ulonglong(seed), rid0,
                           _p(fs, ctypes.c_int), _p(il, ctypes.c_int), _p(itot, ctypes.c_int), _p(fr, ctypes.c_int),
                           order, T, M)
    assert rc == 0, rc
    return dict(fail_step=fs, its_last=il, its_total=itot, reason=fr, dW=dWo)


def emu_step_upwind(ne, Ld, dt, m, u, xi, dW, nsteps, xi_const=0.0, order=0, alphasq=1.0, nupd=1, seed=0, rid0=0,
                    nxi=None, export=False):
    R = u.shape[0]
    if nxi is None:
        nxi = 0 if dW is None else dW.shape[2]
    xarr = 1 if (xi is not None and nxi > 0) else 0
    fs = -np.ones(R, np.int32)
    dWo = np.zeros((nsteps, R, nxi)) if export else None
    L = lib()
    L.pk_emu_step_upwind.restype = ctypes.c_int
    rc = L.pk_emu_step_upwind(ctypes.c_int(ne), ctypes.c_double(Ld), ctypes.c_double(dt), ctypes.c_double(alphasq),
                              xarr, nxi, nupd, ctypes.c_double(xi_const), R, nsteps, 0, _p(m), _p(u), _p(xi), _p(dW), _p(dWo),
                              ctypes.c_ulonglong(seed), rid0, _p(fs, ctypes.c_int), order)
    assert rc == 0, rc
    return dict(fail_step=fs, dW=dWo)
