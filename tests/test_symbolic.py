"""
UFL-free symbolic pin of the hot-path weak form: the integrals of the reference's residual
(/root/reference/peakondrake/equations.py:80-87, hydrodynamic / midpoint) are integrated EXACTLY with sympy on CG1
elements and compared with
  (a) the oracle's 2-point Gauss cell quadrature (oracle/ch_oracle.py, residual + Jacobian), and
  (b) the closed-form element integrals of the CUDA path (CellC in peakondrake_b200/csrc/pk_core.h, executed through the
      emulation harness), including the scalings the kernel stores them with.
Firedrake is not needed: the form is a sum of polynomial integrands per cell; this is what TSFC's generated kernel
integrates (degree 3 -> 2 Gauss points, exact).
"""
import ctypes
import numpy as np
import sympy as sp

from oracle import ch_oracle as O
from emu_util import lib, _p


def exact_cell_terms():
    """N1_a, f_a, g_a (a = 0, 1: test function at the left / right node) of one cell of length h, symbolic."""
    s, h, Dt, a2 = sp.symbols('s h Dt a2', positive=True)
    a0, a1, X0, X1, x0, x1, y0, y1 = sp.symbols('a0 a1 X0 X1 x0 x1 y0 y1', real=True)
    phi = (1 - s, s)
    uh = a0 * phi[0] + a1 * phi[1]
    uh_x = (a1 - a0) / h
    dv = Dt * uh + X0 * phi[0] + X1 * phi[1]                          # equations.py:78
    dXi_x = x0 * phi[0] + x1 * phi[1]
    dXi_xx = y0 * phi[0] + y1 * phi[1]
    I = lambda e: sp.integrate(sp.expand(e) * h, (s, 0, 1))
    N1 = [I(phi[a] * uh_x * dv) for a in (0, 1)]                      # psi uh.dx(0) dvh                    :81
    f = [I(phi[a] * (uh * uh * Dt + sp.Rational(1, 2) * a2 * uh_x * uh_x * Dt)) for a in (0, 1)]     # :85
    g = [I(phi[a] * (2 * uh * dXi_x + a2 * uh_x * dXi_xx)) for a in (0, 1)]                          # :87
    syms = dict(h=h, Dt=Dt, a2=a2, a0=a0, a1=a1, X0=X0, X1=X1, x0=x0, x1=x1, y0=y0, y1=y1)
    return N1, f, g, syms


def test_closed_forms_of_the_cuda_path_equal_exact_integration(built):
    N1, f, g, S = exact_cell_terms()
    L = lib()
    L.pk_emu_cell.restype = None
    rng = np.random.RandomState(7)
    for trial in range(20):
        h, dt, a2 = 40.0 / rng.randint(50, 20001), 10.0 ** rng.uniform(-4, -2), rng.uniform(0.5, 2.0)
        vals = {k: float(rng.randn()) for k in ('a0', 'a1', 'X0', 'X1', 'x0', 'x1', 'y0', 'y1')}
        sub = {S['h']: h, S['Dt']: dt, S['a2']: a2}
        sub.update({S[k]: v for k, v in vals.items()})
        ev = lambda e: float(e.subs(sub))
        A = np.array([vals['a0'], dt * vals['a0'] + vals['X0'], 0, 0, 0, vals['x0'], vals['y0']])
        B = np.array([vals['a1'], dt * vals['a1'] + vals['X1'], 0, 0, 0, vals['x1'], vals['y1']])
        for full in (0, 1):
            out = np.zeros(18)
            L.pk_emu_cell(ctypes.c_double(h), ctypes.c_double(dt), ctypes.c_double(a2), _p(A), _p(B), full, _p(out))
            exact = [ev(N1[0]), ev(N1[1]), 0.5 * ev(f[0]), 0.5 * ev(f[1]), 0.5 * ev(g[0]), 0.5 * ev(g[1])]
            d = lambda e, b: ev(sp.diff(e, S['a%d' % b]))
            exact += [0.5 * d(N1[0], 0), 0.5 * d(N1[0], 1), 0.5 * d(N1[1], 0), 0.5 * d(N1[1], 1)]         # N1'/2
            exact += [0.25 * d(f[0], 0), 0.25 * d(f[0], 1), 0.25 * d(f[1], 0), 0.25 * d(f[1], 1)]         # f'/4
            exact += [0.5 * d(g[0], 0), 0.5 * d(g[0], 1), 0.5 * d(g[1], 0), 0.5 * d(g[1], 1)]             # g'/2
            exact = np.array(exact)
            scale = np.maximum(np.abs(exact), 1e-3 * np.abs(exact).max())
            assert (np.abs(out - exact) <= 2e-13 * scale + 1e-18).all(), (trial, full, out - exact)


def test_oracle_quadrature_equals_exact_integration_of_the_form():
    """The whole residual (all three test-function blocks, equations.py:80-87) and its Jacobian on a small periodic mesh:
    sympy's exact integrals against the oracle's quadrature assembly."""
    ne, Ld, dt, a2 = 6, 3.0, 0.05, 1.3
    h = sp.Rational(1, 2)
    rng = np.random.RandomState(3)
    names = ['u', 'u1', 'dF', 'dG', 'X', 'Xx', 'Xxx']
    num = {n: rng.randn(ne) for n in names}
    sym = {n: sp.symbols('%s_0:%d' % (n, ne), real=True) for n in names}
    s = sp.symbols('s', positive=True)
    Dt, A2 = sp.Rational(1, 20), sp.Rational(13, 10)
    R = {k: [0] * ne for k in ('psi', 'phi', 'zeta')}
    for c in range(ne):
        i0, i1 = c, (c + 1) % ne
        phi = (1 - s, s)
        fld = lambda n: sym[n][i0] * phi[0] + sym[n][i1] * phi[1]
        ddx = lambda n: (sym[n][i1] - sym[n][i0]) / h
        uh, uh_x = (fld('u1') + fld('u')) / 2, (ddx('u1') + ddx('u')) / 2
        dvh = Dt * uh + fld('X')
        for a, node in enumerate((i0, i1)):
            t, tx = phi[a], (-1 if a == 0 else 1) / h
            I = lambda e: sp.integrate(sp.expand(e) * h, (s, 0, 1))
            R['psi'][node] += I(t * (fld('u1') - fld('u')) + t * uh_x * dvh - tx * fld('dF') + t * fld('dG'))
            R['phi'][node] += I(t * fld('dF') + A2 * tx * ddx('dF') - t * uh * uh * Dt - sp.Rational(1, 2) * A2 * t * uh_x * uh_x * Dt)
            R['zeta'][node] += I(t * fld('dG') + A2 * tx * ddx('dG') - 2 * t * uh * fld('Xx') - A2 * t * uh_x * fld('Xxx'))
    sub = {sym[n][i]: float(num[n][i]) for n in names for i in range(ne)}
    exact = np.array([float(e.subs(sub)) for k in ('psi', 'phi', 'zeta') for e in R[k]])
    orc = O.CHOracle(ne, Ld, dt, a2)
    w1 = np.concatenate([num['u1'], num['dF'], num['dG']])
    Rq, Jq = orc.residual(w1, num['u'], num['X'], num['Xx'], num['Xxx'], want_jac=True)
    assert np.abs(Rq - exact).max() <= 1e-13 * max(1.0, np.abs(exact).max())
    unknowns = list(sym['u1']) + list(sym['dF']) + list(sym['dG'])
    allR = [e for k in ('psi', 'phi', 'zeta') for e in R[k]]
    Jx = np.array([[float(sp.diff(e, v).subs(sub)) for v in unknowns] for e in allR])
    assert np.abs(Jq.toarray() - Jx).max() <= 1e-13 * max(1.0, np.abs(Jx).max())
