"""
Known-answer tests that pin the oracle: the reference's tests hold no golden values,
so the oracle is pinned by operator identities, consistency of its own Jacobian, and the structural
properties of the reference's formulation.
"""
import numpy as np
import scipy.sparse.linalg as spla
import pytest

from oracle import ch_oracle as O


def test_operator_identities():
    m = O.Mesh1D(64, 40.0)
    M, K, D = m.mass(), m.stiffness(), m.deriv()
    H = M + 1.3 * K
    assert abs(D + D.T).max() == 0.0                        # D is skew
    assert abs(H @ M - M @ H).max() < 1e-15 and abs(H @ D - D @ H).max() < 1e-15   # circulants commute
    one = np.ones(64)
    assert np.allclose(M @ one, m.h) and np.allclose(K @ one, 0) and np.allclose(D @ one, 0)
    x = m.x_nodes
    # exactness: int phi_i * (linear interpolant of a periodic f) , u^T M u = int u^2
    u = np.sin(2 * np.pi * x / 40)
    exact = 0.5 * 40 * (1 - (2 * np.pi / 64) ** 2 / 6)       # int of the squared P1 interpolant, leading terms
    assert abs(u @ (M @ u) - exact) < 1e-3


def test_jacobian_matches_finite_differences():
    ne = 48
    orc = O.CHOracle(ne, 40.0, 1e-2)
    rs = np.random.RandomState(0)
    u = 0.5 + 0.1 * rs.randn(ne)
    w = 0.1 * rs.randn(3 * ne)
    X, Xx, Xxx = 0.05 * rs.randn(ne), 0.05 * rs.randn(ne), 0.05 * rs.randn(ne)
    R0, J = orc.residual(w, u, X, Xx, Xxx, want_jac=True)
    J = J.toarray()
    eps = 1e-6
    for k in rs.choice(3 * ne, 25, replace=False):
        e = np.zeros(3 * ne); e[k] = eps
        Rp, _ = orc.residual(w + e, u, X, Xx, Xxx)
        Rm, _ = orc.residual(w - e, u, X, Xx, Xxx)
        assert np.allclose((Rp - Rm) / (2 * eps), J[:, k], atol=1e-8)


def test_two_point_gauss_is_exact_for_the_hot_path():
    """All hot-path integrands are cubic per cell: 2 and 4 Gauss points must agree to rounding."""
    ne = 40
    a = O.CHOracle(ne, 40.0, 1e-2)
    b = O.CHOracle(ne, 40.0, 1e-2)
    b.qx, b.qw = O.gauss(4)
    rs = np.random.RandomState(1)
    args = (0.1 * rs.randn(3 * ne), rs.randn(ne), rs.randn(ne), rs.randn(ne), rs.randn(ne))
    Ra, Ja = a.residual(*args, want_jac=True)
    Rb, Jb = b.residual(*args, want_jac=True)
    assert np.allclose(Ra, Rb, rtol=0, atol=1e-13) and abs(Ja - Jb).max() < 1e-13


def test_newton_iteration_counts_and_first_step_from_zero_guess():
    ne, Ld, dt = 200, 40.0, 1e-3
    u0 = O.initial_u(O.Mesh1D(ne, Ld), 'one_peak')
    orc = O.CHOracle(ne, Ld, dt)
    assert not orc.w1.any()                                   # equations.py:71
    u = orc.solve(u0, np.zeros(ne))
    assert orc.last_its == 2 and orc.res_history[0] > 1e-2    # big first residual: w1 = 0 is far from u
    u = orc.solve(u, np.zeros(ne))
    assert orc.last_its == 2 and orc.res_history[0] < 1e-2    # warm start
    assert orc.res_history[-1] <= 1e-8 * orc.res_history[0]


def test_sigma_zero_is_deterministic_and_no_gradient_equals_hydrodynamic():
    ne, Ld, dt, n = 120, 40.0, 1e-3, 5
    u0 = O.initial_u(O.Mesh1D(ne, Ld), 'proper_peak')
    tabs0 = O.noise_tables(O.Mesh1D(ne, Ld), 'sines', 3, 0.0)
    dW = O.reference_dW(7, n, 3, dt)
    a = O.run(ne, Ld, dt, n, u0, dW, tabs0)
    b = O.run(ne, Ld, dt, n, u0, None, None)
    assert np.array_equal(a['u'], b['u'])
    c = O.run(ne, Ld, dt, n, u0, None, None, scheme='no_gradient')
    assert np.allclose(c['u'], b['u'], rtol=0, atol=1e-15)


def test_constant_family_has_zero_derivative_tables_and_zero_dG():
    ne, Ld, dt, n = 100, 40.0, 1e-3, 4
    mesh = O.Mesh1D(ne, Ld)
    xi, xix, xixx = O.noise_tables(mesh, 'constant', 1, 0.3)
    assert np.all(xi == 0.3) and not xix.any() and not xixx.any()     # stochastic_functions.py:114,146-148
    u0 = O.initial_u(mesh, 'periodic_peakon')
    r = O.run(ne, Ld, dt, n, u0, O.reference_dW(1, n, 1, dt), (xi, xix, xixx))
    assert not r['w1'][2 * ne:].any()


def test_noise_tables_weak_derivatives():
    mesh = O.Mesh1D(400, 40.0)
    xi, xix, xixx = O.noise_tables(mesh, 'sines', 4, 0.2)
    x = mesh.x_nodes
    for n in range(4):
        k = 2 * (n + 1) * np.pi / 40.0
        f = np.sin if (n + 1) % 2 == 1 else np.cos
        df = (lambda z: np.cos(z)) if (n + 1) % 2 == 1 else (lambda z: -np.sin(z))
        assert np.allclose(xi[n][1:], 0.2 * f(k * x[1:]), atol=1e-15)
        assert np.allclose(xix[n], 0.2 * k * df(k * x), atol=2e-4)
        assert np.allclose(xixx[n], -0.2 * k * k * f(k * x), atol=2e-4)
    # 'no_gradient' never receives the derivative tables (stochastic_functions.py:121,189,232)
    _, a, b = O.noise_tables(mesh, 'sines', 2, 0.2, scheme='no_gradient')
    assert not a.any() and not b.any()
    with pytest.raises(NotImplementedError):
        O.noise_tables(mesh, 'constant', 2, 0.1)
    with pytest.raises(NotImplementedError):
        O.noise_tables(mesh, 'nonsense', 1, 0.1)


def test_brownian_stream_order():
    """np.random.seed(s) then randn() per step, mode-fastest; nXi_updates sub-draws sum so that a
    coarse run consumes the same stream as the fine one (stochastic_functions.py:211-226)."""
    dt, nxi = 1e-2, 3
    fine = O.reference_dW(11, 40, nxi, dt / 4)
    coarse = O.reference_dW(11, 10, nxi, dt, nXi_updates=4)
    assert np.allclose(coarse, fine.reshape(10, 4, nxi).sum(axis=1), rtol=0, atol=1e-15)
    np.random.seed(11)
    first = np.array([np.random.randn() for _ in range(6)]) * np.sqrt(dt / 4)
    assert np.array_equal(fine[:2].ravel(), first)
    assert O.num_steps(0.01, 0.001) == 10 and O.num_steps(5, 1e-3) == 5000


def test_initial_condition_projection():
    mesh = O.Mesh1D(500, 40.0)
    for ic in ('one_peak', 'proper_peak', 'periodic_peakon', 'gaussian', 'two_peaks'):
        u = O.initial_u(mesh, ic)
        f = O.ic_expression(ic, 40.0, 1.0, 1.0, mesh.h)
        assert np.abs(u - f(mesh.x_nodes))[5:].max() < 5e-3        # projection ~ interpolation + O(h^2)
    pk = O.initial_u(mesh, 'peakon')
    assert abs(pk.max() - 1.0) < 5e-2 and mesh.x_nodes[pk.argmax()] == 20.0


def test_upwind_and_unknown_schemes_raise_like_the_reference():
    with pytest.raises(ValueError):
        O.CHOracle(50, 40.0, 1e-3, scheme='conforming')                 # equations.py:147-148
    with pytest.raises(NotImplementedError):
        O.CHOracle(50, 40.0, 1e-3, alphasq=0.0, gamma=1.0)              # equations.py:150-167


def test_diagnostics_definitions():
    ne, Ld = 300, 40.0
    orc = O.CHOracle(ne, Ld, 1e-3)
    u = O.initial_u(orc.mesh, 'proper_peak')
    m = O.m_field(orc, u)
    d = O.scalar_diagnostics(orc, u, ['energy', 'E_0', 'E_1', 'E_2', 'l2_u', 'h1_u', 'mass_u', 'mass_m', 'max_du',
                                      'min_du', 'max_jump_local', 'max_jump_global', 'q_pde'], m=m)
    assert abs(d['E_0'] - d['E_1']) < 1e-13 and abs(d['E_1'] - 0.5 * d['energy']) < 1e-15
    assert abs(d['E_2'] - d['E_1']) < 1e-12                              # 1/2 int u m = 1/2 int u^2 + a2 u_x^2
    assert abs(d['mass_m'] - d['mass_u']) < 1e-12                        # int m = int u on the periodic domain
    assert d['max_jump_global'] == d['max_du'] - d['min_du']
    assert abs(d['q_pde'] - 10.0) <= orc.h
