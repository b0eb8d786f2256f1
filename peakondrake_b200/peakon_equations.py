"""
The peakon SDE of the reference (`PeakonEquations`, peakon_equations.py:13-65), batched over
realisations.  Two doubles per realisation per step, driven by the SAME Brownian increments as the
PDE (it shares `dW_nums` with StochasticFunctions, peakon_equations.py:15) -- so it consumes the
increments the step kernel used (imported, or exported from the in-kernel Philox stream).

State is tiny (p, q per realisation), so this runs vectorised on the host; the PDE kernel is the
hot path.  `pure_xi`, `pure_xi_x`, `pure_xi_xx` are CG1 fields (stochastic_functions.py:175-182)
evaluated at q by CG1 point evaluation (`.at`, peakon_equations.py:42-44).
"""
import numpy as np
from peakondrake_b200.stochastic_functions import xi_functions, xi_derivative_functions, circulant_tridiag_solve


def _gauss01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def pure_xi_tables(ne, Ld, Xi_family, num_Xis, sigma, nquad=3):
    """
    pure_xi    = interpolate(Xi)                        (stochastic_functions.py:175)
    pure_xi_x  = project(Xi.dx(0))   -- analytic derivative, L2-projected to CG1 (:179);
                 UFL degree estimate 1 + (1+2) = 4 -> 3 Gauss points
    pure_xi_xx = project(pure_xi_x.dx(0))               (:180)  -- exact for a CG1 argument
    All derivative tables are zero for 'constant' (:178).
    """
    h = Ld / ne
    x = np.arange(ne) * h
    fs = xi_functions(Xi_family, num_Xis, sigma, Ld) if num_Xis else []
    dfs = xi_derivative_functions(Xi_family, num_Xis, sigma, Ld) if num_Xis else []
    xi = np.zeros((num_Xis, ne)); xi_x = np.zeros((num_Xis, ne)); xi_xx = np.zeros((num_Xis, ne))
    q, w = _gauss01(nquad)
    for n, f in enumerate(fs):
        v = np.asarray(f(x), dtype=float) * np.ones(ne)
        v[0] = np.asarray(f(np.array([Ld])), dtype=float).ravel()[0]
        xi[n] = v
        if Xi_family == 'constant':
            continue
        bl = np.zeros(ne); br = np.zeros(ne)
        for qq, ww in zip(q, w):
            dfq = np.asarray(dfs[n](x + qq * h), dtype=float)
            bl += ww * h * (1 - qq) * dfq
            br += ww * h * qq * dfq
        xi_x[n] = circulant_tridiag_solve(h / 6.0, 2.0 * h / 3.0, bl + np.roll(br, 1))
        # project the piecewise-constant derivative of the CG1 field: b_i = (g_{i-1} + g_i) h/2
        g = (np.roll(xi_x[n], -1) - xi_x[n]) / h
        xi_xx[n] = circulant_tridiag_solve(h / 6.0, 2.0 * h / 3.0, 0.5 * h * (g + np.roll(g, 1)))
    return xi, xi_x, xi_xx


def cg1_at(table, Ld, q):
    """CG1 point evaluation of table [ne] at positions q [R] on the periodic mesh."""
    ne = table.shape[-1]
    s = np.asarray(q, dtype=float) / (Ld / ne)
    fl = np.floor(s)
    i = fl.astype(np.int64) % ne
    th = s - fl
    return (1 - th) * table[i] + th * table[(i + 1) % ne]


class PeakonEquations(object):
    """Batched (p, q) of the peakon SDE; update(dW) mirrors peakon_equations.py:34-65."""

    def __init__(self, u0, x_nodes, Ld, dt, alphasq, tables, R, peakon_speed=None, method='ito_euler'):
        self.Ld, self.dt, self.alpha, self.method = float(Ld), float(dt), float(np.sqrt(alphasq)), method
        self.xi, self.xi_x, self.xi_xx = tables
        E = np.exp(-self.Ld / self.alpha)
        u0 = np.asarray(u0, dtype=float)
        p0 = np.max(u0) * 2 * (1 - E) / (1 + E)                       # :28
        q0 = x_nodes[np.argmax(u0)]                                   # :29 (CG1 projection of x)
        self.p = np.full(R, peakon_speed if peakon_speed is not None else p0, dtype=float)
        self.q = np.full(R, q0, dtype=float)

    def update(self, dW):
        """dW: [R, num_Xis] increments of this step."""
        E = np.exp(-self.Ld / self.alpha)
        p, q, dt = self.p, self.q, self.dt
        dp = np.zeros_like(p)
        dq = (p / 2) * (1 + E) / (1 - E) * dt
        for k in range(self.xi.shape[0]):
            xi = cg1_at(self.xi[k], self.Ld, q)
            xi_x = cg1_at(self.xi_x[k], self.Ld, q)
            xi_xx = cg1_at(self.xi_xx[k], self.Ld, q)
            w = dW[:, k]
            if self.method == 'ito_euler':
                dp = dp + p / 2 * (xi_x * xi_x - xi * xi_xx) * dt - p * xi_x * w
                dq = dq + 0.5 * xi * xi_x * dt + xi * w
            elif self.method == 'milstein':
                dp = dp + p / 2 * (xi_x * xi_x - xi * xi_xx) * w ** 2 - p * xi_x * w
                dq = dq + 0.5 * xi * xi_x * w ** 2 + xi * w
        self.p = p + dp
        q = q + dq
        q = np.where(q < 0, q + self.Ld, q)
        q = np.where(q > self.Ld, q - self.Ld, q)
        self.q = q
