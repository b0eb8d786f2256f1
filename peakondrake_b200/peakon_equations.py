"""
The peakon SDE of the reference (`PeakonEquations`, peakon_equations.py:13-65), batched over
realisations.  Two doubles per realisation per step, driven by the SAME Brownian increments as the
PDE (it shares `dW_nums` with StochasticFunctions, peakon_equations.py:15) -- so it consumes the
increments the step kernel used (imported, or exported from the in-kernel Philox stream).

The update runs on the device (pk_peakon_sde_step: one thread per realisation, any number of steps per
launch, the reference's arithmetic operation by operation), fed either with the increments the step kernel
exported / imported or -- when no PDE runs (`only_peakons`) -- with the same in-kernel Philox stream.
`pure_xi`, `pure_xi_x`, `pure_xi_xx` are CG1 fields (stochastic_functions.py:175-182) evaluated at q by
CG1 point evaluation (`.at`, peakon_equations.py:42-44); they are built once on the host (set-up).
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


def initial_pq(u0, x_nodes, Ld, alphasq, peakon_speed=None):
    """p0, q0 of peakon_equations.py:28-32."""
    alpha = float(np.sqrt(alphasq))
    E = np.exp(-float(Ld) / alpha)
    u0 = np.asarray(u0, dtype=float)
    p0 = np.max(u0) * 2 * (1 - E) / (1 + E)                           # :28
    q0 = x_nodes[np.argmax(u0)]                                       # :29 (CG1 projection of x)
    return (peakon_speed if peakon_speed is not None else p0), q0


class PeakonEquations(object):
    """Batched (p, q) of the peakon SDE on the device; update(nsteps, dW) replaces nsteps calls of
    peakon_equations.py:34-65 for every realisation of `engine`."""

    def __init__(self, engine, u0, x_nodes, tables, peakon_speed=None, method='ito_euler'):
        import torch
        self.eng, self.method = engine, method
        if method not in ('ito_euler', 'milstein'):
            raise ValueError('peakon_method must be ito_euler or milstein, not %s' % method)
        self.alpha = float(np.sqrt(engine.alphasq))
        self.E = float(np.exp(-engine.Ld / self.alpha))
        p0, q0 = initial_pq(u0, x_nodes, engine.Ld, engine.alphasq, peakon_speed)
        dev = engine.device
        self.p_dev = torch.full((engine.R,), float(p0), dtype=torch.float64, device=dev)
        self.q_dev = torch.full((engine.R,), float(q0), dtype=torch.float64, device=dev)
        self.tables = tuple(torch.from_numpy(np.ascontiguousarray(t, dtype=np.float64)).to(dev) for t in tables)

    def update(self, nsteps, dW=None, step0=0, traj_stride=0):
        """dW: device tensor [nsteps, R, num_Xis] or None (Philox of global steps step0..)."""
        return self.eng.peakon_sde_step(nsteps, self.p_dev, self.q_dev, self.tables, self.E, self.method,
                                        dW=dW, step0=step0, traj_stride=traj_stride)

    @property
    def p(self):
        return self.p_dev.cpu().numpy()

    @property
    def q(self):
        return self.q_dev.cpu().numpy()
