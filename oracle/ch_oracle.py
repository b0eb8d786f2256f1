"""
CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A NumPy/SciPy restatement of the weak forms that peakondrake hands to
Firedrake/PETSc for the stochastic Camassa-Holm timestep.  Only `tests/`,
`__graft_entry__.smoke()` and the CPU legs of `bench.py` may import this file.

PARITY UNPINNED: the reference's tests contain no assertions or golden values
 and Firedrake cannot be installed in this image, so this
oracle cannot be checked against outputs of the reference itself.  It is pinned
instead by (a) operator identities and manufactured known-answer tests in
tests/test_oracle.py and (b) the reference's own pure-NumPy helpers
(`find_min/find_max/find_mus/peakon_diagnostics`, `PeakonEquations.update`)
executed verbatim behind stub modules (tests/golden/make_golden.py).

Everything here is written from the weak forms, by Gauss quadrature on cells,
with a monolithic sparse LU per Newton iteration -- i.e. the way the reference
solves it (reference equations.py:91-95), *not* the way the CUDA path solves it
(closed-form element integrals + block elimination).  Agreement of the two is
therefore a meaningful check.

Reference citations are `file:line` into /root/reference/peakondrake/.
Mesh convention: CG1 node i at x_i = i*h, cell i = [x_i,
x_{i+1}] with nodes (i, i+1 mod ne); DG0 dof i = cell i at x_i + h/2.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# ----------------------------------------------------------------------------
# quadrature
# ----------------------------------------------------------------------------

def gauss(npts):
    """Gauss-Legendre points/weights on the reference cell [0, 1]."""
    x, w = np.polynomial.legendre.leggauss(npts)
    return 0.5 * (x + 1.0), 0.5 * w


class Mesh1D(object):
    """PeriodicIntervalMesh(ne, Ld) as used at experiment.py:171,178,208."""

    def __init__(self, ne, Ld):
        self.ne = int(ne)
        self.Ld = float(Ld)
        self.h = self.Ld / self.ne
        self.x_nodes = np.arange(self.ne) * self.h          # CG1 dof coordinates
        # DG0 coordinates = interpolate(x) into DG0 (simulation.py:265): cell mid-points
        self.x_cells = (np.arange(self.ne) + 0.5) * self.h
        self.left = np.arange(self.ne)
        self.right = (np.arange(self.ne) + 1) % self.ne

    # -- constant CG1 operators, assembled cell by cell --------------------------
    def _assemble(self, Ke):
        ne = self.ne
        rows = np.concatenate([self.left, self.left, self.right, self.right])
        cols = np.concatenate([self.left, self.right, self.left, self.right])
        vals = np.concatenate([np.full(ne, Ke[0, 0]), np.full(ne, Ke[0, 1]),
                               np.full(ne, Ke[1, 0]), np.full(ne, Ke[1, 1])])
        return sp.csr_matrix((vals, (rows, cols)), shape=(ne, ne))

    def mass(self):        # int phi_a phi_b
        return self._assemble(self.h / 6.0 * np.array([[2.0, 1.0], [1.0, 2.0]]))

    def stiffness(self):   # int phi_a' phi_b'
        return self._assemble(1.0 / self.h * np.array([[1.0, -1.0], [-1.0, 1.0]]))

    def deriv(self):       # D[a,b] = int psi_a' phi_b
        return self._assemble(0.5 * np.array([[-1.0, -1.0], [1.0, 1.0]]))


# ----------------------------------------------------------------------------
# noise tables: stochastic_functions.py:66-201
# ----------------------------------------------------------------------------

def xi_expressions(Xi_family, num_Xis, sigma, Ld):
    """List of callables xi_n(x) following stochastic_functions.py:66-118."""
    fs = []
    if Xi_family in ('sines', 'double_sines', 'high_freq_sines'):
        for n in range(num_Xis):
            if Xi_family == 'sines':
                k = 2 * (n + 1) * np.pi / Ld                      # :69,71
            elif Xi_family == 'double_sines':
                k = 4 * (n + 1) * np.pi / Ld                      # :76,78
            else:
                k = (2 * (n + 1) + 10) * np.pi / Ld               # :83,85
            if (n + 1) % 2 == 1:
                fs.append(lambda x, k=k: sigma * np.sin(k * x))
            else:
                fs.append(lambda x, k=k: sigma * np.cos(k * x))
    elif Xi_family == 'gaussians':
        for n in range(num_Xis):                                  # :89
            fs.append(lambda x, n=n: sigma * 0.5 * num_Xis * np.exp(
                -((x - Ld * (n + 1) / (num_Xis + 1.0)) / 2.) ** 2))
    elif Xi_family == 'quadratic':
        if num_Xis > 1:
            raise NotImplementedError('Quadratic Xi not yet implemented for more than one Xi')

        def quad(x):                                              # :95-103 (sigma only on one branch, as written)
            x = np.asarray(x, dtype=float)
            inner = np.where(x < 3 * Ld / 4, sigma * (x - 3 * Ld / 4) ** 2, 0.0)
            b3 = np.where(x > 5 * Ld / 8, inner, (x - Ld / 2) ** 2 + Ld ** 2 / 32)
            b2 = np.where(x > 3 * Ld / 8, b3, (x - Ld / 4) ** 2)
            return 32 / (Ld * Ld) * np.where(x > Ld / 4, b2, 0.0)
        fs.append(quad)
    elif Xi_family == 'proper_peak':
        if num_Xis > 1:
            raise NotImplementedError('Quadratic Xi not yet implemented for more than one Xi')
        fs.append(lambda x: sigma * 0.5 * 2 / (np.exp(x - Ld / 2) + np.exp(-x + Ld / 2)))   # :108
    elif Xi_family == 'constant':
        if num_Xis > 1:
            raise NotImplementedError('Constant Xi not yet implemented for more than one Xi')
        fs.append(lambda x: sigma * (np.sin(0 * np.pi * x / Ld) + 1))                      # :114
    else:
        raise NotImplementedError('Xi_family %s not implemented' % Xi_family)
    return fs


def noise_tables(mesh, Xi_family, num_Xis, sigma, scheme='hydrodynamic', quad_pts=2):
    """
    Returns (xi, xi_x, xi_xx), each [num_Xis, ne].

    xi    : nodal interpolant of the analytic xi_n (Interpolator, :185-187).  On the
            periodic mesh node 0 is written from cell 0 (x=0) and from the last
            cell (x=Ld); we take the last writer (x=Ld) -- SURVEY.md App. A.5.
    xi_x  : CG1 weak derivative  int phi Xi_x + int phi' Xi = 0  with the ANALYTIC Xi
            (:135), integrated with the quadrature UFL would pick (degree 3 -> 2 pts).
    xi_xx : int phi Xi_xx + int phi' Xi_x = 0 with the CG1 Xi_x (:136) -- exact.
    Both zero for 'constant' (:146-148) and for schemes other than hydrodynamic
    (:121,189,232 -- e.g. 'no_gradient' never receives them).
    """
    ne, h = mesh.ne, mesh.h
    fs = xi_expressions(Xi_family, num_Xis, sigma, mesh.Ld)
    xi = np.zeros((num_Xis, ne))
    xi_x = np.zeros((num_Xis, ne))
    xi_xx = np.zeros((num_Xis, ne))
    M = sp.csc_matrix(mesh.mass())
    D = mesh.deriv()
    Mlu = spla.splu(M)
    qx, qw = gauss(quad_pts)
    for n, f in enumerate(fs):
        vals = f(mesh.x_nodes)
        vals = np.array(vals, dtype=float) * np.ones(ne)
        vals[0] = float(f(np.array([mesh.Ld]))[0]) if np.ndim(f(np.array([mesh.Ld]))) else float(f(mesh.Ld))
        xi[n] = vals
        if scheme == 'hydrodynamic' and Xi_family != 'constant':
            # b_a = - int phi_a' Xi   (cell loop, quadrature)
            b = np.zeros(ne)
            for q, w in zip(qx, qw):
                xq = mesh.x_nodes + q * h          # quadrature point in every cell (cell coords, not wrapped)
                fq = f(xq) * w * h
                np.add.at(b, mesh.left, -(-1.0 / h) * fq)
                np.add.at(b, mesh.right, -(1.0 / h) * fq)
            xi_x[n] = Mlu.solve(b)
            xi_xx[n] = Mlu.solve(-(D @ xi_x[n]))
    return xi, xi_x, xi_xx


# ----------------------------------------------------------------------------
# initial conditions: initial_conditions.py:25-55, 83-121
# ----------------------------------------------------------------------------

def ic_expression(ic, Ld, alphasq=1.0, peak_width=1.0, deltax=None):
    alpha = np.sqrt(alphasq)
    w = peak_width
    exp = np.exp
    if ic == 'two_peaks':
        return lambda x: (0.2 * 2 / (exp(x - 403. / 15. * 40. / Ld) + exp(-x + 403. / 15. * 40. / Ld))
                          + 0.5 * 2 / (exp(x - 203. / 15. * 40. / Ld) + exp(-x + 203. / 15. * 40. / Ld)))
    if ic == 'gaussian':
        return lambda x: 0.5 * exp(-((x - 10.) / 2.) ** 2)
    if ic == 'gaussian_narrow':
        return lambda x: 0.5 * exp(-((x - 10.) / 1.) ** 2)
    if ic == 'gaussian_wide':
        return lambda x: 0.5 * exp(-((x - 10.) / 3.) ** 2)
    if ic == 'peakon':
        return lambda x: np.where(x < Ld / 2., exp((x - Ld / 2) / alpha), exp(-(x - Ld / 2) / alpha))
    if ic == 'one_peak':
        return lambda x: 0.5 * 2 / (exp(x - 203. / 15. * 40. / Ld) + exp(-x + 203. / 15. * 40. / Ld))
    if ic == 'proper_peak':
        return lambda x: 0.5 * 2 / (exp(x - Ld / 4) + exp(-x + Ld / 4))
    if ic == 'new_peak':
        return lambda x: 0.5 * 2 / (exp((x - Ld / 4) / w) + exp((-x + Ld / 4) / w))
    if ic == 'flat':
        return lambda x: 2 * np.pi ** 2 / (9 * 40 ** 2) + 0 * x
    if ic == 'fast_flat':
        return lambda x: 0.1 + 0 * x
    if ic == 'coshes':
        return lambda x: (2000 * np.cosh((2000 ** 0.5 / 2) * (x - 0.75)) ** (-2)
                          + 1000 * np.cosh(1000 ** 0.5 / 2 * (x - 0.25)) ** (-2))
    if ic == 'd_peakon':
        return lambda x: exp(-np.sqrt((x - Ld / 2) ** 2 + 1 * deltax ** 2) / alpha)
    if ic == 'zero':
        return lambda x: 0.0 * x
    if ic == 'two_peakons':
        return lambda x: np.where(x < Ld / 4, exp((x - Ld / 4) / alpha) - exp(-(x + Ld / 4) / alpha),
                                  np.where(x < 3 * Ld / 4, exp(-(x - Ld / 4) / alpha) - exp((x - 3 * Ld / 4) / alpha),
                                           exp((x - 5 * Ld / 4) / alpha) - exp(-(x - 3 * Ld / 4) / alpha)))
    if ic == 'twin_peakons':
        return lambda x: np.where(x < Ld / 4, exp((x - Ld / 4) / alpha) + 0.5 * exp((x - Ld / 2) / alpha),
                                  np.where(x < Ld / 2, exp(-(x - Ld / 4) / alpha) + 0.5 * exp((x - Ld / 2) / alpha),
                                           np.where(x < 3 * Ld / 4, exp(-(x - Ld / 4) / alpha) + 0.5 * exp(-(x - Ld / 2) / alpha),
                                                    exp((x - 5 * Ld / 4) / alpha) + 0.5 * exp(-(x - Ld / 2) / alpha))))
    if ic == 'periodic_peakon':
        c = 0.5 / (1 - exp(-Ld / alpha))
        return lambda x: np.where(x < Ld / 2,
                                  c * (exp((x - Ld / 2) / alpha) + exp(-Ld / alpha) * exp(-(x - Ld / 2) / alpha)),
                                  c * (exp(-(x - Ld / 2) / alpha) + exp(-Ld / alpha) * exp((x - Ld / 2) / alpha)))
    if ic == 'cos_bell':
        return lambda x: np.where(x < Ld / 4, (np.cos(np.pi * (x - Ld / 8) / (2 * Ld / 8))) ** 2, 0.0)
    if ic == 'antisymmetric':
        return lambda x: (1 / (exp((x - Ld / 4) / Ld) + exp((-x + Ld / 4) / Ld))
                          - 1 / (exp((Ld - x - Ld / 4) / Ld) + exp((Ld + x + Ld / 4) / Ld)))
    raise KeyError(ic)


def _lagrange_equispaced(deg, s):
    """Values [len(s), deg+1] of the equispaced Lagrange basis of degree `deg` on [0,1]."""
    nodes = np.linspace(0.0, 1.0, deg + 1)
    out = np.ones((len(s), deg + 1))
    for a in range(deg + 1):
        for b in range(deg + 1):
            if a != b:
                out[:, a] *= (s - nodes[b]) / (nodes[a] - nodes[b])
    return out


def initial_u(mesh, ic, alphasq=1.0, peak_width=1.0):
    """
    CG1 initial u for the CG1 schemes (initial_conditions.py:83-118).
      ic == 'peakon' : L2 projection of the kinked expression itself (:113); UFL would
                       integrate phi*expr with its estimated degree (1 + (1+2) = 4 -> 3 pts).
      otherwise      : interpolate into CG5 (equispaced nodes), then L2-project the CG5
                       function to CG1 (:116-118); degree 6 integrand -> 4 Gauss points, exact.
    """
    ne, h, Ld = mesh.ne, mesh.h, mesh.Ld
    f = ic_expression(ic, Ld, alphasq, peak_width, deltax=h)
    M = sp.csc_matrix(mesh.mass())
    b = np.zeros(ne)
    if ic == 'peakon':
        qx, qw = gauss(3)
        for q, w in zip(qx, qw):
            fq = f(mesh.x_nodes + q * h) * w * h
            np.add.at(b, mesh.left, (1 - q) * fq)
            np.add.at(b, mesh.right, q * fq)
    else:
        deg = 5
        s_nodes = np.linspace(0.0, 1.0, deg + 1)
        # CG5 dof values per cell [ne, 6]; shared vertex dofs: last writing cell wins, so the
        # value at vertex i comes from cell i-1 evaluated at its right end (x = i*h, or Ld for i=0).
        xc = mesh.x_nodes[:, None] + s_nodes[None, :] * h
        vals = f(xc) * np.ones_like(xc)
        vals[:, 0] = np.roll(vals[:, -1], 1)
        qx, qw = gauss(4)
        L = _lagrange_equispaced(deg, qx)           # [nq, 6]
        for iq, (q, w) in enumerate(zip(qx, qw)):
            fq = (vals @ L[iq]) * w * h
            np.add.at(b, mesh.left, (1 - q) * fq)
            np.add.at(b, mesh.right, q * fq)
    return spla.splu(M).solve(b)


# ----------------------------------------------------------------------------
# the timestep: equations.py:63-145, 193-203
# ----------------------------------------------------------------------------

class SNESParams(object):
    """PETSc SNES defaults Firedrake leaves untouched (equations.py:93-95 sets only KSP/PC)."""
    rtol = 1e-8
    atol = 1e-50
    stol = 1e-8
    max_it = 50
    divtol = 1e4


class ConvergenceError(Exception):
    pass


class CHOracle(object):
    """
    One realisation of the stochastic CH timestep, scheme in
    {'hydrodynamic', 'no_gradient', 'test'} with 'midpoint' timestepping.
    """

    def __init__(self, ne, Ld, dt, alphasq=1.0, scheme='hydrodynamic', timestepping='midpoint',
                 c0=0.0, gamma=0.0):
        if not (alphasq > 0.0 and gamma == 0.0):
            # equations.py:150-167
            raise NotImplementedError('Only the Camassa-Holm set-up (alphasq>0, gamma=0) is implemented.')
        if not (scheme in ('hydrodynamic', 'no_gradient', 'test') and timestepping == 'midpoint'):
            raise ValueError('Scheme %s and timestepping %s either not compatible or not recognised.'
                             % (scheme, timestepping))           # equations.py:147-148
        self.mesh = Mesh1D(ne, Ld)
        self.ne, self.h, self.dt, self.alphasq = self.mesh.ne, self.mesh.h, float(dt), float(alphasq)
        self.scheme = scheme
        self.M = self.mesh.mass()
        self.K = self.mesh.stiffness()
        self.D = self.mesh.deriv()
        self.H = (self.M + self.alphasq * self.K).tocsr()
        self.nf = 1 if scheme == 'test' else 3
        self.w1 = np.zeros(self.nf * self.ne)     # equations.py:71 -- starts at zero, never reset
        self.last_its = 0
        self.snes = SNESParams()
        self.qx, self.qw = gauss(2)               # all hot-path integrands are cubic per cell

    # -- cell-quadrature residual and Jacobian of the nonlinear parts -------------
    def _nonlinear(self, uh, dXi, dXi_x, dXi_xx, want_jac):
        """
        Quadrature assembly of
          N1 = int psi uh_x dv,  f = Dt int phi (uh^2 + a2/2 uh_x^2),
          g  = int zeta (2 uh dXi_x + a2 uh_x dXi_xx)         (equations.py:81,85,87)
        and (optionally) their derivatives with respect to the nodal values of uh.
        """
        m, h, Dt, a2 = self.mesh, self.h, self.dt, self.alphasq
        L, Rr = m.left, m.right
        a0, a1 = uh[L], uh[Rr]
        X0, X1 = dXi[L], dXi[Rr]
        x0, x1 = dXi_x[L], dXi_x[Rr]
        y0, y1 = dXi_xx[L], dXi_xx[Rr]
        ax = (a1 - a0) / h
        N1 = np.zeros(self.ne); f = np.zeros(self.ne); g = np.zeros(self.ne)
        if want_jac:
            J = {k: [np.zeros(self.ne) for _ in range(4)] for k in ('N1', 'f', 'g')}  # (00,01,10,11) per cell
        for q, w in zip(self.qx, self.qw):
            p0, p1 = 1 - q, q
            wt = w * h
            uq = p0 * a0 + p1 * a1
            vq = Dt * uq + p0 * X0 + p1 * X1
            xq = p0 * x0 + p1 * x1
            yq = p0 * y0 + p1 * y1
            n1 = ax * vq
            fq = Dt * (uq * uq + 0.5 * a2 * ax * ax)
            gq = 2 * uq * xq + a2 * ax * yq
            for test, p in ((L, p0), (Rr, p1)):
                np.add.at(N1, test, wt * p * n1)
                np.add.at(f, test, wt * p * fq)
                np.add.at(g, test, wt * p * gq)
            if want_jac:
                # d/d a_b of the integrands (b = 0,1): d uq = p_b, d ax = -+1/h, d vq = Dt p_b
                for b, (pb, dax) in enumerate(((p0, -1.0 / h), (p1, 1.0 / h))):
                    dn1 = dax * vq + ax * Dt * pb
                    dfq = Dt * (2 * uq * pb + a2 * ax * dax)
                    dgq = 2 * pb * xq + a2 * dax * yq
                    for a_, p in enumerate((p0, p1)):
                        J['N1'][2 * a_ + b] += wt * p * dn1
                        J['f'][2 * a_ + b] += wt * p * dfq
                        J['g'][2 * a_ + b] += wt * p * dgq
        if not want_jac:
            return N1, f, g, None
        rows = np.concatenate([L, L, Rr, Rr]); cols = np.concatenate([L, Rr, L, Rr])
        mats = {k: sp.csr_matrix((np.concatenate(v), (rows, cols)), shape=(self.ne, self.ne))
                for k, v in J.items()}
        return N1, f, g, mats

    def _test_scheme_terms(self, uh, dXi, want_jac):
        """'test' scheme: - int psi uh (dv)_x   (equations.py:140)."""
        m, h, Dt = self.mesh, self.h, self.dt
        L, Rr = m.left, m.right
        a0, a1 = uh[L], uh[Rr]
        v0 = Dt * a0 + dXi[L]; v1 = Dt * a1 + dXi[Rr]
        vx = (v1 - v0) / h
        T = np.zeros(self.ne)
        if want_jac:
            Jc = [np.zeros(self.ne) for _ in range(4)]
        for q, w in zip(self.qx, self.qw):
            p0, p1 = 1 - q, q
            wt = w * h
            uq = p0 * a0 + p1 * a1
            for test, p in ((L, p0), (Rr, p1)):
                np.add.at(T, test, -wt * p * uq * vx)
            if want_jac:
                for b, (pb, dvx) in enumerate(((p0, -Dt / h), (p1, Dt / h))):
                    d = -(pb * vx + uq * dvx)
                    for a_, p in enumerate((p0, p1)):
                        Jc[2 * a_ + b] += wt * p * d
        if not want_jac:
            return T, None
        rows = np.concatenate([L, L, Rr, Rr]); cols = np.concatenate([L, Rr, L, Rr])
        return T, sp.csr_matrix((np.concatenate(Jc), (rows, cols)), shape=(self.ne, self.ne))

    def residual(self, w1, u, dXi, dXi_x, dXi_xx, want_jac=False):
        ne = self.ne
        if self.scheme == 'test':
            u1 = w1
            uh = 0.5 * (u1 + u)
            T, JT = self._test_scheme_terms(uh, dXi, want_jac)
            R = self.M @ (u1 - u) + T
            J = (self.M + 0.5 * JT).tocsc() if want_jac else None
            return R, J
        u1, dF, dG = w1[:ne], w1[ne:2 * ne], w1[2 * ne:]
        uh = 0.5 * (u1 + u)
        N1, f, g, mats = self._nonlinear(uh, dXi, dXi_x, dXi_xx, want_jac)
        # equations.py:80-87 (hydrodynamic) / :115-122 (no_gradient: + psi dF_x instead of - psi_x dF)
        if self.scheme == 'hydrodynamic':
            gradF = -(self.D @ dF)
            JF = -self.D
        else:
            gradF = self.D.T @ dF          # int psi dF_x  = (D^T dF)_a
            JF = self.D.T
        Rpsi = self.M @ (u1 - u) + N1 + gradF + self.M @ dG
        Rphi = self.H @ dF - f
        Rzeta = self.H @ dG - g
        R = np.concatenate([Rpsi, Rphi, Rzeta])
        J = None
        if want_jac:
            J = sp.bmat([[self.M + 0.5 * mats['N1'], JF, self.M],
                         [-0.5 * mats['f'], self.H, None],
                         [-0.5 * mats['g'], None, self.H]], format='csc')
        return R, J

    def solve(self, u, dXi, dXi_x=None, dXi_xx=None):
        """
        Equations.solve (equations.py:193-203): SNES newtonls, basic (full-step) line search,
        preonly+LU, PETSc default tolerances; then u <- u1.  Returns the new u.
        Raises ConvergenceError exactly where SNES would report divergence.
        """
        ne = self.ne
        z = np.zeros(ne)
        dXi_x = z if dXi_x is None else dXi_x
        dXi_xx = z if dXi_xx is None else dXi_xx
        s = self.snes
        w1 = self.w1
        R, J = self.residual(w1, u, dXi, dXi_x, dXi_xx, want_jac=True)
        fnorm = np.linalg.norm(R)
        fnorm0 = fnorm
        self.res_history = [fnorm]
        it = 0
        if not np.isfinite(fnorm):
            raise ConvergenceError('DIVERGED_FNORM_NAN')
        converged = fnorm < s.atol
        while not converged:
            if it >= s.max_it:
                self.last_its = it
                raise ConvergenceError('DIVERGED_MAX_IT')
            try:
                delta = spla.splu(J).solve(-R)
            except RuntimeError:
                raise ConvergenceError('DIVERGED_LINEAR_SOLVE')
            w1 = w1 + delta
            it += 1
            R, J = self.residual(w1, u, dXi, dXi_x, dXi_xx, want_jac=True)
            fnorm = np.linalg.norm(R)
            self.res_history.append(fnorm)
            self.w1 = w1
            if not np.isfinite(fnorm):
                self.last_its = it
                raise ConvergenceError('DIVERGED_FNORM_NAN')
            if fnorm < s.atol or fnorm <= s.rtol * fnorm0:
                converged = True
            elif np.linalg.norm(delta) < s.stol * np.linalg.norm(w1):
                converged = True
            elif fnorm > s.divtol * fnorm0:
                self.last_its = it
                raise ConvergenceError('DIVERGED_DTOL')
        self.w1 = w1
        self.last_its = it
        return w1[:ne].copy()


# ----------------------------------------------------------------------------
# Brownian increments: stochastic_functions.py:41-42, 204-226
# ----------------------------------------------------------------------------

def reference_dW(seed, nsteps, num_Xis, dt, nXi_updates=1):
    """dW[n, i] exactly as the reference draws them from the legacy global MT19937 stream."""
    rs = np.random.RandomState(seed)
    dW = np.zeros((nsteps, num_Xis))
    for n in range(nsteps):
        if nXi_updates > 1:
            for j in range(nXi_updates):
                for i in range(num_Xis):
                    dW[n, i] += rs.randn() * np.sqrt(dt / nXi_updates)
        else:
            for i in range(num_Xis):
                dW[n, i] = rs.randn() * np.sqrt(dt)
    return dW


def num_steps(tmax, dt):
    """Number of passes of `while t < tmax - 0.5*dt: t += dt` (simulation.py:92-93)."""
    t, n = 0.0, 0
    while t < tmax - 0.5 * dt:
        t += dt
        n += 1
    return n


# ----------------------------------------------------------------------------
# diagnostics: diagnostic_equations.py:94-113,236-289 ; outputting.py:99-196
# ----------------------------------------------------------------------------

def du_field(mesh, u):
    """DG0 projection of u_x (diagnostic_equations.py:94-101): (u_{i+1}-u_i)/h on cell i."""
    return (u[mesh.right] - u) / mesh.h


def jump_du_field(mesh, du):
    """|jump(du)| at vertex v between cells v-1 and v (diagnostic_equations.py:103-113)."""
    return np.abs(du - np.roll(du, 1))


def m_field(orc, u):
    """M m = (M + a2 K) u   (diagnostic_equations.py:236-243)."""
    return spla.splu(sp.csc_matrix(orc.M)).solve(orc.H @ u)


def u_sde_field(mesh, p_true, q, alphasq, periodic):
    """
    Nodal interpolant of the peakon profile, diagnostic_equations.py:254-279 with the amplitude
    of :377.  NB the third branch of the periodic expression carries the reference's misplaced
    parenthesis (:270: exp(-Ld/alpha * exp(...))), reproduced as written.
    """
    Ld = mesh.Ld
    a = np.sqrt(alphasq)
    x = mesh.x_nodes.copy()
    p = p_true * 0.5 * (1 + np.exp(-Ld / a)) / (1 - np.exp(-Ld / a))
    E = np.exp(-Ld / a)
    with np.errstate(over='ignore'):
        if periodic:
            b1 = p * ((np.exp(-(x - q + Ld) / a) + E * np.exp((x - q + Ld) / a)) / (1 - E))
            b2 = p * ((np.exp(-np.sqrt((q - x) ** 2 / alphasq)) + E * np.exp(np.sqrt((q - x) ** 2 / alphasq))) / (1 - E))
            b3 = p * ((np.exp(-(q + Ld - x) / a) + np.exp(-Ld / a * np.exp((q + Ld - x) / a))) / (1 - E))
        else:
            b1 = p * np.exp(-(x - q + Ld) / a)
            b2 = p * np.exp(-np.sqrt((q - x) ** 2 / alphasq))
            b3 = p * np.exp(-(q + Ld - x) / a)
    return np.where(x < q - Ld / 2, b1, np.where(x < q + Ld / 2, b2, b3))


def scalar_diagnostics(orc, u, names, m=None, u_sde=None, periodic=False):
    """outputting.py:106-196 for the in-scope names. Returns dict name -> float."""
    mesh, a2 = orc.mesh, orc.alphasq
    M, K = orc.M, orc.K
    du = du_field(mesh, u)
    out = {}
    for d in names:
        if d == 'energy':
            out[d] = float(u @ (M @ u) + a2 * (u @ (K @ u)))                    # :110
        elif d == 'l2_u':
            out[d] = float(np.sqrt(u @ (M @ u)))                                # :114
        elif d == 'h1_u':
            out[d] = float(np.sqrt(u @ (M @ u) + u @ (K @ u)))                  # :116
        elif d == 'l2_m':
            out[d] = float(np.sqrt(m @ (M @ m)))
        elif d == 'h1_m':
            out[d] = float(np.sqrt(m @ (M @ m) + m @ (K @ m)))
        elif d == 'mass_u':
            out[d] = float(np.sum(M @ u))                                       # :124
        elif d == 'mass_m':
            out[d] = float(np.sum(M @ m))
        elif d == 'mass_m2':
            out[d] = float(m @ (M @ m))
        elif d == 'm_max':
            out[d] = float(np.max(m))
        elif d == 'max_du':
            out[d] = float(np.max(du))                                          # :140
        elif d == 'min_du':
            out[d] = float(np.min(du))
        elif d == 'max_du_loc':
            out[d] = float(mesh.x_cells[np.argmax(du)])                         # :132, find_max :254-258
        elif d == 'min_du_loc':
            out[d] = float(mesh.x_cells[np.argmin(du)])
        elif d == 'max_jump_local':
            out[d] = float(np.max(jump_du_field(mesh, du)))                     # :128
        elif d == 'max_jump_global':
            out[d] = float(np.max(du) - np.min(du))                             # :130
        elif d == 'p_pde':
            fac = (1 - np.exp(-mesh.Ld / np.sqrt(a2))) if periodic else 1.0     # :103,156-160
            out[d] = float(np.max(u) * fac)
        elif d == 'q_pde':
            out[d] = float(mesh.x_cells[np.argmax(u)])                          # :162 (DG0 coords indexed by CG1 argmax)
        elif d == 'E_0':
            out[d] = float(0.5 * (u @ (M @ u) + a2 * mesh.h * np.sum(du * du)))  # :168
        elif d == 'E_1':
            out[d] = float(0.5 * (u @ (M @ u) + a2 * (u @ (K @ u))))            # :171
        elif d == 'E_2':
            out[d] = float(0.5 * (u @ (M @ m)))                                 # :174
        elif d == 'u_error_with_sde':
            e = u - u_sde
            out[d] = float(np.sqrt(e @ (M @ e)))                                # :181 errornorm L2
        elif d == 'u_error_weak':
            # u_sde_weak solves M w = M (u - u_sde)  (diagnostic_equations.py:281-289); output ||w||_L2 (:184)
            e = u - u_sde
            out[d] = float(np.sqrt(e @ (M @ e)))
        else:
            raise ValueError('Diagnostic %s not recgonised.' % d)               # :196
    return out


# ----------------------------------------------------------------------------
# a whole run (simulation.py:84-133 semantics), used by tests and the CPU baseline
# ----------------------------------------------------------------------------

def run(ne, Ld, dt, nsteps, u0, dW=None, tables=None, alphasq=1.0, scheme='hydrodynamic',
        allow_fail=True, dump_every=0):
    """
    Advance one realisation nsteps steps. dW: [nsteps, num_Xis] (already scaled by sqrt(dt)),
    tables = (xi, xi_x, xi_xx).  Returns dict(u, w1, its[nsteps], fail_step, dumps[list of u]).
    """
    orc = CHOracle(ne, Ld, dt, alphasq, scheme)
    u = np.array(u0, dtype=float).copy()
    its = np.zeros(nsteps, dtype=np.int32)
    fail_step = -1
    dumps = []
    nX = 0 if dW is None else dW.shape[1]
    for n in range(nsteps):
        if nX:
            dXi = dW[n] @ tables[0]
            dXi_x = dW[n] @ tables[1]
            dXi_xx = dW[n] @ tables[2]
        else:
            dXi = np.zeros(ne); dXi_x = dXi; dXi_xx = dXi
        if fail_step < 0:
            try:
                u = orc.solve(u, dXi, dXi_x, dXi_xx)
                its[n] = orc.last_its
            except ConvergenceError:
                if not allow_fail:
                    raise
                fail_step = n
        if dump_every and (n + 1) % dump_every == 0:
            dumps.append(u.copy())
    return dict(u=u, w1=orc.w1.copy(), its=its, fail_step=fail_step, dumps=dumps, oracle=orc)


# ----------------------------------------------------------------------------
# upwind / ssprk3: equations.py:28-61, 172-185
# ----------------------------------------------------------------------------

class UpwindOracle(object):
    """
    DG1 upwind transport of m with the CG1 velocity us = Dt*u + Xi frozen over three SSPRK3 stages,
    then the Helmholtz solve (M + a2 K) u = P m.  m is stored [ne, 2] = (value at the left node,
    value at the right node) of every cell -- Firedrake's DG1 dof order on an interval.
    Assembled by 2-point Gauss quadrature on cells (integrands are quadratic) and explicit facet terms.
    """

    def __init__(self, ne, Ld, dt, alphasq=1.0):
        self.mesh = Mesh1D(ne, Ld)
        self.ne, self.h, self.dt, self.alphasq = self.mesh.ne, self.mesh.h, float(dt), float(alphasq)
        self.M = self.mesh.mass()
        self.K = self.mesh.stiffness()
        self.H = sp.csc_matrix(self.M + self.alphasq * self.K)
        self.Hlu = spla.splu(self.H)
        self.Me_inv = np.linalg.inv(self.h / 6.0 * np.array([[2.0, 1.0], [1.0, 2.0]]))
        self.qx, self.qw = gauss(2)

    def initial_m(self, u):
        """initial_conditions.py:65-77: M m_CG = (M + a2 K) u in CG1, then interpolate into DG1."""
        m_cg = spla.splu(sp.csc_matrix(self.M)).solve(self.H @ u)
        return np.stack([m_cg, m_cg[self.mesh.right]], axis=1)

    def dm(self, m, s):
        """Solve  psi*dm dx = psi_x*m*us dx - psi*m*us_x dx - jump(psi)*(un+ m+ - un- m-) dS  (:48-51)."""
        h = self.h
        mL, mR = m[:, 0], m[:, 1]
        s0, s1 = s, s[self.mesh.right]
        sx = (s1 - s0) / h
        r = np.zeros((self.ne, 2))
        for q, w in zip(self.qx, self.qw):
            mq = (1 - q) * mL + q * mR
            sq = (1 - q) * s0 + q * s1
            wt = w * h
            r[:, 0] += wt * ((-1.0 / h) * mq * sq - (1 - q) * mq * sx)
            r[:, 1] += wt * ((1.0 / h) * mq * sq - q * mq * sx)
        # upwind flux through vertex v (between cells v-1 and v): un = max(us.n, 0) on either side
        up_left = np.roll(mR, 1)                   # value of m just left of vertex v (cell v-1, right end)
        phi = np.maximum(s, 0.0) * up_left - np.maximum(-s, 0.0) * mL
        r[:, 0] += phi                              # cell v is the right-hand cell of vertex v
        r[:, 1] -= np.roll(phi, -1)                 # and the left-hand cell of vertex v+1
        return r @ self.Me_inv.T

    def solve(self, m, u, dXi):
        """One step (equations.py:172-185). Returns (m_new, u_new)."""
        s = self.dt * u + dXi                       # :42
        m0 = m
        d1 = self.dm(m0, s)
        t1 = m0 + d1                                # :177
        d2 = self.dm(t1, s)
        t2 = 0.75 * m0 + 0.25 * (t1 + d2)           # :179
        d3 = self.dm(t2, s)
        mn = (1.0 / 3.0) * m0 + (2.0 / 3.0) * (t2 + d3)   # :181
        b = self.h / 6.0 * (2 * mn[:, 0] + mn[:, 1]) + np.roll(self.h / 6.0 * (mn[:, 0] + 2 * mn[:, 1]), 1)
        un = self.Hlu.solve(b)                      # :58-61,185
        return mn, un


def run_upwind(ne, Ld, dt, nsteps, u0, dW=None, xi=None, alphasq=1.0, m0=None):
    orc = UpwindOracle(ne, Ld, dt, alphasq)
    u = np.array(u0, dtype=float).copy()
    m = orc.initial_m(u) if m0 is None else np.array(m0, dtype=float).copy()
    for n in range(nsteps):
        dXi = dW[n] @ xi if dW is not None and dW.shape[1] else np.zeros(ne)
        m, u = orc.solve(m, u, dXi)
    return dict(u=u, m=m, oracle=orc)


def dg1_norms(orc, m):
    """L2 / H1 norms and mass of a DG1 field [ne,2] (outputting.py:112,118,120 with m prognostic)."""
    h = orc.h
    mL, mR = m[:, 0], m[:, 1]
    l2sq = np.sum(h / 3.0 * (mL * mL + mL * mR + mR * mR))
    semi = np.sum((mR - mL) ** 2 / h)
    return dict(l2_m=float(np.sqrt(l2sq)), h1_m=float(np.sqrt(l2sq + semi)), mass_m=float(np.sum(h / 2.0 * (mL + mR))),
                m_max=float(max(mL.max(), mR.max())))


# ----------------------------------------------------------------------------
# peakon SDE: peakon_equations.py:13-65 (pinned by tests/golden: the reference's own update() run verbatim)
# ----------------------------------------------------------------------------

def cg1_at(table, Ld, q):
    """CG1 point evaluation on the periodic mesh (Function.at on a CG1 field)."""
    ne = table.shape[-1]
    s = np.asarray(q, dtype=float) / (Ld / ne)
    fl = np.floor(s)
    i = fl.astype(np.int64) % ne
    th = s - fl
    return (1 - th) * table[i] + th * table[(i + 1) % ne]


class PeakonSDE(object):
    """(p, q) of R realisations; update(dW[R, num_Xis]) follows peakon_equations.py:34-65 statement by statement."""

    def __init__(self, p0, q0, Ld, dt, alphasq, tables, R, method='ito_euler'):
        self.Ld, self.dt, self.alpha, self.method = float(Ld), float(dt), float(np.sqrt(alphasq)), method
        self.xi, self.xi_x, self.xi_xx = tables
        self.p = np.full(R, p0, dtype=float)
        self.q = np.full(R, q0, dtype=float)

    def update(self, dW):
        E = np.exp(-self.Ld / self.alpha)
        p, q, dt = self.p, self.q, self.dt
        dp = np.zeros_like(p)
        dq = (p / 2) * (1 + E) / (1 - E) * dt                                   # :37
        for k in range(self.xi.shape[0]):
            xi = cg1_at(self.xi[k], self.Ld, q)
            xi_x = cg1_at(self.xi_x[k], self.Ld, q)
            xi_xx = cg1_at(self.xi_xx[k], self.Ld, q)
            w = dW[:, k]
            if self.method == 'ito_euler':                                      # :45-46
                dp = dp + (p / 2 * (xi_x * xi_x - xi * xi_xx) * dt - p * xi_x * w)
                dq = dq + (0.5 * xi * xi_x * dt + xi * w)
            elif self.method == 'milstein':                                     # :55-56
                dp = dp + (p / 2 * (xi_x * xi_x - xi * xi_xx) * w ** 2 - p * xi_x * w)
                dq = dq + (0.5 * xi * xi_x * w ** 2 + xi * w)
        self.p = p + dp
        q = q + dq
        q = np.where(q < 0, q + self.Ld, q)                                     # :62-65
        q = np.where(q > self.Ld, q - self.Ld, q)
        self.q = q


# ----------------------------------------------------------------------------
# du_smooth: Projector(du, du_smooth) into CG5 (diagnostic_equations.py:115-119), global sparse solve
# ----------------------------------------------------------------------------

def cg5_reference_mass():
    """6x6 mass matrix and load vector of the equispaced Lagrange P5 basis on [0, 1]."""
    from numpy.polynomial import polynomial as P
    t = np.arange(6) / 5.0
    basis = []
    for k in range(6):
        c = np.array([1.0])
        for j in range(6):
            if j != k:
                c = P.polymul(c, np.array([-t[j], 1.0]) / (t[k] - t[j]))
        basis.append(c)
    xq, wq = gauss(6)                                   # exact for degree 10
    B = np.array([P.polyval(xq, b) for b in basis])     # [6, nq]
    return (B * wq) @ B.T, B @ wq


def du_smooth_field(mesh, u):
    """L2 projection of the DG0 slope onto periodic CG5: dof 5 i + k is local node k of cell i (k = 0..4),
    at x = (i + k/5) h.  Monolithic sparse LU of the 5 ne x 5 ne CG5 mass matrix."""
    ne, h = mesh.ne, mesh.h
    Mref, c = cg5_reference_mass()
    du = du_field(mesh, u)
    n = 5 * ne
    rows, cols, vals = [], [], []
    b = np.zeros(n)
    for i in range(ne):
        dofs = [(5 * i + k) % n for k in range(6)]
        for a in range(6):
            b[dofs[a]] += h * c[a] * du[i]
            for bb in range(6):
                rows.append(dofs[a]); cols.append(dofs[bb]); vals.append(h * Mref[a, bb])
    A = sp.csc_matrix((vals, (rows, cols)), shape=(n, n))
    return spla.splu(A).solve(b)


def smooth_locs(mesh, u):
    """(max_du_smooth_loc, min_du_smooth_loc) of outputting.py:135-138 with smooth_coords = CG5 node coordinates."""
    w = du_smooth_field(mesh, u)
    x = (np.arange(5 * mesh.ne) / 5.0) * mesh.h
    return float(x[np.argmax(w)]), float(x[np.argmin(w)]), w
