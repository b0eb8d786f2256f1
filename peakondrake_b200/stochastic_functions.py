"""
Host-side set-up of the noise basis functions -- the part of the reference's
`StochasticFunctions.__init__` (stochastic_functions.py:58-201) that runs once per experiment.
The per-step part (`update`, :204-236) lives in the CUDA step kernel.

Outputs are the three tables xi, xi_x, xi_xx of shape [num_Xis, ne] shared by all realisations.
"""
import numpy as np

VALID_FAMILIES = ('sines', 'double_sines', 'high_freq_sines', 'gaussians', 'quadratic',
                  'proper_peak', 'constant')


def circulant_tridiag_solve(off, diag, b):
    """Solve the periodic system  off*x[i-1] + diag*x[i] + off*x[i+1] = b[i]  (last axis) by FFT."""
    n = b.shape[-1]
    k = np.arange(n // 2 + 1)
    eig = diag + 2.0 * off * np.cos(2.0 * np.pi * k / n)
    return np.fft.irfft(np.fft.rfft(b, axis=-1) / eig, n=n, axis=-1)


def xi_functions(Xi_family, num_Xis, sigma, Ld):
    """Callables xi_n(x); formulas of stochastic_functions.py:66-118 (errors as the reference raises them)."""
    fs = []
    if Xi_family == 'sines':
        ks = [2 * (n + 1) * np.pi / Ld for n in range(num_Xis)]
    elif Xi_family == 'double_sines':
        ks = [4 * (n + 1) * np.pi / Ld for n in range(num_Xis)]
    elif Xi_family == 'high_freq_sines':
        ks = [(2 * (n + 1) + 10) * np.pi / Ld for n in range(num_Xis)]
    else:
        ks = None
    if ks is not None:
        for n, k in enumerate(ks):
            fs.append((lambda x, k=k: sigma * np.sin(k * x)) if (n + 1) % 2 == 1
                      else (lambda x, k=k: sigma * np.cos(k * x)))
    elif Xi_family == 'gaussians':
        for n in range(num_Xis):
            fs.append(lambda x, n=n: sigma * 0.5 * num_Xis
                      * np.exp(-((x - Ld * (n + 1) / (num_Xis + 1.0)) / 2.) ** 2))
    elif Xi_family == 'quadratic':
        if num_Xis > 1:
            raise NotImplementedError('Quadratic Xi not yet implemented for more than one Xi')

        def quad(x):
            x = np.asarray(x, dtype=float)
            v = np.zeros_like(x)
            s1 = (x > Ld / 4) & ~(x > 3 * Ld / 8)
            s2 = (x > 3 * Ld / 8) & ~(x > 5 * Ld / 8)
            s3 = (x > 5 * Ld / 8) & (x < 3 * Ld / 4)
            v[s1] = (x[s1] - Ld / 4) ** 2
            v[s2] = (x[s2] - Ld / 2) ** 2 + Ld ** 2 / 32
            v[s3] = sigma * (x[s3] - 3 * Ld / 4) ** 2
            return 32 / (Ld * Ld) * v
        fs.append(quad)
    elif Xi_family == 'proper_peak':
        if num_Xis > 1:
            raise NotImplementedError('Quadratic Xi not yet implemented for more than one Xi')
        fs.append(lambda x: sigma * 0.5 * 2 / (np.exp(x - Ld / 2) + np.exp(-x + Ld / 2)))
    elif Xi_family == 'constant':
        if num_Xis > 1:
            raise NotImplementedError('Constant Xi not yet implemented for more than one Xi')
        fs.append(lambda x: sigma * (np.sin(0 * np.pi * x / Ld) + 1))
    else:
        raise NotImplementedError('Xi_family %s not implemented' % Xi_family)
    return fs


def xi_derivative_functions(Xi_family, num_Xis, sigma, Ld):
    """Analytic d/dx of the xi_n (what UFL's Xi_function.dx(0) evaluates, stochastic_functions.py:179)."""
    ds = []
    if Xi_family in ('sines', 'double_sines', 'high_freq_sines'):
        for n in range(num_Xis):
            k = {'sines': 2 * (n + 1), 'double_sines': 4 * (n + 1), 'high_freq_sines': 2 * (n + 1) + 10}[Xi_family] * np.pi / Ld
            ds.append((lambda x, k=k: sigma * k * np.cos(k * x)) if (n + 1) % 2 == 1
                      else (lambda x, k=k: -sigma * k * np.sin(k * x)))
    elif Xi_family == 'gaussians':
        for n in range(num_Xis):
            c = Ld * (n + 1) / (num_Xis + 1.0)
            ds.append(lambda x, c=c: sigma * 0.5 * num_Xis * np.exp(-((x - c) / 2.) ** 2) * (-(x - c) / 2.))
    elif Xi_family == 'proper_peak':
        ds.append(lambda x: -sigma * (np.exp(x - Ld / 2) - np.exp(-x + Ld / 2)) / (np.exp(x - Ld / 2) + np.exp(-x + Ld / 2)) ** 2)
    elif Xi_family == 'quadratic':
        def dquad(x):
            x = np.asarray(x, dtype=float)
            v = np.zeros_like(x)
            s1 = (x > Ld / 4) & ~(x > 3 * Ld / 8)
            s2 = (x > 3 * Ld / 8) & ~(x > 5 * Ld / 8)
            s3 = (x > 5 * Ld / 8) & (x < 3 * Ld / 4)
            v[s1] = 2 * (x[s1] - Ld / 4)
            v[s2] = 2 * (x[s2] - Ld / 2)
            v[s3] = sigma * 2 * (x[s3] - 3 * Ld / 4)
            return 32 / (Ld * Ld) * v
        ds.append(dquad)
    elif Xi_family == 'constant':
        ds.append(lambda x: np.zeros_like(np.asarray(x, dtype=float)))
    else:
        raise NotImplementedError('Xi_family %s not implemented' % Xi_family)
    return ds


_G2 = (0.5 - 0.5 / np.sqrt(3.0), 0.5 + 0.5 / np.sqrt(3.0))   # 2-point Gauss on [0,1], weights 1/2


def build_tables(ne, Ld, Xi_family, num_Xis, sigma, scheme='hydrodynamic'):
    """
    xi    [num_Xis, ne]: nodal values (Interpolator of the analytic expression, :185-187);
                         node 0 takes the value written by the last cell, i.e. at x = Ld.
    xi_x  : CG1 weak derivative  M xi_x = -int phi' Xi  with the analytic Xi under the 2-point
            Gauss rule UFL's degree estimate selects (:135);
    xi_xx : M xi_xx = -D xi_x (:136).
    Both zero unless scheme == 'hydrodynamic' (:121) and the family is not 'constant' (:146-148).
    """
    ne = int(ne)
    h = Ld / ne
    x = np.arange(ne) * h
    fs = xi_functions(Xi_family, num_Xis, sigma, Ld) if num_Xis > 0 else []
    xi = np.zeros((num_Xis, ne))
    xi_x = np.zeros((num_Xis, ne))
    xi_xx = np.zeros((num_Xis, ne))
    for n, f in enumerate(fs):
        v = np.asarray(f(x), dtype=float) * np.ones(ne)
        v[0] = np.asarray(f(np.array([Ld])), dtype=float).ravel()[0]
        xi[n] = v
        if scheme == 'hydrodynamic' and Xi_family != 'constant':
            # cell integrals I_c = int_cell Xi dx (2-pt Gauss); b_i = -(I_{i-1} - I_i)/h
            I = 0.5 * h * (np.asarray(f(x + _G2[0] * h)) + np.asarray(f(x + _G2[1] * h)))
            b = -(np.roll(I, 1) - I) / h
            xi_x[n] = circulant_tridiag_solve(h / 6.0, 2.0 * h / 3.0, b)
            # -(D xi_x)_i = -(xi_x[i-1] - xi_x[i+1]) / 2
            b2 = -0.5 * (np.roll(xi_x[n], 1) - np.roll(xi_x[n], -1))
            xi_xx[n] = circulant_tridiag_solve(h / 6.0, 2.0 * h / 3.0, b2)
    return xi, xi_x, xi_xx


def draw_dW(seed, nsteps, num_Xis, dt, nXi_updates=1):
    """
    The reference's Brownian increments for one seed, in its draw order
    (np.random.seed(seed) at :42; per step i-fastest, sub-updates j-outer :213-226).
    Used for path-wise reproduction of reference runs; the throughput path draws with Philox
    on the device instead.
    """
    rs = np.random.RandomState(seed)
    if nXi_updates > 1:
        z = rs.randn(nsteps, nXi_updates, num_Xis) * np.sqrt(dt / nXi_updates)
        dW = np.zeros((nsteps, num_Xis))
        for j in range(nXi_updates):      # same summation order as the reference's += loop
            dW += z[:, j, :]
        return dW
    return rs.randn(nsteps, num_Xis) * np.sqrt(dt)
