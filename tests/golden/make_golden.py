"""
Generates tests/golden/reference_helpers.npz by running the REFERENCE's own pure-NumPy helpers
verbatim (imported from /root/reference, behind stub `firedrake` / `netCDF4` modules):

    peakondrake.outputting.find_min / find_max / find_mus / peakon_diagnostics   (outputting.py:248-450)
    peakondrake.peakon_equations.PeakonEquations.update                           (peakon_equations.py:34-65)

These are the only parts of the reference that can execute in this image.  Run once in
the authoring container:   python tests/golden/make_golden.py
The .npz travels with the repo; nothing at test time reads /root/reference.
"""
import sys
import types
import numpy as np

REF = '/root/reference'


class _Anything(object):
    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def install_stubs():
    fd = types.ModuleType('firedrake')

    class _Mod(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith('__'):
                raise AttributeError(name)
            return _Anything()
    fd = _Mod('firedrake')
    fd.__all__ = []
    fd.exp = np.exp
    fd.sqrt = np.sqrt
    sys.modules['firedrake'] = fd
    nc = _Mod('netCDF4')
    sys.modules['netCDF4'] = nc


class Dat(object):
    def __init__(self, a):
        self.data = np.asarray(a, dtype=float)


class Fn(object):
    """NumPy-backed stand-in for a Firedrake Function: .dat.data and CG1 point evaluation .at()."""

    def __init__(self, a, Ld=None):
        self.dat = Dat(a)
        self.Ld = Ld

    def at(self, xs, tolerance=None):
        scalar = np.isscalar(xs)
        xs = np.atleast_1d(np.asarray(xs, dtype=float))
        n = len(self.dat.data)
        h = self.Ld / n
        out = np.empty(len(xs))
        for k, x in enumerate(xs):
            s = x / h
            i = int(np.floor(s)) % n
            th = s - np.floor(s)
            out[k] = (1 - th) * self.dat.data[i] + th * self.dat.data[(i + 1) % n]
        return out[0] if (scalar or len(out) == 1) else out


def main():
    install_stubs()
    sys.path.insert(0, REF)
    from peakondrake.outputting import find_min, find_max, find_mus, peakon_diagnostics
    from peakondrake.peakon_equations import PeakonEquations

    rng = np.random.RandomState(20261019)
    Ld = 40.0
    cases = {}
    names = []

    def add_case(name, u):
        ne = len(u)
        h = Ld / ne
        x = (np.arange(ne) + 0.5) * h
        du = (np.roll(u, -1) - u) / h
        f, df, xc = Fn(u, Ld), Fn(du), Fn(x)
        fmin, xmin = find_min(df, xc)
        fmax, xmax = find_max(df, xc)
        mus, alt = find_mus(f, df, xc)
        pk = peakon_diagnostics(f, df, xc)
        cases[name + '/u'] = u
        cases[name + '/minmax'] = np.array([fmin, xmin, fmax, xmax])
        cases[name + '/mus'] = np.array(list(mus) + [np.nan] * (4 - len(mus)))[:4] if len(mus) <= 4 else np.array(mus[:4])
        cases[name + '/nmus'] = np.array([len(mus), len(alt)])
        cases[name + '/alt'] = np.array(list(alt) + [np.nan] * (4 - len(alt)))[:4] if len(alt) <= 4 else np.array(alt[:4])
        keys = ['peakon_loc', 'peakon_min_du', 'peakon_max_du', 'peakon_min_du_loc', 'peakon_max_du_loc',
                'peakon_max_u', 'peakon_mu', 'peakon_nu']
        cases[name + '/peakon'] = np.array([float(pk[k]) for k in keys])
        names.append(name)

    ne = 200
    xn = np.arange(ne) * Ld / ne
    sech = lambda c, w=1.0: 0.5 * 2 / (np.exp((xn - c) / w) + np.exp(-(xn - c) / w))
    add_case('single_peak', sech(10.0))
    add_case('steep_peak', sech(10.0, 1 / 6))
    add_case('two_peaks', 0.2 * sech(26.8) + 0.5 * sech(13.5))
    add_case('three_peaks', 0.5 * sech(5.0, .5) + 0.4 * sech(18.0, .5) + 0.3 * sech(31.0, .5))
    add_case('peak_at_wrap', np.roll(sech(10.0), -50))          # maximum at node 0
    add_case('peak_near_end', np.roll(sech(10.0), 148))         # maximum near the last node
    pk = np.exp(-np.abs(xn - 20.0))
    add_case('peakon', pk)
    add_case('noisy_peak', sech(12.0) + 1e-3 * rng.randn(ne))
    add_case('five_peaks', sum(0.1 * (k + 1) * sech(4.0 + 8 * k, .4) for k in range(5)))
    cases['names'] = np.array(names)

    # PeakonEquations.update: drive it with NumPy-backed xi fields
    class PV(object):
        pass
    for method in ('ito_euler', 'milstein'):
        ne = 400
        h = Ld / ne
        x = np.arange(ne) * h
        pv = PV()
        nxi = 2
        k1, k2 = 2 * np.pi / Ld, 4 * np.pi / Ld
        xi = [0.2 * np.sin(k1 * x), 0.2 * np.cos(k2 * x)]
        xi_x = [0.2 * k1 * np.cos(k1 * x), -0.2 * k2 * np.sin(k2 * x)]
        xi_xx = [-0.2 * k1 ** 2 * np.sin(k1 * x), -0.2 * k2 ** 2 * np.cos(k2 * x)]
        pv.dW_nums = np.zeros(nxi)
        pv.pure_xi_list = [Fn(a, Ld) for a in xi]
        pv.pure_xi_x_list = [Fn(a, Ld) for a in xi_x]
        pv.pure_xi_xx_list = [Fn(a, Ld) for a in xi_xx]
        pe = PeakonEquations.__new__(PeakonEquations)
        pe.dWs = pv.dW_nums
        pe.pure_xi_list, pe.pure_xi_x_list, pe.pure_xi_xx_list = pv.pure_xi_list, pv.pure_xi_x_list, pv.pure_xi_xx_list
        pe.dt, pe.Ld, pe.alpha, pe.method = 1e-2, Ld, 1.0, method
        pe.p, pe.q = 1.0, 39.5
        rs = np.random.RandomState(5)
        nst = 200
        dW = rs.randn(nst, nxi) * np.sqrt(pe.dt)
        traj = np.zeros((nst, 2))
        for n in range(nst):
            pv.dW_nums[:] = dW[n]
            pe.update()
            traj[n] = (pe.p, pe.q)
        cases['sde_%s/dW' % method] = dW
        cases['sde_%s/traj' % method] = traj
        cases['sde_%s/xi' % method] = np.array(xi)
        cases['sde_%s/xi_x' % method] = np.array(xi_x)
        cases['sde_%s/xi_xx' % method] = np.array(xi_xx)

    out = __file__.replace('make_golden.py', 'reference_helpers.npz')
    np.savez_compressed(out, **cases)
    print('wrote', out, 'with', len(names), 'field cases')


if __name__ == '__main__':
    main()
