"""
Host-side initial conditions for the CG1 schemes: the reference's `build_initial_conditions`
(initial_conditions.py:8-124).  Set-up only (runs once per resolution); the result is uploaded to
the device and broadcast over the ensemble.

Path (initial_conditions.py:113-118): every ic except 'peakon' is interpolated into CG5 and the CG5
function is L2-projected to CG1; 'peakon' projects the (kinked) expression directly.
"""
import numpy as np
from peakondrake_b200.stochastic_functions import circulant_tridiag_solve

IC_NAMES = ('two_peaks', 'gaussian', 'gaussian_narrow', 'gaussian_wide', 'peakon', 'one_peak',
            'proper_peak', 'new_peak', 'flat', 'fast_flat', 'coshes', 'd_peakon', 'zero',
            'two_peakons', 'twin_peakons', 'periodic_peakon', 'cos_bell', 'antisymmetric')


def ic_function(ic, Ld, alphasq, peak_width, deltax):
    """The 18 expressions of initial_conditions.py:29-55 as NumPy callables."""
    e = np.exp
    sa = np.sqrt(alphasq)
    w = peak_width
    eps = 1
    sech2 = lambda x, c: 2 / (e(x - c) + e(-x + c))
    table = {
        'two_peaks': lambda x: 0.2 * sech2(x, 403. / 15. * 40. / Ld) + 0.5 * sech2(x, 203. / 15. * 40. / Ld),
        'gaussian': lambda x: 0.5 * e(-((x - 10.) / 2.) ** 2),
        'gaussian_narrow': lambda x: 0.5 * e(-((x - 10.) / 1.) ** 2),
        'gaussian_wide': lambda x: 0.5 * e(-((x - 10.) / 3.) ** 2),
        'peakon': lambda x: np.where(x < Ld / 2., e((x - Ld / 2) / sa), e(-(x - Ld / 2) / sa)),
        'one_peak': lambda x: 0.5 * sech2(x, 203. / 15. * 40. / Ld),
        'proper_peak': lambda x: 0.5 * sech2(x, Ld / 4),
        'new_peak': lambda x: 0.5 * 2 / (e((x - Ld / 4) / w) + e((-x + Ld / 4) / w)),
        'flat': lambda x: np.full_like(x, 2 * np.pi ** 2 / (9 * 40 ** 2)),
        'fast_flat': lambda x: np.full_like(x, 0.1),
        'coshes': lambda x: (2000 * np.cosh((2000 ** 0.5 / 2) * (x - 0.75)) ** (-2)
                             + 1000 * np.cosh(1000 ** 0.5 / 2 * (x - 0.25)) ** (-2)),
        'd_peakon': lambda x: e(-np.sqrt((x - Ld / 2) ** 2 + eps * deltax ** 2) / sa),
        'zero': lambda x: np.zeros_like(x),
        'two_peakons': lambda x: np.where(
            x < Ld / 4, e((x - Ld / 4) / sa) - e(-(x + Ld / 4) / sa),
            np.where(x < 3 * Ld / 4, e(-(x - Ld / 4) / sa) - e((x - 3 * Ld / 4) / sa),
                     e((x - 5 * Ld / 4) / sa) - e(-(x - 3 * Ld / 4) / sa))),
        'twin_peakons': lambda x: np.where(
            x < Ld / 4, e((x - Ld / 4) / sa) + 0.5 * e((x - Ld / 2) / sa),
            np.where(x < Ld / 2, e(-(x - Ld / 4) / sa) + 0.5 * e((x - Ld / 2) / sa),
                     np.where(x < 3 * Ld / 4, e(-(x - Ld / 4) / sa) + 0.5 * e(-(x - Ld / 2) / sa),
                              e((x - 5 * Ld / 4) / sa) + 0.5 * e(-(x - Ld / 2) / sa)))),
        'periodic_peakon': lambda x: np.where(
            x < Ld / 2,
            0.5 / (1 - e(-Ld / sa)) * (e((x - Ld / 2) / sa) + e(-Ld / sa) * e(-(x - Ld / 2) / sa)),
            0.5 / (1 - e(-Ld / sa)) * (e(-(x - Ld / 2) / sa) + e(-Ld / sa) * e((x - Ld / 2) / sa))),
        'cos_bell': lambda x: np.where(x < Ld / 4, (np.cos(np.pi * (x - Ld / 8) / (2 * Ld / 8))) ** 2, 0.0),
        'antisymmetric': lambda x: (1 / (e((x - Ld / 4) / Ld) + e((-x + Ld / 4) / Ld))
                                    - 1 / (e((Ld - x - Ld / 4) / Ld) + e((Ld + x + Ld / 4) / Ld))),
    }
    if ic not in table:
        raise KeyError(ic)       # the reference indexes ic_dict[ic] (initial_conditions.py:57)
    return table[ic]


def _gauss01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def _cg5_to_cg1_weights():
    """W[a, k] = int_0^1 phi_a(s) L_k(s) ds for the two CG1 hat halves and the 6 equispaced CG5 nodes."""
    nodes = np.linspace(0.0, 1.0, 6)
    q, w = _gauss01(4)
    L = np.ones((len(q), 6))
    for a in range(6):
        for b in range(6):
            if a != b:
                L[:, a] *= (q - nodes[b]) / (nodes[a] - nodes[b])
    W = np.zeros((2, 6))
    W[0] = (w * (1 - q)) @ L
    W[1] = (w * q) @ L
    return W


def build_initial_u(ne, Ld, ic, alphasq=1.0, peak_width=1.0):
    """CG1 coefficient vector of the initial u in natural node order (node i at i*Ld/ne)."""
    ne = int(ne)
    h = Ld / ne
    f = ic_function(ic, Ld, alphasq, peak_width, h)
    x = np.arange(ne) * h
    if ic == 'peakon':
        q, w = _gauss01(3)
        bl = np.zeros(ne)
        br = np.zeros(ne)
        for qq, ww in zip(q, w):
            fq = np.asarray(f(x + qq * h), dtype=float) * ww * h
            bl += (1 - qq) * fq
            br += qq * fq
    else:
        s = np.linspace(0.0, 1.0, 6)
        vals = np.asarray(f(x[:, None] + s[None, :] * h), dtype=float) * np.ones((ne, 6))
        # the dof shared by cells i-1 and i is written last by ... the convention documented in
        # the cell to the left (so node 0 carries the value at x = Ld).
        vals[:, 0] = np.roll(vals[:, 5], 1)
        W = _cg5_to_cg1_weights()
        bl = h * (vals @ W[0])
        br = h * (vals @ W[1])
    b = bl + np.roll(br, 1)
    return circulant_tridiag_solve(h / 6.0, 2.0 * h / 3.0, b)
