"""
The CUDA kernel's algorithm, executed on the CPU by the emulation harness (tests/emu/pk_emu.cpp runs
the identical `pk_core.h` phases with the threads of a CTA serialised), against the oracle.
This is what keeps the kernel logic under test in the GPU-less container; the GPU tests
(test_gpu_parity.py) repeat the comparison on the real device.
"""
import numpy as np
import pytest

from oracle import ch_oracle as O
from emu_util import emu_step, lib

TOL = 1e-10


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def compare(ne, nsteps=5, ic='proper_peak', fam='sines', nxi=4, sigma=0.2, scheme='hydrodynamic', dt=1e-3,
            Ld=40.0, T=0, M=0, seed=3, order=0):
    mesh = O.Mesh1D(ne, Ld)
    u0 = O.initial_u(mesh, ic)
    tabs = O.noise_tables(mesh, fam, nxi, sigma, scheme=scheme) if nxi else None
    dW = O.reference_dW(seed, nsteps, nxi, dt) if nxi else None
    ref = O.run(ne, Ld, dt, nsteps, u0, dW, tabs, scheme=scheme)
    u = u0[None].copy(); u1 = np.zeros_like(u); dF = np.zeros_like(u); dG = np.zeros_like(u)
    if nxi == 0 or fam == 'constant':
        mode, xc = 0, (tabs[0][0, 1] if nxi else 0.0)
    elif scheme == 'no_gradient':
        mode, xc = 1, 0.0
    else:
        mode, xc = 2, 0.0
    tabs_c = [np.ascontiguousarray(t) for t in tabs] if tabs is not None else None
    r = emu_step(ne, Ld, dt, mode, u, u1, dF, dG, tabs_c, None if dW is None else np.ascontiguousarray(dW[:, None, :]),
                 nsteps, xi_const=xc, T=T, M=M, order=order, nxi=nxi)
    assert r['fail_step'][0] == -1
    assert r['its_total'][0] == ref['its'].sum()
    assert rel(u[0], ref['u']) <= TOL and rel(u1[0], ref['w1'][:ne]) <= TOL
    assert rel(dF[0], ref['w1'][ne:2 * ne]) <= 1e-9
    return u[0], u1[0], dF[0], dG[0]


@pytest.mark.parametrize('ne', [8, 9, 16, 20, 37, 64, 200, 1000, 1001, 1500])
def test_mesh_sizes(ne):
    compare(ne, nsteps=4, nxi=2)


@pytest.mark.parametrize('T,M,ne', [(4, 5, 20), (4, 6, 22), (8, 7, 50), (16, 8, 128), (32, 8, 250), (2, 5, 9)])
def test_forced_layouts(T, M, ne):
    compare(ne, nsteps=3, nxi=3, T=T, M=M)


def test_constant_noise_and_no_noise_and_no_gradient():
    compare(400, fam='constant', nxi=1, ic='periodic_peakon', sigma=1.0)
    compare(300, nxi=0)
    compare(300, scheme='no_gradient', nxi=3)


def test_thread_order_independence():
    """Running the threads of a phase in ascending or descending order must give bit-identical
    results: there is no intra-phase communication (race-freedom of the barrier structure)."""
    a = compare(250, nsteps=4, order=0)
    b = compare(250, nsteps=4, order=1)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_convergence_certificate_changes_nothing():
    """The Newton convergence certificate (run_steps: |R(w_new)| <= cert_k |delta|^2 for the quadratic residual)
    only replaces the last residual EVALUATION: iterates, iteration counts and results are bit-identical with it
    switched off, on meshes where it fires (steady steps) and where it cannot (first step, coarse bound)."""
    L = lib()
    cases = [dict(ne=1000, nsteps=6, nxi=4), dict(ne=200, nsteps=5, nxi=2), dict(ne=400, nsteps=4, fam='constant', nxi=1, ic='periodic_peakon', sigma=1.0),
             dict(ne=300, nsteps=4, nxi=0), dict(ne=2000, nsteps=3, ic='periodic_peakon', fam='constant', nxi=1, sigma=1.0, dt=1e-2)]
    for kw in cases:
        try:
            L.pk_emu_set_certificate(1); a = compare(**kw)
            L.pk_emu_set_certificate(0); b = compare(**kw)
        finally:
            L.pk_emu_set_certificate(1)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_fine_mesh_large_noise():
    # C5-like spacing h = 40/2000 with sigma = 1: P is strongly non diagonally dominant here
    compare(2000, nsteps=3, ic='periodic_peakon', fam='constant', nxi=1, sigma=1.0, dt=1e-2)


def test_failure_latch_and_reason():
    ne, Ld, dt = 64, 40.0, 1e-3
    u0 = O.initial_u(O.Mesh1D(ne, Ld), 'proper_peak')
    tabs = [np.ascontiguousarray(t) for t in O.noise_tables(O.Mesh1D(ne, Ld), 'sines', 1, 1.0)]
    dW = np.zeros((4, 1, 1)); dW[2] = np.nan
    u = u0[None].copy(); u1 = np.zeros_like(u); dF = np.zeros_like(u); dG = np.zeros_like(u)
    r = emu_step(ne, Ld, dt, 2, u, u1, dF, dG, tabs, dW, 4)
    assert r['fail_step'][0] == 2 and r['reason'][0] == 2          # DIVERGED_FNORM_NAN at step 2
    ref = O.run(ne, Ld, dt, 2, u0, dW[:2, 0, :], tabs)
    assert rel(u[0], ref['u']) <= TOL                               # frozen at the last accepted state


def test_philox_generator_statistics_and_reproducibility():
    L = lib()
    z = np.array([L.pk_emu_philox_normal(42, r, s, 0, 0) for r in range(200) for s in range(50)])
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1) < 0.03
    assert abs((z ** 3).mean()) < 0.1 and abs((z ** 4).mean() - 3) < 0.2
    assert L.pk_emu_philox_normal(42, 5, 7, 1, 0) == L.pk_emu_philox_normal(42, 5, 7, 1, 0)
    assert L.pk_emu_philox_normal(42, 5, 7, 1, 0) != L.pk_emu_philox_normal(43, 5, 7, 1, 0)
    # exported in-kernel increments reproduce the path through the oracle; sub-updates sum
    ne, Ld, dt, n = 64, 40.0, 1e-3, 5
    u0 = O.initial_u(O.Mesh1D(ne, Ld), 'proper_peak')
    tabs = [np.ascontiguousarray(t) for t in O.noise_tables(O.Mesh1D(ne, Ld), 'sines', 2, 0.3)]
    u = np.tile(u0, (2, 1)); u1 = np.zeros_like(u); dF = np.zeros_like(u); dG = np.zeros_like(u)
    r = emu_step(ne, Ld, dt, 2, u, u1, dF, dG, tabs, None, n, nxi=2, seed=9, rid0=100, export=True, nupd=4)
    for k in range(2):
        ref = O.run(ne, Ld, dt, n, u0, r['dW'][:, k, :], tabs)
        assert rel(u[k], ref['u']) <= TOL
    w = sum(L.pk_emu_philox_normal(9, 101, 3 * 4 + j, 1, 0) for j in range(4)) * np.sqrt(dt / 4)       # fine step 3*4+j
    assert abs(r['dW'][3, 1, 1] - w) < 1e-16          # kernel scales each sub-draw, then sums
    # a coarse run with nXi_updates = 4 consumes the Brownian path of the fine run (stochastic_functions.py:211-217):
    # its increments are the sums of four consecutive fine increments, bit for bit (dt a power of two: exact scalings)
    dtf = 2.0 ** -10
    mk = lambda: (np.tile(u0, (2, 1)), np.zeros((2, ne)), np.zeros((2, ne)), np.zeros((2, ne)))
    a = mk(); fine = emu_step(ne, Ld, dtf, 2, a[0], a[1], a[2], a[3], tabs, None, 12, nxi=2, seed=9, rid0=100, export=True, nupd=1)['dW']
    b = mk(); coarse = emu_step(ne, Ld, 4 * dtf, 2, b[0], b[1], b[2], b[3], tabs, None, 3, nxi=2, seed=9, rid0=100, export=True, nupd=4)['dW']
    acc = ((fine[0::4] + fine[1::4]) + fine[2::4]) + fine[3::4]
    assert np.array_equal(coarse, acc)


@pytest.mark.parametrize('ne,C,part', [(1000, 2, 1), (1000, 4, 1), (250, 2, 1), (2000, 8, 1), (2000, 16, 1), (128, 8, 1), (256, 8, 1), (200, 2, 1),
                                        (1000, 4, 0), (2000, 8, 0), (1000, 4, 2), (2000, 8, 2), (2000, 16, 2)])
def test_cluster_layout_emulated(ne, C, part):
    """Thread-block-cluster variant (large meshes): shared arrays distributed over C CTAs, foreign
    rows read remotely, every store local -- same decode as SharedCluster in pk_lib.cu.  part=1: Schur solve
    partitioned by CTA (segment-local PCR, the C separators solved redundantly in every CTA: the device default);
    part=2: separators solved by the leading threads; part=0: flat PCR."""
    L = lib()
    try:
        L.pk_emu_set_partitioned(part)
        assert L.pk_emu_set_cluster(C) == 0
        compare(ne, nsteps=3, nxi=2)
        compare(ne, nsteps=2, fam='constant', nxi=1, ic='periodic_peakon', sigma=0.5)
        assert L.pk_emu_set_cluster(1) == 0          # no non-local stores happened
    finally:
        L.pk_emu_set_cluster(1)
        L.pk_emu_set_partitioned(1)


@pytest.mark.parametrize('ne,C', [(1000, 4), (2000, 8), (512, 4), (200, 2)])
def test_cluster_single_row_buffer_emulated(ne, C):
    """The two-CTAs-per-SM cluster kernels keep ONE PCR row buffer (18 exchange slots + separator scratch instead of 30
    slots): a block barrier between the reads and the publish of every segment-local PCR step, the cluster barrier after
    reduce_build.  Same arithmetic as the two-buffer layout: against the oracle, and bit for bit against that layout.
    (Needs at least as many threads per CTA as CTAs per cluster: every CTA solves the separator system itself.)"""
    L = lib()
    try:
        assert L.pk_emu_set_cluster(C) == 0
        L.pk_emu_set_pingpong(0)
        compare(ne, nsteps=3, nxi=2)
        compare(ne, nsteps=2, fam='constant', nxi=1, ic='periodic_peakon', sigma=0.5)
        assert L.pk_emu_set_cluster(C) == 0          # no non-local stores happened
        outs = []
        for pp in (0, 1):
            L.pk_emu_set_pingpong(pp)
            outs.append(compare(ne, nsteps=3, fam='constant', nxi=1, sigma=0.3))
        assert all(np.array_equal(a, b) for a, b in zip(outs[0], outs[1]))
    finally:
        L.pk_emu_set_pingpong(1)
        L.pk_emu_set_cluster(1)


@pytest.mark.parametrize('ne,T,M,C', [(2500, 256, 10, 8), (2200, 256, 9, 4), (2350, 256, 10, 1), (1250, 128, 10, 2)])
def test_wide_chunk_layouts_emulated(ne, T, M, C):
    """9 or 10 rows per thread (choose_layout_wide): what the cluster kernels use where the standard layout would end
    at 5 rows per thread (BASELINE config 5, ne = 20 000 -> 2048 threads x 10 rows, clusters of 8)."""
    L = lib()
    try:
        assert L.pk_emu_set_cluster(C) == 0
        compare(ne, nsteps=3, nxi=2, T=T, M=M)
        compare(ne, nsteps=2, fam='constant', nxi=1, ic='periodic_peakon', sigma=0.5, T=T, M=M)
        assert L.pk_emu_set_cluster(1) == 0
    finally:
        L.pk_emu_set_cluster(1)


def test_wide_layout_choice():
    import ctypes
    L = lib()
    T, M, rem = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert L.pk_emu_layout_wide(20000, ctypes.byref(T), ctypes.byref(M), ctypes.byref(rem)) == 0
    assert (T.value, M.value, rem.value) == (2048, 10, 1568)
    assert L.pk_emu_layout_wide(17000, ctypes.byref(T), ctypes.byref(M), ctypes.byref(rem)) == 0
    assert (T.value, M.value, rem.value) == (2048, 9, 616)
    assert L.pk_emu_layout_wide(18432, ctypes.byref(T), ctypes.byref(M), ctypes.byref(rem)) == 0
    assert (T.value, M.value, rem.value) == (2048, 9, 2048)
    for ne in (16384, 24000, 1000, 12000):       # standard layout has 6..8 rows per thread: no wide variant
        assert L.pk_emu_layout_wide(ne, ctypes.byref(T), ctypes.byref(M), ctypes.byref(rem)) != 0
    for ne in (20000, 17000, 18432, 20480):
        assert L.pk_emu_layout_wide(ne, ctypes.byref(T), ctypes.byref(M), ctypes.byref(rem)) == 0
        assert rem.value * M.value + (T.value - rem.value) * (M.value - 1) == ne and 0 < rem.value <= T.value


# ---- upwind / ssprk3 ---------------------------------------------------------------------------------
@pytest.mark.parametrize('ne,nxi,fam', [(20, 7, 'sines'), (1000, 7, 'sines'), (200, 1, 'constant'), (333, 0, 'sines'), (9, 2, 'sines')])
def test_upwind_kernel_emulated(ne, nxi, fam):
    from emu_util import emu_step_upwind
    Ld, dt, n = 40., 1e-3, 10
    mesh = O.Mesh1D(ne, Ld); u0 = O.initial_u(mesh, 'one_peak')
    xi = O.noise_tables(mesh, fam, nxi, 0.25, scheme='upwind')[0] if nxi else None
    dW = O.reference_dW(0, n, nxi, dt) if nxi else None
    ref = O.run_upwind(ne, Ld, dt, n, u0, dW, xi)
    res = []
    for order in (0, 1):
        m = ref['oracle'].initial_m(u0)[None].copy(); u = u0[None].copy()
        xc = xi[0, 1] if fam == 'constant' else 0.0
        emu_step_upwind(ne, Ld, dt, m, u, None if (fam == 'constant' or not nxi) else np.ascontiguousarray(xi),
                        None if dW is None else np.ascontiguousarray(dW[:, None, :]), n, xi_const=xc, nxi=nxi, order=order)
        assert rel(u[0], ref['u']) <= TOL and rel(m[0], ref['m']) <= TOL
        res.append((u.copy(), m.copy()))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])


def test_upwind_oracle_properties():
    # constants are preserved; int m drifts only weakly; energy drift ~5e-4 over t=2 at ne=1000
    ne, Ld, dt = 200, 40.0, 1e-3
    orc = O.UpwindOracle(ne, Ld, dt)
    m, u = orc.solve(np.full((ne, 2), 0.7), np.full(ne, 0.7), np.zeros(ne))
    assert abs(m - 0.7).max() < 1e-14 and abs(u - 0.7).max() < 1e-12
    u0 = O.initial_u(orc.mesh, 'one_peak')
    r = O.run_upwind(ne, Ld, dt, 200, u0)
    E0 = u0 @ (orc.M @ u0) + u0 @ (orc.K @ u0); E1 = r['u'] @ (orc.M @ r['u']) + r['u'] @ (orc.K @ r['u'])
    assert abs(E1 - E0) / E0 < 5e-3
    assert abs(O.dg1_norms(orc, r['m'])['mass_m'] - O.dg1_norms(orc, orc.initial_m(u0))['mass_m']) < 1e-3


def test_step_blocks_are_bit_identical_to_one_pass():
    """The persistent step kernel hands a launch out in work items of `nblk` steps that different CTAs
    pick up with fresh per-thread state (pk_step_kernel): the Philox window is re-drawn at a block start,
    the failure / iteration counters continue."""
    ne, Ld, dt, nsteps, nxi = 200, 40.0, 1e-3, 11, 4
    mesh = O.Mesh1D(ne, Ld)
    u0 = O.initial_u(mesh, 'proper_peak')
    tabs = [np.ascontiguousarray(t) for t in O.noise_tables(mesh, 'sines', nxi, 0.2)]
    out = []
    for nblk in (0, 1, 3, 4, 11, 20):
        lib().pk_emu_set_block(nblk)
        try:
            u = np.tile(u0, (2, 1)); u1 = np.zeros_like(u); dF = np.zeros_like(u); dG = np.zeros_like(u)
            r = emu_step(ne, Ld, dt, 2, u, u1, dF, dG, tabs, None, nsteps, seed=7, rid0=5, export=True, nxi=nxi)
        finally:
            lib().pk_emu_set_block(0)
        out.append((u, u1, dF, dG, r['dW'], r['its_total'], r['its_last'], r['fail_step']))
    for o in out[1:]:
        for a, b in zip(out[0], o):
            assert np.array_equal(a, b)
    assert out[0][5].min() >= 2 * nsteps and np.abs(out[0][4]).min() > 0


@pytest.mark.parametrize('ne,fam,nxi,sigma', [(200, 'sines', 3, 0.2), (1000, 'sines', 4, 0.2), (300, 'constant', 1, 1.0),
                                             (64, 'sines', 0, 0.0), (37, 'gaussians', 2, 0.3)])
def test_test_scheme(ne, fam, nxi, sigma):
    """Scheme 'test' (equations.py:132-145): one CG1 unknown, M(u1-u) - int psi uh (dv)_x = 0, served by the
    same kernel code with modes TEST_SCALAR / TEST_X (tridiagonal Newton matrix through the pentadiagonal solver)."""
    Ld, dt, nsteps = 40.0, 1e-3, 6
    mesh = O.Mesh1D(ne, Ld)
    u0 = O.initial_u(mesh, 'proper_peak')
    tabs = O.noise_tables(mesh, fam, nxi, sigma, scheme='test') if nxi else None
    dW = O.reference_dW(3, nsteps, nxi, dt) if nxi else None
    ref = O.run(ne, Ld, dt, nsteps, u0, dW, tabs, scheme='test')
    assert rel(ref['u'], u0) > 1e-6
    u = u0[None].copy(); u1 = np.zeros_like(u); dF = np.zeros_like(u); dG = np.zeros_like(u)
    if nxi == 0 or fam == 'constant':
        mode, xc = 3, (tabs[0][0, 1] if nxi else 0.0)
    else:
        mode, xc = 4, 0.0
    tc = [np.ascontiguousarray(t) for t in tabs] if tabs is not None else None
    r = emu_step(ne, Ld, dt, mode, u, u1, dF, dG, tc, None if dW is None else np.ascontiguousarray(dW[:, None, :]),
                 nsteps, xi_const=xc, nxi=nxi)
    assert r['fail_step'][0] == -1 and r['its_total'][0] == ref['its'].sum()
    assert rel(u[0], ref['u']) <= TOL and rel(u1[0], ref['w1'][:ne]) <= TOL
    assert not dF.any() and not dG.any()


@pytest.mark.parametrize('var', [1, 2, 4])
def test_code_variants_are_bit_identical(var):
    """Thread<..., VAR> (pk_core.h): factors in shared memory with rolled loops (1), rolled assembly with register
    factors behind a switch (2), everything unrolled (4) -- same arithmetic in the same order as the default."""
    out = []
    for v in (0, var):
        lib().pk_emu_set_fs(v)
        try:
            res = [compare(ne, nsteps=4, nxi=3, seed=5) for ne in (20, 37, 200, 1000)]
            res.append(compare(300, scheme='no_gradient', nxi=2))
            res.append(compare(400, fam='constant', nxi=1, ic='periodic_peakon', sigma=1.0))
        finally:
            lib().pk_emu_set_fs(0)
        out.append(res)
    for a, b in zip(out[0], out[1]):
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_hybrid_schur_solve_variant():
    """Cyclic reduction compacted onto the leading threads + PCR inside one warp + back-substitution (hyb_* in
    pk_core.h; an experiment build option, slower than plain PCR on B200): same answers to rounding."""
    lib().pk_emu_set_hybrid(1)
    try:
        for ne in (200, 256, 500, 1000, 1500, 4000):
            a = compare(ne, nsteps=3, nxi=2, order=0)
            b = compare(ne, nsteps=3, nxi=2, order=1)
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
    finally:
        lib().pk_emu_set_hybrid(0)


def _emu_mu(u, du, x, Ld):
    import ctypes
    L = lib()
    L.pk_emu_mu.restype = ctypes.c_int
    old4, alt4 = np.zeros(4), np.zeros(4)
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    u, du, x = [np.ascontiguousarray(a, dtype=float) for a in (u, du, x)]
    rc = L.pk_emu_mu(dp(u), dp(du), dp(x), len(u), ctypes.c_double(Ld), dp(old4), dp(alt4))
    return rc, old4, alt4


def test_device_mu_routine_against_reference_golden_vectors():
    """pk_mu.h (the routine pk_mu_kernel runs per realisation) on the arrays behind tests/golden: bit-equal to the
    reference's own find_mus outputs (outputting.py:260-370), including the multi-peak zone logic; NumPy's pairwise
    summation is reproduced exactly."""
    import ctypes, os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_helpers.npz'))
    L = lib(); L.pk_emu_pairwise_sum.restype = ctypes.c_double
    rng = np.random.default_rng(1)
    for n in list(range(1, 200, 7)) + [1000, 2000, 16384, 20000]:
        a = rng.standard_normal(n) * 10 ** rng.uniform(-2, 2)
        assert L.pk_emu_pairwise_sum(a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n) == np.sum(a)
    for name in [str(n) for n in gold['names']]:
        u = gold[name + '/u']; ne = len(u); Ld = 40.0; h = Ld / ne
        x = (np.arange(ne) + 0.5) * h
        du = (np.roll(u, -1) - u) / h
        rc, old4, alt4 = _emu_mu(u, du, x, Ld)
        n_old = int(gold[name + '/nmus'][0])
        assert rc >= 0
        want_old = np.full(4, np.nan); want_alt = np.full(4, np.nan)
        k = min(4, n_old)
        want_old[:k] = gold[name + '/mus'][:k]; want_alt[:k] = gold[name + '/alt'][:k]
        assert np.array_equal(old4, want_old, equal_nan=True) and np.array_equal(alt4, want_alt, equal_nan=True), name


def test_device_mu_routine_equals_host_restatement_on_random_fields():
    from peakondrake_b200 import outputting as out
    rng = np.random.default_rng(7)
    Ld = 40.0
    for trial in range(60):
        ne = int(rng.choice([64, 100, 257, 1000]))
        h = Ld / ne
        xn = np.arange(ne) * h
        npk = int(rng.integers(0, 7))
        u = 0.05 * rng.standard_normal() + np.zeros(ne)
        for _ in range(npk):
            c, a, w = rng.uniform(0, Ld), rng.uniform(0.3, 2.0), rng.uniform(0.5, 3.0)
            d = np.abs(xn - c); d = np.minimum(d, Ld - d)
            u = u + a * np.exp(-d / w)
        u += 1e-3 * rng.standard_normal(ne) * (trial % 3 == 0)
        x = (np.arange(ne) + 0.5) * h
        du = (np.roll(u, -1) - u) / h
        try:
            old, alt = out.mu_diagnostic(u, du, x, Ld)
        except (IndexError, ValueError):
            assert _emu_mu(u, du, x, Ld)[0] == -2
            continue
        rc, old4, alt4 = _emu_mu(u, du, x, Ld)
        assert rc >= 0 or rc == -1
        if rc < 0:
            continue
        k = min(4, len(old))
        assert np.array_equal(old4[:k], np.array(old[:k])) and np.array_equal(alt4[:k], np.array(alt[:k])), trial
        assert np.isnan(old4[k:]).all() and np.isnan(alt4[k:]).all()


def _numpy_dump(u, h, alphasq=1.0):
    """pk_diag_kernel's per-node expressions in NumPy (columns PK_D_ENERGY..PK_D_MAX_U and the four indices)."""
    up = np.roll(u, -1); um = np.roll(u, 1)
    du = (up - u) / h; dum = (u - um) / h; jmp = np.abs(du - dum)
    sM = np.sum(h * (1.0 / 3.0) * (u * u + u * up + up * up)); sK = np.sum((up - u) * (up - u) * (1.0 / h)); sU = np.sum(h * u)
    vals = [sM + alphasq * sK, sM, sK, sU, du.max(), du.min(), jmp.max(), u.max()]
    return np.array(vals), np.array([np.argmax(du), np.argmin(du), np.argmax(u), np.argmax(jmp)])


@pytest.mark.parametrize('ne,C,T,M', [(250, 1, 0, 0), (1000, 1, 0, 0), (37, 1, 0, 0), (1000, 4, 0, 0), (2000, 8, 0, 0), (2500, 8, 256, 10), (200, 2, 0, 0)])
def test_fused_dump_diagnostics_emulated(ne, C, T, M):
    """pk_step_dump: the dump diagnostics evaluated INSIDE the step kernel on the shared-memory state (diag_dump in
    pk_core.h) every `every` steps equal the separate kernels' expressions on the state of that step: extrema, indices
    and the peakon-suite scans bit for bit, sums to rounding; both thread orders agree bit-wise; cluster layouts too."""
    import ctypes
    from emu_util import _p
    L = lib()
    Ld, dt, nsteps, every, nxi, R = 40.0, 1e-3, 6, 2, 2, 2
    mesh = O.Mesh1D(ne, Ld); h = mesh.h
    u0 = O.initial_u(mesh, 'proper_peak')
    tabs = [np.ascontiguousarray(t) for t in O.noise_tables(mesh, 'sines', nxi, 0.3)]
    dW = np.ascontiguousarray(np.stack([O.reference_dW(s, nsteps, nxi, dt) for s in range(R)], axis=1))
    nd = nsteps // every
    outs = []
    try:
        assert L.pk_emu_set_cluster(C) == 0
        for order in (0, 1):
            dg = np.zeros((nd, R, 14)); di = np.zeros((nd, R, 4), np.int32); sv = np.zeros((nd, R, 2)); si = np.zeros((nd, R, 3), np.int32)
            L.pk_emu_set_dump(every, _p(dg), _p(di, ctypes.c_int), _p(sv), _p(si, ctypes.c_int))
            u = np.tile(u0, (R, 1)); z = [np.zeros_like(u) for _ in range(3)]
            emu_step(ne, Ld, dt, 2, u, z[0], z[1], z[2], tabs, dW, nsteps, order=order, T=T, M=M)
            outs.append((dg, di, sv, si, u))
        L.pk_emu_set_dump(0, None, None, None, None)
        for a, b in zip(*outs):
            assert np.array_equal(a, b, equal_nan=True)
        dg, di, sv, si, _ = outs[0]
        # the state at the dump steps, from plain launches
        u = np.tile(u0, (R, 1)); z = [np.zeros_like(u) for _ in range(3)]
        from test_host_logic import _device_scan_restated
        for d in range(nd):
            for n in range(every):      # one step per call keeps the Brownian path aligned with the single launch above
                emu_step(ne, Ld, dt, 2, u, z[0], z[1], z[2], tabs, np.ascontiguousarray(dW[d * every + n:d * every + n + 1]), 1,
                         T=T, M=M, step0=d * every + n)
            for r in range(R):
                v, ix = _numpy_dump(u[r], h)
                assert np.array_equal(di[d, r], ix)
                assert np.array_equal(dg[d, r, 4:8], v[4:8])                        # extrema: bit-identical
                assert np.allclose(dg[d, r, :4], v[:4], rtol=1e-13, atol=0) and np.isnan(dg[d, r, 8:]).all()
                pv, pi = _device_scan_restated(u[r], ne, h)
                assert np.array_equal(sv[d, r], pv, equal_nan=True) and np.array_equal(si[d, r], pi)
        assert np.array_equal(u, outs[0][4])
        assert L.pk_emu_set_cluster(1) == 0              # no non-local stores happened
    finally:
        L.pk_emu_set_dump(0, None, None, None, None)
        L.pk_emu_set_cluster(1)


def test_fused_dump_of_failed_realisation_and_golden_suite():
    """A failed realisation dumps NaN rows / -1 indices (outputting.py:107-108); on the reference helper's golden
    fields the fused peakon-suite scan returns the indices the reference's own peakon_diagnostics resolves."""
    import ctypes, os
    from emu_util import _p
    from peakondrake_b200 import outputting as out
    L = lib()
    GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_helpers.npz'))
    try:
        for name in [str(n) for n in GOLD['names']]:
            u0 = GOLD[name + '/u']; ne = len(u0); Ld = 40.0
            # a zero-length time step leaves the field as it is (dt -> 0: u1 = u is the Newton solution)
            dg = np.zeros((1, 1, 14)); di = np.zeros((1, 1, 4), np.int32); sv = np.zeros((1, 1, 2)); si = np.zeros((1, 1, 3), np.int32)
            L.pk_emu_set_dump(1, _p(dg), _p(di, ctypes.c_int), _p(sv), _p(si, ctypes.c_int))
            u = u0[None].copy(); u1 = u.copy(); dF = np.zeros_like(u); dG = np.zeros_like(u)
            emu_step(ne, Ld, 0.0, 0, u, u1, dF, dG, None, None, 1, nxi=0)
            assert np.array_equal(u[0], u0)
            x = (np.arange(ne) + 0.5) * (Ld / ne)
            pk = out.peakon_suite_from_indices(x, dg[0, 0, 7], sv[0, 0], si[0, 0])
            got = np.array([pk[k] for k in out.PEAKON_DIAGNOSTICS])
            assert np.array_equal(got, GOLD[name + '/peakon'], equal_nan=True), name
        # failure: NaN increments make SNES fail at the first step
        ne = 64; u0 = O.initial_u(O.Mesh1D(ne, 40.0), 'proper_peak')
        dg = np.zeros((2, 1, 14)); di = np.zeros((2, 1, 4), np.int32); sv = np.zeros((2, 1, 2)); si = np.zeros((2, 1, 3), np.int32)
        L.pk_emu_set_dump(1, _p(dg), _p(di, ctypes.c_int), _p(sv), _p(si, ctypes.c_int))
        u = u0[None].copy(); z = [np.zeros_like(u) for _ in range(3)]
        r = emu_step(ne, 40.0, 1e-3, 0, u, z[0], z[1], z[2], None, np.full((2, 1, 1), np.nan), 2, xi_const=0.2, nxi=1)
        assert r['fail_step'][0] == 0 and np.isnan(dg).all() and (di == -1).all() and np.isnan(sv).all() and (si == -1).all()
    finally:
        L.pk_emu_set_dump(0, None, None, None, None)


@pytest.mark.parametrize('ne,C,nxi,fam', [(1000, 4, 7, 'sines'), (2000, 8, 2, 'sines'), (250, 2, 1, 'constant'), (333, 2, 0, 'sines')])
def test_upwind_cluster_kernel_emulated(ne, C, nxi, fam):
    """upwind/ssprk3 on a thread-block cluster (meshes beyond 4096 cells on the device): exchange slots distributed
    over C CTAs, every store local, scans with a cross-CTA level -- same numbers as the oracle, both thread orders
    bit-identical, a blow-up flag raised in one CTA is seen by all."""
    from emu_util import emu_step_upwind
    L = lib()
    Ld, dt, n = 40., 1e-3, 6
    mesh = O.Mesh1D(ne, Ld); u0 = O.initial_u(mesh, 'one_peak')
    xi = O.noise_tables(mesh, fam, nxi, 0.25, scheme='upwind')[0] if nxi else None
    dW = O.reference_dW(0, n, nxi, dt) if nxi else None
    ref = O.run_upwind(ne, Ld, dt, n, u0, dW, xi)
    res = []
    try:
        assert L.pk_emu_set_cluster(C) == 0
        for order in (0, 1):
            m = ref['oracle'].initial_m(u0)[None].copy(); u = u0[None].copy()
            xc = xi[0, 1] if fam == 'constant' else 0.0
            emu_step_upwind(ne, Ld, dt, m, u, None if (fam == 'constant' or not nxi) else np.ascontiguousarray(xi),
                            None if dW is None else np.ascontiguousarray(dW[:, None, :]), n, xi_const=xc, nxi=nxi, order=order)
            assert rel(u[0], ref['u']) <= TOL and rel(m[0], ref['m']) <= TOL
            res.append((u.copy(), m.copy()))
        assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
        if nxi:      # a non-finite increment fails the realisation in every CTA alike
            m = ref['oracle'].initial_m(u0)[None].copy(); u = u0[None].copy()
            bad = np.ascontiguousarray(dW[:, None, :]).copy(); bad[3] = np.inf
            r = emu_step_upwind(ne, Ld, dt, m, u, None if fam == 'constant' else np.ascontiguousarray(xi), bad, n,
                                xi_const=(xi[0, 1] if fam == 'constant' else 0.0), nxi=nxi)
            assert r['fail_step'][0] == 3 and np.isfinite(u).all()
        assert L.pk_emu_set_cluster(1) == 0          # no non-local stores happened
    finally:
        L.pk_emu_set_cluster(1)


@pytest.mark.parametrize('ne,C', [(1000, 4), (2000, 8)])
def test_test_scheme_on_a_cluster_emulated(ne, C):
    """scheme 'test' (equations.py:132-145) through the cluster executor (meshes beyond 4096 cells on the device)."""
    L = lib()
    Ld, dt, n, nxi = 40.0, 1e-3, 4, 2
    mesh = O.Mesh1D(ne, Ld)
    u0 = O.initial_u(mesh, 'proper_peak')
    tabs = O.noise_tables(mesh, 'sines', nxi, 0.2, scheme='test')
    dW = O.reference_dW(1, n, nxi, dt)
    ref = O.run(ne, Ld, dt, n, u0, dW, tabs, scheme='test')
    try:
        assert L.pk_emu_set_cluster(C) == 0
        u = u0[None].copy(); z = [np.zeros_like(u) for _ in range(3)]
        r = emu_step(ne, Ld, dt, 4, u, z[0], z[1], z[2], [np.ascontiguousarray(t) for t in tabs], np.ascontiguousarray(dW[:, None, :]), n)
        assert r['its_total'][0] == ref['its'].sum() and rel(u[0], ref['u']) <= TOL
        assert L.pk_emu_set_cluster(1) == 0
    finally:
        L.pk_emu_set_cluster(1)


def test_philox_many_sub_updates_have_a_fixed_summation_order():
    """nXi_updates >= 64 (the strong-convergence sweep uses int(dt/1e-6) = 100 ... 10000): the sub-draws of one increment
    are shared by the 32 lanes of a warp on the device; the summation order (32 interleaved partial sums, added in lane order) is
    part of the definition, so the emulation, the peakon-SDE kernel and every block size give the same bits."""
    L = lib()
    ne, Ld, dt, nupd, nsteps, R, seed = 64, 40.0, 1e-2, 100, 5, 3, 1234
    u0 = O.initial_u(O.Mesh1D(ne, Ld), 'periodic_peakon')
    outs = []
    for (T, M) in ((0, 0), (8, 8)):         # 32-thread and 8-thread layouts of the same mesh (the latter cannot share draws)
        u = np.tile(u0, (R, 1)); z = [np.zeros_like(u) for _ in range(3)]
        r = emu_step(ne, Ld, dt, 0, u, z[0], z[1], z[2], None, None, nsteps, xi_const=0.5, nupd=nupd, seed=seed, rid0=7,
                     export=True, nxi=1, T=T, M=M + (1 if M else 0))
        outs.append(r['dW'])
    assert np.array_equal(outs[0], outs[1])
    sdt = np.sqrt(dt / nupd)
    for n in (0, 4):
        for rr in range(R):
            pa = [0.0] * 32
            for l in range(32):
                s = 0.0
                for j in range(l, nupd, 32):
                    s += L.pk_emu_philox_normal(seed, 7 + rr, n * nupd + j, 0, 0) * sdt
                pa[l] = s
            tot = pa[0]
            for l in range(1, 32):
                tot += pa[l]
            assert outs[0][n, rr, 0] == tot
    assert abs(np.std(outs[0]) / np.sqrt(dt) - 1.0) < 0.5
