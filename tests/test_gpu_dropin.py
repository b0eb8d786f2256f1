"""
GPU tests of the drop-in surface: experiment()/simulation() with the reference's keyword arguments,
the data.nc schema, the device peakon-suite and ensemble-statistics kernels.
"""
import os
import numpy as np
import pytest

from oracle import ch_oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, 'golden', 'reference_helpers.npz'))


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_experiment_seed_ensemble_matches_oracle_per_seed(tmp_path, monkeypatch):
    """The reference's runnable test parameters (tests/test_dumping.py: 10 steps, dumps every step, 7 sines)
    transplanted to the hydrodynamic scheme, with a seed sweep batched on the GPU."""
    from peakondrake_b200 import experiment, outputting as out
    monkeypatch.chdir(tmp_path)
    Ld, dt, tmax, ne, nxi, sigma = 40., 0.001, 0.01, 200, 7, 0.25
    seeds = [0, 1, 2, 5]
    experiment('test_one_run', Ld, tmax, resolutions=ne, dts=dt, sigmas=sigma, seeds=seeds, schemes='hydrodynamic',
               timesteppings='midpoint', ics='one_peak', num_Xis=nxi, Xi_family='sines', alphasq=1.0, c0=0.,
               gamma=0., diagnostics=['h1_u', 'l2_u', 'energy', 'max_du', 'min_du', 'min_du_loc', 'q_pde', 'u_field',
                                      'mu', 'peakon_suite', 'l2_m', 'E_2'],
               fields_to_output=['du'], ndump=1, field_ndump=1, allow_fail=True)
    d = out.read_data_file('results/test_one_run/data.nc')
    assert d['time'].shape == (11,) and np.allclose(d['time'], np.arange(11) * dt)
    assert d['energy'].shape == (11, 4) and d['u_field'].shape == (11, 4, ne) and list(d['seed']) == seeds
    assert np.allclose(d['x'], (np.arange(ne) + 0.5) * Ld / ne)
    assert np.isnan(d['failed_time']).all() and (d['wallclock_time'][1] >= d['wallclock_time'][0]).all()
    mesh = O.Mesh1D(ne, Ld)
    u0 = O.initial_u(mesh, 'one_peak')
    tabs = O.noise_tables(mesh, 'sines', nxi, sigma)
    orc = O.CHOracle(ne, Ld, dt)
    for k, s in enumerate(seeds):
        dW = O.reference_dW(s, 10, nxi, dt)               # the reference's own stream for this seed
        ref = O.run(ne, Ld, dt, 10, u0, dW, tabs, dump_every=1)
        for t in range(1, 11):
            u = ref['dumps'][t - 1]
            assert rel(d['u_field'][t, k], u) <= 1e-10
            sd = O.scalar_diagnostics(orc, u, ['energy', 'l2_u', 'h1_u', 'max_du', 'min_du', 'min_du_loc', 'q_pde',
                                               'l2_m', 'E_2'], m=O.m_field(orc, u))
            for name, v in sd.items():
                assert abs(d[name][t, k] - v) <= 1e-9 * max(1.0, abs(v)), (name, t, k)
        assert abs(d['energy'][0, k] - (u0 @ (orc.M @ u0) + u0 @ (orc.K @ u0))) < 1e-12
    # field dumps (outputting.py:59-68,224-232): one .pvd collection per realisation, prognostic fields + 'du'
    for k, s in enumerate(seeds):
        pvd = 'results/test_one_run/fields_seed%d_output.pvd' % k
        assert open(pvd).read().count('<DataSet') == 11
        piece = out.read_field_piece(pvd[:-4] + '_10.vtu')
        assert sorted(piece) == sorted(['points', 'u', 'dXi_x', 'dXi_xx', 'dXi', 'du'])
        assert np.array_equal(piece['u'][:, 0], d['u_field'][10, k]) and np.array_equal(piece['u'][:, 1], np.roll(d['u_field'][10, k], -1))
        dW = O.reference_dW(s, 10, nxi, dt)
        assert np.allclose(piece['dXi'][:, 0], dW[9] @ tabs[0], rtol=0, atol=1e-15)
        assert np.allclose(piece['dXi_xx'][:, 0], dW[9] @ tabs[2], rtol=0, atol=1e-13)
        assert np.allclose(piece['du'][:, 0], O.du_field(mesh, d['u_field'][10, k]), rtol=0, atol=1e-12)
        assert not out.read_field_piece(pvd[:-4] + '_0.vtu')['dXi'].any()
    assert np.isfinite(d['mu_0']).all() and np.isnan(d['mu_1']).all()        # single peak
    assert np.isfinite(d['peakon_mu']).all() and (d['peakon_max_u'] > 0.4).all()


def test_simulation_single_realisation_and_convergence_error(tmp_path, monkeypatch):
    from peakondrake_b200 import simulation, ConvergenceError, PeriodicIntervalMesh
    monkeypatch.chdir(tmp_path)
    p = dict(dt=(1e-3,), resolution=(100,), seed=(3,), sigma=(0.2,), scheme=('hydrodynamic',), timestepping=('midpoint',),
             alphasq=(1.0,), c0=(0.,), gamma=(0.,), ic=('proper_peak',), num_Xis=(2,), Xi_family=('sines',),
             dirname=('results/x',), Ld=(40.,), tmax=(0.005,), ndump=(-1,), field_ndump=(-1,), nXi_updates=(1,),
             allow_fail=(False,), smooth_t=(None,), peakon_equations=(False,), only_peakons=(False,),
             true_peakon_data=(None,), true_mean_peakon_data=(None,), peakon_speed=(None,), periodic=(False,),
             fixed_dW=(None,), peak_width=(1.0,), peakon_method=('ito_euler',), store_coordinates=[True],
             mesh=(PeriodicIntervalMesh(100, 40.),), file_name=('none',))
    u = simulation(p, diagnostic_values=None, expected_u=True)
    mesh = O.Mesh1D(100, 40.)
    ref = O.run(100, 40., 1e-3, 5, O.initial_u(mesh, 'proper_peak'), O.reference_dW(3, 5, 2, 1e-3),
                O.noise_tables(mesh, 'sines', 2, 0.2))
    assert rel(u, ref['u']) <= 1e-10
    # a forcing SNES cannot handle: without allow_fail the failure propagates (simulation.py:107-108)
    p2 = dict(p); p2['smooth_t'] = (lambda t: float('nan'),)
    with pytest.raises(ConvergenceError):
        simulation(p2, diagnostic_values=None)
    p3 = dict(p); p3['peakon_equations'] = ('yes',)
    with pytest.raises(ValueError):
        simulation(p3)                                                     # simulation.py:68-69


def test_true_peakon_workflow(tmp_path, monkeypatch):
    """The strong-convergence design of experiments/publication/stochastic (true_peakon_sde.py then
    peakon_strong_dt.py), scaled down: SDE-only 'true' run, then PDE runs measured against it."""
    from peakondrake_b200 import experiment, outputting as out
    monkeypatch.chdir(tmp_path)
    common = dict(resolutions=400, sigmas=1.0, seeds=0, schemes='hydrodynamic', timesteppings='midpoint',
                  ics='periodic_peakon', num_Xis=1, Xi_family='constant', alphasq=1.0, c0=0., gamma=0.,
                  allow_fail=True, periodic=True)
    tmax, dt_true, dt = 0.02, 1e-4, 2e-3
    experiment('true_peakon_data', 40., tmax, dts=dt_true, diagnostics=['l2_u'], fields_to_output=[],
               ndump=int(round(tmax / dt_true)), field_ndump=int(round(tmax / dt_true)), peakon_equations=True, only_peakons=True, **common)
    t = out.read_data_file('results/true_peakon_data/data.nc')
    assert t['p'].shape == (2,) and np.isnan(t['l2_u'][1]) and not np.isnan(t['p'][1])
    # q moves with speed p/2 * coth-like factor plus sigma*W: with sigma=1 the displacement is O(sqrt(tmax))
    assert abs(t['q'][1] - t['q'][0]) < 1.0 and t['p'][1] == t['p'][0]      # constant Xi: dp = 0
    experiment('strong', 40., tmax, dts=dt, diagnostics=['u_error_with_sde', 'u_error_weak', 'p_pde', 'q_pde'],
               fields_to_output=['u_sde', 'u_sde_weak'], ndump=int(round(tmax / dt)), field_ndump=int(round(tmax / dt)),
               true_peakon_data='true_peakon_data', nXi_updates=int(round(dt / dt_true)), **common)
    s = out.read_data_file('results/strong/data.nc')
    assert s['u_error_with_sde'].shape == (2,) and s['u_error_with_sde'][1] < 0.05
    assert s['u_error_with_sde'][1] == s['u_error_weak'][1]
    # only_peakons: the PDE never runs, fields are dumped at t=0 only (simulation.py:88-89,128-131)
    assert open('results/true_peakon_data/fields_output.pvd').read().count('<DataSet') == 1
    piece = out.read_field_piece('results/strong/fields_output_1.vtu')
    us = O.u_sde_field(O.Mesh1D(400, 40.), t['p'][1], t['q'][1], 1.0, True)
    assert np.allclose(piece['u_sde'][:, 0], us, rtol=1e-13)
    err = piece['u_sde_weak'][:, 0]
    assert abs(np.sqrt(err @ (O.Mesh1D(400, 40.).mass() @ err)) - s['u_error_with_sde'][1]) < 1e-12
    # same Brownian path: the PDE peak follows the SDE peakon to within a few cells
    assert abs(s['q_pde'][1] - t['q'][1]) < 0.5


@pytest.mark.parametrize('name', [str(n) for n in GOLD['names']])
def test_device_peakon_suite_against_reference_helper(name):
    import torch
    from peakondrake_b200 import EnsembleIntegrator, outputting as out, _lib
    u = GOLD[name + '/u']; ne = len(u); Ld = 40.0
    eng = EnsembleIntegrator(ne, Ld, 1e-3, 2)
    eng.u.copy_(torch.from_numpy(np.stack([u, np.roll(u, 17)])))
    vals, idx, _ = eng.diagnostics()
    pv, pi = eng.peakon_suite()
    torch.cuda.synchronize()
    x = (np.arange(ne) + 0.5) * (Ld / ne)      # the coordinate array the fixtures were made with
    res = out.peakon_suite_from_indices(x, vals[0, _lib.D_MAX_U].item(), pv[0].cpu().numpy(), pi[0].cpu().numpy())
    got = np.array([res[k] for k in out.PEAKON_DIAGNOSTICS])
    assert np.array_equal(got, GOLD[name + '/peakon'], equal_nan=True)      # bit-exact incl. locations


def test_ensemble_statistics_kernels():
    import torch
    from peakondrake_b200 import EnsembleIntegrator, EnsembleStatistics, build_initial_u
    ne, Ld, R = 300, 40.0, 70
    eng = EnsembleIntegrator(ne, Ld, 1e-3, R, num_Xis=2, sigma=0.3, seed=5)
    eng.set_state(build_initial_u(ne, Ld, 'proper_peak'))
    eng.step(5)
    eng.fail_step[3] = 2                     # a failed realisation is excluded from the statistics
    st = EnsembleStatistics(eng, nbins=16, lo=0.0, hi=Ld, du_lo=-2.0, du_hi=0.0)
    st.update()
    torch.cuda.synchronize()
    u = eng.u.cpu().numpy(); ok = np.ones(R, bool); ok[3] = False
    assert st.count.item() == R - 1
    assert np.allclose(st.mean_u.cpu().numpy(), u[ok].mean(0), rtol=1e-13, atol=1e-15)
    assert np.allclose(st.packed[ne:2 * ne].cpu().numpy(), (u[ok] ** 2).sum(0), rtol=1e-13)
    du = (np.roll(u, -1, axis=1) - u) / (Ld / ne)
    q = (np.argmax(u[ok], axis=1) + 0.5) * Ld / ne
    hq = np.histogram(q, bins=16, range=(0, Ld))[0]
    hd = np.histogram(np.clip(du[ok].min(1), -2.0, -1e-12), bins=16, range=(-2.0, 0.0))[0]
    assert np.array_equal(st.hist_q.cpu().numpy(), hq) and np.array_equal(st.hist_min_du.cpu().numpy(), hd)
    a = st.packed.clone(); st.update(); torch.cuda.synchronize()
    assert torch.equal(a, st.packed)         # deterministic (fixed-order sums, integer histograms)


def test_reference_test_dumping_script_parameters(tmp_path, monkeypatch):
    """The reference's one runnable test, verbatim parameters (tests/test_dumping.py): upwind/ssprk3,
    20 cells, 10 steps, 7 sines, sigma 0.25, the five norms, uscalar/Xiscalar fields."""
    from peakondrake_b200 import experiment, outputting as out
    monkeypatch.chdir(tmp_path)
    code, Ld, dt, tmax = 'test_one_run', 40., 0.001, 0.01
    experiment(code, Ld, tmax, resolutions=20, dts=dt, sigmas=0.25, seeds=0, schemes='upwind', timesteppings='ssprk3',
               ics='one_peak', num_Xis=7, Xi_family='sines', alphasq=1.0, c0=0., gamma=0.,
               diagnostics=['h1_u', 'h1_m', 'l2_u', 'l2_m', 'energy'], fields_to_output=['uscalar', 'Xiscalar'],
               ndump=int(tmax / (10 * dt)), field_ndump=int(tmax / (10 * dt)))
    d = out.read_data_file('results/test_one_run/data.nc')
    assert d['time'].shape == (11,) and d['energy'].shape == (11,)
    mesh = O.Mesh1D(20, Ld)
    u0 = O.initial_u(mesh, 'one_peak')
    xi = O.noise_tables(mesh, 'sines', 7, 0.25, scheme='upwind')[0]
    dW = O.reference_dW(0, 10, 7, dt)
    orc = O.UpwindOracle(20, Ld, dt)
    m, u = orc.initial_m(u0), u0.copy()
    for n in range(11):
        if n > 0:
            m, u = orc.solve(m, u, dW[n - 1] @ xi)
        nm = O.dg1_norms(orc, m)
        l2u = np.sqrt(u @ (orc.M @ u)); h1u = np.sqrt(u @ (orc.M @ u) + u @ (orc.K @ u))
        for name, v in (('l2_u', l2u), ('h1_u', h1u), ('energy', u @ (orc.M @ u) + u @ (orc.K @ u)),
                        ('l2_m', nm['l2_m']), ('h1_m', nm['h1_m'])):
            assert abs(d[name][n] - v) <= 1e-10 * max(1.0, abs(v)), (name, n)
    piece = out.read_field_piece('results/test_one_run/fields_output_10.vtu')
    assert sorted(piece) == sorted(['points', 'u', 'm', 'dXi', 'uscalar', 'Xiscalar'])
    assert np.allclose(piece['m'], m.reshape(20, 2), rtol=1e-10, atol=1e-12) and np.allclose(piece['uscalar'][:, 0], u, atol=1e-12)
    assert np.allclose(piece['Xiscalar'][:, 0], dW[9] @ xi, atol=1e-15)
    with pytest.raises(IOError):          # re-running needs a fresh results directory (experiment.py:25-27)
        experiment(code, Ld, tmax, resolutions=20, dts=dt, sigmas=0.25, seeds=0, schemes='upwind',
                   timesteppings='ssprk3', ics='one_peak', num_Xis=7, Xi_family='sines', alphasq=1.0, c0=0., gamma=0.,
                   diagnostics=['energy'])


def test_experiment_multi_parameter_sweep_schema(tmp_path, monkeypatch):
    """Swept resolutions x seeds x sigmas: dimension order and indexing of data.nc (experiment.py:59-156:
    'resolution' first, then the order of the parameter dictionary)."""
    from peakondrake_b200 import experiment, outputting as out
    monkeypatch.chdir(tmp_path)
    res, seeds, sigmas = [100, 200], [3, 4, 5], [0.0, 0.2]
    experiment('sweep', 40., 0.004, resolutions=res, dts=1e-3, sigmas=sigmas, seeds=seeds, schemes='hydrodynamic',
               timesteppings='midpoint', ics='proper_peak', num_Xis=2, Xi_family='sines', alphasq=1.0, c0=0., gamma=0.,
               diagnostics=['energy', 'max_du', 'q_pde'], ndump=2, allow_fail=True)
    d = out.read_data_file('results/sweep/data.nc')
    assert d['energy'].shape == (3, 2, 3, 2)            # (time, resolution, seed, sigma)
    assert d['failed_time'].shape == (2, 3, 2) and d['wallclock_time'].shape == (2, 2, 3, 2)
    assert np.allclose(d['deltax'], [0.4, 0.2]) and list(d['seed']) == seeds and np.allclose(d['sigma'], sigmas)
    assert d['x'].shape == (100,)                       # swept resolution -> fixed 100-point x axis (:113-115)
    # sigma = 0: every seed gives the same (deterministic) answer; sigma > 0: seeds differ
    assert np.ptp(d['energy'][2, :, :, 0], axis=1).max() == 0.0
    assert np.ptp(d['energy'][2, :, :, 1], axis=1).min() > 0.0
    for ir, ne in enumerate(res):
        mesh = O.Mesh1D(ne, 40.)
        u0 = O.initial_u(mesh, 'proper_peak')
        tabs = O.noise_tables(mesh, 'sines', 2, 0.2)
        for js, s in enumerate(seeds):
            ref = O.run(ne, 40., 1e-3, 4, u0, O.reference_dW(s, 4, 2, 1e-3), tabs)
            orc = ref['oracle']
            e = ref['u'] @ (orc.M @ ref['u']) + ref['u'] @ (orc.K @ ref['u'])
            assert abs(d['energy'][2, ir, js, 1] - e) <= 1e-10 * e


def test_experiment_with_the_test_scheme(tmp_path, monkeypatch):
    """scheme 'test' / 'midpoint' (equations.py:132-145) through the drop-in: seeds batched, u_field and
    energy against the oracle run with the reference's stream per seed."""
    from peakondrake_b200 import experiment, outputting as out
    monkeypatch.chdir(tmp_path)
    Ld, dt, ne, nxi, sigma, seeds = 40., 1e-3, 150, 3, 0.2, [1, 4]
    experiment('tscheme', Ld, 0.006, resolutions=ne, dts=dt, sigmas=sigma, seeds=seeds, schemes='test',
               timesteppings='midpoint', ics='proper_peak', num_Xis=nxi, Xi_family='sines', alphasq=1.0, c0=0.,
               gamma=0., diagnostics=['energy', 'l2_u', 'u_field'], ndump=3)
    d = out.read_data_file('results/tscheme/data.nc')
    mesh = O.Mesh1D(ne, Ld)
    u0 = O.initial_u(mesh, 'proper_peak')
    tabs = O.noise_tables(mesh, 'sines', nxi, sigma, scheme='test')
    orc = O.CHOracle(ne, Ld, dt, scheme='test')
    for k, s in enumerate(seeds):
        ref = O.run(ne, Ld, dt, 6, u0, O.reference_dW(s, 6, nxi, dt), tabs, scheme='test')
        assert rel(d['u_field'][2, k], ref['u']) <= 1e-10
        e = ref['u'] @ (orc.M @ ref['u']) + ref['u'] @ (orc.K @ ref['u'])
        assert abs(d['energy'][2, k] - e) <= 1e-10 * e


def test_device_mu_against_reference_helper_and_host_routine():
    """pk_mu (find_mus on the device, outputting.py:260-370): bit-equal to the reference helper's golden outputs and,
    on evolved stochastic states of a large ensemble, to the host restatement; failed realisations give NaN."""
    import torch
    from peakondrake_b200 import EnsembleIntegrator, build_initial_u, outputting as out
    names = [str(n) for n in GOLD['names']]
    ne, Ld = len(GOLD[names[0] + '/u']), 40.0
    assert all(len(GOLD[n + '/u']) == ne for n in names)
    eng = EnsembleIntegrator(ne, Ld, 1e-3, len(names))
    eng.u.copy_(torch.from_numpy(np.stack([GOLD[n + '/u'] for n in names])).to(eng.device))
    eng.fail_step[-1] = 3
    vals, status = [t.cpu().numpy() for t in eng.mu()]
    for k, name in enumerate(names[:-1]):
        n_old = int(GOLD[name + '/nmus'][0]); kk = min(4, n_old)
        want = np.full(8, np.nan)
        want[:kk] = GOLD[name + '/mus'][:kk]; want[4:4 + kk] = GOLD[name + '/alt'][:kk]
        assert status[k] >= 0 and np.array_equal(vals[k], want, equal_nan=True), name
    assert status[-1] == -3 and np.isnan(vals[-1]).all()
    eng.close()
    # evolved stochastic ensemble, several resolutions (shared-memory and global-memory paths)
    for ne, R, nsteps in ((1000, 300, 40), (16384, 6, 2)):
        eng = EnsembleIntegrator(ne, Ld, 1e-3, R, num_Xis=(4 if ne == 1000 else 1),
                                 Xi_family=('sines' if ne == 1000 else 'constant'), sigma=0.3, seed=2)
        eng.set_state(build_initial_u(ne, Ld, 'two_peaks' if ne == 1000 else 'proper_peak'))
        eng.step(nsteps)
        vals, status = [t.cpu().numpy() for t in eng.mu()]
        u = eng.u.cpu().numpy()
        x = (np.arange(ne) + 0.5) * (Ld / ne)
        assert (status >= 0).all()
        for r in range(min(R, 40)):
            du = (np.roll(u[r], -1) - u[r]) / (Ld / ne)
            old, alt = out.mu_diagnostic(u[r], du, x, Ld)
            kk = min(4, len(old))
            assert np.array_equal(vals[r, :kk], np.array(old[:kk])) and np.array_equal(vals[r, 4:4 + kk], np.array(alt[:kk]))
            assert np.isnan(vals[r, kk:4]).all() and np.isnan(vals[r, 4 + kk:]).all()
        eng.close()


def test_publication_script_patterns_scaled_down(tmp_path, monkeypatch):
    """The parameter patterns of the reference's publication scripts, scaled down in time:
    experiments/publication/stochastic/mu_stochastic_sine.py (resolution x sigma x seed sweep, 6 sines, mu) with a
    resolution that needs the cluster kernel with noise tables, and experiments/publication/deterministic/
    shallow_demo.py (ne=20000, sigma=0 with one 'sines' mode, u_field and slope locations)."""
    from peakondrake_b200 import experiment, outputting as out
    monkeypatch.chdir(tmp_path)
    Ld, dt = 40., 1e-3
    tmax = 4 * dt
    res, sigmas, seeds = [1000, 3000], [0.05, 0.5], range(5, 8)
    experiment('mu_sine', Ld, tmax, resolutions=res, dts=dt, sigmas=sigmas, seeds=seeds, schemes='hydrodynamic',
               timesteppings='midpoint', ics='proper_peak', num_Xis=6, Xi_family='sines', alphasq=1.0, c0=0., gamma=0.,
               diagnostics=['mu', 'max_du', 'min_du'], fields_to_output=['du'], ndump=2, field_ndump=4,
               allow_fail=True, nXi_updates=1)
    d = out.read_data_file('results/mu_sine/data.nc')
    assert d['mu_0'].shape == (3, 2, 3, 2) and d['min_du'].shape == (3, 2, 3, 2) and np.isfinite(d['mu_0']).all()
    assert np.isnan(d['mu_1']).all() and np.isnan(d['failed_time']).all()
    assert os.path.exists('results/mu_sine/fields_resolution1_seed2_sigma1_output_1.vtu')
    for ir, ne in enumerate(res):
        mesh = O.Mesh1D(ne, Ld)
        u0 = O.initial_u(mesh, 'proper_peak')
        orc = O.CHOracle(ne, Ld, dt)
        for jg, sg in enumerate(sigmas):
            tabs = O.noise_tables(mesh, 'sines', 6, sg)
            s = list(seeds)[1]
            ref = O.run(ne, Ld, dt, 4, u0, O.reference_dW(s, 4, 6, dt), tabs)
            sd = O.scalar_diagnostics(orc, ref['u'], ['max_du', 'min_du'])
            for name, v in sd.items():
                assert abs(d[name][2, ir, 1, jg] - v) <= 1e-9 * max(1.0, abs(v)), (name, ne, sg)
    # shallow_demo.py: deterministic, ne = 20000 (clusters of 16)
    ne = 20000
    experiment('shallow_demo', Ld, 3 * dt, resolutions=ne, dts=dt, sigmas=0.0, seeds=0, schemes='hydrodynamic',
               timesteppings='midpoint', ics='new_peak', num_Xis=1, Xi_family='sines', alphasq=1.0, c0=0., gamma=0.,
               diagnostics=['u_field', 'max_du_loc', 'min_du_loc'], fields_to_output=[], ndump=3, field_ndump=3,
               allow_fail=True, nXi_updates=1, peak_width=1.0)
    d = out.read_data_file('results/shallow_demo/data.nc')
    mesh = O.Mesh1D(ne, Ld)
    u0 = O.initial_u(mesh, 'new_peak', peak_width=1.0)
    ref = O.run(ne, Ld, dt, 3, u0, np.zeros((3, 1)), O.noise_tables(mesh, 'sines', 1, 0.0))
    assert rel(d['u_field'][1], ref['u']) <= 1e-10
    du = O.du_field(mesh, ref['u'])
    x = (np.arange(ne) + 0.5) * (Ld / ne)          # the DG0 coordinate array the drop-in looks locations up in
    assert d['max_du_loc'][1] == x[np.argmax(du)] and d['min_du_loc'][1] == x[np.argmin(du)]
