"""
GPU tests of the round-2 additions: the device peakon SDE (bit-exact against the reference's own update()), du_smooth,
the ensemble scalars, and the BASELINE configurations C2-C5 run THROUGH THE DROP-IN (`experiment()`) at their stated
parameters -- mesh, initial condition, noise family, dt, dump cadence and diagnostics -- against the oracle on a
sample of realisations, each with and without the Newton convergence certificate (PK_NO_CERTIFICATE).
The realisation count is the only thing reduced (the oracle needs seconds per realisation at these meshes); the
full-size ensembles are covered by the property tests in test_gpu_parity.py and by bench.py.
"""
import os
import numpy as np
import pytest

from oracle import ch_oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, 'golden', 'reference_helpers.npz'))
TOL = 1e-10


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


# ---- device peakon SDE vs the reference's own PeakonEquations.update (golden vectors) --------------------------
@pytest.mark.parametrize('method', ['ito_euler', 'milstein'])
def test_device_peakon_sde_bit_exact_against_reference_update(method):
    import torch
    from peakondrake_b200 import EnsembleIntegrator
    dW, traj = GOLD['sde_%s/dW' % method], GOLD['sde_%s/traj' % method]
    tabs = [GOLD['sde_%s/%s' % (method, k)] for k in ('xi', 'xi_x', 'xi_xx')]
    ne, nst, R = tabs[0].shape[1], len(dW), 5
    eng = EnsembleIntegrator(ne, 40.0, 1e-2, R, num_Xis=2, Xi_family='sines', sigma=0.2)
    dev = eng.device
    p = torch.full((R,), 1.0, dtype=torch.float64, device=dev)
    q = torch.full((R,), 39.5, dtype=torch.float64, device=dev)
    t_dev = [torch.from_numpy(np.ascontiguousarray(t)).to(dev) for t in tabs]
    dW_dev = torch.from_numpy(np.ascontiguousarray(np.repeat(dW[:, None, :], R, axis=1))).to(dev)
    E = float(np.exp(-40.0 / 1.0))
    tr = eng.peakon_sde_step(nst, p, q, t_dev, E, method, dW=dW_dev, traj_stride=1)
    torch.cuda.synchronize()
    tr = tr.cpu().numpy()
    for r in range(R):
        assert np.array_equal(tr[:, r, :], traj), np.abs(tr[:, r, :] - traj).max()      # bit-exact, 200 steps, q wraps
    assert p.cpu().numpy()[0] == traj[-1, 0] and q.cpu().numpy()[0] == traj[-1, 1]
    # several launches == one launch
    p2 = torch.full((R,), 1.0, dtype=torch.float64, device=dev); q2 = torch.full((R,), 39.5, dtype=torch.float64, device=dev)
    eng.peakon_sde_step(77, p2, q2, t_dev, E, method, dW=dW_dev[:77].contiguous())
    eng.peakon_sde_step(nst - 77, p2, q2, t_dev, E, method, dW=dW_dev[77:].contiguous())
    assert torch.equal(p, p2) and torch.equal(q, q2)
    eng.close()


def test_device_sde_philox_stream_is_the_step_kernels_stream():
    """Without a PDE (`only_peakons`) the SDE kernel draws the increments itself -- the numbers pk_step would draw."""
    import torch
    from peakondrake_b200 import EnsembleIntegrator, build_initial_u, pure_xi_tables
    ne, Ld, dt, R, nxi, nupd = 200, 40.0, 1e-3, 6, 3, 4
    eng = EnsembleIntegrator(ne, Ld, dt, R, num_Xis=nxi, Xi_family='sines', sigma=0.3, nXi_updates=nupd, seed=11, rid_offset=40)
    eng.set_state(build_initial_u(ne, Ld, 'proper_peak'))
    eng.step(3)
    dW = eng.step(9, export_dW=True)           # global steps 3..11
    tabs = [torch.from_numpy(np.ascontiguousarray(t)).to(eng.device) for t in pure_xi_tables(ne, Ld, 'sines', nxi, 0.3)]
    mk = lambda v: torch.full((R,), v, dtype=torch.float64, device=eng.device)
    pa, qa, pb, qb = mk(0.7), mk(11.3), mk(0.7), mk(11.3)
    E = float(np.exp(-Ld))
    eng.peakon_sde_step(9, pa, qa, tabs, E, 'ito_euler', dW=dW)
    eng.peakon_sde_step(9, pb, qb, tabs, E, 'ito_euler', dW=None, step0=3)
    torch.cuda.synchronize()
    assert torch.equal(pa, pb) and torch.equal(qa, qb) and len(set(qa.cpu().numpy())) == R
    # and against the oracle's restatement
    sde = O.PeakonSDE(0.7, 11.3, Ld, dt, 1.0, pure_xi_tables(ne, Ld, 'sines', nxi, 0.3), R)
    for n in range(9):
        sde.update(dW[n].cpu().numpy())
    assert np.allclose(pa.cpu().numpy(), sde.p, rtol=1e-14) and np.allclose(qa.cpu().numpy(), sde.q, rtol=1e-14)
    eng.close()


def test_philox_sub_updates_same_bits_on_every_layout():
    """nXi_updates = 128: the draws of an increment are shared by a warp; the exported increments do not depend on the
    mesh / block size / cluster layout, and the SDE kernel regenerates the same numbers."""
    import torch
    from peakondrake_b200 import EnsembleIntegrator, build_initial_u, pure_xi_tables
    outs = []
    for ne in (64, 1000, 16384):
        eng = EnsembleIntegrator(ne, 40.0, 1e-3, 4, num_Xis=1, Xi_family='constant', sigma=0.3, nXi_updates=128, seed=5, rid_offset=3)
        eng.set_state(build_initial_u(ne, 40.0, 'periodic_peakon'))
        a = eng.step(3, export_dW=True); b = eng.step(4, export_dW=True)
        torch.cuda.synchronize()
        outs.append(torch.cat([a, b]).cpu().numpy())
        if ne == 64:
            tabs = [torch.from_numpy(np.ascontiguousarray(t)).to(eng.device) for t in pure_xi_tables(ne, 40.0, 'constant', 1, 0.3)]
            mk = lambda v: torch.full((4,), v, dtype=torch.float64, device=eng.device)
            pa, qa, pb, qb = mk(0.7), mk(11.3), mk(0.7), mk(11.3)
            eng.peakon_sde_step(7, pa, qa, tabs, float(np.exp(-40.0)), 'ito_euler', dW=torch.cat([a, b]))
            eng.peakon_sde_step(7, pb, qb, tabs, float(np.exp(-40.0)), 'ito_euler', dW=None, step0=0)
            assert torch.equal(qa, qb)
        eng.close()
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2]) and abs(outs[0].std() / np.sqrt(1e-3) - 1) < 0.4
    # a coarse run with nXi_updates = dt/dt0 walks the Brownian path of the fine run (the reference's MT19937 property,
    # stochastic_functions.py:211-217, kept by keying the sub-updates with the fine step index)
    dt0 = 2.0 ** -14
    def path(dt, nupd, nsteps):
        eng = EnsembleIntegrator(64, 40.0, dt, 3, num_Xis=2, Xi_family='sines', sigma=0.1, nXi_updates=nupd, seed=5, rid_offset=3)
        eng.set_state(build_initial_u(64, 40.0, 'proper_peak'))
        w = eng.step(nsteps, export_dW=True).cpu().numpy(); eng.close()
        return w
    fine = path(dt0, 1, 256 * 4)
    for nupd in (4, 256):            # serial and warp-shared summation
        coarse = path(dt0 * nupd, nupd, 1024 // nupd)
        assert np.allclose(coarse, fine.reshape(1024 // nupd, nupd, 3, 2).sum(axis=1), rtol=1e-12, atol=1e-16)


def test_du_smooth_and_ensemble_scalars():
    import torch
    from peakondrake_b200 import EnsembleIntegrator, EnsembleStatistics, build_initial_u, _lib
    for ne in (24, 300):
        Ld, R = 40.0, 9
        eng = EnsembleIntegrator(ne, Ld, 1e-3, R, num_Xis=2, sigma=0.3, seed=5)
        eng.set_state(build_initial_u(ne, Ld, 'proper_peak'))
        eng.step(4)
        eng.fail_step[2] = 1
        outv, ids, fld = eng.du_smooth_loc(want_field=True)
        torch.cuda.synchronize()
        outv, ids, fld = outv.cpu().numpy(), ids.cpu().numpy(), fld.cpu().numpy()
        u = eng.u.cpu().numpy()
        mesh = O.Mesh1D(ne, Ld)
        for r in range(R):
            if r == 2:
                assert np.isnan(outv[r]).all() and (ids[r] == -1).all()
                continue
            xmax, xmin, w = O.smooth_locs(mesh, u[r])
            assert np.abs(fld[r] - w).max() <= 1e-9 * np.abs(w).max()
            assert outv[r, 1] == xmax and outv[r, 3] == xmin and ids[r, 0] == np.argmax(w) and ids[r, 1] == np.argmin(w)
        st = EnsembleStatistics(eng, nbins=8)
        st.update(want_m=True)
        torch.cuda.synchronize()
        d = eng.diag.cpu().numpy(); ok = np.ones(R, bool); ok[2] = False
        assert st.count.item() == R - 1 and st.fail_count.item() == 1
        for col in (_lib.D_ENERGY, _lib.D_MIN_DU, _lib.D_MASS_M):
            mean, var, lo, hi = st.scalar_stats(col)
            assert abs(mean - d[ok, col].mean()) <= 1e-13 * abs(mean) and lo == d[ok, col].min() and hi == d[ok, col].max()
            assert abs(var - d[ok, col].var()) <= 1e-12 * max(abs(mean) ** 2, 1e-30)
        assert np.isnan(st.scalar_stats(_lib.D_ERR_SDE2)[0]) or st.scalar_stats(_lib.D_ERR_SDE2)[0] == 0.0
        eng.close()


# ---- the BASELINE configurations through experiment() ----------------------------------------------------------
def _suite_from_u(u, x_cells):
    """the reference's peakon_diagnostics restated on one field (see test_host_logic: bit-equal to the golden vectors)"""
    from peakondrake_b200 import outputting as out
    ne = len(u); h = 40.0 / ne
    du = (np.roll(u, -1) - u) / h
    pc = int(np.argmax(u)); S = ne // 2
    right = np.nonzero(np.roll(du, -pc)[:S] > np.roll(du, -pc + 1)[:S])[0]
    left_a = np.roll(du[::-1], pc)[:S]              # du[pc-1-k]
    left_b = np.roll(du[::-1], pc + 1)[:S]          # du[pc-k]
    left = np.nonzero(left_a < left_b)[0]
    vals = [np.nan, np.nan]; ids = [pc, -1, -1]
    if len(right):
        D = (pc + right[0]) % ne; ids[1] = (D - 1) % ne; vals[0] = du[ids[1]]
    if len(left):
        D = (pc - 1 - left[0]) % ne; ids[2] = (D - 1) % ne; vals[1] = du[(D + 1) % ne]
    return out.peakon_suite_from_indices(x_cells, u.max(), np.array(vals), np.array(ids))


@pytest.mark.parametrize('nocert', ['0', '1'])
def test_c2_1024_seeds_through_experiment(tmp_path, monkeypatch, nocert):
    """BASELINE config 2: 1024 realisations, ne=1000, 4 sin/cos modes, the reference's per-seed MT19937 streams."""
    from peakondrake_b200 import experiment, outputting as out
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv('PK_NO_CERTIFICATE', nocert)
    ne, Ld, dt, nst = 1000, 40., 1e-3, 60
    experiment('c2', Ld, nst * dt, resolutions=ne, dts=dt, sigmas=0.2, seeds=range(1024), schemes='hydrodynamic',
               timesteppings='midpoint', ics='proper_peak', num_Xis=4, Xi_family='sines', alphasq=1.0, c0=0., gamma=0.,
               diagnostics=['energy', 'min_du', 'max_du', 'q_pde', 'u_field'], ndump=30, allow_fail=True, ensemble_stats=True)
    d = out.read_data_file('results/c2/data.nc')
    assert d['u_field'].shape == (3, 1024, ne) and np.isnan(d['failed_time']).all() and (d['ens_count'] == 1024).all()
    mesh = O.Mesh1D(ne, Ld)
    u0 = O.initial_u(mesh, 'proper_peak'); tabs = O.noise_tables(mesh, 'sines', 4, 0.2); orc = O.CHOracle(ne, Ld, dt)
    for s in (0, 511, 1023):
        ref = O.run(ne, Ld, dt, nst, u0, O.reference_dW(s, nst, 4, dt), tabs, dump_every=30)
        for t in (1, 2):
            assert rel(d['u_field'][t, s], ref['dumps'][t - 1]) <= TOL
            sd = O.scalar_diagnostics(orc, ref['dumps'][t - 1], ['energy', 'min_du', 'max_du', 'q_pde'])
            for k, v in sd.items():
                assert abs(d[k][t, s] - v) <= 1e-9 * max(1.0, abs(v))
    assert np.allclose(d['ens_mean_u'], d['u_field'].mean(axis=1), atol=1e-13)
    assert np.allclose(d['ens_min_du_min'], d['min_du'].min(axis=1))


@pytest.mark.parametrize('nocert', ['0', '1'])
def test_c3_configuration_with_statistics_every_100_steps(tmp_path, monkeypatch, nocert):
    """BASELINE config 3 at its parameters: ne=2000, ic periodic_peakon, 4 sines sigma=0.2, dt=1e-3, moments and the
    q_pde / min_du histograms every 100 steps; 48 realisations of the 8192 per GPU, two of them against the oracle."""
    from peakondrake_b200 import experiment, outputting as out
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv('PK_NO_CERTIFICATE', nocert)
    ne, Ld, dt, nst, R = 2000, 40., 1e-3, 200, 48
    experiment('c3', Ld, nst * dt, resolutions=ne, dts=dt, sigmas=0.2, seeds=range(R), schemes='hydrodynamic',
               timesteppings='midpoint', ics='periodic_peakon', num_Xis=4, Xi_family='sines', alphasq=1.0, c0=0., gamma=0.,
               diagnostics=['q_pde', 'min_du', 'u_field'], ndump=100, allow_fail=True, ensemble_stats=True)
    d = out.read_data_file('results/c3/data.nc')
    assert d['time'].shape == (3,) and d['ens_hist_q_pde'].shape == (3, 64) and (d['ens_hist_q_pde'].sum(axis=1) == R).all()
    mesh = O.Mesh1D(ne, Ld)
    u0 = O.initial_u(mesh, 'periodic_peakon'); tabs = O.noise_tables(mesh, 'sines', 4, 0.2); orc = O.CHOracle(ne, Ld, dt)
    for s in (0, R - 1):
        ref = O.run(ne, Ld, dt, nst, u0, O.reference_dW(s, nst, 4, dt), tabs, dump_every=100)
        for t in (1, 2):
            u = ref['dumps'][t - 1]
            assert rel(d['u_field'][t, s], u) <= TOL
            sd = O.scalar_diagnostics(orc, u, ['q_pde', 'min_du'])
            assert d['q_pde'][t, s] == sd['q_pde'] and abs(d['min_du'][t, s] - sd['min_du']) <= 1e-9 * abs(sd['min_du'])
    # the statistics are those of the rows
    assert np.allclose(d['ens_mean_u'], d['u_field'].mean(axis=1), atol=1e-13)
    assert np.allclose(d['ens_var_u'], d['u_field'].var(axis=1), atol=1e-12)
    hq = np.stack([np.histogram(d['q_pde'][t], bins=64, range=(0, Ld))[0] for t in range(3)])
    assert np.array_equal(d['ens_hist_q_pde'], hq)


@pytest.mark.parametrize('nocert', ['0', '1'])
def test_c4_configuration_per_step_diagnostics(tmp_path, monkeypatch, nocert):
    """BASELINE config 4 at its parameters (steep_stochastic_diagnostics.py:19-31): ne=16384, ic new_peak with
    peak_width=1/6, constant Xi sigma=0.2, dt=5e-4, min_du / max_du / max_jump_local / peakon_suite EVERY step,
    60 steps; 3 realisations of the 4096, two of them against the oracle."""
    from peakondrake_b200 import experiment, outputting as out
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv('PK_NO_CERTIFICATE', nocert)
    ne, Ld, dt, nst, R = 16384, 40., 5e-4, 60, 3
    experiment('c4', Ld, nst * dt, resolutions=ne, dts=dt, sigmas=0.2, seeds=[68, 1, 2], schemes='hydrodynamic',
               timesteppings='midpoint', ics='new_peak', num_Xis=1, Xi_family='constant', alphasq=1.0, c0=0., gamma=0.,
               diagnostics=['max_du', 'min_du', 'max_jump_local', 'peakon_suite', 'u_field'], ndump=1, allow_fail=True,
               nXi_updates=1, peak_width=(1 / 6))
    d = out.read_data_file('results/c4/data.nc')
    assert d['min_du'].shape == (nst + 1, R) and np.isnan(d['failed_time']).all()
    mesh = O.Mesh1D(ne, Ld)
    u0 = O.initial_u(mesh, 'new_peak', peak_width=1 / 6); tabs = O.noise_tables(mesh, 'constant', 1, 0.2)
    orc = O.CHOracle(ne, Ld, dt)
    for k, s in ((0, 68), (2, 2)):
        ref = O.run(ne, Ld, dt, nst, u0, O.reference_dW(s, nst, 1, dt), tabs, dump_every=1)
        for t in range(1, nst + 1):
            u = ref['dumps'][t - 1]
            assert rel(d['u_field'][t, k], u) <= TOL, (t, rel(d['u_field'][t, k], u))
            sd = O.scalar_diagnostics(orc, u, ['max_du', 'min_du', 'max_jump_local'])
            for name, v in sd.items():
                assert abs(d[name][t, k] - v) <= 1e-8 * max(1.0, abs(v)), (name, t)
    # the same run without `u_field` takes the fused in-kernel dumps (one launch for all 60 dumps): same numbers
    experiment('c4f', Ld, nst * dt, resolutions=ne, dts=dt, sigmas=0.2, seeds=[68, 1, 2], schemes='hydrodynamic',
               timesteppings='midpoint', ics='new_peak', num_Xis=1, Xi_family='constant', alphasq=1.0, c0=0., gamma=0.,
               diagnostics=['max_du', 'min_du', 'max_jump_local', 'peakon_suite'], ndump=1, allow_fail=True,
               nXi_updates=1, peak_width=(1 / 6))
    f = out.read_data_file('results/c4f/data.nc')
    for name in ['max_du', 'min_du', 'max_jump_local'] + out.PEAKON_DIAGNOSTICS:
        assert np.array_equal(f[name], d[name], equal_nan=True), name
    # the peakon suite is the reference helper's answer on the dumped field, bit for bit, at every step
    for k in range(R):
        for t in range(0, nst + 1, 7):
            pk = _suite_from_u(d['u_field'][t, k], mesh.x_cells)
            for name in out.PEAKON_DIAGNOSTICS:
                assert d[name][t, k] == pk[name] or (np.isnan(d[name][t, k]) and np.isnan(pk[name])), (name, t, k)


DTS_C5 = [1e-2, 5e-3, 2.5e-3, 1e-3, 5e-4, 2.5e-4, 1e-4]


@pytest.mark.parametrize('nocert', ['0', '1'])
def test_c5_configuration_strong_dt_sweep(tmp_path, monkeypatch, nocert):
    """BASELINE config 5 at its parameters (peakon_strong_dt.py:9-37, true_peakon_sde.py:5-30): ne=20000, periodic
    peakon, constant Xi sigma=1, periodic=True, tmax=0.1; the "true" path is the peakon SDE at dt=1e-6 per seed
    (1e5 Ito-Euler steps on the device, 64 seeds), every dt of the sweep runs with nXi_updates=int(dt/1e-6) -- i.e. on the SAME
    Brownian path -- and reports u_error_with_sde at ne=20000.  Every dt is checked against the oracle on its first
    steps (the oracle needs ~0.3 s per step at this mesh); the errors of the whole sweep give the strong order."""
    from peakondrake_b200 import experiment, outputting as out
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv('PK_NO_CERTIFICATE', nocert)
    ne, Ld, tmax, dt_true = 20000, 40., 0.1, 1e-6
    seeds = list(range(64))
    common = dict(resolutions=ne, sigmas=1.0, seeds=seeds, schemes='hydrodynamic', timesteppings='midpoint',
                  ics='periodic_peakon', num_Xis=1, Xi_family='constant', alphasq=1.0, c0=0., gamma=0.,
                  allow_fail=True, periodic=True)
    nt = int(round(tmax / dt_true))
    experiment('true_peakon_data', Ld, tmax, dts=dt_true, diagnostics=['l2_u'], fields_to_output=[], ndump=nt,
               field_ndump=nt, peakon_equations=True, only_peakons=True, **common)
    tp = out.read_data_file('results/true_peakon_data/data.nc')
    assert tp['p'].shape == (2, 64) and (tp['p'][1] == tp['p'][0]).all()          # constant Xi: dp = 0
    # the true path is q0 + c t + sigma W(t) up to the Euler sum's rounding: check against the host stream
    from peakondrake_b200 import draw_dW
    for k, s in enumerate(seeds[:2]):
        W = draw_dW(s, nt, 1, dt_true).sum()
        E = np.exp(-Ld)
        assert abs((tp['q'][1, k] - tp['q'][0, k]) - (tp['p'][0, k] / 2 * (1 + E) / (1 - E) * tmax + 1.0 * W)) < 1e-9
    mesh = O.Mesh1D(ne, Ld)
    u0 = O.initial_u(mesh, 'periodic_peakon'); tabs = O.noise_tables(mesh, 'constant', 1, 1.0)
    errs = []
    for i, dt in enumerate(DTS_C5):
        nst = int(round(tmax / dt)); nupd = int(round(dt / dt_true))
        experiment('strong_%d' % i, Ld, tmax, dts=dt, diagnostics=['u_error_with_sde', 'u_error_weak', 'p_pde', 'q_pde'],
                   fields_to_output=['u_sde', 'u_sde_weak'], ndump=nst, field_ndump=-1,
                   true_peakon_data='true_peakon_data', nXi_updates=nupd, **common)
        s = out.read_data_file('results/strong_%d/data.nc' % i)
        assert s['u_error_with_sde'].shape == (2, 64) and np.isfinite(s['u_error_with_sde']).all()
        assert np.array_equal(s['u_error_with_sde'], s['u_error_weak'])
        errs.append(s['u_error_with_sde'][1])
        # oracle on the first steps of this dt, same Brownian path (coarse increments = sums of the fine ones)
        ncheck = min(nst, 12)
        experiment('short_%d' % i, Ld, ncheck * dt, dts=dt, diagnostics=['u_field', 'q_pde'], ndump=ncheck,
                   nXi_updates=nupd, **common)
        sh = out.read_data_file('results/short_%d/data.nc' % i)
        for k in (0, 63):
            dW = O.reference_dW(seeds[k], ncheck, 1, dt, nupd)
            ref = O.run(ne, Ld, dt, ncheck, u0, dW, tabs)
            assert rel(sh['u_field'][1, k], ref['u']) <= TOL, (dt, k, rel(sh['u_field'][1, k], ref['u']))
            assert sh['q_pde'][1, k] == mesh.x_cells[np.argmax(ref['u'])]
    errs = np.array(errs)                       # [dt, seed]
    rms = np.sqrt((errs ** 2).mean(axis=1))
    # Strong convergence, fitted like plotting/peakon_strong_dt_plotter.py:37-39 (log error against log dt).  Expected
    # order 1/2: for constant Xi the SPDE solution is the deterministic peakon shifted by sigma W(t), and the coarse
    # increments are sums of the fine ones, so the only error is the midpoint rule's phase error of pure transport,
    # theta^3/12 per step and Fourier mode (theta = k * displacement).  Summed over the steps it is k^3 S / 12 with
    # S = sum_n d_n^3 / 1 = O(dt) (3 c sigma^2 T dt in the mean); a peakon's spectrum decays like k^-2 (the kink), so
    # ||error||_L2^2 ~ int min(k^6 S^2, 4) k^-4 dk ~ S, i.e. ||error|| ~ dt^(1/2).
    slope = np.polyfit(np.log(DTS_C5), np.log(rms), 1)[0]
    print('C5 rms u_error_with_sde per dt:', dict(zip(DTS_C5, rms)), 'fitted strong order: %.3f' % slope)
    assert (np.diff(rms) < 0).all() and 0.35 <= slope <= 0.65


@pytest.mark.parametrize('ne,fam,nxi,force', [(1000, 'sines', 4, 0), (200, 'sines', 2, 0), (20, 'sines', 1, 0), (2000, 'sines', 4, 0),
                                               (1000, 'sines', 2, 4), (16384, 'constant', 1, 0), (20000, 'constant', 1, 0),
                                               (4000, 'constant', 1, 0)])
def test_fused_dump_equals_separate_kernels(ne, fam, nxi, force, monkeypatch):
    """pk_step_dump (diagnostics + peakon suite evaluated in the step kernel on the shared-memory state, several dumps per
    launch) against pk_step + pk_diagnostics + pk_peakon_suite after every dump interval: identical states, bit-identical
    extrema / indices / suite, sums to rounding.  Single-CTA, forced-cluster and the two cluster meshes of C4 / C5."""
    import torch
    from peakondrake_b200 import EnsembleIntegrator, build_initial_u, _lib
    if force:
        monkeypatch.setenv('PK_FORCE_CLUSTER', str(force))
    R, every, nd = 5, 3, 4
    ic = 'new_peak' if ne == 16384 else 'proper_peak'
    mk = lambda: EnsembleIntegrator(ne, 40.0, 5e-4, R, num_Xis=nxi, Xi_family=fam, sigma=0.2, seed=9)
    a, b = mk(), mk()
    u0 = build_initial_u(ne, 40.0, ic, peak_width=1 / 6)
    a.set_state(u0); b.set_state(u0)
    a.fail_step[3] = 0; b.fail_step[3] = 0            # a failed realisation dumps NaN rows
    na = a.launch_count
    _, dg, di, sv, si = a.step_dump(every * nd, every, suite=True)
    torch.cuda.synchronize()
    dg, di, sv, si = [t.cpu().numpy() for t in (dg, di, sv, si)]
    # third engine: state snapshots + batched diagnostics over all dumps (pk_step_snap + pk_diagnostics_rows + ..._suite_rows)
    c = mk(); c.set_state(u0); c.fail_step[3] = 0
    nc = c.launch_count
    _, cg_, ci, cv, cs, snaps, sflag = c.step_snap(every * nd, every, suite=True, want_m=True)
    torch.cuda.synchronize()
    assert c.launch_count - nc == 3
    cg_, ci, cv, cs, snaps, sflag = [t.cpu().numpy() for t in (cg_, ci, cv, cs, snaps, sflag)]
    n0 = b.launch_count
    for d in range(nd):
        b.step(every)
        v, ix, _ = b.diagnostics(want_m=True)
        pv, pi = b.peakon_suite()
        torch.cuda.synchronize()
        v, ix, pv, pi = [t.cpu().numpy() for t in (v, ix, pv, pi)]
        # the batched rows are the separate kernels' output bit for bit (same kernels, same state)
        assert np.array_equal(cg_[d][:, :13], v[:, :13], equal_nan=True) and np.array_equal(ci[d], ix)
        assert np.array_equal(cv[d], pv, equal_nan=True) and np.array_equal(cs[d], pi)
        assert np.array_equal(snaps[d][np.arange(R) != 3], b.u.cpu().numpy()[np.arange(R) != 3]) and sflag[d][3] == 0 and (sflag[d][np.arange(R) != 3] == -1).all()
        assert np.array_equal(di[d], ix) and np.array_equal(si[d], pi) and np.array_equal(sv[d], pv, equal_nan=True)
        assert np.array_equal(dg[d][:, 4:8], v[:, 4:8], equal_nan=True)
        ok = np.arange(R) != 3
        assert np.allclose(dg[d][ok, :4], v[ok, :4], rtol=1e-12, atol=0) and np.isnan(dg[d][3]).all() and np.isnan(dg[d][:, 8:]).all()
    assert torch.equal(a.u, b.u) and torch.equal(a.its_total, b.its_total)
    assert b.launch_count - n0 == 3 * nd and a.launch_count - na == 1     # one kernel for all dumps of the chunk
    assert torch.equal(c.u, b.u)
    a.close(); b.close(); c.close()
