"""
Host logic of the drop-in surface on CPU: experiment() / simulation() / run_batch() driven through the emulated
engine (tests/emu_engine.py: the CUDA timestep source executed on the CPU, NumPy diagnostics).  What is tested here
is everything AROUND the kernels: the seed axis as ensemble axis (no 1000-seed cap), one array write per diagnostic
and dump, records appended to data.nc as the run goes, the ensemble statistics variables, the t = 0 dump of
`only_peakons` runs, `simulation()` appending to `file_name`, and the sharding of the seeds over the ranks of a
torch.distributed job (gloo, world_size 2) with rank 0 writing the file.
"""
import os
import sys
import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from oracle import ch_oracle as O          # noqa: E402


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


KW = dict(resolutions=40, dts=1e-3, sigmas=0.3, schemes='hydrodynamic', timesteppings='midpoint', ics='proper_peak',
          num_Xis=2, Xi_family='sines', alphasq=1.0, c0=0., gamma=0.)
DIAGS = ['energy', 'max_du', 'min_du', 'q_pde', 'E_3', 'u_field', 'peakon_suite', 'mu', 'max_du_smooth_loc']


def _check_against_oracle(d, seeds, nsteps, ndump, ne=40, Ld=40., dt=1e-3, nxi=2, sigma=0.3):
    mesh = O.Mesh1D(ne, Ld)
    u0 = O.initial_u(mesh, 'proper_peak')
    tabs = O.noise_tables(mesh, 'sines', nxi, sigma)
    orc = O.CHOracle(ne, Ld, dt)
    for k, s in enumerate(seeds):
        ref = O.run(ne, Ld, dt, nsteps, u0, O.reference_dW(s, nsteps, nxi, dt), tabs, dump_every=ndump)
        for t, u in enumerate([u0] + ref['dumps']):
            assert rel(d['u_field'][t, k], u) <= 1e-10
            sd = O.scalar_diagnostics(orc, u, ['energy', 'max_du', 'min_du', 'q_pde'])
            for name, v in sd.items():
                assert abs(d[name][t, k] - v) <= 1e-9 * max(1.0, abs(v)), (name, t, k)
            assert abs(d['E_3'][t, k] - 0.5 * (u @ (orc.M @ u))) < 1e-12                 # u_xx == 0 in the reference
            assert d['max_du_smooth_loc'][t, k] == O.smooth_locs(mesh, u)[0]


def test_experiment_ensemble_axis_records_and_statistics(tmp_path, monkeypatch, built):
    from emu_engine import EmuEngine
    from peakondrake_b200 import experiment, outputting as out
    monkeypatch.chdir(tmp_path)
    seeds = [3, 0, 7, 11, 5]
    experiment('ens', 40., 0.006, seeds=seeds, diagnostics=DIAGS, ndump=2, allow_fail=True, ensemble_stats=True,
               engine_cls=EmuEngine, **KW)
    d = out.read_data_file('results/ens/data.nc')
    assert d['time'].shape == (4,) and d['energy'].shape == (4, 5) and d['u_field'].shape == (4, 5, 40)
    assert list(d['seed']) == seeds and d['failed_time'].shape == (5,) and np.isnan(d['failed_time']).all()
    assert d['wallclock_time'].shape == (2, 5) and (d['wallclock_time'][1] >= d['wallclock_time'][0]).all()
    _check_against_oracle(d, seeds, 6, 2)
    # ensemble statistics of every dump
    assert d['ens_count'].shape == (4,) and (d['ens_count'] == 5).all() and (d['ens_fail_count'] == 0).all()
    assert np.allclose(d['ens_mean_u'], d['u_field'].mean(axis=1), rtol=0, atol=1e-14)
    assert np.allclose(d['ens_var_u'], d['u_field'].var(axis=1), rtol=0, atol=1e-13)
    assert np.allclose(d['ens_min_du_min'], d['min_du'].min(axis=1)) and np.allclose(d['ens_max_du_max'], d['max_du'].max(axis=1))
    assert np.allclose(d['ens_energy_mean'], d['energy'].mean(axis=1)) and np.allclose(d['ens_energy_var'], d['energy'].var(axis=1), atol=1e-14)
    assert d['ens_hist_q_pde'].shape == (4, 64) and (d['ens_hist_q_pde'].sum(axis=1) == 5).all()
    assert (d['ens_hist_min_du'].sum(axis=1) == 5).all()


def test_batched_dump_paths_equal_separate_kernels(tmp_path, monkeypatch, built):
    """Several dumps per launch: state snapshots + one batched diagnostics call (pk_step_snap, the default), or -- when the
    snapshot budget is too small -- the dump fused into the step kernel (pk_step_dump); `mu` in the list forces the plain
    path (one launch and separate kernels per dump).  Same file every way (scalars to rounding of the sums, extrema /
    locations / peakon suite / u_field / p, q bit for bit)."""
    from emu_engine import EmuEngine
    from peakondrake_b200 import experiment, outputting as out
    monkeypatch.chdir(tmp_path)
    diags = ['energy', 'l2_u', 'h1_u', 'mass_u', 'max_du', 'min_du', 'max_jump_local', 'max_jump_global', 'max_du_loc',
             'min_du_loc', 'p_pde', 'q_pde', 'E_0', 'E_1', 'E_3', 'peakon_suite']
    calls = []
    for name in ('step_dump', 'step_snap'):
        def spy(self, *a, _orig=getattr(EmuEngine, name), _name=name, **k):
            calls.append((_name,) + a[:2])
            return _orig(self, *a, **k)
        monkeypatch.setattr(EmuEngine, name, spy)
    kw = dict(seeds=[5, 6, 7], ndump=2, allow_fail=True, peakon_equations=True, engine_cls=EmuEngine, **KW)
    experiment('snap', 40., 0.009, diagnostics=diags + ['u_field', 'l2_m', 'E_2'], **kw)     # 9 steps: 4 dumps batched + 1 trailing step
    monkeypatch.setenv('PK_SNAP_BYTES', '0')
    experiment('fused', 40., 0.009, diagnostics=diags, **kw)
    experiment('plain', 40., 0.009, diagnostics=diags + ['u_field', 'l2_m', 'E_2', 'mu'], **kw)
    assert calls == [('step_snap', 8, 2), ('step_dump', 8, 2)]
    a, b, c = [out.read_data_file('results/%s/data.nc' % n) for n in ('snap', 'fused', 'plain')]
    assert a['time'].shape == (5,) and np.array_equal(a['time'], c['time']) and np.array_equal(b['time'], c['time'])
    for x in (a, b):
        for k in x:
            if k in ('wallclock_time',):
                continue
            if k in ('energy', 'l2_u', 'h1_u', 'mass_u', 'E_0', 'E_1', 'E_3', 'l2_m', 'E_2'):
                assert np.allclose(x[k], c[k], rtol=1e-12, atol=0), k
            else:
                assert np.array_equal(x[k], c[k], equal_nan=True), k
    monkeypatch.delenv('PK_SNAP_BYTES')
    # a failure inside a batched chunk: NaN rows from the failing dump on, failed_time recorded
    blow = lambda t: float('nan') if t > 0.0045 else 0.01
    experiment('fail', 40., 0.008, diagnostics=['energy', 'min_du', 'u_field'], seeds=[1, 2], ndump=2, allow_fail=True, smooth_t=blow,
               engine_cls=EmuEngine, **KW)
    f = out.read_data_file('results/fail/data.nc')
    assert np.isfinite(f['energy'][:3]).all() and np.isnan(f['energy'][3:]).all() and np.allclose(f['failed_time'], 0.005)
    assert np.isnan(f['u_field'][3:]).all() and np.isfinite(f['u_field'][:3]).all()


def test_seed_axis_is_not_capped(tmp_path, monkeypatch, built):
    """experiment.py:164-166 refuses more than 1000 simulation() calls; the seeds of one combination are ONE batched
    launch here, so 1024 (BASELINE config 2) and more are legal."""
    from emu_engine import EmuEngine
    from peakondrake_b200 import experiment, outputting as out
    monkeypatch.chdir(tmp_path)
    kw = dict(KW); kw['resolutions'] = 16; kw['num_Xis'] = 1
    experiment('many', 40., 0.001, seeds=range(1100), diagnostics=['energy'], ndump=1, rng='philox', engine_cls=EmuEngine, **kw)
    d = out.read_data_file('results/many/data.nc')
    assert d['energy'].shape == (2, 1100) and np.isfinite(d['energy']).all() and len(set(d['energy'][1])) > 1000


def test_swept_parameters_and_fixed_seed(tmp_path, monkeypatch, built):
    from emu_engine import EmuEngine
    from peakondrake_b200 import experiment, outputting as out
    monkeypatch.chdir(tmp_path)
    kw = dict(KW); kw['sigmas'] = [0.1, 0.3]; kw['resolutions'] = [24, 40]
    experiment('sweep', 40., 0.002, seeds=[4, 9, 2], diagnostics=['energy', 'q_pde'], ndump=1, engine_cls=EmuEngine, **kw)
    d = out.read_data_file('results/sweep/data.nc')
    assert d['energy'].shape == (3, 2, 3, 2) and np.isfinite(d['energy']).all()        # time, resolution, seed, sigma
    assert np.allclose(d['deltax'], [40. / 24, 40. / 40]) and d['failed_time'].shape == (2, 3, 2)
    mesh = O.Mesh1D(24, 40.)
    ref = O.run(24, 40., 1e-3, 2, O.initial_u(mesh, 'proper_peak'), O.reference_dW(9, 2, 2, 1e-3),
                O.noise_tables(mesh, 'sines', 2, 0.1))
    orc = O.CHOracle(24, 40., 1e-3)
    assert abs(d['energy'][2, 0, 1, 0] - O.scalar_diagnostics(orc, ref['u'], ['energy'])['energy']) < 1e-10
    # no swept parameter at all: scalars are written without dimensions, like the reference's file
    kw1 = dict(KW)
    experiment('single', 40., 0.002, seeds=6, diagnostics=['energy'], ndump=1, engine_cls=EmuEngine, **kw1)
    s = out.read_data_file('results/single/data.nc')
    assert s['energy'].shape == (3,) and s['failed_time'].shape == () and s['wallclock_time'].shape == (2,)


def test_dumps_survive_a_convergence_error(tmp_path, monkeypatch, built):
    """The reference appends to data.nc as it goes, so the dumps made before a ConvergenceError are on disk
    (simulation.py:107-108 with allow_fail=False)."""
    from emu_engine import EmuEngine
    from peakondrake_b200 import experiment, ConvergenceError, outputting as out
    monkeypatch.chdir(tmp_path)
    blow = lambda t: float('nan') if t > 0.0035 else 0.01
    with pytest.raises(ConvergenceError):
        experiment('boom', 40., 0.006, seeds=[1, 2], diagnostics=['energy'], ndump=1, smooth_t=blow, allow_fail=False,
                   engine_cls=EmuEngine, **KW)
    d = out.read_data_file('results/boom/data.nc')
    assert d['energy'].shape == (4, 2) and np.isfinite(d['energy']).all()      # t = 0 and the three good steps


def test_only_peakons_dumps_initial_diagnostics_and_device_sde(tmp_path, monkeypatch, built):
    """simulation.py:85-88: `failed` is True from the start, but the dump at t = 0 is made with failed=False."""
    from emu_engine import EmuEngine
    from peakondrake_b200 import experiment, outputting as out
    monkeypatch.chdir(tmp_path)
    kw = dict(KW); kw['ics'] = 'periodic_peakon'; kw['Xi_family'] = 'sines'
    experiment('sde', 40., 0.004, seeds=[0, 1], diagnostics=['l2_u', 'energy'], ndump=2, peakon_equations=True,
               only_peakons=True, periodic=True, engine_cls=EmuEngine, **kw)
    d = out.read_data_file('results/sde/data.nc')
    mesh = O.Mesh1D(40, 40.)
    u0 = O.initial_u(mesh, 'periodic_peakon')
    orc = O.CHOracle(40, 40., 1e-3)
    assert np.allclose(d['l2_u'][0], np.sqrt(u0 @ (orc.M @ u0))) and np.isnan(d['l2_u'][1:]).all()
    assert np.isnan(d['failed_time']).all() and d['p'].shape == (3, 2) and np.isfinite(d['p']).all()
    # (p, q) follow the reference's update with the reference's increments for each seed
    from peakondrake_b200.peakon_equations import pure_xi_tables, initial_pq
    tabs = pure_xi_tables(40, 40., 'sines', 2, 0.3)
    from peakondrake_b200 import build_initial_u
    p0, q0 = initial_pq(build_initial_u(40, 40., 'periodic_peakon'), mesh.x_nodes, 40., 1.0)
    for k, s in enumerate([0, 1]):
        sde = O.PeakonSDE(p0, q0, 40., 1e-3, 1.0, tabs, 1)
        dW = O.reference_dW(s, 4, 2, 1e-3)
        for n in range(4):
            sde.update(dW[n][None, :])
        assert d['p'][2, k] == sde.p[0] and d['q'][2, k] == sde.q[0]


def test_simulation_appends_to_file_name(tmp_path, monkeypatch, built):
    """Called with the reference's signature (no data_file), simulation() appends to simulation_parameters['file_name']
    (outputting.py:29-32) -- and raises if experiment() has not created the file."""
    from emu_engine import EmuEngine
    from peakondrake_b200 import outputting as out, PeriodicIntervalMesh
    import peakondrake_b200
    sim_mod = sys.modules['peakondrake_b200.simulation']
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(sim_mod, 'EnsembleIntegrator', EmuEngine)
    os.makedirs('results/x')
    fn = 'results/x/data.nc'
    df = out.DataFile(fn)
    df.createDimension('time', None); df.createDimension('x', 40); df.createDimension('wallclock_times', 2)
    df.createDimension('sigma', 2)
    for n, dims in (('time', ('time',)), ('x', ('x',)), ('energy', ('time', 'sigma')), ('wallclock_time', ('wallclock_times', 'sigma')),
                    ('failed_time', ('sigma',))):
        df.createVariable(n, float, dims)
    df.close()
    p = dict(dt=(1e-3,), resolution=(40,), seed=(3,), sigma=(1, 0.2), scheme=('hydrodynamic',), timestepping=('midpoint',),
             alphasq=(1.0,), c0=(0.,), gamma=(0.,), ic=('proper_peak',), num_Xis=(2,), Xi_family=('sines',),
             dirname=('results/x',), Ld=(40.,), tmax=(0.003,), ndump=(1,), field_ndump=(-1,), nXi_updates=(1,),
             allow_fail=(False,), smooth_t=(None,), peakon_equations=(False,), only_peakons=(False,),
             true_peakon_data=(None,), true_mean_peakon_data=(None,), peakon_speed=(None,), periodic=(False,),
             fixed_dW=(None,), peak_width=(1.0,), peakon_method=('ito_euler',), store_coordinates=[True],
             mesh=(PeriodicIntervalMesh(40, 40.),), file_name=(fn,))
    sim_mod.simulation(p, diagnostic_values=['energy'])
    d = out.read_data_file(fn)
    assert d['energy'].shape == (4, 2) and np.isnan(d['energy'][:, 0]).all() and np.isfinite(d['energy'][:, 1]).all()
    p['sigma'] = (0, 0.1)
    sim_mod.simulation(p, diagnostic_values='energy')                      # a bare string, outputting.py:35-36
    d = out.read_data_file(fn)
    assert np.isfinite(d['energy']).all() and np.allclose(d['x'], (np.arange(40) + 0.5))
    p['file_name'] = ('results/x/nothing.nc',)
    with pytest.raises((IOError, OSError)):
        sim_mod.simulation(p, diagnostic_values=['energy'])


def test_mean_peakon_diagnostics_follow_the_reference_wiring(tmp_path, monkeypatch, built):
    """u_error_with_sde_mean (diagnostic_equations.py:291-313): raises as the reference does when its update_constants
    cannot work, and otherwise reproduces which expression follows the true-peakon file."""
    from emu_engine import EmuEngine
    from peakondrake_b200 import experiment, outputting as out
    monkeypatch.chdir(tmp_path)
    kw = dict(KW); kw['ics'] = 'peakon'; kw['num_Xis'] = 1; kw['Xi_family'] = 'constant'
    experiment('true_peakon_data', 40., 0.004, seeds=0, diagnostics=['l2_u'], ndump=2, peakon_equations=True,
               only_peakons=True, engine_cls=EmuEngine, **kw)
    with pytest.raises(AttributeError):
        experiment('m1', 40., 0.004, seeds=0, diagnostics=['u_error_with_sde_mean'], ndump=2, engine_cls=EmuEngine, **kw)
    with pytest.raises(AttributeError):
        experiment('m2', 40., 0.004, seeds=0, diagnostics=['u_error_with_sde_mean'], ndump=2,
                   true_peakon_data='true_peakon_data', engine_cls=EmuEngine, **kw)
    with pytest.raises(NotImplementedError):
        experiment('m3', 40., 0.004, seeds=0, diagnostics=['u_error_with_sde', 'u_error_with_sde_mean'], ndump=2,
                   true_peakon_data='true_peakon_data', periodic=True, engine_cls=EmuEngine, **kw)
    experiment('m4', 40., 0.004, seeds=0, diagnostics=['u_error_with_sde', 'u_error_with_sde_mean', 'u_field'], ndump=2,
               true_peakon_data='true_peakon_data', engine_cls=EmuEngine, **kw)
    d = out.read_data_file('results/m4/data.nc')
    t = out.read_data_file('results/true_peakon_data/data.nc')
    mesh = O.Mesh1D(40, 40.); M = mesh.mass()
    for j in range(3):
        u = d['u_field'][j]
        # u_sde_mean was registered last: it follows the file; u_sde keeps its initial constants (p = amp, q = Ld/2)
        e_mean = u - O.u_sde_field(mesh, t['p'][j], t['q'][j], 1.0, False)
        e_sde = u - O.u_sde_field(mesh, 1.0, 20.0, 1.0, False)
        assert abs(d['u_error_with_sde_mean'][j] - np.sqrt(e_mean @ (M @ e_mean))) < 1e-12
        assert abs(d['u_error_with_sde'][j] - np.sqrt(e_sde @ (M @ e_sde))) < 1e-12


# ---- two ranks (gloo): seeds sharded, rank 0 writes ----------------------------------------------------------
def _rank(rank, world, port, cwd):
    import torch.distributed as dist
    os.chdir(cwd)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, HERE)
    from emu_engine import EmuEngine
    from peakondrake_b200 import experiment
    experiment('two', 40., 0.004, seeds=[3, 0, 7, 11, 5], diagnostics=DIAGS, ndump=2, allow_fail=True, ensemble_stats=True,
               engine_cls=EmuEngine, **KW)
    experiment('two_philox', 40., 0.004, seeds=range(5), diagnostics=['energy', 'u_field'], ndump=2, rng='philox',
               philox_seed=99, ensemble_stats=True, expected_u=True, engine_cls=EmuEngine, **KW)
    try:
        experiment('two', 40., 0.004, seeds=[1, 2], diagnostics=['energy'], engine_cls=EmuEngine, **KW)
        raise SystemExit(3)
    except IOError:
        pass                                         # every rank sees the existing directory (experiment.py:26-27)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_write_the_same_file_as_one(tmp_path, monkeypatch, built):
    from emu_engine import EmuEngine
    from emu_util import lib
    from peakondrake_b200 import experiment, outputting as out
    lib()
    monkeypatch.chdir(tmp_path)
    os.makedirs('one'); os.makedirs('par')
    os.chdir('one')
    experiment('two', 40., 0.004, seeds=[3, 0, 7, 11, 5], diagnostics=DIAGS, ndump=2, allow_fail=True, ensemble_stats=True,
               engine_cls=EmuEngine, **KW)
    mean1 = experiment('two_philox', 40., 0.004, seeds=range(5), diagnostics=['energy', 'u_field'], ndump=2, rng='philox',
                       philox_seed=99, ensemble_stats=True, expected_u=True, engine_cls=EmuEngine, **KW)
    os.chdir('..')
    ctx = mp.get_context('spawn')
    port = 29500 + (os.getpid() % 2000) + 7
    procs = [ctx.Process(target=_rank, args=(r, 2, port, str(tmp_path / 'par'))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    for code in ('two', 'two_philox'):
        a = out.read_data_file('one/results/%s/data.nc' % code)
        b = out.read_data_file('par/results/%s/data.nc' % code)
        assert sorted(a) == sorted(b)
        for k in a:
            if k == 'wallclock_time':
                continue
            if k.startswith('ens_') and not k.startswith('ens_hist') and 'count' not in k and not k.endswith(('_min', '_max')):
                assert np.allclose(a[k], b[k], rtol=1e-13, atol=1e-15, equal_nan=True), k     # sums in a different order
            else:
                assert np.array_equal(a[k], b[k], equal_nan=True), k                           # per-realisation rows: identical
    assert np.allclose(np.load('par/results/two_philox/expected_u.npy'), mean1, rtol=0, atol=1e-15)
