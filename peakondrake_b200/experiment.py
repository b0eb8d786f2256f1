"""
`experiment()` -- drop-in for the reference's sweep driver (peakondrake/experiment.py:10-182).

Same keyword arguments, same parameter plumbing (lists -> swept dimensions, scalars -> fixed), same
`results/<code>/data[_dtI].nc` schema, `.pvd` field dumps (`fields_to_output`, `field_ndump`) and the same
error behaviour (IOError for an existing results directory :25-27, ValueError for too many simulations
:164-166).  What changes is the execution: where the reference runs one `simulation()` per parameter
combination, serially, all SEEDS of a combination are integrated here as one ensemble on the GPU
(`run_batch`), and in a torch.distributed job (one process per GPU) the seeds are sharded over the ranks.

Deviations, all additions:
  * the 1000-simulation cap (:164-166) counts the swept combinations WITHOUT the seed axis: the seeds are the
    ensemble axis of one batched launch, not 1000 separate runs, so `seeds=range(65536)` is legal;
  * `rng` ('mt19937' = the reference's per-seed NumPy streams, 'philox' = in-kernel counter-based generator),
    `device`, `ensemble_stats` (also write the ensemble moments / histograms / extrema / fail counts of every
    dump) and `philox_seed`;
  * multi-GPU: when torch.distributed is initialised with more than one rank, every rank calls experiment() with
    the same arguments; rank r integrates the contiguous block ensemble.partition(len(seeds), world, r) of the
    seeds, rank 0 creates `results/<code>` and writes the file (per-dump rows are gathered, statistics all-reduced);
  * the `expected_u` path is broken in the reference (experiment.py:231 raises TypeError); here it returns the
    ensemble mean of u and saves it as expected_u.npy.
"""
from collections import OrderedDict
from itertools import product
from os import path, makedirs
import numpy as np

from peakondrake_b200.simulation import (run_batch, check_diagnostics, PeriodicIntervalMesh, ensemble_variables,
                                         ENSEMBLE_NBINS)
from peakondrake_b200 import outputting as out

KNOWN_FIELDS = ('uscalar', 'Xiscalar', 'pure_xi', 'pure_xi_x', 'pure_xi_xx', 'pure_xi_xxx', 'pure_xi_xxxx',
                'Euscalar', 'du', 'du_smooth', 'jump_du', 'F', 'u2_flux', 'a', 'b', 'kdv_1', 'kdv_2', 'kdv_3',
                'm', 'u_xx', 'u_sde', 'u_sde_weak', 'u_sde_mean', 'u_sde_weak_mean')


def experiment(code, Ld, tmax, resolutions=[],
               dts=[], sigmas=[], seeds=[],
               schemes=[], timesteppings=[],
               alphasq=[], c0=[], gamma=[],
               ics=[], num_Xis=[], Xi_family=[],
               diagnostics=None, fields_to_output=None,
               ndump=-1, field_ndump=-1, nXi_updates=1, allow_fail=False,
               t_kick=[], sigma_kick=0.0, smooth_t=None,
               peakon_equations=False, only_peakons=False,
               true_peakon_data=None, true_mean_peakon_data=None,
               peakon_speed=None, expected_u=False, periodic=False,
               peak_width=1.0, peakon_method='ito_euler',
               fixed_dW=None, rng='mt19937', device=None, ensemble_stats=False, philox_seed=0,
               engine_cls=None):

    from peakondrake_b200.ensemble import Comm
    comm = Comm.world(device=None)
    if comm is not None and comm.backend == 'nccl':
        import torch
        comm.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
    root = comm is None or comm.rank == 0

    dirname = 'results/' + code
    exists = path.exists(dirname) if root else False
    if comm is not None:
        exists = comm.any(exists)
    if exists:
        raise IOError("results directory '%s' already exists" % dirname)
    if root:
        makedirs(dirname)
    if comm is not None:
        comm.barrier()

    for field in fields_to_output or []:
        if field not in KNOWN_FIELDS:
            raise ValueError('Output field %s not recognised.' % field)          # simulation.py:295
        if field in ('uscalar', 'Xiscalar'):
            for sch in ([schemes] if isinstance(schemes, str) else list(schemes)):
                if sch not in ('upwind', 'LASCH'):                                  # simulation.py:272-276
                    raise ValueError('Output field %s only usable with upwind scheme, not %s' % (field, sch))
    check_diagnostics(diagnostics)

    expmt = OrderedDict([('dt', (float, dts)), ('resolution', (float, resolutions)), ('seed', (int, seeds)),
                         ('sigma', (float, sigmas)), ('scheme', ('S1', schemes)),
                         ('timestepping', ('S1', timesteppings)), ('alphasq', (float, alphasq)),
                         ('c0', (float, c0)), ('gamma', (float, gamma)), ('ic', ('S1', ics)),
                         ('num_Xis', (int, num_Xis)), ('Xi_family', ('S1', Xi_family))])
    for key, (typ, val) in list(expmt.items()):
        if isinstance(val, str):
            expmt[key] = (typ, [val])
        else:
            try:
                expmt[key] = (typ, list(val))
            except TypeError:
                expmt[key] = (typ, [val])

    variable = OrderedDict()
    params = {}
    if len(expmt['resolution'][1]) > 1:
        variable['resolution'] = None            # resolution loops outermost (experiment.py:63-65)
    for key, (typ, vals) in expmt.items():
        if len(vals) > 1:
            variable[key] = (typ, vals)
            params[key] = None
        elif len(vals) == 1:
            params[key] = (vals[0],)
        else:
            raise ValueError('experiment: parameter %s has no value' % key)
    for key, value in zip(['dirname', 'Ld', 'tmax', 'ndump', 'field_ndump', 'nXi_updates', 'allow_fail', 'smooth_t',
                           'peakon_equations', 'only_peakons', 'true_peakon_data', 'true_mean_peakon_data',
                           'peakon_speed', 'periodic', 'fixed_dW', 'peak_width', 'peakon_method'],
                          [dirname, Ld, tmax, ndump, field_ndump, nXi_updates, allow_fail, smooth_t, peakon_equations,
                           only_peakons, true_peakon_data, true_mean_peakon_data, peakon_speed, periodic, fixed_dW,
                           peak_width, peakon_method]):
        params[key] = (value,)

    keys = list(variable.keys())
    if keys:
        # the reference caps the number of simulation() calls at 1000 (:160-166); here one call integrates all seeds
        # of a combination, so the cap applies to the combinations of the other swept parameters
        sizes = [len(variable[k][1]) for k in keys if k != 'seed']
        num_sims = int(np.prod(sizes)) if sizes else 1
        if num_sims > 1000:
            raise ValueError('You are asking to do %d simulations! Please reduce your loops.' % num_sims)
    seed_vals = expmt['seed'][1]
    n_seeds = len(seed_vals)
    if comm is not None and n_seeds < comm.size:
        raise ValueError('experiment: %d seeds cannot be sharded over %d ranks' % (n_seeds, comm.size))

    out_dims = ('time',) + tuple(keys)
    stat_dims = ('time',) + tuple(k for k in keys if k != 'seed')
    mean_u = None

    for i, dt in enumerate(expmt['dt'][1]):
        file_name = dirname + ('/data_dt%d.nc' % i if len(expmt['dt'][1]) > 1 else '/data.nc')
        params['file_name'] = (file_name,)
        df = None
        fixed_resolution = 'resolution' not in variable
        params['store_coordinates'] = [fixed_resolution]
        if root:
            df = out.DataFile(file_name)
            df.createDimension('time', None)
            df.createVariable('time', float, ('time',))
            if not fixed_resolution:
                res_list = variable['resolution'][1]
                df.createDimension('resolution', len(res_list))
                df.createVariable('deltax', float, ('resolution',))
                df.data['deltax'][:] = [Ld / r for r in res_list]
                df.createDimension('x', 100)
                df.createVariable('x', float, ('x',))
                df.data['x'][:] = np.linspace(0, Ld, num=100, endpoint=False)
            else:
                npts = int(params['resolution'][-1])
                df.createDimension('x', npts)
                df.createVariable('x', float, ('x',))
            for key, spec in variable.items():
                if key != 'resolution':
                    typ, vals = spec
                    df.createDimension(key, len(vals))
                    if typ == 'S1':
                        df.createVariable(key, 'S1', (key,))
                        df.data[key][:] = [str(v)[:1].encode() for v in vals]
                    else:
                        df.createVariable(key, typ, (key,))
                        df.data[key][:] = vals
            for name in diagnostics or []:
                if name == 'mu':
                    for j in range(4):
                        df.createVariable('mu_%d' % j, float, out_dims)
                        df.createVariable('alt_mu_%d' % j, float, out_dims)
                elif name in ('a', 'b', 'u_field'):
                    df.createVariable(name, float, out_dims + ('x',))
                elif name == 'peakon_suite':
                    for pd in out.PEAKON_DIAGNOSTICS:
                        df.createVariable(pd, float, out_dims)
                else:
                    df.createVariable(name, float, out_dims)
            if peakon_equations:
                df.createVariable('p', float, out_dims)
                df.createVariable('q', float, out_dims)
            if ensemble_stats:
                df.createDimension('ens_bin', ENSEMBLE_NBINS)
                for name, tail in ensemble_variables(diagnostics, fixed_resolution).items():
                    df.createVariable(name, float, stat_dims + tail)
            df.createDimension('wallclock_times', 2)
            df.createVariable('wallclock_time', float, ('wallclock_times',) + tuple(keys))
            df.createVariable('failed_time', float, tuple(keys))

        # every combination of the swept parameters except the seed is one batch over all seeds
        loop_keys = [k for k in keys if k != 'seed']
        loop_vals = [expmt[k][1] for k in loop_keys]
        if comm is not None:
            off, cnt = comm.shard(n_seeds)
        else:
            off, cnt = 0, n_seeds
        try:
            for combo in product(*[list(enumerate(v)) for v in loop_vals]):
                p = dict(params)
                combo_index = {}
                for k, (idx, val) in zip(loop_keys, combo):
                    p[k] = (idx, val)
                    combo_index[k] = idx
                if 'dt' not in variable:
                    p['dt'] = (dt,)
                p['mesh'] = (PeriodicIntervalMesh(int(p['resolution'][-1]), Ld),)
                if not fixed_resolution and 'u_field' in (diagnostics or []):
                    raise ValueError("'u_field' needs a fixed resolution (the file's x dimension is fixed to 100 points)")
                res = run_batch(p, seed_vals[off:off + cnt], list(range(off, off + cnt)), df, diagnostics,
                                combo_index=combo_index, rng=rng, device=device, return_state=expected_u,
                                fields_to_output=fields_to_output, dim_names=keys, comm=comm, n_total=n_seeds,
                                ensemble_stats=ensemble_stats, philox_seed=philox_seed, engine_cls=engine_cls)
                if expected_u:
                    su = res['u'].sum(axis=0)
                    if comm is not None:
                        import torch
                        t = torch.from_numpy(su).to(comm.device)
                        comm.allreduce(t)
                        su = t.cpu().numpy()
                    mean_u = su / n_seeds
        finally:
            if df is not None:
                df.close()
    if expected_u:
        if root:
            np.save(dirname + '/expected_u.npy', mean_u)
        return mean_u
