"""
`experiment()` -- drop-in for the reference's sweep driver (peakondrake/experiment.py:10-182).

Same keyword arguments, same parameter plumbing (lists -> swept dimensions, scalars -> fixed), same
`results/<code>/data[_dtI].nc` schema and the same error behaviour (IOError for an existing results
directory :25-27, ValueError for more than 1000 simulations :164-166).  What changes is the
execution: where the reference runs one `simulation()` per parameter combination, serially, all
SEEDS of a combination are integrated here as one ensemble on the GPU (`run_batch`).

Two keyword arguments are additions: `rng` ('mt19937' = the reference's per-seed NumPy streams,
'philox' = in-kernel counter-based generator) and `device`.
Not carried over: `.pvd` field dumps (`fields_to_output`, `field_ndump`; accepted and validated,
fields can be requested through the 'u_field' diagnostic instead) and the broken `expected_u` path
(experiment.py:231 raises TypeError in the reference; here it returns the ensemble mean of u).
"""
from collections import OrderedDict
from itertools import product
from os import path, makedirs
import numpy as np

from peakondrake_b200.simulation import run_batch, check_diagnostics, PeriodicIntervalMesh
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
               fixed_dW=None, rng='mt19937', device=None):

    dirname = 'results/' + code
    if path.exists(dirname):
        raise IOError("results directory '%s' already exists" % dirname)
    makedirs(dirname)

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

    out_dims = ('time',) + tuple(variable.keys())
    mean_u = None

    for i, dt in enumerate(expmt['dt'][1]):
        file_name = dirname + ('/data_dt%d.nc' % i if len(expmt['dt'][1]) > 1 else '/data.nc')
        params['file_name'] = (file_name,)
        df = out.DataFile(file_name)
        df.createDimension('time', None)
        df.createVariable('time', float, ('time',))
        if 'resolution' in variable:
            res_list = variable['resolution'][1]
            df.createDimension('resolution', len(res_list))
            df.createVariable('deltax', float, ('resolution',))
            df.data['deltax'][:] = [Ld / r for r in res_list]
            df.createDimension('x', 100)
            df.createVariable('x', float, ('x',))
            df.data['x'][:] = np.linspace(0, Ld, num=100, endpoint=False)
            params['store_coordinates'] = [False]
        else:
            npts = int(params['resolution'][-1])
            df.createDimension('x', npts)
            df.createVariable('x', float, ('x',))
            params['store_coordinates'] = [True]
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
        df.createDimension('wallclock_times', 2)
        df.createVariable('wallclock_time', float, ('wallclock_times',) + tuple(variable.keys()))
        df.createVariable('failed_time', float, tuple(variable.keys()))

        keys = list(variable.keys())
        if keys:
            sizes = [len(variable[k][1]) if variable[k] is not None else len(expmt[k][1]) for k in keys]
            num_sims = int(np.prod(sizes))
            if num_sims > 1000:
                raise ValueError('You are asking to do %d simulations! Please reduce your loops.' % num_sims)
        # every combination of the swept parameters except the seed is one batch over all seeds
        loop_keys = [k for k in keys if k != 'seed']
        loop_vals = [expmt[k][1] for k in loop_keys]
        seed_vals = expmt['seed'][1]
        for combo in product(*[list(enumerate(v)) for v in loop_vals]):
            p = dict(params)
            combo_index = {}
            for k, (idx, val) in zip(loop_keys, combo):
                p[k] = (idx, val)
                combo_index[k] = idx
            if 'dt' not in variable:
                p['dt'] = (dt,)
            p['mesh'] = (PeriodicIntervalMesh(int(p['resolution'][-1]), Ld),)
            if p['store_coordinates'][-1] is False and 'u_field' in (diagnostics or []):
                raise ValueError("'u_field' needs a fixed resolution (the file's x dimension is fixed to 100 points)")

            def index_of(si, combo_index=combo_index):
                return tuple(si if k == 'seed' else combo_index[k] for k in keys)

            res = run_batch(p, seed_vals, list(range(len(seed_vals))), df, diagnostics, index_of,
                            rng=rng, device=device, return_state=expected_u,
                            fields_to_output=fields_to_output, dim_names=keys)
            if expected_u:
                mean_u = res['u'].mean(axis=0)
            df.save()
        df.save()
    if expected_u:
        np.save(dirname + '/expected_u.npy', mean_u)
        return mean_u
