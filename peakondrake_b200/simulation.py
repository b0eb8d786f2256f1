"""
`simulation()` -- drop-in for the reference's peakondrake/simulation.py:11-138, plus the batched
runner `run_batch()` that `experiment()` uses to integrate all seeds of a parameter combination as
ONE ensemble on the GPU.

Loop semantics kept from the reference (simulation.py:81-135):
  * step count: `while t < tmax - 0.5*dt: t += dt`;
  * Brownian increments are drawn every step, before the solve, also after a failure (:96);
  * `allow_fail`: a SNES failure latches the realisation, its diagnostics become NaN and its
    `failed_time` is stored (:99-106, outputting.py:107-108,234-243); without `allow_fail` the failure
    raises ConvergenceError;
  * `only_peakons`: the PDE is skipped from the start (`failed=True`, :52,85);
  * diagnostics are dumped at t=0 and then every `ndump` steps (:88,114-125).
"""
import time
from math import gcd
from collections import OrderedDict
import numpy as np

from peakondrake_b200.engine import EnsembleIntegrator, validate_scheme
from peakondrake_b200.initial_conditions import build_initial_u
from peakondrake_b200.stochastic_functions import build_tables, draw_dW
from peakondrake_b200.peakon_equations import PeakonEquations, pure_xi_tables
from peakondrake_b200 import outputting as out
from peakondrake_b200 import _lib


class ConvergenceError(Exception):
    """Raised where the reference lets Firedrake's ConvergenceError propagate (simulation.py:108)."""


class PeriodicIntervalMesh(object):
    """Stand-in for firedrake.PeriodicIntervalMesh(ne, Ld): carries what this path needs."""

    def __init__(self, ne, Ld):
        self.ne, self.Ld = int(ne), float(Ld)

    def x_nodes(self):
        return np.arange(self.ne) * (self.Ld / self.ne)

    def x_cells(self):
        return (np.arange(self.ne) + 0.5) * (self.Ld / self.ne)


SCALAR_DIAGNOSTICS = ('energy', 'l2_m', 'l2_u', 'h1_u', 'h1_m', 'mass_m', 'mass_m2', 'mass_u',
                      'max_jump_local', 'max_jump_global', 'max_du_loc', 'min_du_loc', 'max_du', 'min_du',
                      'p_pde', 'q_pde', 'm_max', 'E_0', 'E_1', 'E_2', 'u_error_with_sde', 'u_error_weak')
# names the reference accepts but whose branches are broken/dead at this snapshot
# or need CG5 / KdV machinery outside the hot path
REFERENCE_ONLY = ('min_u', 'max_du_smooth_loc', 'min_du_smooth_loc', 'a', 'b', 'l2_kdv_1', 'l2_kdv_2',
                  'l2_kdv_3', 'E_3', 'u_error_with_sde_mean', 'u_error_weak_mean')
M_DIAGNOSTICS = ('l2_m', 'h1_m', 'mass_m', 'mass_m2', 'm_max', 'E_2')
SDE_DIAGNOSTICS = ('u_error_with_sde', 'u_error_weak')


def check_diagnostics(diagnostics):
    for d in diagnostics or []:
        if d in SCALAR_DIAGNOSTICS or d in ('mu', 'u_field', 'peakon_suite'):
            continue
        if d in REFERENCE_ONLY:
            raise NotImplementedError('Diagnostic %s is not part of the B200 hot path (DESIGN.md, out of scope)' % d)
        raise ValueError('Diagnostic %s not recgonised.' % d)      # outputting.py:196


def num_steps(tmax, dt):
    t, n = 0.0, 0
    while t < tmax - 0.5 * dt:
        t += dt
        n += 1
    return n


def _val(p, key):
    return p[key][-1]


SUPPORTED_FIELDS = ('du', 'jump_du', 'm', 'u_sde', 'u_sde_weak', 'uscalar', 'Xiscalar')


def run_batch(params, seeds, seed_indices, data_file, diagnostics, index_of, rng='mt19937', device=None,
              philox_seed=0, max_chunk=1000, return_state=False, fields_to_output=None, dim_names=()):
    """
    Integrate `len(seeds)` realisations that share every parameter but the seed.

      params      : dict-of-tuples `simulation_parameters` (experiment.py:61-83); read with [-1]
      seeds       : list of seed values; seed_indices: their indices along the 'seed' dimension
      data_file   : outputting.DataFile (or None); index_of(seed_index) -> tuple of indices of the swept
                    dimensions for that realisation (the reference's index_slices, outputting.py:43-50)
      fields_to_output, dim_names : diagnostic fields of the .pvd dumps (written every field_ndump steps when
                    field_ndump > 0, outputting.py:59-68,224-232) and the names of the swept dimensions that
                    index_of() addresses (they go into the file name)
      rng         : 'mt19937' reproduces the reference's per-seed NumPy stream (host-drawn, uploaded);
                    'philox' draws in-kernel (throughput path)
    """
    ne = int(_val(params, 'resolution'))
    Ld, dt, tmax = float(_val(params, 'Ld')), float(_val(params, 'dt')), float(_val(params, 'tmax'))
    scheme, timestepping = _val(params, 'scheme'), _val(params, 'timestepping')
    alphasq, c0, gamma = float(_val(params, 'alphasq')), float(_val(params, 'c0')), float(_val(params, 'gamma'))
    ic, num_Xis, fam = _val(params, 'ic'), int(_val(params, 'num_Xis')), _val(params, 'Xi_family')
    sigma = float(_val(params, 'sigma')) if num_Xis != 0 else 0.0
    ndump = int(_val(params, 'ndump'))
    field_ndump = int(_val(params, 'field_ndump')) if 'field_ndump' in params else -1
    fields_to_output = list(fields_to_output or [])
    for fld in fields_to_output:
        if fld in ('uscalar', 'Xiscalar') and scheme != 'upwind':
            raise ValueError('Output field %s only usable with upwind scheme, not %s' % (fld, scheme))   # simulation.py:272-276
        if field_ndump > 0 and fld not in SUPPORTED_FIELDS:
            raise NotImplementedError('Output field %s is not part of the B200 hot path (DESIGN.md, out of scope)' % fld)
        if field_ndump > 0 and fld == 'm' and scheme == 'upwind':
            raise ValueError("'m' is a prognostic field of the upwind scheme and is always dumped")
    nupd = int(_val(params, 'nXi_updates'))
    allow_fail = bool(_val(params, 'allow_fail'))
    only_peakons = bool(_val(params, 'only_peakons'))
    periodic = bool(_val(params, 'periodic'))
    smooth_t, fixed_dW = _val(params, 'smooth_t'), _val(params, 'fixed_dW')
    peakon_flag = _val(params, 'peakon_equations')
    true_peakon_data = _val(params, 'true_peakon_data')
    if _val(params, 'true_mean_peakon_data') is not None:
        raise NotImplementedError('true_mean_peakon_data (u_sde_mean) is outside the B200 hot path')
    if peakon_flag not in (True, False):
        raise ValueError('peakon_equations must be True or False, not %s' % peakon_flag)   # simulation.py:68-69
    if smooth_t is not None and nupd > 1:
        raise ValueError('Prescribing forcing and including multiple Xi updates are not compatible.')
    allow_fail = True if only_peakons else allow_fail
    validate_scheme(scheme, timestepping, alphasq, gamma)
    check_diagnostics(diagnostics)
    diagnostics = list(diagnostics or [])

    R = len(seeds)
    mesh = PeriodicIntervalMesh(ne, Ld)
    x_cells, x_nodes = mesh.x_cells(), mesh.x_nodes()
    u0 = build_initial_u(ne, Ld, ic, alphasq, float(_val(params, 'peak_width')))
    tables = build_tables(ne, Ld, fam, num_Xis, sigma, scheme)
    if scheme == 'upwind' and peakon_flag:
        raise NotImplementedError('peakon_equations with the upwind scheme is outside the B200 hot path')
    eng = EnsembleIntegrator(ne, Ld, dt, R, alphasq=alphasq, scheme=scheme, timestepping=timestepping,
                             num_Xis=num_Xis, Xi_family=fam, sigma=sigma, nXi_updates=nupd,
                             seed=philox_seed, device=device, tables=tables)
    eng.set_state(u0)
    nsteps = num_steps(tmax, dt)

    peq = None
    if peakon_flag:
        peq = PeakonEquations(u0, x_nodes, Ld, dt, alphasq, pure_xi_tables(ne, Ld, fam, num_Xis, sigma), R,
                              peakon_speed=_val(params, 'peakon_speed'), method=_val(params, 'peakon_method'))
    true_pq = None
    if true_peakon_data is not None:
        tp = out.read_data_file('results/' + true_peakon_data + '/data.nc')
        expect = int(tmax / (ndump * dt)) + 1
        if len(tp['time']) != expect:                                      # diagnostic_equations.py:53-56
            raise ValueError('If reading in true peakon data, the dump frequency must be the same as that used '
                             'for the true peakon data. Length of true peakon data as %i, but proposed length is %i'
                             % (len(tp['time']), expect))
        true_pq = (np.asarray(tp['p'], dtype=float), np.asarray(tp['q'], dtype=float))
    if any(d in SDE_DIAGNOSTICS for d in diagnostics) and true_pq is None:
        raise ValueError('u_error_with_sde / u_error_weak need true_peakon_data')
    want_m = any(d in M_DIAGNOSTICS for d in diagnostics)
    want_du = False       # `mu` runs on the device (pk_mu); the host routine is only the fallback for flagged realisations
    pfac = (1 - np.exp(-Ld / np.sqrt(alphasq))) if periodic else 1.0

    # Brownian increments for the whole run when reproducing the reference's streams
    host_dW = None
    if num_Xis > 0:
        if smooth_t is not None:
            host_dW = np.empty((nsteps, R, num_Xis))
            t = 0.0
            for n in range(nsteps):
                t += dt
                host_dW[n] = smooth_t(t)                                   # stochastic_functions.py:221-222
        elif fixed_dW is not None:
            host_dW = np.full((nsteps, R, num_Xis), fixed_dW * np.sqrt(dt))  # :223-224
        elif rng == 'mt19937':
            host_dW = np.stack([draw_dW(int(s), nsteps, num_Xis, dt, nupd) for s in seeds], axis=1)
        elif rng != 'philox':
            raise ValueError("rng must be 'mt19937' or 'philox'")
    if only_peakons:
        eng.fail_step.fill_(0)

    wall0 = time.time()
    if data_file is not None:
        for si in seed_indices:
            data_file.put('wallclock_time', (0,) + index_of(si), wall0)
        if _val(params, 'store_coordinates'):
            data_file.data['x'][:] = x_cells                                # outputting.py:79-80

    def dump(t_idx, t):
        pq = None
        if true_pq is not None:
            pq = np.tile(np.array([[true_pq[0][t_idx], true_pq[1][t_idx]]]), (R, 1))
        vals, idx, du = eng.diagnostics(pq=pq, periodic=periodic, want_m=want_m, want_du=want_du)
        pk_v = pk_i = None
        if 'peakon_suite' in diagnostics:
            pv, pi = eng.peakon_suite()
            pk_v, pk_i = pv.cpu().numpy(), pi.cpu().numpy()
        vals, idx = vals.cpu().numpy(), idx.cpu().numpy()
        mu_vals = mu_status = None
        if 'mu' in diagnostics:
            mv, ms = eng.mu()
            mu_vals, mu_status = mv.cpu().numpy(), ms.cpu().numpy()
        u_host = eng.u.cpu().numpy() if ('u_field' in diagnostics or
                                         (mu_status is not None and ((mu_status == -1) | (mu_status == -2)).any())) else None
        failed = eng.fail_step.cpu().numpy() >= 0
        if data_file is None:
            return vals, idx
        data_file.put('time', t_idx, t)
        for k, si in enumerate(seed_indices):
            ix = (t_idx,) + index_of(si)
            v, ii = vals[k], idx[k]
            for d in diagnostics:
                if d == 'mu':
                    if not failed[k] and mu_status[k] in (-1, -2):
                        # more peaks than the device routine holds, or a field on which the reference itself raises:
                        # the host restatement of find_mus decides (and raises like the reference)
                        du_k = (np.roll(u_host[k], -1) - u_host[k]) / (Ld / ne)
                        old, alt = out.mu_diagnostic(u_host[k], du_k, x_cells, Ld)
                        for j in range(4):
                            data_file.put('mu_%d' % j, ix, old[j] if j < len(old) else np.nan)
                            data_file.put('alt_mu_%d' % j, ix, alt[j] if j < len(old) else np.nan)
                        continue
                    for j in range(4):
                        data_file.put('mu_%d' % j, ix, mu_vals[k, j])
                        data_file.put('alt_mu_%d' % j, ix, mu_vals[k, 4 + j])
                    continue
                if d == 'peakon_suite':
                    if failed[k]:
                        res = {n: np.nan for n in out.PEAKON_DIAGNOSTICS}
                    else:
                        res = out.peakon_suite_from_indices(x_cells, v[_lib.D_MAX_U], pk_v[k], pk_i[k])
                    for n in out.PEAKON_DIAGNOSTICS:
                        data_file.put(n, ix, res[n])
                    continue
                if d == 'u_field':
                    data_file.put('u_field', ix, np.full(ne, np.nan) if failed[k] else u_host[k])
                    continue
                if failed[k]:
                    o = np.nan
                elif d == 'energy':
                    o = v[_lib.D_ENERGY]
                elif d == 'l2_u':
                    o = np.sqrt(v[_lib.D_L2U2])
                elif d == 'h1_u':
                    o = np.sqrt(v[_lib.D_L2U2] + v[_lib.D_H1SEMI2])
                elif d == 'l2_m':
                    o = np.sqrt(v[_lib.D_L2M2])
                elif d == 'h1_m':
                    o = np.sqrt(v[_lib.D_L2M2] + v[_lib.D_H1SEMI_M2])
                elif d == 'mass_m':
                    o = v[_lib.D_MASS_M]
                elif d == 'mass_m2':
                    o = v[_lib.D_L2M2]
                elif d == 'mass_u':
                    o = v[_lib.D_MASS_U]
                elif d == 'max_jump_local':
                    o = v[_lib.D_MAX_JUMP_LOCAL]
                elif d == 'max_jump_global':
                    o = v[_lib.D_MAX_DU] - v[_lib.D_MIN_DU]
                elif d == 'max_du_loc':
                    o = x_cells[ii[_lib.I_ARGMAX_DU]]
                elif d == 'min_du_loc':
                    o = x_cells[ii[_lib.I_ARGMIN_DU]]
                elif d == 'max_du':
                    o = v[_lib.D_MAX_DU]
                elif d == 'min_du':
                    o = v[_lib.D_MIN_DU]
                elif d == 'p_pde':
                    o = v[_lib.D_MAX_U] * pfac
                elif d == 'q_pde':
                    o = x_cells[ii[_lib.I_ARGMAX_U]]
                elif d == 'm_max':
                    o = v[_lib.D_M_MAX]
                elif d in ('E_0', 'E_1'):
                    o = 0.5 * (v[_lib.D_L2U2] + alphasq * v[_lib.D_H1SEMI2])
                elif d == 'E_2':
                    o = 0.5 * v[_lib.D_UMU]
                elif d in SDE_DIAGNOSTICS:
                    o = np.sqrt(v[_lib.D_ERR_SDE2])
                data_file.put(d, ix, o)
            if peq is not None:
                data_file.put('p', ix, peq.p[k])
                data_file.put('q', ix, peq.q[k])
        return vals, idx

    # ---- field dumps (.pvd collection per realisation) ----
    field_files = None
    if field_ndump > 0:
        if ('u_sde' in fields_to_output or 'u_sde_weak' in fields_to_output) and true_pq is None:
            raise ValueError('Output fields u_sde / u_sde_weak need true_peakon_data')
        dirname = _val(params, 'dirname')
        field_files = [out.FieldFile(out.field_file_name(dirname, dim_names, index_of(si)), ne, Ld)
                       for si in seed_indices]
    last_dW = [None]

    def dump_fields(t, t_idx_now):
        """prognostic fields in the order PrognosticVariables registers them (simulation.py:154-249), then the
        requested diagnostic fields (outputting.py:66-68); nothing once a realisation has failed (simulation.py:129)"""
        u_host = eng.u.cpu().numpy()
        m_host = eng.m.cpu().numpy() if eng.upwind else None
        failed = eng.fail_step.cpu().numpy() >= 0
        xi, xi_x, xi_xx = eng.tables_host
        for k, ff in enumerate(field_files):
            if failed[k] and t > 0:
                continue
            w = np.zeros(num_Xis) if last_dW[0] is None else last_dW[0][k]
            f = OrderedDict()
            f['u'] = (u_host[k], 'CG1')
            if scheme == 'upwind':
                f['m'] = (m_host[k], 'DG1')
            elif scheme == 'hydrodynamic':
                f['dXi_x'] = (w @ xi_x if num_Xis else np.zeros(ne), 'CG1')
                f['dXi_xx'] = (w @ xi_xx if num_Xis else np.zeros(ne), 'CG1')
            f['dXi'] = (w @ xi if num_Xis else np.zeros(ne), 'CG1')
            for fld in fields_to_output:
                du = (np.roll(u_host[k], -1) - u_host[k]) * (ne / Ld)            # diagnostic_equations.py:94-101
                if fld == 'du':
                    f[fld] = (du, 'DG0')
                elif fld == 'jump_du':
                    f[fld] = (np.abs(du - np.roll(du, 1)), 'CG1')                # :103-113
                elif fld == 'm':
                    f[fld] = (out.m_from_u(u_host[k], Ld, alphasq), 'CG1')       # :236-243
                elif fld == 'uscalar':
                    f[fld] = (u_host[k], 'CG1')                                  # :79-80
                elif fld == 'Xiscalar':
                    f[fld] = (f['dXi'][0], 'CG1')                                # :81-82
                elif fld in ('u_sde', 'u_sde_weak'):
                    j = min(t_idx_now, len(true_pq[0]) - 1)
                    us = out.u_sde_field(x_nodes, Ld, alphasq, true_pq[0][j], true_pq[1][j], periodic)   # :254-279
                    # u_sde_weak: M w = M (u - u_sde) (:281-289), i.e. w = u - u_sde nodally
                    f[fld] = (us if fld == 'u_sde' else u_host[k] - us, 'CG1')
            ff.write(f, t)

    dump(0, 0.0)
    if field_files is not None:
        dump_fields(0.0, 0)                                                     # simulation.py:88-89
    t, done, t_idx = 0.0, 0, 1
    chunk = ndump if ndump > 0 else min(nsteps, max_chunk)
    if field_ndump > 0:
        chunk = gcd(chunk, field_ndump) if ndump > 0 else field_ndump
    while done < nsteps:
        c = min(chunk, nsteps - done)
        dW_chunk = None if host_dW is None else host_dW[done:done + c]
        need_export = (peq is not None or field_files is not None) and host_dW is None and num_Xis > 0
        exported = eng.step(c, dW=dW_chunk, export_dW=need_export)
        used = None
        if peq is not None or field_files is not None:
            used = dW_chunk if dW_chunk is not None else (exported.cpu().numpy() if exported is not None
                                                          else np.zeros((c, R, 0)))
            last_dW[0] = used[-1]
        if peq is not None:
            for n in range(c):
                peq.update(used[n])
        done += c
        for _ in range(c):
            t += dt
        if not allow_fail:
            fs = eng.fail_step.cpu().numpy()
            if (fs >= 0).any():
                k = int(np.argmax(fs >= 0))
                raise ConvergenceError('Nonlinear solve failed to converge for seed %s at step %d'
                                       % (seeds[k], fs[k]))
        if ndump > 0 and done % ndump == 0:
            dump(t_idx, t)
            t_idx += 1
        if field_files is not None and done % field_ndump == 0:
            # p, q of u_sde are read at the diagnostics index current at that moment (diagnostic_equations.py:373-378)
            dump_fields(t, t_idx - 1 if (ndump > 0 and done % ndump == 0) else t_idx)   # simulation.py:128-131

    fs = eng.fail_step.cpu().numpy()
    if data_file is not None:
        wall1 = time.time()
        for k, si in enumerate(seed_indices):
            ix = index_of(si)
            data_file.put('wallclock_time', (1,) + ix, wall1)
            ft = np.nan if (fs[k] < 0 or only_peakons) else (fs[k] + 1) * dt      # simulation.py:105
            data_file.put('failed_time', ix if len(ix) else (0,), ft)
    result = dict(fail_step=fs, its_total=eng.its_total.cpu().numpy(), nsteps=nsteps)
    if return_state:
        result['u'] = eng.u.cpu().numpy()
    eng.close()
    return result


def simulation(simulation_parameters, diagnostic_values=None, fields_to_output=None, expected_u=False,
               data_file=None, rng='mt19937', device=None):
    """
    One realisation, same argument meaning as the reference's simulation() (simulation.py:11-14).
    `simulation_parameters` is the dict of tuples experiment() builds; swept parameters are
    (index, value) pairs.  Results go to `data_file` (an outputting.DataFile experiment() created).
    Returns u (NumPy array, natural node order) if `expected_u`.
    """
    for key in ['alphasq', 'c0', 'gamma']:
        if len(simulation_parameters[key]) not in (1, 2):
            raise ValueError('The value of simulation_parameters %s seems to be of length %d. It should be either 1 or 2.'
                             % (key, len(simulation_parameters[key])))
    swept = [k for k, v in simulation_parameters.items() if len(v) > 1 and k != 'mesh']
    dims = []
    if data_file is not None and diagnostic_values:
        first = diagnostic_values if isinstance(diagnostic_values, str) else diagnostic_values[0]
        first = 'mu_0' if first == 'mu' else ('peakon_loc' if first == 'peakon_suite' else first)
        dims = [d for d in data_file.vars[first][1] if d not in ('time', 'x')]
    fixed_index = tuple(simulation_parameters[d][0] for d in dims)
    seed = simulation_parameters['seed'][-1]
    res = run_batch(simulation_parameters, [seed], [0], data_file, diagnostic_values,
                    index_of=lambda si: fixed_index, rng=rng, device=device, return_state=expected_u,
                    fields_to_output=fields_to_output, dim_names=dims)
    if data_file is not None:
        data_file.save()
    if expected_u:
        return res['u'][0]
