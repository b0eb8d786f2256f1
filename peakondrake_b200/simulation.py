"""
`simulation()` -- drop-in for the reference's peakondrake/simulation.py:11-138, plus the batched
runner `run_batch()` that `experiment()` uses to integrate all seeds of a parameter combination as
ONE ensemble on the GPU (sharded over the GPUs of a torch.distributed job).

Loop semantics kept from the reference (simulation.py:81-135):
  * step count: `while t < tmax - 0.5*dt: t += dt`;
  * Brownian increments are drawn every step, before the solve, also after a failure (:96);
  * `allow_fail`: a SNES failure latches the realisation, its diagnostics become NaN and its
    `failed_time` is stored (:99-106, outputting.py:107-108,234-243); without `allow_fail` the failure
    raises ConvergenceError;
  * `only_peakons`: the PDE is skipped from the start (`failed=True`, :52,85); the dump at t=0 is still
    made with the initial fields (:88 calls dump_diagnostics(t) with failed=False);
  * diagnostics are dumped at t=0 and then every `ndump` steps (:88,114-125).

Every dump is ONE array assignment per diagnostic (the realisation axis of the batch is a slice of the
file's 'seed' dimension) and one record appended to `data.nc` (outputting.DataFile).
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
                      'p_pde', 'q_pde', 'm_max', 'E_0', 'E_1', 'E_2', 'E_3', 'u_error_with_sde', 'u_error_weak',
                      'max_du_smooth_loc', 'min_du_smooth_loc', 'u_error_with_sde_mean', 'u_error_weak_mean')
# names the reference accepts but whose branches are broken/dead at this snapshot (SURVEY.md App. B:
# undefined names in outputting.py:126 / diagnostic_equations.py:150,177,181) or belong to the KdV set-up
REFERENCE_ONLY = ('min_u', 'a', 'b', 'l2_kdv_1', 'l2_kdv_2', 'l2_kdv_3')
M_DIAGNOSTICS = ('l2_m', 'h1_m', 'mass_m', 'mass_m2', 'm_max', 'E_2')
SDE_DIAGNOSTICS = ('u_error_with_sde', 'u_error_weak')
SDE_MEAN_DIAGNOSTICS = ('u_error_with_sde_mean', 'u_error_weak_mean')
SMOOTH_DIAGNOSTICS = ('max_du_smooth_loc', 'min_du_smooth_loc')
# diagnostics column behind each scalar (for the ensemble statistics of a dump)
STAT_COLUMN = {'energy': _lib.D_ENERGY, 'mass_u': _lib.D_MASS_U, 'max_du': _lib.D_MAX_DU, 'min_du': _lib.D_MIN_DU,
               'max_jump_local': _lib.D_MAX_JUMP_LOCAL, 'mass_m': _lib.D_MASS_M, 'mass_m2': _lib.D_L2M2,
               'm_max': _lib.D_M_MAX}


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


SUPPORTED_FIELDS = ('du', 'du_smooth', 'jump_du', 'm', 'u_xx', 'u_sde', 'u_sde_weak', 'uscalar', 'Xiscalar')


def diagnostic_columns(diagnostics, vals, idx, failed, x_cells, alphasq, pfac, extra):
    """
    The per-realisation values of one dump as arrays: OrderedDict name -> [R] (or [R, ne] for u_field), computed
    from the device tables `vals` [R, NDIAG], `idx` [R, NIDX] (outputting.py:106-196).  `extra` carries what other
    kernels produced (mu, peakon suite, smooth locations, the *_mean error, u).  Failed rows are NaN (:107-108).
    """
    cols = OrderedDict()
    v = lambda c: vals[:, c]
    for d in diagnostics:
        if d == 'mu':
            mu = extra['mu']
            for j in range(4):
                cols['mu_%d' % j] = mu[:, j]
                cols['alt_mu_%d' % j] = mu[:, 4 + j]
            continue
        if d == 'peakon_suite':
            for n, a in extra['peakon_suite'].items():
                cols[n] = a
            continue
        if d == 'u_field':
            cols['u_field'] = extra['u']
            continue
        if d == 'energy':
            o = v(_lib.D_ENERGY)
        elif d == 'l2_u':
            o = np.sqrt(v(_lib.D_L2U2))
        elif d == 'h1_u':
            o = np.sqrt(v(_lib.D_L2U2) + v(_lib.D_H1SEMI2))
        elif d == 'l2_m':
            o = np.sqrt(v(_lib.D_L2M2))
        elif d == 'h1_m':
            o = np.sqrt(v(_lib.D_L2M2) + v(_lib.D_H1SEMI_M2))
        elif d == 'mass_m':
            o = v(_lib.D_MASS_M)
        elif d == 'mass_m2':
            o = v(_lib.D_L2M2)
        elif d == 'mass_u':
            o = v(_lib.D_MASS_U)
        elif d == 'max_jump_local':
            o = v(_lib.D_MAX_JUMP_LOCAL)
        elif d == 'max_jump_global':
            o = v(_lib.D_MAX_DU) - v(_lib.D_MIN_DU)
        elif d == 'max_du_loc':
            o = x_cells[np.maximum(idx[:, _lib.I_ARGMAX_DU], 0)]
        elif d == 'min_du_loc':
            o = x_cells[np.maximum(idx[:, _lib.I_ARGMIN_DU], 0)]
        elif d == 'max_du':
            o = v(_lib.D_MAX_DU)
        elif d == 'min_du':
            o = v(_lib.D_MIN_DU)
        elif d == 'p_pde':
            o = v(_lib.D_MAX_U) * pfac
        elif d == 'q_pde':
            o = x_cells[np.maximum(idx[:, _lib.I_ARGMAX_U], 0)]
        elif d == 'm_max':
            o = v(_lib.D_M_MAX)
        elif d in ('E_0', 'E_1'):
            o = 0.5 * (v(_lib.D_L2U2) + alphasq * v(_lib.D_H1SEMI2))
        elif d == 'E_2':
            o = 0.5 * v(_lib.D_UMU)
        elif d == 'E_3':
            # the reference's u_xx solves  phi u_xx + phi_x (u_xx)_x = 0  (diagnostic_equations.py:249),
            # i.e. u_xx == 0, so E_3 = assemble(0.5 u (u + alphasq * 0) dx)  (outputting.py:175-178)
            o = 0.5 * v(_lib.D_L2U2)
        elif d in SDE_DIAGNOSTICS:
            o = np.sqrt(v(_lib.D_ERR_SDE2))
        elif d in SDE_MEAN_DIAGNOSTICS:
            o = np.sqrt(extra['err_mean2'])
        elif d == 'max_du_smooth_loc':
            o = extra['smooth'][:, 1]
        elif d == 'min_du_smooth_loc':
            o = extra['smooth'][:, 3]
        cols[d] = np.where(failed, np.nan, o)
    return cols


def peakon_suite_columns(x_cells, max_u, pk_vals, pk_ids, failed):
    """Vectorised outputting.peakon_suite_from_indices: the eight numbers of peakon_diagnostics
    (outputting.py:381-399,437-448) for every realisation, coordinates looked up, never recomputed."""
    L = np.max(x_cells) + np.min(x_cells)
    pc = pk_ids[:, 0]
    ok = ~failed
    pc = np.where(ok, pc, 0)
    xm, xp = x_cells[pc - 1], x_cells[pc]              # pc-1 == -1 wraps to the last cell, as in the reference
    peak_loc = np.where(xm > xp, 0.5 * (xm + xp - L), 0.5 * (xm + xp))
    min_du, max_du = pk_vals[:, 0], pk_vals[:, 1]
    x_min = np.where(pk_ids[:, 1] >= 0, x_cells[np.maximum(pk_ids[:, 1], 0)], np.nan)
    x_max = np.where(pk_ids[:, 2] >= 0, x_cells[np.maximum(pk_ids[:, 2], 0)], np.nan)
    with np.errstate(invalid='ignore', divide='ignore'):
        mu = np.where(x_max < x_min, x_min - x_max, L - x_min + x_max)
        nu = (max_du - min_du) / mu
    res = OrderedDict([('peakon_loc', peak_loc), ('peakon_min_du', min_du), ('peakon_max_du', max_du),
                       ('peakon_min_du_loc', x_min), ('peakon_max_du_loc', x_max), ('peakon_max_u', max_u),
                       ('peakon_mu', mu), ('peakon_nu', nu)])
    return OrderedDict((k, np.where(ok, a, np.nan)) for k, a in res.items())


ENSEMBLE_NBINS = 64


def ensemble_variables(diagnostics, fixed_resolution):
    """Names and trailing dimensions of the ensemble-statistics variables `experiment(..., ensemble_stats=True)`
    adds to data.nc: counts, per-x moments of u, the two histograms, and mean / variance / min / max
    of every requested scalar diagnostic that is a column of the device table."""
    names = OrderedDict([('ens_count', ()), ('ens_fail_count', ()),
                         ('ens_hist_q_pde', ('ens_bin',)), ('ens_hist_min_du', ('ens_bin',))])
    if fixed_resolution:
        names['ens_mean_u'] = ('x',)
        names['ens_var_u'] = ('x',)
    for d in diagnostics or []:
        if d in STAT_COLUMN:
            for s in ('mean', 'var', 'min', 'max'):
                names['ens_%s_%s' % (d, s)] = ()
    return names


def run_batch(params, seeds, seed_indices, data_file, diagnostics, combo_index=None, rng='mt19937', device=None,
              philox_seed=0, max_chunk=1000, return_state=False, fields_to_output=None, dim_names=(),
              comm=None, n_total=None, ensemble_stats=False, engine_cls=None):
    """
    Integrate `len(seeds)` realisations that share every parameter but the seed.

      params      : dict-of-tuples `simulation_parameters` (experiment.py:61-83); read with [-1]
      seeds       : seed values of THIS rank's realisations; seed_indices: their indices along the 'seed' dimension
      data_file   : outputting.DataFile (rank 0) or None
      dim_names, combo_index : names of the swept dimensions of the file's variables (after 'time') and the index of
                    this batch along each of them except 'seed' (the reference's index_slices, outputting.py:43-50)
      comm, n_total : ensemble.Comm of a multi-GPU run and the number of seeds over all ranks: realisation ids are
                    global (Philox counters, rows of the file), the per-realisation rows of a dump are gathered to
                    rank 0, the ensemble statistics are all-reduced
      fields_to_output : diagnostic fields of the .pvd dumps (written every field_ndump steps when field_ndump > 0,
                    outputting.py:59-68,224-232)
      rng         : 'mt19937' reproduces the reference's per-seed NumPy stream (host-drawn, uploaded);
                    'philox' draws in-kernel (throughput path)
      ensemble_stats : also write the ensemble statistics of every dump (ensemble_variables)
    """
    import torch
    from peakondrake_b200.ensemble import EnsembleStatistics
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
    true_mean_peakon_data = _val(params, 'true_mean_peakon_data')
    if peakon_flag not in (True, False):
        raise ValueError('peakon_equations must be True or False, not %s' % peakon_flag)   # simulation.py:68-69
    if smooth_t is not None and nupd > 1:
        raise ValueError('Prescribing forcing and including multiple Xi updates are not compatible.')
    allow_fail = True if only_peakons else allow_fail
    validate_scheme(scheme, timestepping, alphasq, gamma)
    check_diagnostics(diagnostics)
    diagnostics = list(diagnostics or [])

    R = len(seeds)
    n_total = R if n_total is None else int(n_total)
    seed_indices = np.asarray(seed_indices, dtype=np.int64)
    rid_offset = int(seed_indices[0]) if (comm is not None and R > 0) else 0
    root = comm is None or comm.rank == 0
    mesh = PeriodicIntervalMesh(ne, Ld)
    x_cells, x_nodes = mesh.x_cells(), mesh.x_nodes()
    u0 = build_initial_u(ne, Ld, ic, alphasq, float(_val(params, 'peak_width')))
    tables = build_tables(ne, Ld, fam, num_Xis, sigma, scheme)
    if scheme == 'upwind' and peakon_flag:
        raise NotImplementedError('peakon_equations with the upwind scheme is outside the B200 hot path')
    engine_cls = engine_cls or EnsembleIntegrator
    eng = engine_cls(ne, Ld, dt, R, alphasq=alphasq, scheme=scheme, timestepping=timestepping,
                     num_Xis=num_Xis, Xi_family=fam, sigma=sigma, nXi_updates=nupd,
                     seed=philox_seed, rid_offset=rid_offset, device=device, tables=tables)
    eng.set_state(u0)
    nsteps = num_steps(tmax, dt)

    peq = None
    if peakon_flag:
        peq = PeakonEquations(eng, u0, x_nodes, pure_xi_tables(ne, Ld, fam, num_Xis, sigma),
                              peakon_speed=_val(params, 'peakon_speed'), method=_val(params, 'peakon_method'))

    # ---- "true" peakon data (diagnostic_equations.py:45-72,373-378) ----
    true_pq = None
    if true_peakon_data is not None:
        tp = out.read_data_file('results/' + true_peakon_data + '/data.nc')
        expect = int(tmax / (ndump * dt)) + 1
        if len(tp['time']) != expect:                                      # diagnostic_equations.py:53-56
            raise ValueError('If reading in true peakon data, the dump frequency must be the same as that used '
                             'for the true peakon data. Length of true peakon data as %i, but proposed length is %i'
                             % (len(tp['time']), expect))
        # :57-59 wants one path, shape (expect,); the ensemble extension also takes one path PER SEED, shape
        # (expect, number of seeds), as an `only_peakons` experiment over the same seeds writes it
        if np.shape(tp['p']) not in ((expect,), (expect, n_total)):
            raise ValueError('True peakon data shape %s must be the same shape as proposed data %s'
                             % ((expect,), np.shape(tp['p'])))
        true_pq = (np.asarray(tp['p'], dtype=float), np.asarray(tp['q'], dtype=float))
    if true_mean_peakon_data is not None:                                  # :62-72 (the file is only checked, never read again)
        tm = out.read_data_file('results/' + true_mean_peakon_data + '/data.nc')
        expect = int(tmax / (ndump * dt))
        if len(tm['time']) != expect:
            raise ValueError('If reading in true peakon data, the dump frequency must be the same as that used for the true peakon data.')
        if np.shape(tm['p']) != (expect,):
            raise ValueError('True peakon data must have same shape as proposed data!')
    want_sde = any(d in SDE_DIAGNOSTICS for d in diagnostics) or any(f in ('u_sde', 'u_sde_weak') for f in fields_to_output)
    want_mean = any(d in SDE_MEAN_DIAGNOSTICS for d in diagnostics)
    if want_mean:
        # The reference's u_sde_mean branch (diagnostic_equations.py:291-303) creates fresh Constants p, q but takes Ld,
        # alphasq and the data file from the u_sde branch: update_constants (:373-378) raises AttributeError unless
        # `u_sde` is also being computed and true_peakon_data is given; periodic raises NotImplementedError (:296-297).
        if periodic:
            raise NotImplementedError('u_sde_mean not yet implemented for periodic peakon')
        if true_pq is None:
            raise AttributeError("'DiagnosticEquations' object has no attribute 'true_peakon_file'")
        if not want_sde:
            raise AttributeError("'DiagnosticEquations' object has no attribute 'Ld'")
    if want_sde and true_pq is None:
        raise ValueError('u_error_with_sde / u_error_weak need true_peakon_data')
    # which of the two expressions owns the Constants update_constants() assigns (the LAST branch to run in the order
    # DiagnosticVariables registers the fields: fields_to_output first, then the diagnostics' required fields,
    # simulation.py:269-362); the other one keeps its initial constants for ever
    mean_follows_file = False
    if want_mean:
        order = [f for f in fields_to_output if f in ('u_sde', 'u_sde_weak')]
        for d in diagnostics:
            order += {'u_error_with_sde': ['u_sde'], 'u_error_weak': ['u_sde', 'u_sde_weak'],
                      'u_error_with_sde_mean': ['u_sde_mean'], 'u_error_weak_mean': ['u_sde_mean', 'u_sde_weak_mean']}.get(d, [])
        first = OrderedDict((f, None) for f in order if f in ('u_sde', 'u_sde_mean'))
        mean_follows_file = list(first)[-1] == 'u_sde_mean'

    def true_at(t_idx, k=None):
        """(p, q) of the true peakon at dump index t_idx: [R, 2], or the pair of realisation k"""
        p_, q_ = true_pq
        if p_.ndim == 1:
            one = np.array([[p_[t_idx], q_[t_idx]]])
            return np.tile(one, (R, 1)) if k is None else one[0]
        rows = np.stack([p_[t_idx, seed_indices], q_[t_idx, seed_indices]], axis=1)
        return rows if k is None else rows[k]

    want_m = any(d in M_DIAGNOSTICS for d in diagnostics)
    want_smooth = any(d in SMOOTH_DIAGNOSTICS for d in diagnostics)
    pfac = (1 - np.exp(-Ld / np.sqrt(alphasq))) if periodic else 1.0
    amp = 0.5 * (1 + np.exp(-Ld / np.sqrt(alphasq))) / (1 - np.exp(-Ld / np.sqrt(alphasq)))

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

    # ---- where this batch lives in the file ----
    dim_names = tuple(dim_names)
    combo_index = dict(combo_index or {})
    all_idx = np.arange(n_total) if comm is not None else seed_indices

    def block(*lead):
        """index of the whole batch in a variable with dims lead + dim_names (+ trailing)"""
        ix = tuple(lead) + tuple(all_idx if d == 'seed' else combo_index[d] for d in dim_names)
        return ix

    def put_rows(name, t_idx, rows):
        """rows: [n, ...] per-realisation values in global order (rank 0)"""
        lead = () if t_idx is None else (t_idx,)
        if 'seed' in dim_names:
            data_file.put_block(name, block(*lead), rows)
        else:
            data_file.put(name, block(*lead) if (lead or dim_names) else (), rows[0])

    def gather(rows):
        return rows if comm is None else comm.gather_rows(rows, n_total)

    stats = EnsembleStatistics(eng, nbins=ENSEMBLE_NBINS, lo=0.0, hi=Ld) if ensemble_stats else None
    ens_names = ensemble_variables(diagnostics, bool(_val(params, 'store_coordinates'))) if ensemble_stats else {}

    wall0 = time.time()
    if data_file is not None and root:
        put_rows('wallclock_time', 0, np.full(n_total, wall0))
        if _val(params, 'store_coordinates'):
            data_file.data['x'][:] = x_cells                                # outputting.py:79-80

    def dump(t_idx, t, initial=False):
        # at t = 0 the reference dumps with failed=False whatever `only_peakons` says (simulation.py:88)
        fs_dev = None if initial else eng.fail_step
        pq = None
        if true_pq is not None and want_sde:
            if want_mean and mean_follows_file:
                pq = np.tile(np.array([[1.0, Ld / 2]]), (R, 1))            # u_sde keeps its initial Constants (:258-259)
            else:
                pq = true_at(t_idx)
        extra = {}
        if want_mean:
            # u_sde_mean: the non-periodic profile with its own Constants (diagnostic_equations.py:291-303)
            pm = true_at(t_idx) if mean_follows_file else np.tile(np.array([[1.0 / amp, Ld / 2]]), (R, 1))
            vm, _, _ = eng.diagnostics(pq=pm, periodic=False, fail_step=fs_dev)
            extra['err_mean2'] = vm[:, _lib.D_ERR_SDE2].cpu().numpy().copy()
        vals, idx, _ = eng.diagnostics(pq=pq, periodic=periodic, want_m=want_m, fail_step=fs_dev)
        if stats is not None:
            stats.update(comm=comm, diagnostics_done=True)
        failed = np.zeros(R, dtype=bool) if initial else (eng.fail_step.cpu().numpy() >= 0)
        vals, idx = vals.cpu().numpy(), idx.cpu().numpy()
        if 'peakon_suite' in diagnostics:
            pv, pi = eng.peakon_suite(fail_step=fs_dev)
            extra['peakon_suite'] = peakon_suite_columns(x_cells, vals[:, _lib.D_MAX_U], pv.cpu().numpy(),
                                                         pi.cpu().numpy(), failed)
        if want_smooth:
            sm, _, _ = eng.du_smooth_loc(fail_step=fs_dev)
            extra['smooth'] = sm.cpu().numpy()
        u_host = None
        if 'mu' in diagnostics:
            mv, ms = eng.mu(fail_step=fs_dev)
            mu_vals, mu_status = mv.cpu().numpy().copy(), ms.cpu().numpy()
            flagged = np.nonzero(~failed & ((mu_status == -1) | (mu_status == -2)))[0]
            if len(flagged):
                # more peaks than the device routine holds, or a field on which the reference itself raises:
                # the host restatement of find_mus decides (and raises like the reference)
                u_host = eng.u.cpu().numpy()
                for k in flagged:
                    du_k = (np.roll(u_host[k], -1) - u_host[k]) / (Ld / ne)
                    old, alt = out.mu_diagnostic(u_host[k], du_k, x_cells, Ld)
                    for j in range(4):
                        mu_vals[k, j] = old[j] if j < len(old) else np.nan
                        mu_vals[k, 4 + j] = alt[j] if j < len(old) else np.nan
            extra['mu'] = mu_vals
        if 'u_field' in diagnostics:
            u_host = eng.u.cpu().numpy() if u_host is None else u_host
            extra['u'] = np.where(failed[:, None], np.nan, u_host)
        cols = diagnostic_columns(diagnostics, vals, idx, failed, x_cells, alphasq, pfac, extra)
        if peq is not None:
            cols['p'] = peq.p
            cols['q'] = peq.q
        write_columns(t_idx, t, cols)
        return vals, idx

    def write_columns(t_idx, t, cols, sync=True):
        """one dump: gather the per-realisation columns to rank 0 and write the record (+ the ensemble statistics)"""
        if data_file is None and comm is None:
            return
        # one gather per dump: the scalar columns packed side by side; u_field on its own
        scal = [k for k in cols if k != 'u_field']
        packed = gather(np.stack([cols[k] for k in scal], axis=1)) if scal else None
        ufield = gather(cols['u_field']) if 'u_field' in cols else None
        if root and data_file is not None:
            data_file.put('time', t_idx, t)
            for j, k in enumerate(scal):
                put_rows(k, t_idx, packed[:, j])
            if ufield is not None:
                put_rows('u_field', t_idx, ufield)
            if stats is not None:
                ne_ = eng.ne
                pk = stats.packed.cpu().numpy()
                lead = (t_idx,) + tuple(combo_index[d] for d in dim_names if d != 'seed')
                cnt = pk[2 * ne_]
                put1 = lambda n, v: data_file.put(n, lead, v)
                put1('ens_count', cnt)
                put1('ens_fail_count', float(stats.fail_count.item()))
                put1('ens_hist_q_pde', pk[2 * ne_ + 1:2 * ne_ + 1 + stats.nbins])
                put1('ens_hist_min_du', pk[2 * ne_ + 1 + stats.nbins:])
                if 'ens_mean_u' in ens_names:
                    with np.errstate(invalid='ignore', divide='ignore'):
                        mean = pk[:ne_] / cnt
                        put1('ens_mean_u', mean)
                        put1('ens_var_u', pk[ne_:2 * ne_] / cnt - mean * mean)
                for d in diagnostics:
                    if d in STAT_COLUMN:
                        for s, val in zip(('mean', 'var', 'min', 'max'), stats.scalar_stats(STAT_COLUMN[d])):
                            put1('ens_%s_%s' % (d, s), val)
            if sync:
                data_file.sync()                                            # the record is on disk from here on

    # ---- field dumps (.pvd collection per realisation) ----
    field_files = None
    if field_ndump > 0:
        if ('u_sde' in fields_to_output or 'u_sde_weak' in fields_to_output) and true_pq is None:
            raise ValueError('Output fields u_sde / u_sde_weak need true_peakon_data')
        dirname = _val(params, 'dirname')
        field_files = [out.FieldFile(out.field_file_name(dirname, dim_names,
                                                         tuple(int(si) if d == 'seed' else combo_index[d] for d in dim_names)),
                                     ne, Ld) for si in seed_indices]
    last_dW = [None]

    def dump_fields(t, t_idx_now):
        """prognostic fields in the order PrognosticVariables registers them (simulation.py:154-249), then the
        requested diagnostic fields (outputting.py:66-68); nothing once a realisation has failed (simulation.py:129)"""
        u_host = eng.u.cpu().numpy()
        m_host = eng.m.cpu().numpy() if eng.upwind else None
        failed = eng.fail_step.cpu().numpy() >= 0
        smooth = eng.du_smooth_loc(want_field=True)[2].cpu().numpy() if 'du_smooth' in fields_to_output else None
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
                elif fld == 'du_smooth':
                    # CG5 field sampled at the two ends of every cell (the output space is DG1, outputting.FieldFile)
                    f[fld] = (np.stack([smooth[k, 0::5], np.roll(smooth[k, 0::5], -1)], axis=1).ravel(), 'DG1')
                elif fld == 'jump_du':
                    f[fld] = (np.abs(du - np.roll(du, 1)), 'CG1')                # :103-113
                elif fld == 'm':
                    f[fld] = (out.m_from_u(u_host[k], Ld, alphasq), 'CG1')       # :236-243
                elif fld == 'u_xx':
                    f[fld] = (np.zeros(ne), 'CG1')                               # :245-252 (u_xx == 0)
                elif fld == 'uscalar':
                    f[fld] = (u_host[k], 'CG1')                                  # :79-80
                elif fld == 'Xiscalar':
                    f[fld] = (f['dXi'][0], 'CG1')                                # :81-82
                elif fld in ('u_sde', 'u_sde_weak'):
                    j = min(t_idx_now, len(true_pq[0]) - 1)
                    pj, qj = true_at(j, k)
                    us = out.u_sde_field(x_nodes, Ld, alphasq, pj, qj, periodic)   # :254-279
                    # u_sde_weak: M w = M (u - u_sde) (:281-289), i.e. w = u - u_sde nodally
                    f[fld] = (us if fld == 'u_sde' else u_host[k] - us, 'CG1')
            ff.write(f, t)

    result = None
    try:
        dump(0, 0.0, initial=True)
        if only_peakons:
            eng.fail_step.fill_(0)
        if field_files is not None:
            dump_fields(0.0, 0)                                                     # simulation.py:88-89
        t, done, t_idx = 0.0, 0, 1
        chunk = ndump if ndump > 0 else min(nsteps, max_chunk)
        if field_ndump > 0:
            chunk = gcd(chunk, field_ndump) if ndump > 0 else field_ndump
        if host_dW is not None:
            # bound the upload of one launch (imported increments travel as [steps, R, num_Xis] doubles)
            chunk = max(1, min(chunk, (1 << 26) // max(1, R * num_Xis)))
        pde_runs = not only_peakons
        # Several dumps per launch.  Preferred: state SNAPSHOTS (pk_step_snap: the kernel stores u at every dump step, ONE
        # batched streaming call computes the diagnostics of all dumps of the launch) -- a dump then costs the step kernel
        # ~1 % of a step; if the snapshots of even one dump exceed the memory budget (PK_SNAP_BYTES, default 8 GiB of the 180:
        # BASELINE config 4 on one GPU then advances 14 steps per launch instead of one): the
        # dump FUSED into the step kernel (pk_step_dump, no extra memory, ~11 % of a cluster-kernel step per dump).
        import os
        batch_ok = (pde_runs and ndump > 0 and not eng.upwind and field_files is None and stats is None
                    and not (want_mean or want_smooth) and 'mu' not in diagnostics)
        ksnap = int(os.environ.get('PK_SNAP_BYTES', 8 << 30)) // max(1, R * ne * 8) if hasattr(eng, 'step_snap') else 0
        fused_ok = (batch_ok and hasattr(eng, 'step_dump') and not (want_m or want_sde) and 'u_field' not in diagnostics)
        mode = 'snap' if (batch_ok and ksnap >= 1) else ('fused' if fused_ok else None)
        kmax = max(1, min(256, (1 << 24) // max(1, R * ndump * max(1, num_Xis)))) if host_dW is not None else 256
        kmax = max(1, min(kmax, (1 << 23) // max(1, R * 23)))
        if mode == 'snap':
            kmax = max(1, min(kmax, ksnap))
        want_suite = 'peakon_suite' in diagnostics
        while mode is not None and nsteps - done >= ndump:
            k = min(kmax, (nsteps - done) // ndump)
            c = k * ndump
            dW_dev = None
            if host_dW is not None:
                dW_dev = torch.from_numpy(np.ascontiguousarray(host_dW[done:done + c])).to(eng.device)
            exp_dW = (peq is not None and dW_dev is None and num_Xis > 0)
            snap_u = None
            if mode == 'snap':
                pq_blk = np.stack([true_at(t_idx + j) for j in range(k)]) if (true_pq is not None and want_sde) else None
                exported, dg, di, sv, si, snap_u, snap_flag = eng.step_snap(c, ndump, dW=dW_dev, export_dW=exp_dW, pq=pq_blk,
                                                                            periodic=periodic, want_m=want_m, suite=want_suite)
            else:
                exported, dg, di, sv, si = eng.step_dump(c, ndump, dW=dW_dev, export_dW=exp_dW, suite=want_suite)
            traj = None
            if peq is not None:
                traj = peq.update(c, dW=dW_dev if dW_dev is not None else exported, step0=done, traj_stride=ndump).cpu().numpy()
            dg, di = dg.cpu().numpy(), di.cpu().numpy()
            if want_suite:
                sv, si = sv.cpu().numpy(), si.cpu().numpy()
            u_snaps = snap_u.cpu().numpy() if (snap_u is not None and 'u_field' in diagnostics) else None
            fs = eng.fail_step.cpu().numpy()
            for j in range(k):
                for _ in range(ndump):
                    t += dt
                failed = di[j, :, 0] < 0
                if not allow_fail and failed.any():
                    break
                extra = {}
                if want_suite:
                    extra['peakon_suite'] = peakon_suite_columns(x_cells, dg[j][:, _lib.D_MAX_U], sv[j], si[j], failed)
                if u_snaps is not None:
                    extra['u'] = np.where(failed[:, None], np.nan, u_snaps[j])
                cols = diagnostic_columns(diagnostics, dg[j], di[j], failed, x_cells, alphasq, pfac, extra)
                if traj is not None:
                    cols['p'] = traj[j, :, 0]
                    cols['q'] = traj[j, :, 1]
                write_columns(t_idx, t, cols, sync=(j == k - 1))
                t_idx += 1
            done += c
            if not allow_fail:
                bad = (fs >= 0).any()
                if comm is not None:
                    bad = comm.any(bad)
                if bad:
                    kk = int(np.argmax(fs >= 0)) if (fs >= 0).any() else -1
                    raise ConvergenceError('Nonlinear solve failed to converge for seed %s at step %s'
                                           % (seeds[kk] if kk >= 0 else '(other rank)', fs[kk] if kk >= 0 else '?'))
        while done < nsteps:
            c = min(chunk, nsteps - done)
            dW_dev = None
            if host_dW is not None:
                dW_dev = torch.from_numpy(np.ascontiguousarray(host_dW[done:done + c])).to(eng.device)
            need_dW = (peq is not None or field_files is not None) and num_Xis > 0
            exported = None
            if pde_runs:
                exported = eng.step(c, dW=dW_dev, export_dW=(need_dW and dW_dev is None))
            used = dW_dev if dW_dev is not None else exported
            if peq is not None:
                # the SAME increments as the PDE (peakon_equations.py:15): imported, exported by the step kernel, or --
                # when no PDE runs -- the identical in-kernel Philox stream of these global steps
                peq.update(c, dW=used, step0=done)
            if field_files is not None and num_Xis > 0 and used is not None:
                last_dW[0] = used[-1].cpu().numpy()       # (no PDE -> no field dumps after t = 0, simulation.py:128-131)
            done += c
            for _ in range(c):
                t += dt
            if not allow_fail:
                fs = eng.fail_step.cpu().numpy()
                bad = (fs >= 0).any()
                if comm is not None:
                    bad = comm.any(bad)
                if bad:
                    k = int(np.argmax(fs >= 0)) if (fs >= 0).any() else -1
                    raise ConvergenceError('Nonlinear solve failed to converge for seed %s at step %s'
                                           % (seeds[k] if k >= 0 else '(other rank)', fs[k] if k >= 0 else '?'))
            if ndump > 0 and done % ndump == 0:
                dump(t_idx, t)
                t_idx += 1
            if field_files is not None and done % field_ndump == 0:
                # p, q of u_sde are read at the diagnostics index current at that moment (diagnostic_equations.py:373-378)
                dump_fields(t, t_idx - 1 if (ndump > 0 and done % ndump == 0) else t_idx)   # simulation.py:128-131

        fs = eng.fail_step.cpu().numpy()
        ft = np.where((fs < 0) | only_peakons, np.nan, (fs + 1) * dt)               # simulation.py:105
        ft_all = gather(ft)
        if data_file is not None and root:
            put_rows('wallclock_time', 1, np.full(n_total, time.time()))
            put_rows('failed_time', None, ft_all)
        result = dict(fail_step=fs, its_total=eng.its_total.cpu().numpy(), nsteps=nsteps)
        if return_state:
            result['u'] = eng.u.cpu().numpy()
    finally:
        if data_file is not None and root:
            data_file.sync()          # whatever was dumped before an exception stays on disk, like the reference's 'a' file
        eng.close()
    return result


def simulation(simulation_parameters, diagnostic_values=None, fields_to_output=None, expected_u=False,
               data_file=None, rng='mt19937', device=None):
    """
    One realisation, same argument meaning as the reference's simulation() (simulation.py:11-14).
    `simulation_parameters` is the dict of tuples experiment() builds; swept parameters are
    (index, value) pairs.  Results are appended to `simulation_parameters['file_name']` (outputting.py:29-32; the
    file experiment() created -- pass its DataFile as `data_file` to keep it open across calls).
    Returns u (NumPy array, natural node order) if `expected_u`.
    """
    for key in ['alphasq', 'c0', 'gamma']:
        if len(simulation_parameters[key]) not in (1, 2):
            raise ValueError('The value of simulation_parameters %s seems to be of length %d. It should be either 1 or 2.'
                             % (key, len(simulation_parameters[key])))
    dims = []
    own_file = False
    if diagnostic_values:
        if data_file is None:
            # the reference opens simulation_parameters['file_name'] in append mode (outputting.py:29-32) and fails if
            # experiment() has not created it
            fn = simulation_parameters['file_name'][-1]
            data_file = out.DataFile.reopen(fn)                           # raises IOError/FileNotFoundError like Dataset(fn, 'a')
            own_file = True
        first = diagnostic_values if isinstance(diagnostic_values, str) else diagnostic_values[0]
        first = 'mu_0' if first == 'mu' else ('peakon_loc' if first == 'peakon_suite' else first)
        dims = [d for d in data_file.vars[first][1] if d not in ('time', 'x')]
    if isinstance(diagnostic_values, str):
        diagnostic_values = [diagnostic_values]
    combo = {d: simulation_parameters[d][0] for d in dims}
    seed = simulation_parameters['seed'][-1]
    seed_index = combo.get('seed', 0)
    try:
        res = run_batch(simulation_parameters, [seed], [seed_index], data_file, diagnostic_values,
                        combo_index=combo, rng=rng, device=device, return_state=expected_u,
                        fields_to_output=fields_to_output, dim_names=dims)
    finally:
        if data_file is not None:
            data_file.sync()
            if own_file:
                data_file.close()
    if expected_u:
        return res['u'][0]
