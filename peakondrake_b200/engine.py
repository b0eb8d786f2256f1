"""
EnsembleIntegrator -- the batched replacement of the reference's per-step operator protocol

    Xis.update(t); equations.solve()                       (simulation.py:96-108)
    diagnostic_equations.solve(); outputting.dump_diagnostics(t, failed)   (simulation.py:118-125)

for R independent realisations on one GPU.  PyTorch is used only to own device memory and the
CUDA stream; every numerical kernel is in libpeakon_b200.so (include/peakon_b200.h).
"""
import ctypes
import numpy as np
import torch

from peakondrake_b200 import _lib
from peakondrake_b200.stochastic_functions import build_tables, circulant_tridiag_solve

CG1_MIDPOINT_SCHEMES = ('hydrodynamic', 'no_gradient', 'test')


def validate_scheme(scheme, timestepping, alphasq=1.0, gamma=0.0):
    """The checks of Equations.__init__ (equations.py:26-28,63,97,147-167), same exceptions."""
    if alphasq > 0.0 and gamma == 0.0:
        if scheme in CG1_MIDPOINT_SCHEMES and timestepping == 'midpoint':
            return
        if scheme == 'upwind' and timestepping == 'ssprk3':
            return
        raise ValueError('Scheme %s and timestepping %s either not compatible or not recognised.'
                         % (scheme, timestepping))
    elif alphasq == 0.0 and gamma > 0.0:
        if scheme in ('upwind', 'hydrodynamic') and timestepping in ('ssprk3', 'midpoint') \
                and (scheme, timestepping) != ('hydrodynamic', 'ssprk3'):
            raise NotImplementedError('Scheme %s and timestepping %s not yet implemented.' % (scheme, timestepping))
        raise ValueError('Scheme %s and timestepping %s either not compatible or not recognised.'
                         % (scheme, timestepping))
    raise NotImplementedError('Schemes for your values of alpha squared %.3f and gamma %.3f are not yet implemented.'
                              % (alphasq, gamma))


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class EnsembleIntegrator(object):
    """
    R realisations of the stochastic Camassa-Holm equation on PeriodicIntervalMesh(ne, Ld).

    State (device, fp64, [R, ne], natural node order): u and the Newton unknown (u1, dF, dG) the
    reference carries as warm start (equations.py:71,89).
    """

    def __init__(self, ne, Ld, dt, R, alphasq=1.0, scheme='hydrodynamic', timestepping='midpoint',
                 num_Xis=0, Xi_family='sines', sigma=0.0, nXi_updates=1, seed=0, rid_offset=0,
                 device=None, tables=None, c0=0.0, gamma=0.0):
        validate_scheme(scheme, timestepping, alphasq, gamma)
        if not torch.cuda.is_available():
            raise RuntimeError('peakondrake_b200 needs a CUDA device (B200); there is no CPU fallback.')
        self.lib = _lib.load()
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        self.ne, self.Ld, self.dt, self.R = int(ne), float(Ld), float(dt), int(R)
        self.alphasq, self.scheme = float(alphasq), scheme
        self.upwind = (scheme == 'upwind')
        self.num_Xis, self.Xi_family, self.sigma = int(num_Xis), Xi_family, float(sigma)
        self.nXi_updates = int(nXi_updates)
        if tables is None:
            tables = build_tables(self.ne, self.Ld, Xi_family, self.num_Xis, self.sigma, scheme)
        xi, xi_x, xi_xx = [np.ascontiguousarray(t, dtype=np.float64) for t in tables]
        self.tables_host = (xi, xi_x, xi_xx)
        spatially_constant = self.num_Xis == 0 or (self.num_Xis == 1 and np.all(xi[0] == xi[0, 0])
                                                   and not xi_x.any() and not xi_xx.any())
        if spatially_constant:
            mode = _lib.PK_NOISE_SCALAR
        elif not xi_x.any() and not xi_xx.any():
            mode = _lib.PK_NOISE_X
        else:
            mode = _lib.PK_NOISE_FULL
        self.noise_mode = mode
        cfg = _lib.PkConfig(ne=self.ne, R=self.R, Ld=self.Ld, dt=self.dt, alphasq=self.alphasq,
                            noise_mode=mode, num_xis=self.num_Xis, nxi_updates=self.nXi_updates,
                            device=self.device.index, seed=int(seed), rid_offset=int(rid_offset), max_it=0,
                            rtol=0.0, atol=0.0, stol=0.0, divtol=0.0,
                            xi_const=float(xi[0, 0]) if (spatially_constant and self.num_Xis) else 0.0,
                            scheme=1 if self.upwind else (2 if scheme == 'test' else 0))
        ctx = ctypes.c_void_p()
        rc = self.lib.pk_create(ctypes.byref(cfg), ctypes.byref(ctx))
        if rc != 0:
            raise _lib.PkError(self.lib.pk_last_error(None).decode())
        self.ctx = ctx
        dev = self.device
        self.xi = torch.from_numpy(xi).to(dev)
        self.xi_x = torch.from_numpy(xi_x).to(dev)
        self.xi_xx = torch.from_numpy(xi_xx).to(dev)
        if self.num_Xis > 0 and mode != _lib.PK_NOISE_SCALAR:
            _lib.check(self.lib, ctx, self.lib.pk_set_tables(ctx, _ptr(self.xi), _ptr(self.xi_x), _ptr(self.xi_xx)))
        z = lambda: torch.zeros(self.R, self.ne, dtype=torch.float64, device=dev)
        self.u, self.u1, self.dF, self.dG = z(), z(), z(), z()
        # upwind scheme: prognostic DG1 m, (left-node, right-node) value per cell (simulation.py:154-160)
        self.m = torch.zeros(self.R, self.ne, 2, dtype=torch.float64, device=dev) if self.upwind else None
        self.fail_step = torch.full((self.R,), -1, dtype=torch.int32, device=dev)
        self.its_last = torch.zeros(self.R, dtype=torch.int32, device=dev)
        self.its_total = torch.zeros(self.R, dtype=torch.int32, device=dev)
        self.diag = torch.zeros(self.R, _lib.NDIAG, dtype=torch.float64, device=dev)
        self.didx = torch.zeros(self.R, _lib.NIDX, dtype=torch.int32, device=dev)

    # -- life cycle ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, 'ctx', None) is not None:
            self.lib.pk_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def layout(self):
        T, M, rem, smem = (ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32())
        c = self.lib.pk_get_layout(self.ctx, ctypes.byref(T), ctypes.byref(M), ctypes.byref(rem), ctypes.byref(smem))
        d = dict(threads=T.value, rows_per_thread=M.value, rem=rem.value, smem_bytes=smem.value,
                 ctas_per_realisation=max(1, c))
        g, nb, nblk, grid = (ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32())
        self.lib.pk_get_work_split(self.ctx, ctypes.byref(g), ctypes.byref(nb), ctypes.byref(nblk), ctypes.byref(grid))
        if g.value:     # persistent CTAs + work queue: the split of the last launch
            d.update(resident_ctas=g.value, step_blocks=nb.value, steps_per_block=nblk.value, ctas=grid.value)
        return d

    @property
    def launch_count(self):
        """kernels of libpeakon_b200 launched through this context so far"""
        return int(self.lib.pk_launch_count(self.ctx))

    @property
    def step_counter(self):
        return int(self.lib.pk_get_step_counter(self.ctx))

    # -- state --------------------------------------------------------------------------------
    def set_state(self, u0, w1=None):
        """u0: [ne] (broadcast over realisations) or [R, ne]; w1 = (u1, dF, dG) or None for the
        reference's zero start (equations.py:71)."""
        u0 = torch.as_tensor(np.asarray(u0), dtype=torch.float64)
        self.u.copy_(u0.to(self.device).expand(self.R, self.ne))
        if w1 is None:
            self.u1.zero_(); self.dF.zero_(); self.dG.zero_()
        else:
            for dst, src in zip((self.u1, self.dF, self.dG), w1):
                dst.copy_(torch.as_tensor(np.asarray(src), dtype=torch.float64).to(self.device).expand(self.R, self.ne))
        self.fail_step.fill_(-1); self.its_last.zero_(); self.its_total.zero_()
        self.lib.pk_set_step_counter(self.ctx, 0)
        if self.upwind:
            # initial m: M m_CG = (M + alphasq K) u in CG1, interpolated into DG1 (initial_conditions.py:65-77)
            h = self.Ld / self.ne
            uu = self.u.cpu().numpy()
            b = (h / 6.0 - self.alphasq / h) * (np.roll(uu, 1, axis=1) + np.roll(uu, -1, axis=1)) \
                + (2.0 * h / 3.0 + 2.0 * self.alphasq / h) * uu
            m_cg = circulant_tridiag_solve(h / 6.0, 2.0 * h / 3.0, b)
            m = np.stack([m_cg, np.roll(m_cg, -1, axis=1)], axis=2)
            self.m.copy_(torch.from_numpy(np.ascontiguousarray(m)))

    # -- the hot path -------------------------------------------------------------------------
    def step(self, nsteps, dW=None, export_dW=False):
        """
        Advance every realisation `nsteps` steps in ONE kernel launch.
        dW: None (draw in-kernel with Philox) or array/tensor [nsteps, R, num_Xis] of increments
        already scaled by sqrt(dt).  Returns the exported increments (device tensor) if asked.
        """
        dW_dev = None
        if dW is not None and self.num_Xis > 0:
            dW_dev = torch.as_tensor(dW, dtype=torch.float64).to(self.device).contiguous()
            if tuple(dW_dev.shape) != (nsteps, self.R, self.num_Xis):
                raise ValueError('dW must have shape (nsteps, R, num_Xis) = %s, got %s'
                                 % ((nsteps, self.R, self.num_Xis), tuple(dW_dev.shape)))
        out = None
        if export_dW and self.num_Xis > 0:
            out = torch.empty(nsteps, self.R, self.num_Xis, dtype=torch.float64, device=self.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        if self.upwind:
            rc = self.lib.pk_step_upwind(self.ctx, int(nsteps), _ptr(self.m), _ptr(self.u), _ptr(dW_dev), _ptr(out),
                                         _ptr(self.fail_step), stream)
        else:
            rc = self.lib.pk_step(self.ctx, int(nsteps), _ptr(self.u), _ptr(self.u1), _ptr(self.dF), _ptr(self.dG),
                                  _ptr(dW_dev), _ptr(out), _ptr(self.fail_step), _ptr(self.its_last),
                                  _ptr(self.its_total), stream)
        _lib.check(self.lib, self.ctx, rc)
        self._keep = (dW_dev, out)      # keep inputs alive until the stream has consumed them
        return out

    def step_dump(self, nsteps, dump_every, dW=None, export_dW=False, suite=False):
        """
        `nsteps` steps in ONE launch with the dump diagnostics fused into the step kernel (pk_step_dump): returns
        (exported dW or None, diag [nsteps//dump_every, R, NDIAG], idx [.., R, NIDX], suite_vals [.., R, 2] or None,
        suite_ids [.., R, 3] or None) as device tensors.  Midpoint schemes only; the m- and SDE-based columns are NaN.
        """
        if self.upwind:
            raise NotImplementedError('step_dump is provided for the midpoint schemes')
        dW_dev = None
        if dW is not None and self.num_Xis > 0:
            dW_dev = torch.as_tensor(dW, dtype=torch.float64).to(self.device).contiguous()
            if tuple(dW_dev.shape) != (nsteps, self.R, self.num_Xis):
                raise ValueError('dW must have shape (nsteps, R, num_Xis) = %s, got %s'
                                 % ((nsteps, self.R, self.num_Xis), tuple(dW_dev.shape)))
        out = None
        if export_dW and self.num_Xis > 0:
            out = torch.empty(nsteps, self.R, self.num_Xis, dtype=torch.float64, device=self.device)
        nd = nsteps // dump_every
        diag = torch.empty(nd, self.R, _lib.NDIAG, dtype=torch.float64, device=self.device)
        idx = torch.empty(nd, self.R, _lib.NIDX, dtype=torch.int32, device=self.device)
        sv = torch.empty(nd, self.R, 2, dtype=torch.float64, device=self.device) if suite else None
        si = torch.empty(nd, self.R, 3, dtype=torch.int32, device=self.device) if suite else None
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        rc = self.lib.pk_step_dump(self.ctx, int(nsteps), _ptr(self.u), _ptr(self.u1), _ptr(self.dF), _ptr(self.dG),
                                   _ptr(dW_dev), _ptr(out), _ptr(self.fail_step), _ptr(self.its_last),
                                   _ptr(self.its_total), int(dump_every), _ptr(diag), _ptr(idx), _ptr(sv), _ptr(si), stream)
        _lib.check(self.lib, self.ctx, rc)
        self._keep = (dW_dev, out)
        return out, diag, idx, sv, si

    def step_snap(self, nsteps, snap_every, dW=None, export_dW=False, pq=None, periodic=False, want_m=False, suite=False):
        """
        `nsteps` steps in ONE launch; after every `snap_every`-th step the kernel stores a snapshot of u, and the dump
        diagnostics of all snapshots are computed by one batched call (pk_step_snap + pk_diagnostics_rows
        [+ pk_peakon_suite_rows]).  pq: None or [nsnap, R, 2] true-peakon (p, q) per dump and realisation.
        Returns (exported dW or None, diag [nsnap, R, NDIAG], idx [nsnap, R, NIDX], suite_vals or None, suite_ids or
        None, snap_u [nsnap, R, ne], snap_flag [nsnap, R]) as device tensors.  Midpoint schemes only.
        """
        if self.upwind:
            raise NotImplementedError('step_snap is provided for the midpoint schemes')
        dW_dev = None
        if dW is not None and self.num_Xis > 0:
            dW_dev = torch.as_tensor(dW, dtype=torch.float64).to(self.device).contiguous()
            if tuple(dW_dev.shape) != (nsteps, self.R, self.num_Xis):
                raise ValueError('dW must have shape (nsteps, R, num_Xis) = %s, got %s'
                                 % ((nsteps, self.R, self.num_Xis), tuple(dW_dev.shape)))
        out = None
        if export_dW and self.num_Xis > 0:
            out = torch.empty(nsteps, self.R, self.num_Xis, dtype=torch.float64, device=self.device)
        ns = nsteps // snap_every
        dev = self.device
        snap_u = torch.empty(ns, self.R, self.ne, dtype=torch.float64, device=dev)
        snap_flag = torch.empty(ns, self.R, dtype=torch.int32, device=dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = self.lib.pk_step_snap(self.ctx, int(nsteps), _ptr(self.u), _ptr(self.u1), _ptr(self.dF), _ptr(self.dG),
                                   _ptr(dW_dev), _ptr(out), _ptr(self.fail_step), _ptr(self.its_last),
                                   _ptr(self.its_total), int(snap_every), _ptr(snap_u), _ptr(snap_flag), stream)
        _lib.check(self.lib, self.ctx, rc)
        diag = torch.empty(ns, self.R, _lib.NDIAG, dtype=torch.float64, device=dev)
        idx = torch.empty(ns, self.R, _lib.NIDX, dtype=torch.int32, device=dev)
        pq_dev = None if pq is None else torch.as_tensor(pq, dtype=torch.float64).to(dev).contiguous()
        rc = self.lib.pk_diagnostics_rows(self.ctx, _ptr(snap_u), ns * self.R, _ptr(pq_dev), int(bool(periodic)),
                                          int(bool(want_m)), _ptr(diag), _ptr(idx), None, _ptr(snap_flag), stream)
        _lib.check(self.lib, self.ctx, rc)
        sv = si = None
        if suite:
            sv = torch.empty(ns, self.R, 2, dtype=torch.float64, device=dev)
            si = torch.empty(ns, self.R, 3, dtype=torch.int32, device=dev)
            rc = self.lib.pk_peakon_suite_rows(self.ctx, _ptr(snap_u), ns * self.R, _ptr(idx), _ptr(snap_flag), _ptr(sv), _ptr(si), stream)
            _lib.check(self.lib, self.ctx, rc)
        self._keep = (dW_dev, out, pq_dev)
        return out, diag, idx, sv, si, snap_u, snap_flag

    def last_kernel_ms(self):
        return float(self.lib.pk_last_step_kernel_ms(self.ctx))

    def kernel_ms_since(self):
        """(total ms, launches) of the step kernels since the last call (CUDA events around each kernel; one sync)."""
        n = ctypes.c_int32()
        ms = float(self.lib.pk_step_kernel_ms_since(self.ctx, ctypes.byref(n)))
        return ms, n.value

    def step_host(self, nsteps, u, u1, dF, dG, dW, fail_step, its_last, its_total, dW_out=None):
        """pk_step_host on (pinned or pageable) host NumPy arrays, in place. Synchronous.
        u1 = dF = dG = None: the Newton warm start stays resident in the context, as it does inside the
        reference's Equations object (equations.py:71,89); `reset_warm_start()` zeroes it."""
        if self.upwind:
            raise NotImplementedError('step_host is provided for the midpoint schemes')
        p = lambda a: None if a is None else ctypes.c_void_p(a.ctypes.data)
        rc = self.lib.pk_step_host(self.ctx, int(nsteps), p(u), p(u1), p(dF), p(dG), p(dW), p(dW_out),
                                   p(fail_step), p(its_last), p(its_total))
        _lib.check(self.lib, self.ctx, rc)

    def reset_warm_start(self):
        _lib.check(self.lib, self.ctx, self.lib.pk_reset_warm_start(self.ctx))

    # -- diagnostics --------------------------------------------------------------------------
    def diagnostics(self, pq=None, periodic=False, want_m=False, want_du=False, fail_step='own', u=None):
        """
        Fused per-realisation reductions. Returns (vals [R, NDIAG] float64, idx [R, NIDX] int32,
        du [R, ne] or None) as device tensors; columns are _lib.D_* / _lib.I_*.
        Failed realisations give NaN rows (outputting.py:107-108); fail_step=None computes every row (the dump
        at t = 0, simulation.py:88).
        """
        fs = self.fail_step if isinstance(fail_step, str) else fail_step
        pq_dev = None
        if pq is not None:
            pq_dev = torch.as_tensor(pq, dtype=torch.float64).to(self.device).contiguous()
        du = torch.empty(self.R, self.ne, dtype=torch.float64, device=self.device) if want_du else None
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        rc = self.lib.pk_diagnostics(self.ctx, _ptr(self.u if u is None else u), _ptr(pq_dev), int(bool(periodic)),
                                     int(bool(want_m) and not self.upwind),
                                     _ptr(self.diag), _ptr(self.didx), _ptr(du), _ptr(fs), stream)
        _lib.check(self.lib, self.ctx, rc)
        if self.upwind:      # m is prognostic (DG1): its norms come from the state, not from a Helmholtz solve
            rc = self.lib.pk_diagnostics_dg1(self.ctx, _ptr(self.m), _ptr(self.u), _ptr(fs),
                                             _ptr(self.diag), stream)
            _lib.check(self.lib, self.ctx, rc)
        self._keep_d = pq_dev
        return self.diag, self.didx, du

    def mu(self, fail_step='own'):
        """The `mu` diagnostic (find_mus, outputting.py:260-370) on the device: ([R, 8] mu_0..3, alt_mu_0..3; [R] status)."""
        if getattr(self, '_x_cells', None) is None:
            xc = (np.arange(self.ne) + 0.5) * (self.Ld / self.ne)          # DiagnosticVariables.coords (simulation.py:265)
            self._x_cells = torch.from_numpy(xc).to(self.device)
            self._mu_out = torch.empty(self.R, 8, dtype=torch.float64, device=self.device)
            self._mu_status = torch.empty(self.R, dtype=torch.int32, device=self.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        fs = self.fail_step if isinstance(fail_step, str) else fail_step
        rc = self.lib.pk_mu(self.ctx, _ptr(self.u), _ptr(self._x_cells), _ptr(fs), _ptr(self._mu_out),
                            _ptr(self._mu_status), stream)
        _lib.check(self.lib, self.ctx, rc)
        return self._mu_out, self._mu_status

    def peakon_suite(self, fail_step='own'):
        """Device part of peakon_diagnostics (outputting.py:373-450); call after diagnostics().
        Returns (vals [R,2] = (min_du, max_du), ids [R,3] = (argmax u, min_du_idx, max_du_idx))."""
        if not hasattr(self, 'pk_vals'):
            self.pk_vals = torch.zeros(self.R, 2, dtype=torch.float64, device=self.device)
            self.pk_ids = torch.zeros(self.R, 3, dtype=torch.int32, device=self.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        fs = self.fail_step if isinstance(fail_step, str) else fail_step
        rc = self.lib.pk_peakon_suite(self.ctx, _ptr(self.u), _ptr(self.didx), _ptr(fs),
                                      _ptr(self.pk_vals), _ptr(self.pk_ids), stream)
        _lib.check(self.lib, self.ctx, rc)
        return self.pk_vals, self.pk_ids

    def du_smooth_loc(self, want_field=False, fail_step='own'):
        """du_smooth (CG5 L2 projection of du, diagnostic_equations.py:115-119) and the locations of its extrema
        (outputting.py:135-138): (out [R,4] = max, x_max, min, x_min; idx [R,2] CG5 dof indices; field [R,5 ne] or None)."""
        from peakondrake_b200.outputting import cg5_condensed
        if not hasattr(self, '_sm_out'):
            self._sm_consts = np.ascontiguousarray(cg5_condensed(), dtype=np.float64)
            self._sm_out = torch.empty(self.R, 4, dtype=torch.float64, device=self.device)
            self._sm_idx = torch.empty(self.R, 2, dtype=torch.int32, device=self.device)
        field = torch.empty(self.R, 5 * self.ne, dtype=torch.float64, device=self.device) if want_field else None
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        fs = self.fail_step if isinstance(fail_step, str) else fail_step
        rc = self.lib.pk_du_smooth_loc(self.ctx, _ptr(self.u), _ptr(fs),
                                       ctypes.c_void_p(self._sm_consts.ctypes.data), _ptr(self._sm_out),
                                       _ptr(self._sm_idx), _ptr(field), stream)
        _lib.check(self.lib, self.ctx, rc)
        return self._sm_out, self._sm_idx, field

    # -- ensemble statistics (local part; ensemble.EnsembleStatistics owns the buffers and the all-reduce) -----
    def ensemble_stats_into(self, stats, vals, idx, stream=None, u=None, fail_step=None):
        if stream is None:
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        u = self.u if u is None else u
        fs = self.fail_step if fail_step is None else fail_step
        rc = self.lib.pk_ensemble_stats(self.ctx, _ptr(u), _ptr(fs), _ptr(vals), _ptr(idx), stats.nbins,
                                        stats.q_lo, stats.q_hi, stats.d_lo, stats.d_hi, _ptr(stats.packed), stream)
        _lib.check(self.lib, self.ctx, rc)
        rc = self.lib.pk_ensemble_scalars(self.ctx, _ptr(vals), _ptr(fs), _ptr(stats.sums), _ptr(stats.ext), stream)
        _lib.check(self.lib, self.ctx, rc)

    # -- peakon SDE (PeakonEquations.update, peakon_equations.py:34-65) ------------------------------------------
    def peakon_sde_step(self, nsteps, p, q, tables, E, method='ito_euler', dW=None, step0=0, traj_stride=0):
        """Advance the (p, q) of every realisation `nsteps` SDE steps on the device.
        p, q: device tensors [R] (in/out); tables: device tensors (pure_xi, pure_xi_x, pure_xi_xx) [num_Xis, ne];
        dW: device tensor [nsteps, R, num_Xis] (e.g. what step(export_dW=True) returned) or None to draw the
        Philox increments of global steps step0.. in-kernel; traj_stride > 0 returns (p, q) after every
        traj_stride steps as a device tensor [nsteps // traj_stride, R, 2]."""
        if method not in ('ito_euler', 'milstein'):
            raise ValueError('peakon_method must be ito_euler or milstein, not %s' % method)
        traj = None
        if traj_stride > 0:
            traj = torch.empty(nsteps // traj_stride, self.R, 2, dtype=torch.float64, device=self.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        rc = self.lib.pk_peakon_sde_step(self.ctx, int(nsteps), _ptr(p), _ptr(q), _ptr(dW), int(step0),
                                         _ptr(tables[0]), _ptr(tables[1]), _ptr(tables[2]),
                                         0 if method == 'ito_euler' else 1, float(E), _ptr(traj),
                                         int(max(traj_stride, 1)), stream)
        _lib.check(self.lib, self.ctx, rc)
        self._keep_sde = (dW, traj)
        return traj
