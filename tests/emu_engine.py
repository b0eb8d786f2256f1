"""
TEST INFRASTRUCTURE: a CPU stand-in for peakondrake_b200.EnsembleIntegrator with the same interface, so that the
host logic of run_batch()/experiment() -- dump vectorisation, the NetCDF writer, sharding over ranks, the gather to
rank 0, the statistics all-reduce -- runs in the not-gpu tests (gloo, world_size 2).  The timestep is the CUDA
source itself executed by the emulation harness (tests/emu); the diagnostics are NumPy restatements.  Nothing in
the product imports this file; experiment(..., engine_cls=EmuEngine) injects it.
"""
import ctypes
import numpy as np
import torch

from emu_util import lib, _p
from oracle import ch_oracle as O
from peakondrake_b200 import _lib as L
from peakondrake_b200.stochastic_functions import build_tables


class EmuEngine(object):
    def __init__(self, ne, Ld, dt, R, alphasq=1.0, scheme='hydrodynamic', timestepping='midpoint', num_Xis=0,
                 Xi_family='sines', sigma=0.0, nXi_updates=1, seed=0, rid_offset=0, device=None, tables=None,
                 c0=0.0, gamma=0.0):
        assert scheme in ('hydrodynamic', 'no_gradient')
        self.ne, self.Ld, self.dt, self.R, self.alphasq = int(ne), float(Ld), float(dt), int(R), float(alphasq)
        self.scheme, self.upwind, self.num_Xis, self.nupd = scheme, False, int(num_Xis), int(nXi_updates)
        self.seed, self.rid_offset = int(seed), int(rid_offset)
        self.device = torch.device('cpu')
        if tables is None:
            tables = build_tables(ne, Ld, Xi_family, num_Xis, sigma, scheme)
        self.tables_host = tuple(np.ascontiguousarray(t, dtype=np.float64) for t in tables)
        xi, xi_x, xi_xx = self.tables_host
        const = self.num_Xis == 0 or (self.num_Xis == 1 and np.all(xi[0] == xi[0, 0]) and not xi_x.any() and not xi_xx.any())
        self.mode = 0 if const else (1 if (not xi_x.any() and not xi_xx.any()) else 2)
        self.xi_const = float(xi[0, 0]) if (const and self.num_Xis) else 0.0
        z = lambda: torch.zeros(self.R, self.ne, dtype=torch.float64)
        self.u, self.u1, self.dF, self.dG = z(), z(), z(), z()
        self.m = None
        self.fail_step = torch.full((self.R,), -1, dtype=torch.int32)
        self.its_last = torch.zeros(self.R, dtype=torch.int32)
        self.its_total = torch.zeros(self.R, dtype=torch.int32)
        self.diag = torch.zeros(self.R, L.NDIAG, dtype=torch.float64)
        self.didx = torch.zeros(self.R, L.NIDX, dtype=torch.int32)
        self.stepc = 0
        self.mesh = O.Mesh1D(self.ne, self.Ld)
        self.orc = O.CHOracle(self.ne, self.Ld, self.dt, self.alphasq)

    def close(self):
        pass

    def set_state(self, u0, w1=None):
        self.u.copy_(torch.as_tensor(np.asarray(u0), dtype=torch.float64).expand(self.R, self.ne))
        self.u1.zero_(); self.dF.zero_(); self.dG.zero_()
        self.fail_step.fill_(-1); self.its_last.zero_(); self.its_total.zero_()
        self.stepc = 0

    def step(self, nsteps, dW=None, export_dW=False):
        n = lambda t: t.numpy()
        dWn = None
        if dW is not None and self.num_Xis > 0:
            dWn = np.ascontiguousarray(torch.as_tensor(dW, dtype=torch.float64).numpy())
        dWo = np.zeros((nsteps, self.R, self.num_Xis)) if (export_dW and self.num_Xis > 0) else None
        fr = np.zeros(self.R, np.int32)
        xi, xix, xixx = self.tables_host if self.num_Xis else (None, None, None)
        rc = lib().pk_emu_step(ctypes.c_int(self.ne), ctypes.c_double(self.Ld), ctypes.c_double(self.dt),
                               ctypes.c_double(self.alphasq), self.mode, self.num_Xis, self.nupd,
                               ctypes.c_double(self.xi_const), self.R, int(nsteps), self.stepc,
                               _p(n(self.u)), _p(n(self.u1)), _p(n(self.dF)), _p(n(self.dG)), _p(xi), _p(xix), _p(xixx),
                               _p(dWn), _p(dWo), ctypes.c_ulonglong(self.seed), self.rid_offset,
                               _p(n(self.fail_step), ctypes.c_int), _p(n(self.its_last), ctypes.c_int),
                               _p(n(self.its_total), ctypes.c_int), _p(fr, ctypes.c_int), 0, 0, 0)
        assert rc == 0, rc
        self.stepc += nsteps
        return None if dWo is None else torch.from_numpy(dWo)

    def step_dump(self, nsteps, dump_every, dW=None, export_dW=False, suite=False):
        nd = nsteps // dump_every
        dg = np.zeros((nd, self.R, L.NDIAG)); di = np.zeros((nd, self.R, L.NIDX), np.int32)
        sv = np.zeros((nd, self.R, 2)) if suite else None
        si = np.zeros((nd, self.R, 3), np.int32) if suite else None
        lib().pk_emu_set_dump(int(dump_every), _p(dg), _p(di, ctypes.c_int), _p(sv), _p(si, ctypes.c_int))
        try:
            out = self.step(nsteps, dW=dW, export_dW=export_dW)
        finally:
            lib().pk_emu_set_dump(0, None, None, None, None)
        t = torch.from_numpy
        return out, t(dg), t(di), (t(sv) if suite else None), (t(si) if suite else None)

    def step_snap(self, nsteps, snap_every, dW=None, export_dW=False, pq=None, periodic=False, want_m=False, suite=False):
        ns = nsteps // snap_every
        su = np.zeros((ns, self.R, self.ne)); sf = np.zeros((ns, self.R), np.int32)
        lib().pk_emu_set_snap(int(snap_every), _p(su), _p(sf, ctypes.c_int))
        try:
            out = self.step(nsteps, dW=dW, export_dW=export_dW)
        finally:
            lib().pk_emu_set_snap(0, None, None)
        # the batched diagnostics kernels, restated: the per-realisation routines on every snapshot row
        keep = (self.u.clone(), self.fail_step.clone())
        dg = np.zeros((ns, self.R, L.NDIAG)); di = np.zeros((ns, self.R, L.NIDX), np.int32)
        sv = np.zeros((ns, self.R, 2)); si = np.zeros((ns, self.R, 3), np.int32)
        for d in range(ns):
            self.u.copy_(torch.from_numpy(su[d])); self.fail_step.copy_(torch.from_numpy(sf[d]))
            v, ix, _ = self.diagnostics(pq=None if pq is None else pq[d], periodic=periodic, want_m=want_m)
            dg[d] = v.numpy(); di[d] = ix.numpy()
            if suite:
                a, b = self.peakon_suite()
                sv[d] = a.numpy(); si[d] = b.numpy()
        self.u.copy_(keep[0]); self.fail_step.copy_(keep[1])
        t = torch.from_numpy
        return out, t(dg), t(di), (t(sv) if suite else None), (t(si) if suite else None), t(su), t(sf)

    # ---- NumPy restatements of the device diagnostics (column layout of include/peakon_b200.h) ----
    def _failed(self, fail_step):
        fs = self.fail_step if isinstance(fail_step, str) else fail_step
        return np.zeros(self.R, bool) if fs is None else (fs.numpy() >= 0)

    def diagnostics(self, pq=None, periodic=False, want_m=False, want_du=False, fail_step='own', u=None):
        failed = self._failed(fail_step)
        u = (self.u if u is None else u).numpy()
        h, a2 = self.Ld / self.ne, self.alphasq
        M, K = self.orc.M, self.orc.K
        d = np.full((self.R, L.NDIAG), np.nan); ix = np.full((self.R, L.NIDX), -1, np.int32)
        for r in range(self.R):
            if failed[r]:
                continue
            ur = u[r]; du = (np.roll(ur, -1) - ur) / h; jmp = np.abs(du - np.roll(du, 1))
            l2, h1 = ur @ (M @ ur), ur @ (K @ ur)
            d[r, L.D_ENERGY] = l2 + a2 * h1; d[r, L.D_L2U2] = l2; d[r, L.D_H1SEMI2] = h1; d[r, L.D_MASS_U] = h * ur.sum()
            d[r, L.D_MAX_DU] = du.max(); d[r, L.D_MIN_DU] = du.min(); d[r, L.D_MAX_JUMP_LOCAL] = jmp.max(); d[r, L.D_MAX_U] = ur.max()
            ix[r] = [np.argmax(du), np.argmin(du), np.argmax(ur), np.argmax(jmp)]
            if want_m:
                m = O.m_field(self.orc, ur)
                d[r, L.D_L2M2] = m @ (M @ m); d[r, L.D_H1SEMI_M2] = m @ (K @ m); d[r, L.D_MASS_M] = np.sum(M @ m)
                d[r, L.D_M_MAX] = m.max(); d[r, L.D_UMU] = ur @ (M @ m)
            if pq is not None:
                e = ur - O.u_sde_field(self.mesh, pq[r][0], pq[r][1], a2, periodic)
                d[r, L.D_ERR_SDE2] = e @ (M @ e)
        self.diag.copy_(torch.from_numpy(d)); self.didx.copy_(torch.from_numpy(ix))
        return self.diag, self.didx, None

    def peakon_suite(self, fail_step='own'):
        failed = self._failed(fail_step)
        u = self.u.numpy(); ne = self.ne; h = self.Ld / ne
        vals = np.full((self.R, 2), np.nan); ids = np.full((self.R, 3), -1, np.int32)
        for r in range(self.R):
            if failed[r]:
                continue
            ur = u[r]; du = (np.roll(ur, -1) - ur) / h
            pc = int(np.argmax(ur)); S = ne // 2
            kmin = next((k for k in range(S) if du[(pc + k) % ne] > du[(pc + k - 1) % ne]), None)
            kmax = next((k for k in range(S) if du[(pc - 1 - k) % ne] < du[(pc - k) % ne]), None)
            ids[r, 0] = pc
            if kmin is not None:
                D = (pc + kmin) % ne; ids[r, 1] = (D - 1) % ne; vals[r, 0] = du[ids[r, 1]]
            if kmax is not None:
                D = (pc - 1 - kmax) % ne; ids[r, 2] = (D - 1) % ne; vals[r, 1] = du[(D + 1) % ne]
        return torch.from_numpy(vals), torch.from_numpy(ids)

    def mu(self, fail_step='own'):
        failed = self._failed(fail_step)
        outv = np.full((self.R, 8), np.nan); st = np.full(self.R, -3, np.int32)
        u = self.u.numpy(); x = self.mesh.x_cells
        lib().pk_emu_mu.restype = ctypes.c_int
        for r in range(self.R):
            if failed[r]:
                continue
            ur = np.ascontiguousarray(u[r]); du = np.ascontiguousarray((np.roll(ur, -1) - ur) / (self.Ld / self.ne))
            o4, a4 = np.zeros(4), np.zeros(4)
            st[r] = lib().pk_emu_mu(_p(ur), _p(du), _p(np.ascontiguousarray(x)), self.ne, ctypes.c_double(self.Ld), _p(o4), _p(a4))
            outv[r, :4] = o4; outv[r, 4:] = a4
        return torch.from_numpy(outv), torch.from_numpy(st)

    def du_smooth_loc(self, want_field=False, fail_step='own'):
        failed = self._failed(fail_step)
        outv = np.full((self.R, 4), np.nan); ids = np.full((self.R, 2), -1, np.int32)
        fld = np.zeros((self.R, 5 * self.ne))
        for r in range(self.R):
            if failed[r]:
                continue
            xmax, xmin, w = O.smooth_locs(self.mesh, self.u.numpy()[r])
            outv[r] = [w.max(), xmax, w.min(), xmin]; ids[r] = [np.argmax(w), np.argmin(w)]; fld[r] = w
        return torch.from_numpy(outv), torch.from_numpy(ids), (torch.from_numpy(fld) if want_field else None)

    def ensemble_stats_into(self, stats, vals, idx, stream=None, u=None, fail_step=None):
        u = (self.u if u is None else u).numpy(); ok = (self.fail_step if fail_step is None else fail_step).numpy() < 0
        ne, nb = self.ne, stats.nbins
        v, ix = vals.numpy(), idx.numpy()
        pk = np.zeros(stats.n_packed)
        pk[:ne] = u[ok].sum(0); pk[ne:2 * ne] = (u[ok] ** 2).sum(0); pk[2 * ne] = ok.sum()

        def hist(x, lo, hi):
            b = np.clip(np.floor((x - lo) / (hi - lo) * nb).astype(int), 0, nb - 1)
            return np.bincount(b, minlength=nb).astype(float)
        pk[2 * ne + 1:2 * ne + 1 + nb] = hist((ix[ok, L.I_ARGMAX_U] + 0.5) * (self.Ld / ne), stats.q_lo, stats.q_hi)
        pk[2 * ne + 1 + nb:] = hist(v[ok, L.D_MIN_DU], stats.d_lo, stats.d_hi)
        stats.packed.copy_(torch.from_numpy(pk))
        nd = L.NDIAG
        s = np.zeros(2 * nd + 2); e = np.full(2 * nd, np.inf)
        for k in range(nd):
            col = v[ok, k]; col = col[~np.isnan(col)]
            s[k] = col.sum(); s[nd + k] = (col * col).sum()
            if len(col):
                e[k] = col.min(); e[nd + k] = -col.max()
        s[2 * nd] = ok.sum(); s[2 * nd + 1] = (~ok).sum()
        stats.sums.copy_(torch.from_numpy(s)); stats.ext.copy_(torch.from_numpy(e))

    def peakon_sde_step(self, nsteps, p, q, tables, E, method='ito_euler', dW=None, step0=0, traj_stride=0):
        tabs = tuple(t.numpy() for t in tables)
        sde = O.PeakonSDE(0.0, 0.0, self.Ld, self.dt, self.alphasq, tabs, self.R, method)
        sde.p, sde.q = p.numpy().copy(), q.numpy().copy()
        traj = []
        for n in range(nsteps):
            if dW is not None:
                w = dW[n].numpy()
            else:
                w = np.zeros((self.R, self.num_Xis))
                for r in range(self.R):
                    for k in range(self.num_Xis):
                        for j in range(self.nupd):
                            w[r, k] += lib().pk_emu_philox_normal(self.seed, self.rid_offset + r, (step0 + n) * self.nupd + j, k, 0) * np.sqrt(self.dt / self.nupd)
            sde.update(w)
            if traj_stride and (n + 1) % traj_stride == 0:
                traj.append(np.stack([sde.p, sde.q], 1))
        p.copy_(torch.from_numpy(sde.p)); q.copy_(torch.from_numpy(sde.q))
        return torch.from_numpy(np.array(traj)) if traj_stride else None
