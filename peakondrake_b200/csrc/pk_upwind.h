// pk_upwind.h -- thread-level phases of the 'upwind'/'ssprk3' timestep for an ensemble.
//
// Replaces Equations.solve() of the reference for scheme 'upwind', timestepping 'ssprk3'
// (/root/reference/peakondrake/equations.py:28-61 set-up, :172-185 step):
//   m in DG1, u in CG1;  us = Dt*u + Xi frozen over the step (:42)
//   three SSPRK3 stages of  dm = -M_e^-1 r(m_trial)  (:48-55, block-Jacobi+ILU is exact for DG)
//   then  (M + a2 K) u = P m  (:58-61,185).
// Per cell (nodes s0,s1 of us; mL,mR of m):
//   Q  = (2 mL s0 + mL s1 + mR s0 + 2 mR s1)/6,   sx = (s1-s0)/h,
//   Phi_v = s_v * (s_v > 0 ? mR_{v-1} : mL_v)                       (upwind flux at vertex v)
//   rhs0 = -Q - sx h/6 (2mL+mR) + Phi_i ,  rhs1 = +Q - sx h/6 (mL+2mR) - Phi_{i+1},  dm = M_e^-1 rhs.
// The volume terms collapse exactly (expand 6Q +- (s1-s0)(..)):
//   rhs0 = Phi_i - s1 (mL+mR)/2 ,   rhs1 = s0 (mL+mR)/2 - Phi_{i+1},
// which is what the kernel evaluates (a third of the flops; each vertex flux Phi is computed once).
// With the factor 2/h of M_e^-1 folded into the frozen velocity, s' = (2/h) us, Phi' = s' * (upwind value):
//   dm0 = (2 Phi'_i + Phi'_{i+1}) - (mL+mR) (s'1 + s'0/2),   dm1 = (mL+mR) (s'0 + s'1/2) - (2 Phi'_{i+1} + Phi'_i),
// i.e. 6 fp64 instructions per cell and stage before the stage combination; the upwind side is picked from the sign
// bit of s' and the blow-up test reads the exponent bits (integer pipe): the kernel is fp64-issue-bound.
//
// Layout: one CTA per realisation, the same chunk geometry as the midpoint kernel (thread t owns
// cells/nodes [s_t, s_t+m_t)), but the scheme is explicit and local, so the WHOLE state of a thread
// (m of its cells, u of its nodes) lives in registers for all timesteps of a launch; only the
// chunk-edge values travel through a handful of shared-memory exchange slots.  The Helmholtz solve
// reuses the recursive-filter + cyclic affine scan of pk_core.h.
#pragma once
#include <cmath>
#include "pk_core.h"

namespace pk {

// exchange slots (of Shared::xb, slot c of thread t at c*T + t)
enum { UW_S0 = 0, UW_A_L = 1, UW_A_R = 2, UW_B_L = 3, UW_B_R = 4, UW_H_L = 5, UW_H_R = 6, UW_SLOTS = 7 };

struct UpwindIO {
  double* m;              // [R, ne, 2]  DG1 m: (left-node value, right-node value) per cell
  double* u;              // [R, ne]
  const double* dW_in;    // [nsteps, R, nxi] or null -> Philox
  double* dW_out;
  const double* xi;       // [nxi, T*M] chunk-interleaved table (permute_table)
  int R, r;
  unsigned rid;
  unsigned long long seed;
  int step0, nsteps;
  int* fail_step;         // [R] in/out
};

// sign bit clear.  Equals (x > 0) except at x = +0 and NaN, where the flux s * (either side) is 0 / NaN anyway.
PK_HD bool uw_positive(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x) >= 0;
#else
  return !std::signbit(x);
#endif
}
// finite: exponent bits not all ones
PK_HD bool uw_finite(double x) {
#if defined(__CUDA_ARCH__)
  return (__double2hiint(x) & 0x7ff00000) != 0x7ff00000;
#else
  return std::isfinite(x);
#endif
}

template <int M, int XARR, int TT = 0>
struct UpwindThread {
  int t, m, s, tl, tr, active, ext;
  double mL[M], mR[M];       // m0: the state
  double aL[M], aR[M];       // trial
  double u[M];               // nodal u of my nodes
  double sv[M];              // s' = (2/h) us at my nodes
  double sR;                 // s' at the first node of the right neighbour
  double hh[M];              // Helmholtz work
  double dw_hold;            // my pre-drawn Brownian increment (draw_step)
  DrawSlot dw_slot;
  static PK_HD int TN(const Params& p) { return TT > 0 ? TT : p.T; }

  PK_HD void init(const Params& p, int tid) {
    t = tid; active = (tid < p.T);
    const int tt = active ? tid : 0;
    t = tt;
    m = chunk_rows(p, tt); s = chunk_start(p, tt); ext = (m == M);
    tl = (tt + p.T - 1) & (p.T - 1); tr = (tt + 1) & (p.T - 1);
  }
  // cell / node j of my chunk exists: 0..M-2 always (folds at compile time), M-1 only in full chunks
  PK_HD bool has_row(int j) const { return j < M - 1 || ext; }

  PK_HD void load(const Params& p, const UpwindIO& io) {
    const long long off = (long long)io.r * p.ne;
    PK_UNROLL
    for (int j = 0; j < M; ++j) {
      mL[j] = mR[j] = u[j] = 0.0; aL[j] = aR[j] = 0.0; hh[j] = 0.0; sv[j] = 0.0;
      if (j < m) { mL[j] = io.m[2 * (off + s + j)]; mR[j] = io.m[2 * (off + s + j) + 1]; u[j] = io.u[off + s + j]; }
    }
    sR = 0.0; dw_hold = 0.0;
  }
  PK_HD void store(const Params& p, const UpwindIO& io) const {
    const long long off = (long long)io.r * p.ne;
    PK_UNROLL
    for (int j = 0; j < M; ++j)
      if (j < m) { io.m[2 * (off + s + j)] = mL[j]; io.m[2 * (off + s + j) + 1] = mR[j]; io.u[off + s + j] = u[j]; }
  }

  // us = Dt*u + Xi at my nodes; publish my first node's us and the edges of m0 (trial of stage 1)
  template <class SH>
  PK_HD void begin_step(const Params& p, const SH& sh, const double* xi) {
    const int T = TN(p);
    const double kdt = p.ih2 * p.dt;
    const double Xc = (!XARR && p.nxi > 0) ? p.ih2 * (sh.dw(0) * p.xi_const) : 0.0;
    PK_UNROLL
    for (int j = 0; j < M; ++j) {
      sv[j] = kdt * u[j] + Xc;
      aL[j] = mL[j]; aR[j] = mR[j];
    }
    if (XARR) {
      for (int k = 0; k < p.nxi; ++k) {
        const double w = p.ih2 * sh.dw(k);
        PK_UNROLL
        for (int j = 0; j < M; ++j) sv[j] += w * xi[k * T * M + j * T + t];      // padding rows of the table are 0
      }
    }
    sh.st(sh.xb, UW_S0 * T + t, sv[0]);
    sh.st(sh.xb, UW_A_L * T + t, aL[0]);
    sh.st(sh.xb, UW_A_R * T + t, ext ? aR[M - 1] : aR[M - 2]);
  }

  // one SSPRK3 stage: d = dm(trial) and the stage combination, edges published to the other slot set.
  // Short chunks (M-1 cells) carry a GHOST cell M-1: a copy of the right neighbour's first cell (trial value and
  // velocity), refreshed at the start of every stage, so that the loop below is the same straight-line code for
  // every thread (no per-row existence tests or selects); what the ghost computes is never published or stored.
  template <class SH>
  PK_HD void stage(const Params& p, const SH& sh, int k) {
    const int T = TN(p);
    const int inL = (k == 2) ? UW_B_L : UW_A_L, inR = (k == 2) ? UW_B_R : UW_A_R;
    if (k == 1) sR = sh.ld(sh.xb, UW_S0 * T + tr);     // s' at the node right of my last cell
    const double upL = sh.ld(sh.xb, inR * T + tl);   // trial value just left of my first vertex
    const double dnR = sh.ld(sh.xb, inL * T + tr);   // trial value just right of my last vertex
    if (!ext) { aL[M - 1] = dnR; if (k == 1) sv[M - 1] = sR; }
    // upwind flux at my first vertex; the flux at the right vertex of cell j is the left flux of cell j+1
    double phi0 = sv[0] * (uw_positive(sv[0]) ? upL : aL[0]);
    PK_UNROLL
    for (int j = 0; j < M; ++j) {
      const bool last = (j == M - 1);
      const double s0 = sv[j], s1 = last ? sR : sv[j + 1 < M ? j + 1 : M - 1];
      const double a = aL[j], b = aR[j];
      const double right = last ? dnR : aL[j + 1 < M ? j + 1 : M - 1];
      const double phi1 = s1 * (uw_positive(s1) ? b : right);
      const double ab = a + b;
      const double c0 = 0.5 * s0 + s1, c1 = 0.5 * s1 + s0;
      const double t0 = 2.0 * phi0 + phi1, t1 = 2.0 * phi1 + phi0;
      // base + dm with dm0 = t0 - ab c0, dm1 = ab c1 - t1; base = m0 (stage 1) or the trial value (stages 2, 3)
      const double e0 = (t0 + (k == 1 ? mL[j] : a)) - ab * c0;
      const double e1 = ((k == 1 ? mR[j] : b) - t1) + ab * c1;
      // in place: cell j+1 reads only its own trial values and the carried flux
      if (k == 1) { aL[j] = e0; aR[j] = e1; }
      else if (k == 2) { aL[j] = 0.75 * mL[j] + 0.25 * e0; aR[j] = 0.75 * mR[j] + 0.25 * e1; }
      else { aL[j] = (1.0 / 3.0) * mL[j] + (2.0 / 3.0) * e0; aR[j] = (1.0 / 3.0) * mR[j] + (2.0 / 3.0) * e1; }
      phi0 = phi1;
    }
    const double lastL = ext ? aL[M - 1] : aL[M - 2], lastR = ext ? aR[M - 1] : aR[M - 2];
    if (k == 1) { sh.st(sh.xb, UW_B_L * T + t, aL[0]); sh.st(sh.xb, UW_B_R * T + t, lastR); }
    else if (k == 2) { sh.st(sh.xb, UW_A_L * T + t, aL[0]); sh.st(sh.xb, UW_A_R * T + t, lastR); }
    else {
      PK_UNROLL
      for (int j = 0; j < M; ++j) { mL[j] = aL[j]; mR[j] = aR[j]; }            // m0 <- m  (:181-182)
      sh.st(sh.xb, UW_H_L * T + t, lastL); sh.st(sh.xb, UW_H_R * T + t, lastR);
    }
  }

  // Helmholtz right-hand side b = P m and the causal local pass (the ghost row of a short chunk is computed and ignored)
  template <class SH>
  PK_HD void helm_causal(const Params& p, const SH& sh, double& A, double& c0, double& c1) {
    const int T = TN(p);
    double pL = sh.ld(sh.xb, UW_H_L * T + tl), pR = sh.ld(sh.xb, UW_H_R * T + tl);
    double y = 0.0, yS = 0.0;
    PK_UNROLL
    for (int j = 0; j < M; ++j) {
      const double b = (pL + 2.0 * pR) + (2.0 * mL[j] + mR[j]);      // (P m)_j * 6/h
      y = b * p.h6k + p.rho * y; hh[j] = y;
      pL = mL[j]; pR = mR[j];
      if (j == M - 2) yS = y;
    }
    A = ext ? p.rpow[M] : p.rpow[M - 1]; c0 = ext ? y : yS; c1 = 0.0;
  }
  PK_HD void helm_anticausal(const Params& p, double in0, double& A, double& c0, double& c1) {
    const double xg = hh[M - 1] + p.rpow[M] * in0;
    double x = ext ? xg : 0.0;
    hh[M - 1] = x;
    PK_UNROLL
    for (int j = M - 2; j >= 0; --j) { x = (hh[j] + p.rpow[j + 1] * in0) + p.rho * x; hh[j] = x; }
    A = ext ? p.rpow[M] : p.rpow[M - 1]; c0 = x; c1 = 0.0;
  }
  // returns true if every new u of this thread is finite
  PK_HD bool helm_finish(const Params& p, double in0) {
    if (!ext) in0 *= p.irho;              // short chunks: fold the missing factor rho into the carry
    bool ok = true;
    PK_UNROLL
    for (int j = M - 1; j >= 0; --j) {
      u[j] = hh[j] + p.rpow[M - j] * in0;
      ok = ok && (uw_finite(u[j]) || (j == M - 1 && !ext));
    }
    return ok;
  }
};

// COOP: warp-shared Philox sub-updates compiled in (nXi_updates >= PK_COOP_NUPD); the lean instantiation serves the rest
template <int M, int XARR, int TT, class Exec, class Th, class SH, bool COOP = false>
PK_HD void run_upwind_steps(Exec& ex, const Params& p, const SH& sh, const UpwindIO& io, int* flag) {
  int failed = io.fail_step[io.r];
  DrawIO dio;
  dio.dW_in = io.dW_in; dio.dW_out = io.dW_out; dio.R = io.R; dio.r = io.r; dio.step0 = io.step0;
  dio.nsteps = io.nsteps; dio.rid = io.rid; dio.seed = io.seed;
  ex.each([&](Th& th) { th.load(p, io); });
  ex.one([&]() { *flag = 0; });
  ex.each_tid([&](Th& th, int tid, int nth, bool, int, int) { th.dw_slot = draw_slot(p, tid); if (dio.per == 0) draw_begin<COOP>(p, dio, nth); });
  for (int n = 0; n < io.nsteps; ++n) {
    draw_advance(dio, n);
    ex.each_tid([&](Th& th, int tid, int nth, bool leader, int, int) {
      draw_step<COOP>(p, sh, dio, n, tid, nth, leader, th.dw_hold, th.dw_slot);
    });
    ex.sync();
    if (failed >= 0) continue;
    // the reference's linear solves raise on a non-finite residual: a non-finite increment fails the
    // step before the state is touched; a blow-up of the state is caught one step after it happened
    bool bad = ex.any_flag(flag);
    { double sw = 0.0; for (int k = 0; k < p.nxi; ++k) sw += sh.dw(k); bad = bad || !uw_finite(sw); }   // any non-finite increment makes the sum non-finite
    if (bad) { failed = ex.any_flag(flag) ? io.step0 + n - 1 : io.step0 + n; continue; }
    ex.each([&](Th& th) { th.begin_step(p, sh, io.xi); });
    ex.sync();
    ex.each([&](Th& th) { th.stage(p, sh, 1); });
    ex.sync();
    ex.each([&](Th& th) { th.stage(p, sh, 2); });
    ex.sync();
    ex.each([&](Th& th) { th.stage(p, sh, 3); });
    ex.sync();
    ex.scan_fwd([&](Th& th, double& A, double& c0, double& c1) { th.helm_causal(p, sh, A, c0, c1); },
                [&](Th& th, double in0, double, double& A, double& c0, double& c1) { th.helm_anticausal(p, in0, A, c0, c1); });
    ex.scan_rev([&](Th& th, double in0, double) { if (!th.helm_finish(p, in0)) *flag = 1; });
  }
  ex.sync();
  if (failed < 0 && ex.any_flag(flag)) failed = io.step0 + io.nsteps - 1;
  ex.each([&](Th& th) { th.store(p, io); });
  ex.one([&]() { io.fail_step[io.r] = failed; });
}

}  // namespace pk
