// pk_core.h -- thread-level phases of the batched stochastic Camassa-Holm timestep
// (scheme 'hydrodynamic'/'no_gradient', timestepping 'midpoint').
//
// Replaces, for an ensemble of realisations, what the reference does once per call in
//   Equations.solve()            /root/reference/peakondrake/equations.py:193-199  (SNES Newton, LU)
//   StochasticFunctions.update() /root/reference/peakondrake/stochastic_functions.py:204-236
// The weak form is equations.py:80-87; element integrals are the closed forms of SURVEY.md
// App. A.1, the Newton system is reduced exactly (App. A.3) to
//   P du = -(H Rpsi + D Rphi - M Rzeta),  P = H(M + N1'/2) - D f'/2 + M g'/2   (cyclic pentadiagonal)
//   dF_new = H^-1 ( f(uh) + f'/2 du ),  dG_new = H^-1 ( g(uh) + g'/2 du )        (constant coefficients)
//
// Parallel layout: one CTA per realisation, T = 2^k threads, thread t owns the contiguous row
// chunk [s_t, s_t + m_t), m_t in {M-1, M}.  Rows 0,1 of a chunk are SEPARATOR rows, the rest is
// the chunk INTERIOR.  The pentadiagonal solve is a two-level nested dissection:
//   level 1 (per thread, registers): LU of the interior block, carrying the fill of the
//            coupling to the own separators; back-substitution of 5 columns gives the "tips";
//   level 2 (block): cyclic block-tridiagonal (2x2 blocks) Schur system on the separators,
//            solved by parallel cyclic reduction;
// the Helmholtz solves are two first-order recursive filters (H = kappa (1-rho S)(1-rho S^-1))
// evaluated as chunk-local recurrences plus a cyclic affine scan over threads.
//
// Everything here is __host__ __device__ so that the identical code is executed by the CUDA
// kernels (pk_kernels.cu) and by the CPU emulation harness the not-gpu tests use
// (tests/emu/pk_emu.cpp).  Per-thread state that lives across a barrier is a member of Thread.
#pragma once
#include "pk_params.h"

#if defined(__CUDA_ARCH__)
#define PK_RCP(x) __drcp_rn(x)     // IEEE round-to-nearest reciprocal == 1.0/x
#else
#define PK_RCP(x) (1.0 / (x))
#endif

namespace pk {
// Reciprocal of a pivot / determinant.  FAST (the single-CTA kernels with a compile-time thread count): MUFU.RCP64H seed
// (20 bits) + two Newton steps -- 1 ulp instead of correctly rounded, no range checks, no slow path: half the
// instructions of __drcp_rn and one dependent FMA less on the critical chain of every elimination row and PCR step
// (+3.5 % on the headline kernel, profiles/r2d_variants.log; the cluster kernels lose 3 % with it and keep __drcp_rn).
// The arguments are O(1) (hm0 = h^2/2 + alphasq); NaN propagates, which is what the failure detection needs.
template <bool FAST>
#if defined(__CUDACC__)
__host__ __device__ __forceinline__
#else
inline
#endif
double rcp_t(double x) {
#if defined(__CUDA_ARCH__)
  if (FAST) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0); r = fma(r, e, r);
    e = fma(-x, r, 1.0); return fma(r, e, r);
  }
  return __drcp_rn(x);
#else
  return 1.0 / x;
#endif
}
}  // namespace pk

namespace pk {

// Shared (per-realisation) storage. Field arrays are [M*T] doubles, interleaved: row j of chunk t
// lives at j*T + t (conflict-free when all threads touch their row j).
struct Shared {
  double *U0, *U1, *F, *G, *X, *Xx, *Xxx;
  double *xb;   // [18*T] exchange slots / PCR rows: slot c of thread t at c*T + t
  double *dW;   // [max(nxi,1)]
  double *fac;  // FS variant only: [6*(M-2)*T] level-1 factors, component c of interior row k of thread t at (c*(M-2)+k)*T + t
  // flat index i = j*T + t (row j of chunk t).  One CTA per realisation: plain shared memory.
  PK_HD double ld(const double* a, int i) const { return a[i]; }
  PK_HD void st(double* a, int i, double v) const { a[i] = v; }
  // slot j of the CALLING thread t (T threads per realisation): same flat index here; the cluster variant turns
  // it into a block-local address without decoding (pk_kernels.cuh, SharedCluster)
  PK_HD double ldo(const double* a, int j, int T, int t) const { return a[j * T + t]; }
  PK_HD void sto(double* a, int j, int T, int t, double v) const { a[j * T + t] = v; }
  // slot j of chunk t, a chunk of the calling thread's own CTA (any thread of it)
  PK_HD double ldc(const double* a, int j, int T, int t) const { return a[j * T + t]; }
  // slot c of the LEFT / RIGHT neighbour thread q = tl / tr of the calling thread (the cluster variant reads these
  // through two per-thread pointers that already name the owning CTA)
  PK_HD double ldL(const double* a, int c, int T, int q) const { return a[c * T + q]; }
  PK_HD double ldR(const double* a, int c, int T, int q) const { return a[c * T + q]; }
  // the tips a chunk publishes for the separator rows of its right neighbour (exchange slots 0..13), and the separator
  // solution (slots 14, 15) of the right neighbour: plain slots here; the two-CTAs-per-SM cluster kernels route the
  // accesses of their edge threads through CTA-local copies, which turns two cluster barriers per Newton iteration
  // into block barriers (pk_kernels.cuh, SharedClusterT)
  PK_HD void stTip(int c, int T, int t, double v) const { xb[c * T + t] = v; }
  PK_HD double ldTip(int c, int T, int q) const { return xb[c * T + q]; }
  PK_HD double ldXsR(int c, int T, int q) const { return xb[(14 + c) * T + q]; }
  // pairs of doubles (16-byte shared-memory accesses): pair slot q of thread tt lives at doubles 2 (q T + tt), +1.
  // Used for the PCR rows, which travel as whole 2x2 blocks.
  PK_HD void ld2(const double* a, int q, int T, int tt, double& v0, double& v1) const {
#if defined(__CUDA_ARCH__)
    const double2 v = reinterpret_cast<const double2*>(a)[q * T + tt]; v0 = v.x; v1 = v.y;
#else
    v0 = a[2 * (q * T + tt)]; v1 = a[2 * (q * T + tt) + 1];
#endif
  }
  PK_HD void ld2l(const double* a, int q, int T, int tt, double& v0, double& v1) const { ld2(a, q, T, tt, v0, v1); }
  // raw offsets into the calling CTA's own copy of an array (scratch use of dead exchange slots)
  PK_HD double ldl(const double* a, int off) const { return a[off]; }
  PK_HD void stl(double* a, int off, double v) const { a[off] = v; }
  PK_HD void st2o(double* a, int q, int T, int t, double v0, double v1) const {
#if defined(__CUDA_ARCH__)
    reinterpret_cast<double2*>(a)[q * T + t] = make_double2(v0, v1);
#else
    a[2 * (q * T + t)] = v0; a[2 * (q * T + t) + 1] = v1;
#endif
  }
  PK_HD double dw(int k) const { return dW[k]; }
  PK_HD void set_dw(int k, double v) const { dW[k] = v; }
};

// windowed drawing of the Brownian increments (draw_step): the (step, mode) pair a thread owns inside a window
struct DrawSlot { int mine = 0, k = 0; };

enum { XB_SLOTS = 18, XB_SLOTS_PP = 30, XB_PP_BASE = 16 };   // _PP: room for a second PCR row buffer (slots 16..29;
                                                             // 16, 17 are dead while the PCR runs, 14, 15 receive its result)
static_assert(XB_SLOTS == 18 && XB_SLOTS_PP == 30, "cluster_xb_doubles() in pk_params.h sizes the exchange area with these");

// natural row index i in [0, ne)  ->  interleaved slot j*T + t
PK_HD int row_slot(const Params& p, int i) {
  const int big = p.rem * p.M;
  int t, j;
  if (i < big) { t = i / p.M; j = i - t * p.M; }
  else { const int k = i - big; const int q = k / (p.M - 1); t = p.rem + q; j = k - q * (p.M - 1); }
  return j * p.T + t;
}

// same with a compile-time M (constant divisions)
template <int M, int TT = 0>
PK_HD int row_slot_c(const Params& p, int i) {
  const int big = p.rem * M;
  int t, j;
  if (i < big) { t = i / M; j = i - t * M; }
  else { const int k = i - big; const int q = k / (M - 1); t = p.rem + q; j = k - q * (M - 1); }
  return j * (TT > 0 ? TT : p.T) + t;
}

// same, as (row j, chunk t)
template <int M>
PK_HD void row_jt_c(const Params& p, int i, int& j, int& t) {
  const int big = p.rem * M;
  if (i < big) { t = i / M; j = i - t * M; }
  else { const int k = i - big; const int q = k / (M - 1); t = p.rem + q; j = k - q * (M - 1); }
}

// chunk-interleaved copy of a [nxi, ne] table: dst[k*T*M + row_slot(i)] = src[k*ne + i], padding rows 0
inline void permute_table_host(const Params& p, const double* src, double* dst) {
  const int TM = p.T * p.M;
  for (int k = 0; k < p.nxi; ++k) {
    for (int i = 0; i < TM; ++i) dst[k * TM + i] = 0.0;
    for (int i = 0; i < p.ne; ++i) dst[k * TM + row_slot(p, i)] = src[k * p.ne + i];
  }
}

struct NodeS { double a, v, w, F, G, x, y; };

// Contributions of one cell (left node A, right node B) to its two nodes, and their derivatives
// with respect to the two nodal values of uh.  SURVEY.md App. A.1.
template <int MODE>
struct CellC {
  double N1L, N1R;
  double fL, fR, gL, gR;          // HALF the cell contributions of f and g (node_packet doubles them in an FMA)
  double dN1L0, dN1L1, dN1R0, dN1R1;
  double dfL0, dfL1, dfR0, dfR1;
  double dgL0, dgL1, dgR0, dgR1;
  // values only (residual evaluation)
  PK_HD void compute_res(const Params& p, const NodeS& A, const NodeS& B) {
    if (PK_IS_TEST(MODE)) {
      // 'test' scheme (equations.py:140): - int psi uh (dv)_x, (dv)_x constant on the cell
      const double dv6 = (B.v - A.v) * (1.0 / 6.0);
      N1L = -dv6 * (2.0 * A.a + B.a); N1R = -dv6 * (A.a + 2.0 * B.a);
      fL = fR = gL = gR = 0.0;
      return;
    }
    const double da = B.a - A.a;
    const double sL = 2.0 * A.v + B.v, sR = A.v + 2.0 * B.v;
    const double da6 = da * (1.0 / 6.0);
    N1L = da6 * sL; N1R = da6 * sR;
    const double a0 = A.a, a1 = B.a;
    const double a00 = a0 * a0, a01 = a0 * a1, a11 = a1 * a1;
    const double cda2 = p.c2h * da * da;
    fL = p.c1h * (3.0 * a00 + 2.0 * a01 + a11) + cda2;
    fR = p.c1h * (a00 + 2.0 * a01 + 3.0 * a11) + cda2;
    if (MODE == NOISE_FULL) {
      const double x0 = A.x, x1 = B.x, y0 = A.y, y1 = B.y;
      const double xs = x0 + x1;
      gL = p.h12 * (a0 * (3.0 * x0 + x1) + a1 * xs) + p.a212 * da * (2.0 * y0 + y1);
      gR = p.h12 * (a0 * xs + a1 * (x0 + 3.0 * x1)) + p.a212 * da * (y0 + 2.0 * y1);
    } else {
      gL = gR = 0.0;
    }
  }
  // values AND derivatives (Newton pass).  f is a homogeneous quadratic form and g is linear in uh, so by Euler's
  // theorem their values follow from the derivative entries computed for the Jacobian anyway:
  //    f/2 = (f'/4) . 2 uh / 2 ... i.e.  fL/2 = dfL0 a0 + dfL1 a1,   gL/2 = dgL0 a0 + dgL1 a1   (df = f'/4, dg = g'/2).
  PK_HD void compute_full(const Params& p, const NodeS& A, const NodeS& B) {
    compute_jac(p, A, B);
    if (PK_IS_TEST(MODE)) { compute_res(p, A, B); return; }
    const double da6 = (B.a - A.a) * (1.0 / 6.0);
    N1L = da6 * (2.0 * A.v + B.v); N1R = da6 * (A.v + 2.0 * B.v);
    fL = dfL0 * A.a + dfL1 * B.a; fR = dfR0 * A.a + dfR1 * B.a;
    if (MODE == NOISE_FULL) { gL = dgL0 * A.a + dgL1 * B.a; gR = dgR0 * A.a + dgR1 * B.a; }
    else { gL = gR = 0.0; }
  }
  // derivatives with respect to the two nodal values of uh, PRE-SCALED the way the Newton matrix uses them
  // (u1 enters through uh = (u1+u)/2):  dN1* = N1'/2,  df* = f'/4 (the matrix has -+ D f'/2 with D = +-1/2),
  // dg* = g'/2.  All scalings are by powers of two, i.e. exact.
  PK_HD void compute_jac(const Params& p, const NodeS& A, const NodeS& B) {
    if (PK_IS_TEST(MODE)) {
      const double dv12 = (B.v - A.v) * (1.0 / 12.0);
      const double sa = p.dt12 * (2.0 * A.a + B.a), sb = p.dt12 * (A.a + 2.0 * B.a);
      dN1L0 = sa - 2.0 * dv12; dN1L1 = -sa - dv12;
      dN1R0 = sb - dv12;       dN1R1 = -sb - 2.0 * dv12;
      dfL0 = dfL1 = dfR0 = dfR1 = 0.0; dgL0 = dgL1 = dgR0 = dgR1 = 0.0;
      return;
    }
    const double da = B.a - A.a;
    const double sL = 2.0 * A.v + B.v, sR = A.v + 2.0 * B.v;
    const double dd = da * p.dt12;
    dN1L0 = -sL * (1.0 / 12.0) + 2.0 * dd; dN1L1 = sL * (1.0 / 12.0) + dd;
    dN1R0 = -sR * (1.0 / 12.0) + dd;       dN1R1 = sR * (1.0 / 12.0) + 2.0 * dd;
    const double a0 = A.a, a1 = B.a;
    const double k = p.c2h * da;
    const double s2 = p.c1h * (a0 + a1);
    dfL0 = p.c1q * (6.0 * a0 + 2.0 * a1) - k; dfL1 = s2 + k;
    dfR0 = s2 - k;                            dfR1 = p.c1q * (2.0 * a0 + 6.0 * a1) + k;
    if (MODE == NOISE_FULL) {
      const double x0 = A.x, x1 = B.x, y0 = A.y, y1 = B.y;
      const double yL = 2.0 * y0 + y1, yR = y0 + 2.0 * y1;
      const double xs = x0 + x1;
      dgL0 = p.h12 * (3.0 * x0 + x1) - p.a212 * yL; dgL1 = p.h12 * xs + p.a212 * yL;
      dgR0 = p.h12 * xs - p.a212 * yR;              dgR1 = p.h12 * (x0 + 3.0 * x1) + p.a212 * yR;
    } else {
      dgL0 = dgL1 = dgR0 = dgR1 = 0.0;
    }
  }
};

// What node l contributes to pentadiagonal rows l-1 (Cp,Qp), l (C0,Q0) and l+1 (Cm,Qm).
struct Packet { double Cm[3], C0[3], Cp[3], Qm, Q0, Qp; };

template <bool FAST = false>
PK_HD void inv2(const double* D, double* Di) {
  const double det = D[0] * D[3] - D[1] * D[2];
  const double id = rcp_t<FAST>(det);
  Di[0] = D[3] * id; Di[1] = -D[1] * id; Di[2] = -D[2] * id; Di[3] = D[0] * id;
}
// C = A*B (2x2 row-major)
PK_HD void mm2(const double* A, const double* B, double* C) {
  C[0] = A[0] * B[0] + A[1] * B[2]; C[1] = A[0] * B[1] + A[1] * B[3];
  C[2] = A[2] * B[0] + A[3] * B[2]; C[3] = A[2] * B[1] + A[3] * B[3];
}
// C += A*B   (written as accumulation chains: two fused multiply-adds per entry)
PK_HD void mma2(const double* A, const double* B, double* C) {
  C[0] = A[1] * B[2] + (A[0] * B[0] + C[0]); C[1] = A[1] * B[3] + (A[0] * B[1] + C[1]);
  C[2] = A[3] * B[2] + (A[2] * B[0] + C[2]); C[3] = A[3] * B[3] + (A[2] * B[1] + C[3]);
}
// y += A*x
PK_HD void mva2(const double* A, const double* x, double* y) {
  y[0] = A[1] * x[1] + (A[0] * x[0] + y[0]); y[1] = A[3] * x[1] + (A[2] * x[0] + y[1]);
}

// TT > 0: number of threads per realisation known at compile time (constant-folds the shared-memory
// address arithmetic); TT == 0: read it from Params.
// FSV ("factors in shared memory"): the level-1 factors and the solution increment live in shared memory instead
// of per-thread register arrays, so the row loops of the assembly / elimination, of the tip back-substitution and
// of the interior solve are ROLLED (a few KB of code that stays in the instruction caches) instead of unrolled
// over the M rows of a thread.  Costs 6(M-2) doubles of shared memory per thread.
// VAR = 0 (default): fully unrolled Newton pass with register-resident factors; the residual-only pass (no factors
//          involved) runs as a rolled loop (VAR = 4 unrolls that too).  Measured on B200:
//          VAR 0 beats 4 by 6 %, and beats the rolled variants 1 and 2 by 24 % and 29 %.
// VAR = 2 ("rolled assembly"): factors stay in registers, but the per-row assembly code is shared by a rolled loop
// and only the elimination step of each interior row (static register indices) is selected by a switch.
template <int M, int MODE, int TT = 0, int VAR = 0>
struct Thread {
  enum { NI = M - 2, FS = (VAR == 1) ? 1 : 0, ROLL = (VAR == 2) ? 1 : 0, ROLL_LITE = (VAR != 4) ? 1 : 0,
#ifdef PK_EXP_FASTR_ALL
         FASTR = 1 };
#else
         FASTR = (TT > 0) ? 1 : 0 };      // fast reciprocal in the tuned single-CTA kernels (rcp_t)
#endif
  double h1[6], h2[6];            // FS: (r, e, f, zb, zw1, zw2) of the two previously eliminated interior rows
  static PK_HD int TN(const Params& p) { return TT > 0 ? TT : p.T; }
  // geometry
  int t, m, s, ext, tl, tr, ml, active;
  // level-1 factors of the interior block (interior row k <-> chunk row k+2)
  double r[NI], e[NI], f[NI], zb[NI], zw1[NI], zw2[NI];
  // separator rows (a,b,c,e,f) and right-hand sides
  double sp0[5], sp1[5], sb0, sb1;
  // streaming window of the assembly pass
  NodeS nm, n0, np;
  CellC<MODE> pc;                 // previous cell
  double wCm2[3], wQm2;           // Cm,Qm of node n-2
  double wCm1[3], wQm1;           // Cm,Qm of node n-1
  double wC01[3], wQ01;           // C0,Q0 of node n-1
  // tips of the interior solution columns (y, W1, W2, V1, V2) at interior rows 0,1 ("F2")
  double F2a[5], F2b[5];
  // reduced block row
  double Lb[4], Db[4], Ub[4], rb[2];
  // results
  double xs0, xs1;                // separator solution
  double du[M];
  double hF[M], hG[M];            // Helmholtz work (rhs -> y -> x)
  double fn2, sn2, xn2;           // partial sums: |R|^2, |delta|^2, |w1|^2
  double Xs;                      // NOISE_SCALAR: dXi value of this step
  double dw_hold;                 // my pre-drawn Brownian increment (draw_step)
  DrawSlot dw_slot;               // which (step, mode) of a window it belongs to

  PK_HD void init(const Params& p, int tid) {
    t = tid; active = (tid < TN(p));
    const int tt = active ? tid : 0;
    m = chunk_rows(p, tt); s = chunk_start(p, tt); ext = (m == M);
    tl = (tt + TN(p) - 1) & (TN(p) - 1); tr = (tt + 1) & (TN(p) - 1);
    ml = chunk_rows(p, tl);
    fn2 = sn2 = xn2 = 0.0; Xs = 0.0; dw_hold = 0.0;
  }

  // ---- liveness helpers ---------------------------------------------------------------------
  // Several per-thread arrays are written under predicates (optional last row, k < nI).  Without an
  // unconditional definition the compiler must keep their previous values alive around the Newton
  // and time-step loops (~190 registers); these "kills" end those live ranges explicitly.
  PK_HD void kill_factors() {
    PK_UNROLL
    for (int c = 0; c < 6; ++c) { h1[c] = 0.0; h2[c] = 0.0; }
    if (FS) return;
    PK_UNROLL
    for (int k = 0; k < NI; ++k) { r[k] = 0.0; e[k] = 0.0; f[k] = 0.0; zb[k] = 0.0; zw1[k] = 0.0; zw2[k] = 0.0; }
  }
  PK_HD void kill_all() {
    kill_factors();
    PK_UNROLL
    for (int j = 0; j < M; ++j) { if (!FS) du[j] = 0.0; hF[j] = 0.0; hG[j] = 0.0; }
    PK_UNROLL
    for (int c = 0; c < 5; ++c) { sp0[c] = 0.0; sp1[c] = 0.0; F2a[c] = 0.0; F2b[c] = 0.0; }
    PK_UNROLL
    for (int c = 0; c < 4; ++c) { Lb[c] = 0.0; Db[c] = 0.0; Ub[c] = 0.0; }
    PK_UNROLL
    for (int c = 0; c < 3; ++c) { wCm2[c] = 0.0; wCm1[c] = 0.0; wC01[c] = 0.0; }
    rb[0] = rb[1] = 0.0; sb0 = sb1 = 0.0; xs0 = xs1 = 0.0; wQm2 = wQm1 = wQ01 = 0.0;
    NodeS z; z.a = z.v = z.w = z.F = z.G = z.x = z.y = 0.0;
    nm = z; n0 = z; np = z;
    pc.N1L = pc.N1R = pc.fL = pc.fR = pc.gL = pc.gR = 0.0;
    pc.dN1L0 = pc.dN1L1 = pc.dN1R0 = pc.dN1R1 = 0.0;
    pc.dfL0 = pc.dfL1 = pc.dfR0 = pc.dfR1 = 0.0;
    pc.dgL0 = pc.dgL1 = pc.dgR0 = pc.dgR1 = 0.0;
  }

  // row j of my chunk exists: rows 0..M-2 always do (compile-time true), row M-1 only in full chunks
  PK_HD bool has_row(int j) const { return j < M - 1 || ext; }

  // ---- smem index helpers -------------------------------------------------------------
  PK_HD int idx(const Params& p, int j) const { return j * TN(p) + t; }
  // node with local index j in [-1, m]: own rows, last row of the left chunk, first row of the right chunk
  PK_HD int nidx(const Params& p, int j) const {
    return j < 0 ? (ml - 1) * TN(p) + tl : ((j < M - 1 || (j == M - 1 && ext)) ? j * TN(p) + t : tr);
  }

  // field value at node j of my chunk (j in [-1, m]): own rows through the block-local accessor, the two halo
  // nodes (last row of the left chunk, first row of the right chunk) through the generic one
  PK_HD bool own_node(int j) const { return j >= 0 && (j < M - 1 || (j == M - 1 && ext)); }
  template <class SH>
  PK_HD double ldn(const Params& p, const SH& sh, const double* a, int j) const {
    return own_node(j) ? sh.ldo(a, j, TN(p), t) : (j < 0 ? sh.ldL(a, ml - 1, TN(p), tl) : sh.ldR(a, 0, TN(p), tr));
  }
  template <class SH>
  PK_HD void load_node(const Params& p, const SH& sh, int j, NodeS& n) const {
    const double u0 = ldn(p, sh, sh.U0, j), u1 = ldn(p, sh, sh.U1, j);
    n.a = 0.5 * (u1 + u0);
    n.w = u1 - u0;                       // + dG below (NOISE_FULL): R_psi = M (u1 - u + dG) + ...
    n.F = PK_IS_TEST(MODE) ? 0.0 : ldn(p, sh, sh.F, j);
    if (MODE == NOISE_FULL) { n.G = ldn(p, sh, sh.G, j); n.x = ldn(p, sh, sh.Xx, j); n.y = ldn(p, sh, sh.Xxx, j); n.w += n.G; }
    else { n.G = 0.0; n.x = 0.0; n.y = 0.0; }
    const double X = (PK_NOISE_OF(MODE) == NOISE_SCALAR) ? Xs : ldn(p, sh, sh.X, j);
    n.v = p.dt * n.a + X;
  }

  // Residual and Jacobian rows of node n0 (neighbours nm, np; cells pc = (nm,n0), cc = (n0,np)).
  PK_HD void node_packet(const Params& p, const CellC<MODE>& cc, Packet& k, bool lite) {
    if (PK_IS_TEST(MODE)) {
      // one unknown: R = M (u1 - u) + N1 (N1 = -T), Newton matrix M + N1'/2: row n0 only
      const double R = p.mb * (nm.w + np.w) + p.ma * n0.w + (pc.N1R + cc.N1L);
      fn2 += R * R;
      if (lite) return;
      k.C0[0] = pc.dN1R0 + p.hm1; k.C0[1] = (pc.dN1R1 + cc.dN1L0) + p.hm0; k.C0[2] = cc.dN1L1 + p.hm1;   // + constant part (M)
      PK_UNROLL
      for (int c = 0; c < 3; ++c) { k.Cm[c] = 0.0; k.Cp[c] = 0.0; }
      k.Q0 = R; k.Qm = 0.0; k.Qp = 0.0;
      return;
    }
    const double N1 = pc.N1R + cc.N1L, ff = pc.fR + cc.fL;
    // NodeS::w = (u1 - u) + dG: the mass matrix acts on both in R_psi.  Accumulation chains (one FMA per term).
    const double Rpsi = p.ma * n0.w + (p.mb * (nm.w + np.w) + (0.5 * (np.F - nm.F) + N1));
    const double Rphi = p.ha * n0.F + (p.hb * (nm.F + np.F) - 2.0 * ff);          // ff, gg: half values (CellC)
    double Rzeta = 0.0;
    if (MODE == NOISE_FULL) Rzeta = p.ha * n0.G + (p.hb * (nm.G + np.G) - 2.0 * (pc.gR + cc.gL));
    fn2 = Rzeta * Rzeta + (Rphi * Rphi + (Rpsi * Rpsi + fn2));
    if (lite) return;                    // residual norm only (uniform across the block)
    // tridiagonal rows of N1'/2, f'/4, g'/2 at node n0 (compute_jac delivers them pre-scaled)
    const double B[3] = {pc.dN1R0, pc.dN1R1 + cc.dN1L0, cc.dN1L1};
    const double Fq[3] = {pc.dfR0, pc.dfR1 + cc.dfL0, cc.dfL1};
    double Gm[3] = {0.0, 0.0, 0.0};
    if (MODE == NOISE_FULL) { Gm[0] = pc.dgR0; Gm[1] = pc.dgR1 + cc.dgL0; Gm[2] = cc.dgL1; }
    PK_UNROLL
    for (int c = 0; c < 3; ++c) {
      const double S = p.hb * B[c] + p.mb * Gm[c];
      k.Cm[c] = S - Fq[c];
      k.Cp[c] = S + Fq[c];
      // the node's own row also carries the constant part H M of the Newton matrix (offsets -1, 0, +1)
      k.C0[c] = p.ha * B[c] + (p.ma * Gm[c] + (c == 1 ? p.hm0 : p.hm1));
    }
    const double hr = p.hb * Rpsi - p.mb * Rzeta;
    k.Qm = hr + 0.5 * Rphi;
    k.Qp = hr - 0.5 * Rphi;
    k.Q0 = p.ha * Rpsi - p.ma * Rzeta;
  }

  // LU step of interior row k (static after inlining) with the three carried columns.
  PK_HD void elim(int k, double a, double b, double c, double ee, double ff, double rhs) {
    double d, z, w1, w2;
    if (k == 0) {
      d = c; z = rhs; w1 = a; w2 = b;
    } else if (k == 1) {
      const double l1 = b * r[0];
      d = c - l1 * e[0]; ee = ee - l1 * f[0];
      z = rhs - l1 * zb[0]; w1 = -l1 * zw1[0]; w2 = a - l1 * zw2[0];
    } else {
      const double l2 = a * r[k - 2];
      const double bb = b - l2 * e[k - 2];
      const double l1 = bb * r[k - 1];
      d = c - l2 * f[k - 2] - l1 * e[k - 1];
      ee = ee - l1 * f[k - 1];
      z = rhs - l2 * zb[k - 2] - l1 * zb[k - 1];
      w1 = -l2 * zw1[k - 2] - l1 * zw1[k - 1];
      w2 = -l2 * zw2[k - 2] - l1 * zw2[k - 1];
    }
    r[k] = rcp_t<FASTR != 0>(d); e[k] = ee; f[k] = ff; zb[k] = z; zw1[k] = w1; zw2[k] = w2;
  }

  // Row (n-1) is complete once packet n is known: separator rows are stored, interior rows eliminated.
  // cp/qp = Cp,Qp of node n (own or, for the last row, the right neighbour's first node).
  PK_HD void finish_row(const Params& p, int row, const double* cp, double qp) {
    if (row == 0) {
      sp0[2] += cp[0]; sp0[3] += cp[1]; sp0[4] += cp[2]; sb0 -= qp;
      return;
    }
    const double a = p.hm2 + wCm2[0];
    const double b = wCm2[1] + wC01[0];            // (C0 carries the constants of offsets -1, 0, +1)
    const double c = wCm2[2] + wC01[1] + cp[0];
    const double ee = wC01[2] + cp[1];
    const double ff = p.hm2 + cp[2];
    const double rhs = -(wQm2 + wQ01 + qp);
    if (row == 1) {
      sp1[0] = a; sp1[1] = b; sp1[2] = c; sp1[3] = ee; sp1[4] = ff; sb1 = rhs;
    } else {
      elim(row - 2, a, b, c, ee, ff, rhs);
    }
  }

  PK_HD void push_packet(const Packet& k) {
    PK_UNROLL
    for (int c = 0; c < 3; ++c) { wCm2[c] = wCm1[c]; wCm1[c] = k.Cm[c]; wC01[c] = k.C0[c]; }
    wQm2 = wQm1; wQm1 = k.Qm; wQ01 = k.Q0;
  }

  // One streaming iteration q >= 1: cell (q-1, q); finalises node q-1 and row q-2.
  template <class SH>
  PK_HD void stream(const Params& p, const SH& sh, int q, bool lite) {
    load_node(p, sh, q, np);
    CellC<MODE> cc;
    if (lite) cc.compute_res(p, n0, np); else cc.compute_full(p, n0, np);
    Packet k; node_packet(p, cc, k, lite);     // node q-1
    if (lite) { nm = n0; n0 = np; pc.N1R = cc.N1R; pc.fR = cc.fR; pc.gR = cc.gR; return; }
    const int n = q - 1;
    if (n == 0) {
      // separator row 0 so far: H*M part, own C0/Q0 (Cm of the left neighbour is added in reduce_build)
      sp0[0] = p.hm2; sp0[1] = k.C0[0]; sp0[2] = k.C0[1]; sp0[3] = k.C0[2]; sp0[4] = p.hm2;
      sb0 = -k.Q0;
      // publish what the left neighbour's last row needs from my first node
      sh.sto(sh.xb, 14, TN(p), t, k.Cp[0]); sh.sto(sh.xb, 15, TN(p), t, k.Cp[1]); sh.sto(sh.xb, 16, TN(p), t, k.Cp[2]);
      sh.sto(sh.xb, 17, TN(p), t, k.Qp);
    } else {
      finish_row(p, n - 1, k.Cp, k.Qp);
    }
    push_packet(k);
    nm = n0; n0 = np; pc = cc;
  }

  // ================= FS variant: rolled row loops, factors in shared memory =================================
  PK_HD int fidx(const Params& p, int c, int k) const { return (c * NI + k) * TN(p) + t; }
  // solution increment of my row j: registers (du[]) or, FS, exchange slot j (free between the PCR and the update)
  template <class SH>
  PK_HD double get_du(const Params& p, const SH& sh, int j) const {
    if (FS) return sh.ldo(sh.xb, j, TN(p), t);
    return du[j < M ? j : M - 1];
  }

  // LU step of interior row k (runtime k).  With a zero history the generic recurrence reproduces the k = 0, 1
  // special cases of elim(): there the sub-diagonal entries couple to the separators and seed the spike columns.
  template <class SH>
  PK_HD void elim_fs(const Params& p, const SH& sh, int k, double a, double b, double c, double ee, double ff, double rhs) {
    const double l2 = a * h2[0];
    const double bb = b - l2 * h2[1];
    const double l1 = bb * h1[0];
    const double d = c - l2 * h2[2] - l1 * h1[1];
    ee = ee - l1 * h1[2];
    const double z = rhs - l2 * h2[3] - l1 * h1[3];
    double w1 = -l2 * h2[4] - l1 * h1[4];
    double w2 = -l2 * h2[5] - l1 * h1[5];
    if (k == 0) { w1 += a; w2 += b; }
    if (k == 1) w2 += a;
    const double rr = rcp_t<FASTR != 0>(d);
    sh.st(sh.fac, fidx(p, 0, k), rr); sh.st(sh.fac, fidx(p, 1, k), ee); sh.st(sh.fac, fidx(p, 2, k), ff);
    sh.st(sh.fac, fidx(p, 3, k), z); sh.st(sh.fac, fidx(p, 4, k), w1); sh.st(sh.fac, fidx(p, 5, k), w2);
    PK_UNROLL
    for (int q = 0; q < 6; ++q) h2[q] = h1[q];
    h1[0] = rr; h1[1] = ee; h1[2] = ff; h1[3] = z; h1[4] = w1; h1[5] = w2;
  }

  template <class SH>
  PK_HD void finish_row_fs(const Params& p, const SH& sh, int row, const double* cp, double qp) {
    if (row == 0) {
      sp0[2] += cp[0]; sp0[3] += cp[1]; sp0[4] += cp[2]; sb0 -= qp;
      return;
    }
    const double a = p.hm2 + wCm2[0];
    const double b = wCm2[1] + wC01[0];            // (C0 carries the constants of offsets -1, 0, +1)
    const double c = wCm2[2] + wC01[1] + cp[0];
    const double ee = wC01[2] + cp[1];
    const double ff = p.hm2 + cp[2];
    const double rhs = -(wQm2 + wQ01 + qp);
    if (row == 1) {
      sp1[0] = a; sp1[1] = b; sp1[2] = c; sp1[3] = ee; sp1[4] = ff; sb1 = rhs;
    } else if (FS) {
      elim_fs(p, sh, row - 2, a, b, c, ee, ff, rhs);
    } else {
      // register-resident factors: the elimination step with static indices, picked by the (thread-uniform) row
      switch (row - 2) {
        case 0: elim(0, a, b, c, ee, ff, rhs); break;
        case 1: if (NI > 1) elim(NI > 1 ? 1 : 0, a, b, c, ee, ff, rhs); break;
        case 2: if (NI > 2) elim(NI > 2 ? 2 : 0, a, b, c, ee, ff, rhs); break;
        case 3: if (NI > 3) elim(NI > 3 ? 3 : 0, a, b, c, ee, ff, rhs); break;
        case 4: if (NI > 4) elim(NI > 4 ? 4 : 0, a, b, c, ee, ff, rhs); break;
        default: if (NI > 5) elim(NI > 5 ? 5 : 0, a, b, c, ee, ff, rhs); break;
      }
    }
  }

  // streaming iteration q >= 2 (runtime q): cell (q-1, q); finalises node q-1 and row q-2
  template <class SH>
  PK_HD void stream_fs(const Params& p, const SH& sh, int q, bool lite) {
    load_node(p, sh, q, np);
    CellC<MODE> cc;
    if (lite) cc.compute_res(p, n0, np); else cc.compute_full(p, n0, np);
    Packet k; node_packet(p, cc, k, lite);     // node q-1
    if (lite) { nm = n0; n0 = np; pc.N1R = cc.N1R; pc.fR = cc.fR; pc.gR = cc.gR; return; }
    finish_row_fs(p, sh, q - 2, k.Cp, k.Qp);
    push_packet(k);
    nm = n0; n0 = np; pc = cc;
  }

  // residual-only pass (no Jacobian, no factors): one rolled loop over the rows of the chunk
  template <class SH>
  PK_HD void residual_finish_rolled(const Params& p, const SH& sh) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int q = 2; q <= m; ++q) {
      load_node(p, sh, q, np);
      CellC<MODE> cc; cc.compute_res(p, n0, np);
      Packet k; node_packet(p, cc, k, true);
      nm = n0; n0 = np; pc.N1R = cc.N1R; pc.fR = cc.fR; pc.gR = cc.gR;
    }
  }

  template <class SH>
  PK_HD void assemble_finish_fs(const Params& p, const SH& sh, bool lite) {
    if (lite) { residual_finish_rolled(p, sh); return; }
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int q = 2; q <= m; ++q) stream_fs(p, sh, q, false);
    // last row (m-1): Cp,Qp come from the right neighbour's first node
    double cp[3];
    cp[0] = sh.ldR(sh.xb, 14, TN(p), tr); cp[1] = sh.ldR(sh.xb, 15, TN(p), tr); cp[2] = sh.ldR(sh.xb, 16, TN(p), tr);
    const double qp = sh.ldR(sh.xb, 17, TN(p), tr);
    finish_row_fs(p, sh, m - 1, cp, qp);
    sh.sto(sh.xb, 10, TN(p), t, wCm1[0]); sh.sto(sh.xb, 11, TN(p), t, wCm1[1]); sh.sto(sh.xb, 12, TN(p), t, wCm1[2]);
    sh.sto(sh.xb, 13, TN(p), t, wQm1);
    if (!FS) { tips_registers(p, sh); return; }

    // tips: back-substitute the five columns (y, W1, W2, V1, V2) through the interior rows
    const int nI = m - 2;
    double xp1[5] = {0.0, 0.0, 0.0, 1.0, 0.0};
    double xp2[5] = {0.0, 0.0, 0.0, 0.0, 1.0};
    double L2a[5], L2b[5];
    PK_UNROLL
    for (int c = 0; c < 5; ++c) { L2a[c] = L2b[c] = 0.0; F2a[c] = F2b[c] = 0.0; }
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int k = nI - 1; k >= 0; --k) {
      const double rk = sh.ld(sh.fac, fidx(p, 0, k)), ek = sh.ld(sh.fac, fidx(p, 1, k)), fk = sh.ld(sh.fac, fidx(p, 2, k));
      const double zc[5] = {sh.ld(sh.fac, fidx(p, 3, k)), -sh.ld(sh.fac, fidx(p, 4, k)), -sh.ld(sh.fac, fidx(p, 5, k)), 0.0, 0.0};
      PK_UNROLL
      for (int c = 0; c < 5; ++c) {
        const double x = rk * (zc[c] - ek * xp1[c] - fk * xp2[c]);
        xp2[c] = xp1[c]; xp1[c] = x;
        if (k == nI - 1) L2b[c] = x;
        if (k == nI - 2) L2a[c] = x;
        if (k == 1) F2b[c] = x;
        if (k == 0) F2a[c] = x;
      }
    }
    sh.sto(sh.xb, 0, TN(p), t, L2a[0]); sh.sto(sh.xb, 1, TN(p), t, L2b[0]);
    sh.sto(sh.xb, 2, TN(p), t, L2a[1]); sh.sto(sh.xb, 3, TN(p), t, L2a[2]);
    sh.sto(sh.xb, 4, TN(p), t, L2b[1]); sh.sto(sh.xb, 5, TN(p), t, L2b[2]);
    sh.sto(sh.xb, 6, TN(p), t, L2a[3]); sh.sto(sh.xb, 7, TN(p), t, L2a[4]);
    sh.sto(sh.xb, 8, TN(p), t, L2b[3]); sh.sto(sh.xb, 9, TN(p), t, L2b[4]);
  }

  // interior solution with known separators; the increment of row j goes to exchange slot j
  template <class SH>
  PK_HD void interior_solve_fs(const Params& p, const SH& sh) {
    const int T = TN(p);
    const int nI = m - 2;
    double xp1 = sh.ldR(sh.xb, 14, T, tr), xp2 = sh.ldR(sh.xb, 15, T, tr);
    sh.sto(sh.xb, 0, T, t, xs0); sh.sto(sh.xb, 1, T, t, xs1);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int k = nI - 1; k >= 0; --k) {
      const double z = sh.ld(sh.fac, fidx(p, 3, k)) - sh.ld(sh.fac, fidx(p, 4, k)) * xs0 - sh.ld(sh.fac, fidx(p, 5, k)) * xs1;
      const double x = sh.ld(sh.fac, fidx(p, 0, k)) * (z - sh.ld(sh.fac, fidx(p, 1, k)) * xp1 - sh.ld(sh.fac, fidx(p, 2, k)) * xp2);
      xp2 = xp1; xp1 = x;
      sh.sto(sh.xb, (k + 2), T, t, x);
      if (k == nI - 1) sh.sto(sh.xb, 16, T, t, x);     // my last row, for the right neighbour
    }
  }

  // ---- phase A1: first two cells, publish first-node packet --------------------------------
  template <class SH>
  // lite: evaluate the residual norm only (no Jacobian, no elimination) -- the same code, so that the
  // convergence check does not cost a second copy of the assembly in the instruction cache
  PK_HD void assemble_begin(const Params& p, const SH& sh, bool lite) {
    fn2 = 0.0;
    kill_factors();
    load_node(p, sh, -1, nm);
    load_node(p, sh, 0, n0);
    if (lite) pc.compute_res(p, nm, n0); else pc.compute_full(p, nm, n0);
    PK_UNROLL
    for (int c = 0; c < 3; ++c) { wCm1[c] = 0.0; wCm2[c] = 0.0; wC01[c] = 0.0; }
    wQm1 = wQm2 = wQ01 = 0.0;
    stream(p, sh, 1, lite);
  }

  // ---- phase A2: rest of the chunk, interior LU, tips; publishes tips + last packet --------
  template <class SH>
  PK_HD void assemble_finish(const Params& p, const SH& sh, bool lite) {
    PK_UNROLL
    for (int q = 2; q <= M - 1; ++q) stream(p, sh, q, lite);
    if (ext) stream(p, sh, M, lite);
    if (lite) return;
    // last row (m-1): Cp,Qp come from the right neighbour's first node
    double cp[3];
    cp[0] = sh.ldR(sh.xb, 14, TN(p), tr); cp[1] = sh.ldR(sh.xb, 15, TN(p), tr); cp[2] = sh.ldR(sh.xb, 16, TN(p), tr);
    const double qp = sh.ldR(sh.xb, 17, TN(p), tr);
    if (ext) finish_row(p, M - 1, cp, qp); else finish_row(p, M - 2, cp, qp);
    // publish my last node's Cm,Qm for the right neighbour's separator row 0
    sh.stTip(10, TN(p), t, wCm1[0]); sh.stTip(11, TN(p), t, wCm1[1]); sh.stTip(12, TN(p), t, wCm1[2]);
    sh.stTip(13, TN(p), t, wQm1);

    tips_registers(p, sh);
  }

  // tips of the interior solution columns from the register-resident factors
  template <class SH>
  PK_HD void tips_registers(const Params& p, const SH& sh) {
    // back-substitute the five columns  x_I = y + Wn1 xs0 + Wn2 xs1 + Vn1 xR0 + Vn2 xR1
    // (a chunk has NI or NI-1 interior rows: the top row is skipped by short chunks, and the tips at the last
    // two interior rows are picked once at the end instead of per row and column)
    double xp1[5] = {0.0, 0.0, 0.0, 1.0, 0.0};
    double xp2[5] = {0.0, 0.0, 0.0, 0.0, 1.0};
    double xa[5], xb_[5], xc[5];          // columns at interior rows NI-1, NI-2, NI-3
    PK_UNROLL
    for (int c = 0; c < 5; ++c) { xa[c] = xb_[c] = xc[c] = 0.0; F2a[c] = F2b[c] = 0.0; }
    PK_UNROLL
    for (int k = NI - 1; k >= 0; --k) {
      if (k < NI - 1 || ext) {
        const double zc[5] = {zb[k], -zw1[k], -zw2[k], 0.0, 0.0};
        PK_UNROLL
        for (int c = 0; c < 5; ++c) {
          const double x = r[k] * (zc[c] - e[k] * xp1[c] - f[k] * xp2[c]);
          xp2[c] = xp1[c]; xp1[c] = x;
          if (k == NI - 1) xa[c] = x;
          if (k == NI - 2) xb_[c] = x;
          if (k == NI - 3) xc[c] = x;
          if (k == 1) F2b[c] = x;
          if (k == 0) F2a[c] = x;
        }
      }
    }
    double L2a[5], L2b[5];
    PK_UNROLL
    for (int c = 0; c < 5; ++c) { L2b[c] = ext ? xa[c] : xb_[c]; L2a[c] = ext ? xb_[c] : xc[c]; }
    // slots 0..9: tips at the LAST two interior rows, read by the right neighbour
    sh.stTip(0, TN(p), t, L2a[0]); sh.stTip(1, TN(p), t, L2b[0]);
    sh.stTip(2, TN(p), t, L2a[1]); sh.stTip(3, TN(p), t, L2a[2]);
    sh.stTip(4, TN(p), t, L2b[1]); sh.stTip(5, TN(p), t, L2b[2]);
    sh.stTip(6, TN(p), t, L2a[3]); sh.stTip(7, TN(p), t, L2a[4]);
    sh.stTip(8, TN(p), t, L2b[3]); sh.stTip(9, TN(p), t, L2b[4]);
  }

  // ---- phase B1: build the reduced (Schur) block row of my separator pair -------------------
  template <class SH>
  PK_HD void reduce_build(const Params& p, const SH& sh) {
    const int T = TN(p);
    const double yL[2] = {sh.ldTip(0, T, tl), sh.ldTip(1, T, tl)};
    const double WL[4] = {sh.ldTip(2, T, tl), sh.ldTip(3, T, tl), sh.ldTip(4, T, tl), sh.ldTip(5, T, tl)};
    const double VL[4] = {sh.ldTip(6, T, tl), sh.ldTip(7, T, tl), sh.ldTip(8, T, tl), sh.ldTip(9, T, tl)};
    sp0[0] += sh.ldTip(10, T, tl); sp0[1] += sh.ldTip(11, T, tl); sp0[2] += sh.ldTip(12, T, tl);
    sb0 -= sh.ldTip(13, T, tl);
    const double E[4] = {sp0[0], sp0[1], 0.0, sp1[0]};
    const double C[4] = {sp0[2], sp0[3], sp1[1], sp1[2]};
    const double Gc[4] = {sp0[4], 0.0, sp1[3], sp1[4]};
    const double WF[4] = {F2a[1], F2a[2], F2b[1], F2b[2]};
    const double VF[4] = {F2a[3], F2a[4], F2b[3], F2b[4]};
    const double yF[2] = {F2a[0], F2b[0]};
    mm2(E, WL, Lb);
    PK_UNROLL
    for (int c = 0; c < 4; ++c) Db[c] = C[c];
    mma2(E, VL, Db); mma2(Gc, WF, Db);
    mm2(Gc, VF, Ub);
    double tmp[2] = {0.0, 0.0};
    mva2(E, yL, tmp); mva2(Gc, yF, tmp);
    rb[0] = sb0 - tmp[0]; rb[1] = sb1 - tmp[1];
  }

  // write my block row (L, U, D^-1, r) to the exchange slots 0..13
  template <class SH>
  PK_HD void reduce_publish(const Params& p, const SH& sh, int base = 0, bool rawD = false) {
    const int T = TN(p);
    // seven 16-byte stores: (L0 L1)(L2 L3)(U0 U1)(U2 U3)(D^-1 ...)(r0 r1), pair slots base/2 .. base/2 + 6
    const int b2 = base >> 1;
    sh.st2o(sh.xb, b2 + 0, T, t, Lb[0], Lb[1]); sh.st2o(sh.xb, b2 + 1, T, t, Lb[2], Lb[3]);
    sh.st2o(sh.xb, b2 + 2, T, t, Ub[0], Ub[1]); sh.st2o(sh.xb, b2 + 3, T, t, Ub[2], Ub[3]);
    if (rawD) {                                                                    // separator rows travel with D itself
      sh.st2o(sh.xb, b2 + 4, T, t, Db[0], Db[1]); sh.st2o(sh.xb, b2 + 5, T, t, Db[2], Db[3]);
    } else {
      double Di[4];
      inv2<FASTR != 0>(Db, Di);
      sh.st2o(sh.xb, b2 + 4, T, t, Di[0], Di[1]); sh.st2o(sh.xb, b2 + 5, T, t, Di[2], Di[3]);
    }
    sh.st2o(sh.xb, b2 + 6, T, t, rb[0], rb[1]);
  }
  // one published row (pair layout of reduce_publish) of thread q
  template <class SH>
  PK_HD void pcr_load_row(const Params& p, const SH& sh, int b2, int q, double* L_, double* U_, double* Di_, double* r_) const {
    const int T = TN(p);
    sh.ld2(sh.xb, b2 + 0, T, q, L_[0], L_[1]); sh.ld2(sh.xb, b2 + 1, T, q, L_[2], L_[3]);
    sh.ld2(sh.xb, b2 + 2, T, q, U_[0], U_[1]); sh.ld2(sh.xb, b2 + 3, T, q, U_[2], U_[3]);
    sh.ld2(sh.xb, b2 + 4, T, q, Di_[0], Di_[1]); sh.ld2(sh.xb, b2 + 5, T, q, Di_[2], Di_[3]);
    sh.ld2(sh.xb, b2 + 6, T, q, r_[0], r_[1]);
  }

  // same for a row of my own CTA (cluster variant: block-local address, no owner decode)
  template <class SH>
  PK_HD void pcr_load_row_local(const Params& p, const SH& sh, int b2, int q, double* L_, double* U_, double* Di_, double* r_) const {
    const int T = TN(p);
    sh.ld2l(sh.xb, b2 + 0, T, q, L_[0], L_[1]); sh.ld2l(sh.xb, b2 + 1, T, q, L_[2], L_[3]);
    sh.ld2l(sh.xb, b2 + 2, T, q, U_[0], U_[1]); sh.ld2l(sh.xb, b2 + 3, T, q, U_[2], U_[3]);
    sh.ld2l(sh.xb, b2 + 4, T, q, Di_[0], Di_[1]); sh.ld2l(sh.xb, b2 + 5, T, q, Di_[2], Di_[3]);
    sh.ld2l(sh.xb, b2 + 6, T, q, r_[0], r_[1]);
  }

  // one parallel-cyclic-reduction step with stride d (reads only)
  template <class SH>
  // base: first exchange slot of the row buffer to read (0, or XB_PP_BASE when the rows ping-pong between two buffers)
  PK_HD void pcr_step(const Params& p, const SH& sh, int d, int base = 0) {
    const int T = TN(p);
    const int ql = (t + T - d) & (T - 1), qr = (t + d) & (T - 1);
#ifdef PK_EXP_PCR_STEP_WIDE      // A/B: both partner rows live at once (the form up to round 2)
    double Ll[4], Ul[4], Dl[4], rl[2], Lr[4], Ur[4], Dr[4], rr[2];
    pcr_load_row(p, sh, base >> 1, ql, Ll, Ul, Dl, rl);
    pcr_load_row(p, sh, base >> 1, qr, Lr, Ur, Dr, rr);
    double al[4], ga[4];
    mm2(Lb, Dl, al); mm2(Ub, Dr, ga);
    PK_UNROLL
    for (int c = 0; c < 4; ++c) { al[c] = -al[c]; ga[c] = -ga[c]; }
    mm2(al, Ll, Lb);
    mm2(ga, Ur, Ub);
    mma2(al, Ul, Db); mma2(ga, Lr, Db);
    mva2(al, rl, rb); mva2(ga, rr, rb);
#else
    // left partner, then right partner: one neighbour row live at a time; the contributions to D and r are added in the
    // same order as before (bit-identical)
    {
      double Ll[4], Ul[4], Dl[4], rl[2], al[4];
      pcr_load_row(p, sh, base >> 1, ql, Ll, Ul, Dl, rl);
      mm2(Lb, Dl, al);
      PK_UNROLL
      for (int c = 0; c < 4; ++c) al[c] = -al[c];
      mma2(al, Ul, Db); mva2(al, rl, rb);
      mm2(al, Ll, Lb);
    }
    {
      double Lr[4], Ur[4], Dr[4], rr[2], ga[4];
      pcr_load_row(p, sh, base >> 1, qr, Lr, Ur, Dr, rr);
      mm2(Ub, Dr, ga);
      PK_UNROLL
      for (int c = 0; c < 4; ++c) ga[c] = -ga[c];
      mma2(ga, Lr, Db); mva2(ga, rr, rb);
      mm2(ga, Ur, Ub);
    }
#endif
  }

  // ---- partitioned Schur solve (thread-block-cluster variant) ------------------------------------------------
  // The T block rows are cut into C segments of Tc rows, one per CTA.  Row 0 of a segment is the segment's
  // SEPARATOR s_k; rows 1..Tc-1 are reduced by a PCR that never leaves the segment: a row whose partner at
  // distance d would lie outside keeps its coupling, which from then on is the coupling to the separator on that
  // side (left: s_k, the own CTA's; right: s_{k+1}).  A row is never coupled to an in-segment row and to the
  // separator on the same side at once, so the L / U slots carry either meaning and no extra storage or flops
  // are needed.  After ceil(log2(Tc-1)) steps (block-local barriers, block-local loads) every interior row reads
  //     L s_k + D x + U s_{k+1} = r .
  // The separators then eliminate their two neighbours (an ordinary PCR step at distance 1) and form a cyclic
  // block-tridiagonal system of C rows, solved by PCR among the C leading threads (strides Tc, 2Tc, ...).
  template <class SH>
  PK_HD void pcr_step_local(const Params& p, const SH& sh, int d, int base, int Tc) {
    const int tl_ = t & (Tc - 1);
    const bool hl = tl_ - d >= 1, hr = tl_ != 0 && tl_ + d <= Tc - 1;
    // (left partner first, then right: the contributions to D and r are added in that order)
    if (hl) {
      double Ll[4], Ul[4], Dl[4], rl[2], al[4];
      pcr_load_row_local(p, sh, base >> 1, t - d, Ll, Ul, Dl, rl);
      mm2(Lb, Dl, al);
      PK_UNROLL
      for (int c = 0; c < 4; ++c) al[c] = -al[c];
      mma2(al, Ul, Db); mva2(al, rl, rb);
      mm2(al, Ll, Lb);
    }
    if (hr) {
      double Lr[4], Ur[4], Dr[4], rr[2], ga[4];
      pcr_load_row_local(p, sh, base >> 1, t + d, Lr, Ur, Dr, rr);
      mm2(Ub, Dr, ga);
      PK_UNROLL
      for (int c = 0; c < 4; ++c) ga[c] = -ga[c];
      mma2(ga, Lr, Db); mva2(ga, rr, rb);
      mm2(ga, Ur, Ub);
    }
  }
  // interior rows of a segment once the separators are known (slots 14, 15 of the leading threads)
  template <class SH>
  PK_HD void part_back(const Params& p, const SH& sh, int Tc) {
    const int T = TN(p);
    const int lead = t - (t & (Tc - 1)), nxt = (lead + Tc) & (T - 1);
    const double sL[2] = {sh.ld(sh.xb, 14 * T + lead), sh.ld(sh.xb, 15 * T + lead)};
    const double sR[2] = {sh.ld(sh.xb, 14 * T + nxt), sh.ld(sh.xb, 15 * T + nxt)};
    double tmp[2] = {0.0, 0.0}, Di[4];
    mva2(Lb, sL, tmp); mva2(Ub, sR, tmp);
    const double r0 = rb[0] - tmp[0], r1 = rb[1] - tmp[1];
    inv2<FASTR != 0>(Db, Di);
    xs0 = Di[0] * r0 + Di[1] * r1;
    xs1 = Di[2] * r0 + Di[3] * r1;
    sh.sto(sh.xb, 14, T, t, xs0); sh.sto(sh.xb, 15, T, t, xs1);
  }

  // ---- separator system solved REDUNDANTLY by the first C threads of every CTA (no cluster barrier inside) ----
  // Thread j < C of each CTA builds separator row j from three published rows (separator j with its raw diagonal
  // block, the last interior row of segment j-1, the first interior row of segment j; DSMEM loads), then the C rows
  // are reduced by PCR through a scratch area in the CTA's own shared memory (the row buffer that is not being
  // read: 2 x C rows of 14 doubles, then the C solutions), with block barriers only.  The separator row lives in
  // Lb/Ub/Db/rb of the solving thread, which reloads its own (published) row afterwards: no extra per-thread state.
  static PK_HD int sep_scratch_doubles(int C) { return 30 * C; }
  template <class SH>
  PK_HD void sep_store(const SH& sh, int scr_off, int half, int C, int j) {
    double Di[4]; inv2<FASTR != 0>(Db, Di);
    const int o = scr_off + (half * C + j) * 14;
    PK_UNROLL
    for (int c = 0; c < 4; ++c) { sh.stl(sh.xb, o + c, Lb[c]); sh.stl(sh.xb, o + 4 + c, Ub[c]); sh.stl(sh.xb, o + 8 + c, Di[c]); }
    sh.stl(sh.xb, o + 12, rb[0]); sh.stl(sh.xb, o + 13, rb[1]);
  }
  template <class SH>
  PK_HD void sep_load(const SH& sh, int scr_off, int half, int C, int j, double* L_, double* U_, double* Di_, double* r_) const {
    const int o = scr_off + (half * C + j) * 14;
    PK_UNROLL
    for (int c = 0; c < 4; ++c) { L_[c] = sh.ldl(sh.xb, o + c); U_[c] = sh.ldl(sh.xb, o + 4 + c); Di_[c] = sh.ldl(sh.xb, o + 8 + c); }
    r_[0] = sh.ldl(sh.xb, o + 12); r_[1] = sh.ldl(sh.xb, o + 13);
  }
  template <class SH>
  PK_HD void sep_absorb(const double* Ll, const double* Ul, const double* Dl, const double* rl,
                        const double* Lr, const double* Ur, const double* Dr, const double* rr) {
    double al[4], ga[4], Ln[4], Un[4];
    mm2(Lb, Dl, al); mm2(Ub, Dr, ga);
    PK_UNROLL
    for (int c = 0; c < 4; ++c) { al[c] = -al[c]; ga[c] = -ga[c]; }
    mm2(al, Ll, Ln); mm2(ga, Ur, Un);
    mma2(al, Ul, Db); mma2(ga, Lr, Db);
    mva2(al, rl, rb); mva2(ga, rr, rb);
    PK_UNROLL
    for (int c = 0; c < 4; ++c) { Lb[c] = Ln[c]; Ub[c] = Un[c]; }
  }
  template <class SH>
  PK_HD void sep_build(const Params& p, const SH& sh, int base, int Tc, int C, int scr_off) {
    const int T = TN(p);
    const int j = t & (Tc - 1), lead = j * Tc;
    double Ll[4], Ul[4], Dl[4], rl[2], Lr[4], Ur[4], Dr[4], rr[2];
    pcr_load_row(p, sh, base >> 1, lead, Lb, Ub, Db, rb);                     // raw D (reduce_publish rawD)
    pcr_load_row(p, sh, base >> 1, (lead + T - 1) & (T - 1), Ll, Ul, Dl, rl);
    pcr_load_row(p, sh, base >> 1, lead + 1, Lr, Ur, Dr, rr);
    sep_absorb<SH>(Ll, Ul, Dl, rl, Lr, Ur, Dr, rr);
    sep_store(sh, scr_off, 0, C, j);
  }
  template <class SH>
  PK_HD void sep_step(const SH& sh, int d, int half, int Tc, int C, int scr_off) {
    const int j = t & (Tc - 1);
    double Ll[4], Ul[4], Dl[4], rl[2], Lr[4], Ur[4], Dr[4], rr[2];
    sep_load(sh, scr_off, half, C, (j + C - d) & (C - 1), Ll, Ul, Dl, rl);
    sep_load(sh, scr_off, half, C, (j + d) & (C - 1), Lr, Ur, Dr, rr);
    sep_absorb<SH>(Ll, Ul, Dl, rl, Lr, Ur, Dr, rr);
    sep_store(sh, scr_off, half ^ 1, C, j);
  }
  template <class SH>
  PK_HD void sep_final(const SH& sh, int half, int Tc, int C, int scr_off) {
    const int j = t & (Tc - 1);
    double Lq[4], Uq[4], Dq[4], rq[2];
    sep_load(sh, scr_off, half, C, (j + C / 2) & (C - 1), Lq, Uq, Dq, rq);
    double S[4], Sq[4], SD[4], A[4], Ai[4], rhs[2];
    PK_UNROLL
    for (int c = 0; c < 4; ++c) { S[c] = Lb[c] + Ub[c]; Sq[c] = Lq[c] + Uq[c]; A[c] = Db[c]; }
    mm2(S, Dq, SD);
    PK_UNROLL
    for (int c = 0; c < 4; ++c) SD[c] = -SD[c];
    mma2(SD, Sq, A);
    rhs[0] = rb[0]; rhs[1] = rb[1];
    mva2(SD, rq, rhs);
    inv2<FASTR != 0>(A, Ai);
    sh.stl(sh.xb, scr_off + 28 * C + 2 * j, Ai[0] * rhs[0] + Ai[1] * rhs[1]);
    sh.stl(sh.xb, scr_off + 28 * C + 2 * j + 1, Ai[2] * rhs[0] + Ai[3] * rhs[1]);
  }
  // every thread: its separator-pair solution from the separators of its own and the next segment
  template <class SH>
  PK_HD void part_back_local(const Params& p, const SH& sh, int Tc, int C, int scr_off, int base) {
    const int T = TN(p);
    const int k = t / Tc, kn = (k + 1) & (C - 1);
    const double sL_[2] = {sh.ldl(sh.xb, scr_off + 28 * C + 2 * k), sh.ldl(sh.xb, scr_off + 28 * C + 2 * k + 1)};
    if ((t & (Tc - 1)) == 0) { xs0 = sL_[0]; xs1 = sL_[1]; }
    else {
      const double sR_[2] = {sh.ldl(sh.xb, scr_off + 28 * C + 2 * kn), sh.ldl(sh.xb, scr_off + 28 * C + 2 * kn + 1)};
      double tmp[2] = {0.0, 0.0}, Di[4];
      // the threads that solved a separator row get their own row back from the published copy (L, U, D^-1, r)
      if ((t & (Tc - 1)) < C) pcr_load_row_local(p, sh, base >> 1, t, Lb, Ub, Di, rb);
      else inv2<FASTR != 0>(Db, Di);
      mva2(Lb, sL_, tmp); mva2(Ub, sR_, tmp);
      const double r0 = rb[0] - tmp[0], r1 = rb[1] - tmp[1];
      xs0 = Di[0] * r0 + Di[1] * r1;
      xs1 = Di[2] * r0 + Di[3] * r1;
    }
    sh.sto(sh.xb, 14, T, t, xs0); sh.sto(sh.xb, 15, T, t, xs1);
  }

  // last level: rows t and t+T/2 form a closed 2-block system
  template <class SH>
  PK_HD void pcr_final(const Params& p, const SH& sh, int base = 0) {
    const int T = TN(p);
    const int q = (t + T / 2) & (T - 1);
    double Lq[4], Uq[4], Dq[4], rq[2];
    pcr_load_row(p, sh, base >> 1, q, Lq, Uq, Dq, rq);
    double S[4], Sq[4], SD[4], A[4], Ai[4], rhs[2];
    PK_UNROLL
    for (int c = 0; c < 4; ++c) { S[c] = Lb[c] + Ub[c]; Sq[c] = Lq[c] + Uq[c]; A[c] = Db[c]; }
    mm2(S, Dq, SD);
    PK_UNROLL
    for (int c = 0; c < 4; ++c) SD[c] = -SD[c];
    mma2(SD, Sq, A);
    rhs[0] = rb[0]; rhs[1] = rb[1];
    mva2(SD, rq, rhs);
    inv2<FASTR != 0>(A, Ai);
    xs0 = Ai[0] * rhs[0] + Ai[1] * rhs[1];
    xs1 = Ai[2] * rhs[0] + Ai[3] * rhs[1];
    sh.sto(sh.xb, 14, T, t, xs0); sh.sto(sh.xb, 15, T, t, xs1);
  }

  // =========== hybrid Schur solve: cyclic reduction down to 32 block rows, PCR inside ONE warp, back-substitution ===========
  // Plain PCR keeps all T threads busy for log2(T) steps (O(T log T) work).  Here the rows are halved by cyclic
  // reduction with the work COMPACTED onto the leading threads (row 2 s k of level s is handled by thread k, so
  // whole warps drop out), the remaining 32 rows are solved by warp 0 with warp-level synchronisation only, and
  // the eliminated rows are recovered level by level.  Rows live in the exchange slots: L 0-3, U 4-7, D^-1 8-11,
  // r 12-13, D 14-17; the solution goes to slots 14-15 (D is dead by then).
  template <class SH>
  PK_HD void hyb_publish(const Params& p, const SH& sh) {
    const int T = TN(p);
    double Di[4]; inv2<FASTR != 0>(Db, Di);
    PK_UNROLL
    for (int c = 0; c < 4; ++c) {
      sh.st(sh.xb, c * T + t, Lb[c]); sh.st(sh.xb, (4 + c) * T + t, Ub[c]); sh.st(sh.xb, (8 + c) * T + t, Di[c]);
      sh.st(sh.xb, (14 + c) * T + t, Db[c]);
    }
    sh.st(sh.xb, 12 * T + t, rb[0]); sh.st(sh.xb, 13 * T + t, rb[1]);
  }
  // one elimination: row i absorbs its neighbours il, ir (rows given as smem indices); result in L_, D_, U_, r_
  template <class SH>
  PK_HD void hyb_absorb(const Params& p, const SH& sh, int il, int ir, double* L_, double* D_, double* U_, double* r_) const {
    const int T = TN(p);
    double Ll[4], Ul[4], Dl[4], rl[2], Lr[4], Ur[4], Dr[4], rr[2];
    PK_UNROLL
    for (int c = 0; c < 4; ++c) {
      Ll[c] = sh.ld(sh.xb, c * T + il); Ul[c] = sh.ld(sh.xb, (4 + c) * T + il); Dl[c] = sh.ld(sh.xb, (8 + c) * T + il);
      Lr[c] = sh.ld(sh.xb, c * T + ir); Ur[c] = sh.ld(sh.xb, (4 + c) * T + ir); Dr[c] = sh.ld(sh.xb, (8 + c) * T + ir);
    }
    rl[0] = sh.ld(sh.xb, 12 * T + il); rl[1] = sh.ld(sh.xb, 13 * T + il);
    rr[0] = sh.ld(sh.xb, 12 * T + ir); rr[1] = sh.ld(sh.xb, 13 * T + ir);
    double al[4], ga[4], Ln[4], Un[4];
    mm2(L_, Dl, al); mm2(U_, Dr, ga);
    PK_UNROLL
    for (int c = 0; c < 4; ++c) { al[c] = -al[c]; ga[c] = -ga[c]; }
    mm2(al, Ll, Ln);
    mm2(ga, Ur, Un);
    mma2(al, Ul, D_); mma2(ga, Lr, D_);
    mva2(al, rl, r_); mva2(ga, rr, r_);
    PK_UNROLL
    for (int c = 0; c < 4; ++c) { L_[c] = Ln[c]; U_[c] = Un[c]; }
  }
  // cyclic-reduction level with stride s: thread k (called for k < T/(2s)) rewrites row 2 s k in place
  template <class SH>
  PK_HD void hyb_down(const Params& p, const SH& sh, int s_) {
    const int T = TN(p);
    const int i = 2 * s_ * t, il = (i - s_) & (T - 1), ir = (i + s_) & (T - 1);
    double L_[4], D_[4], U_[4], r_[2];
    PK_UNROLL
    for (int c = 0; c < 4; ++c) { L_[c] = sh.ld(sh.xb, c * T + i); U_[c] = sh.ld(sh.xb, (4 + c) * T + i); D_[c] = sh.ld(sh.xb, (14 + c) * T + i); }
    r_[0] = sh.ld(sh.xb, 12 * T + i); r_[1] = sh.ld(sh.xb, 13 * T + i);
    hyb_absorb(p, sh, il, ir, L_, D_, U_, r_);
    double Di[4]; inv2<FASTR != 0>(D_, Di);
    PK_UNROLL
    for (int c = 0; c < 4; ++c) {
      sh.st(sh.xb, c * T + i, L_[c]); sh.st(sh.xb, (4 + c) * T + i, U_[c]); sh.st(sh.xb, (8 + c) * T + i, Di[c]);
      sh.st(sh.xb, (14 + c) * T + i, D_[c]);
    }
    sh.st(sh.xb, 12 * T + i, r_[0]); sh.st(sh.xb, 13 * T + i, r_[1]);
  }
  // warp 0, lane t: the 32 remaining rows S*t.  step 0: load my row; odd steps 2j+1: PCR step with distance 2^j
  // (reads), even steps 2j+2: publish (writes); last step: closed 2-block solve with the row 16 lanes away.
  template <class SH>
  PK_HD void hyb_warp_step(const Params& p, const SH& sh, int S_, int step) {
    const int T = TN(p);
    const int i = S_ * t;
    if (step == 0) {
      PK_UNROLL
      for (int c = 0; c < 4; ++c) { Lb[c] = sh.ld(sh.xb, c * T + i); Ub[c] = sh.ld(sh.xb, (4 + c) * T + i); Db[c] = sh.ld(sh.xb, (14 + c) * T + i); }
      rb[0] = sh.ld(sh.xb, 12 * T + i); rb[1] = sh.ld(sh.xb, 13 * T + i);
      return;
    }
    if (step == 9) {
      const int q = S_ * ((t + 16) & 31);
      double Lq[4], Uq[4], Dq[4], rq[2];
      PK_UNROLL
      for (int c = 0; c < 4; ++c) {
        Lq[c] = sh.ld(sh.xb, c * T + q); Uq[c] = sh.ld(sh.xb, (4 + c) * T + q); Dq[c] = sh.ld(sh.xb, (8 + c) * T + q);
      }
      rq[0] = sh.ld(sh.xb, 12 * T + q); rq[1] = sh.ld(sh.xb, 13 * T + q);
      double Sm[4], Sq[4], SD[4], A[4], Ai[4], rhs[2];
      PK_UNROLL
      for (int c = 0; c < 4; ++c) { Sm[c] = Lb[c] + Ub[c]; Sq[c] = Lq[c] + Uq[c]; A[c] = Db[c]; }
      mm2(Sm, Dq, SD);
      PK_UNROLL
      for (int c = 0; c < 4; ++c) SD[c] = -SD[c];
      mma2(SD, Sq, A);
      rhs[0] = rb[0]; rhs[1] = rb[1];
      mva2(SD, rq, rhs);
      inv2<FASTR != 0>(A, Ai);
      sh.st(sh.xb, 14 * T + i, Ai[0] * rhs[0] + Ai[1] * rhs[1]);
      sh.st(sh.xb, 15 * T + i, Ai[2] * rhs[0] + Ai[3] * rhs[1]);
      return;
    }
    if (step & 1) {
      const int d = 1 << ((step - 1) >> 1);
      hyb_absorb(p, sh, S_ * ((t - d) & 31), S_ * ((t + d) & 31), Lb, Db, Ub, rb);
    } else {
      double Di[4]; inv2<FASTR != 0>(Db, Di);
      PK_UNROLL
      for (int c = 0; c < 4; ++c) {
        sh.st(sh.xb, c * T + i, Lb[c]); sh.st(sh.xb, (4 + c) * T + i, Ub[c]); sh.st(sh.xb, (8 + c) * T + i, Di[c]);
      }
      sh.st(sh.xb, 12 * T + i, rb[0]); sh.st(sh.xb, 13 * T + i, rb[1]);
    }
  }
  // back-substitution level with stride s: thread k (k < T/(2s)) recovers row 2 s k + s from its two neighbours
  template <class SH>
  PK_HD void hyb_up(const Params& p, const SH& sh, int s_) {
    const int T = TN(p);
    const int i = 2 * s_ * t + s_, il = i - s_, ir = (i + s_) & (T - 1);
    double L_[4], U_[4], Di[4];
    PK_UNROLL
    for (int c = 0; c < 4; ++c) { L_[c] = sh.ld(sh.xb, c * T + i); U_[c] = sh.ld(sh.xb, (4 + c) * T + i); Di[c] = sh.ld(sh.xb, (8 + c) * T + i); }
    const double xl[2] = {sh.ld(sh.xb, 14 * T + il), sh.ld(sh.xb, 15 * T + il)};
    const double xr[2] = {sh.ld(sh.xb, 14 * T + ir), sh.ld(sh.xb, 15 * T + ir)};
    double v[2] = {sh.ld(sh.xb, 12 * T + i), sh.ld(sh.xb, 13 * T + i)};
    v[0] -= L_[0] * xl[0] + L_[1] * xl[1] + U_[0] * xr[0] + U_[1] * xr[1];
    v[1] -= L_[2] * xl[0] + L_[3] * xl[1] + U_[2] * xr[0] + U_[3] * xr[1];
    sh.st(sh.xb, 14 * T + i, Di[0] * v[0] + Di[1] * v[1]);
    sh.st(sh.xb, 15 * T + i, Di[2] * v[0] + Di[3] * v[1]);
  }
  template <class SH>
  PK_HD void hyb_load_xs(const Params& p, const SH& sh) {
    xs0 = sh.ld(sh.xb, 14 * TN(p) + t); xs1 = sh.ld(sh.xb, 15 * TN(p) + t);
  }

  // ---- phase C: interior solution with known separators --------------------------------------
  template <class SH>
  PK_HD void interior_solve(const Params& p, const SH& sh) {
    const int T = TN(p);
    double xp1 = sh.ldXsR(0, T, tr), xp2 = sh.ldXsR(1, T, tr);
    du[0] = xs0; du[1] = xs1;
    PK_UNROLL
    for (int k = NI - 1; k >= 0; --k) {
      if (k < NI - 1 || ext) {
        const double z = zb[k] - zw1[k] * xs0 - zw2[k] * xs1;
        const double x = r[k] * (z - e[k] * xp1 - f[k] * xp2);
        xp2 = xp1; xp1 = x;
        du[k + 2] = x;
      } else {
        du[k + 2] = 0.0;
      }
    }
    sh.sto(sh.xb, 16, T, t, ext ? du[M - 1] : du[M - 2]);   // my last row, for the right neighbour
  }

  // ---- phase D1: Helmholtz right-hand sides f + f'/2 du, g + g'/2 du; causal local pass ------
  // returns the chunk aggregate (value with zero carry-in) in cF, cG and the multiplier A = rho^m
  template <class SH>
  PK_HD void helm_causal_local(const Params& p, const SH& sh, double& A, double& cF, double& cG) {
    // Right-hand sides  f(uh) + f'(uh) e  and  g(uh) + g'(uh) e  with e = du/2 (the change of uh).
    // f is a homogeneous quadratic form, f(a) = B(a, a) with B symmetric bilinear, and g is linear in uh, hence
    //    f(uh) + f'(uh) e = B(uh, uh + 2e) = B(uh, uh + du),        g(uh) + g'(uh) e = g(uh + e),
    // with  B_L(a, s) = c1 (a0 (3 s0 + s1) + a1 (s0 + s1)) + c2 (a1 - a0)(s1 - s0)  on a cell (B_R mirrored).
    const int T = TN(p);
    double a0, s0, b0, x0 = 0.0, y0 = 0.0;
    {
      a0 = 0.5 * (ldn(p, sh, sh.U0, -1) + ldn(p, sh, sh.U1, -1));
      if (MODE == NOISE_FULL) { x0 = ldn(p, sh, sh.Xx, -1); y0 = ldn(p, sh, sh.Xxx, -1); }
      const double d0 = sh.ldL(sh.xb, 16, T, tl);          // increment of the left neighbour's last row
      s0 = a0 + d0; b0 = a0 + 0.5 * d0;
    }
    double prevFR = 0.0, prevGR = 0.0;
    PK_UNROLL
    for (int j = 0; j < M; ++j) { hF[j] = 0.0; hG[j] = 0.0; }
    PK_UNROLL
    for (int q = 0; q <= M; ++q) {
      if (q < M || ext) {
        const double a1 = 0.5 * (ldn(p, sh, sh.U0, q) + ldn(p, sh, sh.U1, q));
        double x1 = 0.0, y1 = 0.0;
        if (MODE == NOISE_FULL) { x1 = ldn(p, sh, sh.Xx, q); y1 = ldn(p, sh, sh.Xxx, q); }
        const double d1 = (q < M - 1 || (q == M - 1 && ext)) ? get_du(p, sh, q) : sh.ldXsR(0, T, tr);
        const double s1 = a1 + d1, b1 = a1 + 0.5 * d1;      // uh + du and uh + e at the right node of the cell
        const double ss = s0 + s1;
        const double cdd = p.c2k * (a1 - a0) * (s1 - s0);        // (constants carry the 1/kappa of the filter)
        const double fLv = p.c1k * (a0 * (2.0 * s0 + ss) + a1 * ss) + cdd;
        const double fRv = p.c1k * (a0 * ss + a1 * (2.0 * s1 + ss)) + cdd;
        double gLv = 0.0, gRv = 0.0;
        if (MODE == NOISE_FULL) {
          const double xs = x0 + x1, db = b1 - b0;
          gLv = p.h6k * (b0 * (3.0 * x0 + x1) + b1 * xs) + p.a26k * db * (2.0 * y0 + y1);
          gRv = p.h6k * (b0 * xs + b1 * (x0 + 3.0 * x1)) + p.a26k * db * (y0 + 2.0 * y1);
        }
        if (q >= 1) { hF[q - 1] = prevFR + fLv; hG[q - 1] = prevGR + gLv; }
        prevFR = fRv; prevGR = gRv;
        a0 = a1; s0 = s1; b0 = b1; x0 = x1; y0 = y1;
      }
    }
    // causal recurrence with zero carry-in
    double yF = 0.0, yG = 0.0;
    PK_UNROLL
    for (int j = 0; j < M; ++j) {
      if (has_row(j)) {
        yF = hF[j] + p.rho * yF; hF[j] = yF;
        if (MODE == NOISE_FULL) { yG = hG[j] + p.rho * yG; hG[j] = yG; }
      }
    }
    A = ext ? p.rpow[M] : p.rpow[M - 1]; cF = yF; cG = yG;
  }

  // ---- phase D2: apply causal carry, anticausal local pass -----------------------------------
  PK_HD void helm_anticausal_local(const Params& p, double inF, double inG, double& A, double& cF, double& cG) {
    // carry-in of the causal filter: y_j += rho^(j+1) * in, then the anticausal recurrence with zero carry-in
    double xF = 0.0, xG = 0.0;
    PK_UNROLL
    for (int j = M - 1; j >= 0; --j) {
      if (has_row(j)) {
        xF = (hF[j] + p.rpow[j + 1] * inF) + p.rho * xF; hF[j] = xF;
        if (MODE == NOISE_FULL) { xG = (hG[j] + p.rpow[j + 1] * inG) + p.rho * xG; hG[j] = xG; }
      }
    }
    A = ext ? p.rpow[M] : p.rpow[M - 1]; cF = xF; cG = xG;
  }

  // ---- phase D3: apply anticausal carry, write the new iterate, accumulate |delta|^2, |w1|^2 ---
  template <class SH>
  PK_HD void helm_finish_update(const Params& p, const SH& sh, double inF, double inG) {
    sn2 = 0.0; xn2 = 0.0;
    // x_j = xl_j + rho^(m-j) * carry
    // weights rho^(m-j): for short chunks (m = M-1) fold the missing factor rho into the carry
    if (!ext) { inF *= p.irho; inG *= p.irho; }
    double pwv[M];
    PK_UNROLL
    for (int j = 0; j < M; ++j) pwv[j] = p.rpow[M - j];
    PK_UNROLL
    for (int j = 0; j < M; ++j) {
      if (has_row(j)) {
        const int T = TN(p);
        const double duj = get_du(p, sh, j);
        const double u1n = sh.ldo(sh.U1, j, T, t) + duj;
        const double Fn = hF[j] + pwv[j] * inF;
        const double dFv = Fn - sh.ldo(sh.F, j, T, t);
        sh.sto(sh.U1, j, T, t, u1n); sh.sto(sh.F, j, T, t, Fn);
        sn2 = dFv * dFv + (duj * duj + sn2); xn2 = Fn * Fn + (u1n * u1n + xn2);     // accumulation chains (FMAs)
        if (MODE == NOISE_FULL) {
          const double Gn = hG[j] + pwv[j] * inG;
          const double dGv = Gn - sh.ldo(sh.G, j, T, t);
          sh.sto(sh.G, j, T, t, Gn);
          sn2 = dGv * dGv + sn2; xn2 = Gn * Gn + xn2;
        }
      }
    }
  }

  // ---- scheme 'test': the unknown is u1 alone, no Helmholtz relations to update ------------------
  template <class SH>
  PK_HD void test_update(const Params& p, const SH& sh) {
    sn2 = 0.0; xn2 = 0.0;
    PK_UNROLL
    for (int j = 0; j < M; ++j) {
      if (has_row(j)) {
        const double duj = get_du(p, sh, j);
        const double u1n = sh.ldo(sh.U1, j, TN(p), t) + duj;
        sh.sto(sh.U1, j, TN(p), t, u1n);
        sn2 = duj * duj + sn2; xn2 = u1n * u1n + xn2;
      }
    }
  }

  // ---- accept the step: u <- u1 (equations.py:195) ----------------------------------------------
  template <class SH>
  PK_HD void accept(const Params& p, const SH& sh) {
    PK_UNROLL
    for (int j = 0; j < M; ++j)
      if (has_row(j)) sh.sto(sh.U0, j, TN(p), t, sh.ldo(sh.U1, j, TN(p), t));
  }

  // ---- noise fields of this step: dXi = sum_k dW_k xi_k etc. (stochastic_functions.py:229-234) ----
  // The tables arrive in the chunk-interleaved order of the shared arrays (pk_set_tables permutes
  // them once): table k holds row j of chunk t at k*(T*M) + j*T + t, so the loads of a warp are
  // contiguous and every thread fills exactly its own rows.
  template <class SH>
  PK_HD void fill_noise(const Params& p, const SH& sh, const double* xi, const double* xi_x, const double* xi_xx) {
    if (PK_NOISE_OF(MODE) == NOISE_SCALAR) return;
    const int T = TN(p), TM = T * M;
    double X[M], Xx[M], Xxx[M];
    PK_UNROLL
    for (int j = 0; j < M; ++j) { X[j] = 0.0; Xx[j] = 0.0; Xxx[j] = 0.0; }
#if defined(__CUDA_ARCH__)
#ifdef PK_EXP_NOISE_UNROLL4
#pragma unroll 4
#else
#pragma unroll 2
#endif
#endif
    for (int k = 0; k < p.nxi; ++k) {
      const double w = sh.dw(k);
      PK_UNROLL
      for (int j = 0; j < M; ++j) {
        const int i = k * TM + j * T + t;
        X[j] += w * xi[i];
        if (MODE == NOISE_FULL) { Xx[j] += w * xi_x[i]; Xxx[j] += w * xi_xx[i]; }
      }
    }
    PK_UNROLL
    for (int j = 0; j < M; ++j) {
      if (has_row(j)) {
        sh.sto(sh.X, j, T, t, X[j]);
        if (MODE == NOISE_FULL) { sh.sto(sh.Xx, j, T, t, Xx[j]); sh.sto(sh.Xxx, j, T, t, Xxx[j]); }
      }
    }
  }

  // ---- fused dump diagnostics on the state in shared memory -----------------------------------------
  // What pk_diag_kernel computes from global memory (outputting.py:106-142,156-162 on the fields of
  // diagnostic_equations.py:94-113), evaluated by every thread on its own rows, with the SAME per-node expressions
  // (du = (u_{i+1} - u_i)/h, jump = |du_i - du_{i-1}|), so extrema and indices are bit-identical to the separate
  // kernel; sums differ only in summation order.  av/ai: candidates of  max du, max -du, max u, max |jump|
  // (ties -> lowest node index, NumPy argmax / argmin).
  // x / h, correctly rounded, without a division: q = x * RN(1/h), one fused correction with the exact remainder
  // (Markstein: with the correctly rounded reciprocal and r = x - q h computed exactly by the FMA, q + r * RN(1/h) rounds
  // to RN(x / h)).  Three instructions instead of the ~40 of a double division; the same bits as pk_diag_kernel's `/ h`
  // (tests/test_gpu_configs.py::test_fused_dump_equals_separate_kernels compares them bit for bit).
  static PK_HD double div_h(const Params& p, double x) {
#if defined(__CUDA_ARCH__)
    const double q = x * p.ih;
    const double r = fma(-q, p.h, x);
    const double res = fma(r, p.ih, q);
    return res == 0.0 ? q : res;           // (-0) / h = -0: the correction loses the sign of a zero
#else
    return x / p.h;
#endif
  }
  template <class SH>
  PK_HD double ldn2(const Params& p, const SH& sh, const double* a, int j) const {      // j in [-1, m + 1]
    const int mm = ext ? M : M - 1;
    return j <= mm ? ldn(p, sh, a, j) : sh.ldR(a, 1, TN(p), tr);                           // row 1 of the right neighbour
  }
  template <class SH>
  PK_HD void diag_local(const Params& p, const SH& sh, double* s3, double* av, int* ai) const {
    const double h = p.h, ih = p.ih, h3 = p.h * (1.0 / 3.0);
    double sM = 0.0, sK = 0.0, sU = 0.0;
    PK_UNROLL
    for (int c = 0; c < 4; ++c) { av[c] = -1.7976931348623157e308 * 2.0; ai[c] = 0x7fffffff; }
    double um = ldn(p, sh, sh.U0, -1), u0 = ldn(p, sh, sh.U0, 0);
    PK_UNROLL
    for (int j = 0; j < M; ++j) {
      if (has_row(j)) {
        const double up = ldn(p, sh, sh.U0, j + 1);
        const int i = s + j;
        const double du = div_h(p, up - u0), dum = div_h(p, u0 - um);
        const double jmp = fabs(du - dum);
        sM += h3 * (u0 * u0 + u0 * up + up * up);
        sK += (up - u0) * (up - u0) * ih;
        sU += h * u0;
        const double cand[4] = {du, -du, u0, jmp};
        PK_UNROLL
        for (int c = 0; c < 4; ++c)
          if (cand[c] > av[c] || (cand[c] == av[c] && i < ai[c])) { av[c] = cand[c]; ai[c] = i; }
        um = u0; u0 = up;
      }
    }
    s3[0] = sM; s3[1] = sK; s3[2] = sU;
  }
  // the two scans of peakon_diagnostics (outputting.py:401-435) as "smallest k" searches (pk_peakon_suite_kernel):
  //   k2[0]: smallest k in [0, ne/2) with du[D] > du[D-1], D = pc + k;   k2[1]: ... with du[D] < du[D+1], D = pc - 1 - k
  template <class SH>
  PK_HD void suite_local(const Params& p, const SH& sh, int pc, int* k2) const {
    const int ne = p.ne, S = ne / 2;
    k2[0] = k2[1] = 0x7fffffff;
    double u0 = ldn2(p, sh, sh.U0, 0), up = ldn2(p, sh, sh.U0, 1);
    double dum = div_h(p, u0 - ldn2(p, sh, sh.U0, -1)), du0 = div_h(p, up - u0);
    PK_UNROLL
    for (int j = 0; j < M; ++j) {
      if (has_row(j)) {
        const double upp = ldn2(p, sh, sh.U0, j + 2);
        const double dup = div_h(p, upp - up);
        const int i = s + j;
        if (du0 > dum) { int k = i - pc; if (k < 0) k += ne; if (k < S && k < k2[0]) k2[0] = k; }
        if (du0 < dup) { int k = pc - 1 - i; if (k < 0) k += ne; if (k < S && k < k2[1]) k2[1] = k; }
        dum = du0; du0 = dup; up = upp;
      }
    }
  }

  // ---- NOISE_SCALAR: the dXi value of this step ------------------------------------------------
  template <class SH>
  PK_HD void set_scalar_noise(const Params& p, const SH& sh) {
    Xs = (p.nxi > 0) ? sh.dw(0) * p.xi_const : 0.0;
  }
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator + Box-Muller. Counter =
// (global realisation id, step, mode, sub-update), key = 64-bit seed: results do not depend on how
// the ensemble is sharded over GPUs.
// ------------------------------------------------------------------------------------------------
PK_HD void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned* out) {
  PK_UNROLL
  for (int i = 0; i < 10; ++i) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
    const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0;
    const unsigned n1 = (unsigned)p1;
    const unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1;
    const unsigned n3 = (unsigned)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

PK_HD double philox_normal(unsigned long long seed, unsigned rid, unsigned step, unsigned mode, unsigned sub) {
  unsigned o[4];
  philox4x32_10(rid, step, mode, sub, (unsigned)seed, (unsigned)(seed >> 32), o);
  const double two53 = 1.0 / 9007199254740992.0;
  const double u1 = ((double)(((unsigned long long)(o[0] >> 5) << 26) | (o[1] >> 6)) + 0.5) * two53;
  const double u2 = ((double)(((unsigned long long)(o[2] >> 5) << 26) | (o[3] >> 6)) + 0.5) * two53;
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}


// ------------------------------------------------------------------------------------------------
// Brownian increments of step n into sh.dW[0..nxi): read from dW_in or drawn with Philox.  With
// nxi <= nth every thread of a block owns one (step, mode) pair of a window of nth/nxi steps and keeps
// its value in a register (`hold`), so the generator runs once per window on all threads instead of
// once per step on nxi threads.  Every block of a cluster draws the same numbers; the leader exports.
// ------------------------------------------------------------------------------------------------
// One Brownian increment of the Philox path: the sum of the nupd sub-update draws (stochastic_functions.py:213-217), each
// scaled by sqrt(dt/nupd) before it is added.  Sub-update j of step n is the draw of FINE step n * nupd + j (counter word 1),
// so a run at dt with nXi_updates = dt/dt0 consumes the same Brownian path as a run at dt0 with nXi_updates = 1 -- the
// property of the reference's sequential MT19937 stream that its strong-convergence study relies on (:211-217).  Summation order (part of the definition: every code path -- step kernels,
// peakon-SDE kernel, emulation -- produces the same bits):
//   nupd <  PK_COOP_NUPD : j = 0, 1, ..., nupd-1 in sequence;
//   nupd >= PK_COOP_NUPD : 32 interleaved partial sums (lane l adds j = l, l+32, ... in sequence), added up in lane
//                          order 0, 1, ..., 31 -- what a warp produces when its 32 lanes share the draws.
enum { PK_COOP_NUPD = 64 };
PK_HD double philox_partial(unsigned long long seed, unsigned rid, unsigned step, unsigned k, int nupd, double sdt, int lane) {
  double s = 0.0;
  for (int j = lane; j < nupd; j += 32) s += philox_normal(seed, rid, step * (unsigned)nupd + (unsigned)j, k, 0u) * sdt;
  return s;
}
PK_HD double philox_increment(unsigned long long seed, unsigned rid, unsigned step, unsigned k, int nupd, double sdt) {
  if (nupd < PK_COOP_NUPD) {
    double w = 0.0;
    for (int j = 0; j < nupd; ++j) w += philox_normal(seed, rid, step * (unsigned)nupd + (unsigned)j, k, 0u) * sdt;
    return w;
  }
  double w = philox_partial(seed, rid, step, k, nupd, sdt, 0);
  for (int l = 1; l < 32; ++l) w += philox_partial(seed, rid, step, k, nupd, sdt, l);
  return w;
}

struct DrawIO {
  const double* dW_in; double* dW_out;
  int R, r, step0, nsteps;
  unsigned rid;
  unsigned long long seed;
  int n0 = 0;             // first step (launch-local index) this call sequence executes: a window is (re)drawn there too
  // windowed drawing (nxi <= threads): per = steps per window, w0 = first step of the current window.  Set by
  // draw_begin() once and advanced by draw_advance() every step (no integer division in the step loop).
  int per = 0, w0 = 0;
  int coop = 0;           // many sub-updates: a WARP owns a step of the window and shares the draws of each increment
};
template <bool COOP = true>
PK_HD void draw_begin(const Params& p, DrawIO& d, int nth) {
  if (p.nxi <= 0 || p.nxi > nth) return;
  d.coop = (COOP && !d.dW_in && p.nupd >= PK_COOP_NUPD && p.nxi <= 32 && nth >= 32) ? 1 : 0;
  d.per = d.coop ? (nth >> 5) : nth / p.nxi;
  d.w0 = d.n0 - (d.n0 % d.per);
}
PK_HD void draw_advance(DrawIO& d, int n) { if (d.per > 0 && n - d.w0 >= d.per) d.w0 += d.per; }
PK_HD DrawSlot draw_slot(const Params& p, int tid) {
  DrawSlot s;
  if (p.nxi > 0) { s.mine = tid / p.nxi; s.k = tid - s.mine * p.nxi; }
  return s;
}
PK_HD double draw_one(const Params& p, const DrawIO& d, int n, int k) {
  const long long o = ((long long)n * d.R + d.r) * p.nxi + k;
  if (d.dW_in) return d.dW_in[o];
  return philox_increment(d.seed, d.rid, (unsigned)(d.step0 + n), (unsigned)k, p.nupd, p.sdt);
}
template <bool COOP = true, class SH>
PK_HD void draw_step(const Params& p, const SH& sh, const DrawIO& d, int n, int tid, int nth, bool leader, double& hold, const DrawSlot& slot) {
  if (p.nxi <= 0) return;
  if (p.nxi > nth) {                       // more modes than threads: plain loop
    for (int k = tid; k < p.nxi; k += nth) {
      const double w = draw_one(p, d, n, k);
      sh.set_dw(k, w);
      if (leader && d.dW_out) d.dW_out[((long long)n * d.R + d.r) * p.nxi + k] = w;
    }
    return;
  }
  const int per = d.per, w0 = d.w0;        // steps per window, first step of the current window
  if (COOP && d.coop) {
    // warp w of the block owns step w0 + w of the window; its lanes share the nupd draws of every increment and lane k
    // keeps the increment of mode k (philox_increment's summation order)
    const int w = tid >> 5, lane = tid & 31;
    if ((n == w0 || n == d.n0) && w0 + w < d.nsteps) {
      for (int k = 0; k < p.nxi; ++k) {
#if defined(__CUDA_ARCH__)
        const double part = philox_partial(d.seed, d.rid, (unsigned)(d.step0 + w0 + w), (unsigned)k, p.nupd, p.sdt, lane);
        double s = __shfl_sync(0xffffffffu, part, 0);
        for (int l = 1; l < 32; ++l) s += __shfl_sync(0xffffffffu, part, l);
        if (lane == k) hold = s;
#else
        if (lane == k) hold = philox_increment(d.seed, d.rid, (unsigned)(d.step0 + w0 + w), (unsigned)k, p.nupd, p.sdt);
#endif
      }
    }
    if (w == n - w0 && lane < p.nxi) {
      sh.set_dw(lane, hold);
      if (leader && d.dW_out) d.dW_out[((long long)n * d.R + d.r) * p.nxi + lane] = hold;
    }
    return;
  }
  const int mine = slot.mine, k = slot.k;
  if ((n == w0 || n == d.n0) && mine < per && w0 + mine < d.nsteps) hold = draw_one(p, d, w0 + mine, k);
  if (mine == n - w0 && mine < per) {
    sh.set_dw(k, hold);
    if (leader && d.dW_out) d.dW_out[((long long)n * d.R + d.r) * p.nxi + k] = hold;
  }
}

// ------------------------------------------------------------------------------------------------
// Block-level orchestration, templated on an executor that provides
//   each(f)            run f(Thread&) on every thread of the realisation
//   sync()             barrier
//   allreduce3(f, out) out[0..2] = sum over threads of the three partials f(th, v) fills
//   scan_fwd(get, put) cyclic exclusive scan of the affine maps x -> A x + c (two channels sharing A)
//                      that get() returns, in thread order 0..T-1; put() receives the value entering
//                      its thread and returns the maps of the following reverse scan
//   scan_rev(fin)      the reverse cyclic scan (order T-1..0); fin() receives the entering value
//   each_tid(f)        run f(th, tid, nthreads, leader, g0, g1) on every thread of every block working on the
//                      realisation: block-local thread id / count, whether the block is the cluster
//                      leader, and the range of mesh rows the block owns
//   one(f)             run f() once (thread 0)
//   pingpong()         true if the exchange area has XB_SLOTS_PP slots (second PCR row buffer)
//   hybrid(T)          true if the executor supports the compacted Schur solve for T threads per realisation:
//   each_first(n, f)     run f on the threads t < n only (the others fall through to the next barrier)
//   warp0_steps(n, f)    threads t < 32: for step = 0..n-1 { f(th, step); warp-level barrier }
// ------------------------------------------------------------------------------------------------
struct StepIO {
  const double* dW_in;    // [nsteps, R, nxi] or null -> Philox
  double* dW_out;         // [nsteps, R, nxi] or null
  const double *xi, *xi_x, *xi_xx;   // [nxi, T*M] chunk-interleaved copies of the [nxi, ne] tables (permute_table)
  int R;                  // realisations in this launch (stride of dW_in/out)
  int r;                  // this realisation (local index)
  unsigned rid;           // global realisation id (Philox counter)
  unsigned long long seed;
  int step0;              // global index of the first step of this launch
  int nsteps;             // steps of the whole launch (extent of dW_in / dW_out)
  int n0, n1;             // launch-local range [n0, n1) of steps this call advances (one work item)
  int* fail_step;         // [R] in/out: -1 = healthy, else global step index at which SNES diverged
  int* its_last;          // [R] out: Newton iterations of the last attempted step
  int* its_total;         // [R] in/out: accumulated Newton iterations
  int* fail_reason;       // [R] out: 0 ok, 1 max_it, 2 nan, 3 dtol
  // fused dumps: after every diag_every-th step of the launch (launch-local step n with (n+1) % diag_every == 0) the
  // diagnostics of the state are written to row d = (n+1)/diag_every - 1 of these arrays (0 = off)
  int diag_every = 0;
  double* diag_out = nullptr;     // [ndumps, R, DIAG_NCOL]
  int* diag_idx = nullptr;        // [ndumps, R, DIAG_NIDX]
  double* suite_vals = nullptr;   // [ndumps, R, 2] or null
  int* suite_ids = nullptr;       // [ndumps, R, 3] or null
  // state snapshots: after every snap_every-th step of the launch, u goes to row d = (n+1)/snap_every - 1 of snap_u
  // (natural node order) and the realisation's fail state to snap_flag (-1 healthy, else the failing step): the dump
  // diagnostics are then computed for all dumps of the launch by ONE batched streaming kernel (pk_diagnostics_rows)
  int snap_every = 0;
  double* snap_u = nullptr;       // [nsnap, R, ne]
  int* snap_flag = nullptr;       // [nsnap, R]
};

// column layout of the fused dump == include/peakon_b200.h (PK_D_*, PK_I_*); static_assert'ed in pk_lib.cu
enum { DIAG_NCOL = 14, DIAG_NIDX = 4,
       DG_ENERGY = 0, DG_L2U2 = 1, DG_H1SEMI2 = 2, DG_MASS_U = 3, DG_MAX_DU = 4, DG_MIN_DU = 5, DG_MAX_JUMP = 6, DG_MAX_U = 7,
       DGI_ARGMAX_DU = 0, DGI_ARGMIN_DU = 1, DGI_ARGMAX_U = 2, DGI_ARGMAX_JUMP = 3 };
PK_HD double nan_quiet() {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(0x7ff8000000000000LL);
#else
  return std::nan("");
#endif
}

// Snapshot of u (natural node order) for the batched dump diagnostics: every block stores the mesh rows it owns.
template <int M, int TT, class Exec, class Th, class SH>
PK_HD void snapshot(Exec& ex, const Params& p, const SH& sh, const StepIO& io, int d, int failed) {
  const long long row = (long long)d * io.R + io.r;
  double* dst = io.snap_u + row * p.ne;
  ex.each_tid([&](Th&, int tid, int nth, bool, int g0, int g1) {
    // (rows [g0, g1) belong to the calling CTA: block-local slot, no decoding of a flat index)
    for (int i = g0 + tid; i < g1; i += nth) { int j, t; row_jt_c<M>(p, i, j, t); dst[i] = sh.ldc(sh.U0, j, TT > 0 ? TT : p.T, t); }
  });
  ex.one([&]() { io.snap_flag[row] = failed; });
}

// One fused dump of the realisation whose state is resident in `sh` (executor collectives: argmax4sum3, minint2).
template <int M, int MODE, int TT, class Exec, class Th, class SH>
PK_HD void diag_dump(Exec& ex, const Params& p, const SH& sh, const StepIO& io, int d, bool failed) {
  const long long row = (long long)d * io.R + io.r;
  double* out = io.diag_out + row * DIAG_NCOL;
  int* idx = io.diag_idx + row * DIAG_NIDX;
  if (failed) {          // outputting.py:107-108
    ex.one([&]() {
      for (int k = 0; k < DIAG_NCOL; ++k) out[k] = nan_quiet();
      for (int k = 0; k < DIAG_NIDX; ++k) idx[k] = -1;
      if (io.suite_vals) { io.suite_vals[2 * row] = nan_quiet(); io.suite_vals[2 * row + 1] = nan_quiet(); }
      if (io.suite_ids) { io.suite_ids[3 * row] = -1; io.suite_ids[3 * row + 1] = -1; io.suite_ids[3 * row + 2] = -1; }
    });
    return;
  }
  double sums[3], av[4];
  int ai[4];
  // one pass over the rows, ONE collective for the four arg-reductions and the three sums
  ex.argmax4sum3([&](Th& th, double* v, int* i, double* s3) { th.diag_local(p, sh, s3, v, i); }, av, ai, sums, sh.xb);
  int k2[2] = {0x7fffffff, 0x7fffffff};
  if (io.suite_vals) ex.minint2([&](Th& th, int* k) { th.suite_local(p, sh, ai[2], k); }, k2, sh.xb);
  ex.one([&]() {
    out[DG_ENERGY] = sums[0] + p.alphasq * sums[1];
    out[DG_L2U2] = sums[0]; out[DG_H1SEMI2] = sums[1]; out[DG_MASS_U] = sums[2];
    out[DG_MAX_DU] = av[0]; out[DG_MIN_DU] = -av[1]; out[DG_MAX_JUMP] = av[3]; out[DG_MAX_U] = av[2];
    for (int k = DG_MAX_U + 1; k < DIAG_NCOL; ++k) out[k] = nan_quiet();      // m- and SDE-based columns: pk_diagnostics
    idx[DGI_ARGMAX_DU] = ai[0]; idx[DGI_ARGMIN_DU] = ai[1]; idx[DGI_ARGMAX_U] = ai[2]; idx[DGI_ARGMAX_JUMP] = ai[3];
    if (io.suite_vals) {
      const int ne = p.ne, pc = ai[2];
      auto uat = [&](int i) { return sh.ld(sh.U0, row_slot_c<M, TT>(p, i)); };
      auto du = [&](int i) { i %= ne; if (i < 0) i += ne; const int ip = (i + 1 == ne) ? 0 : i + 1; return (uat(ip) - uat(i)) / p.h; };
      double* vals = io.suite_vals + 2 * row;
      int* ids = io.suite_ids + 3 * row;
      ids[0] = pc;
      if (k2[0] != 0x7fffffff) {
        const int D = (pc + k2[0]) % ne; int j = D - 1; if (j < 0) j += ne;
        vals[0] = du(j); ids[1] = j;
      } else { vals[0] = nan_quiet(); ids[1] = -1; }
      if (k2[1] != 0x7fffffff) {
        int D = (pc - 1 - k2[1]) % ne; if (D < 0) D += ne;
        int j = D - 1; if (j < 0) j += ne;
        vals[1] = du(D + 1); ids[2] = j;
      } else { vals[1] = nan_quiet(); ids[2] = -1; }
    }
  });
  // (the exchange area served as scratch; the next step does not write it before its first barrier)
}

// SNES convergence test (SNESConvergedDefault + divergence tolerance).
// returns 0 = keep iterating, 1 = converged, 2 = diverged (reason set)
// The comparisons are made on the squared 2-norms (no square roots on the device); fnorm0sq keeps |F_0|^2.
PK_HD int snes_test(const Params& p, int it, const double* nrm, double& fnorm0sq, int& reason) {
  const double f2 = nrm[0];
  if (!(f2 <= 1.7976931348623157e308)) { reason = 2; return 2; }      // NaN or Inf
  if (it == 0) {
    fnorm0sq = f2;
    if (f2 < p.atol2) return 1;
  } else {
    if (f2 < p.atol2 || f2 <= p.rtol2 * fnorm0sq) return 1;
    if (nrm[1] < p.stol2 * nrm[2]) return 1;
    if (f2 > p.divtol2 * fnorm0sq) { reason = 3; return 2; }
  }
  if (it >= p.max_it) { reason = 1; return 2; }
  return 0;
}

#if defined(__CUDA_ARCH__)
#define PK_LD_COHERENT(ptr) __ldcg(ptr)      // L2 (coherent across SMs inside one launch)
#else
#define PK_LD_COHERENT(ptr) (*(ptr))
#endif

// FEAT: optional features compiled into the time loop.  bit 0: fused dump diagnostics (diag_dump); bit 1: warp-shared
// Philox sub-updates (nupd >= PK_COOP_NUPD).  The headline launches (no dumps inside the launch, nXi_updates < 64) use the
// lean instantiation: the mere presence of the two blocks in the loop costs 4 % of the single-CTA kernel (code placement /
// scheduling; measured, profiles/r2c_variants.log), although neither is executed.
enum { FEAT_DUMP = 1, FEAT_COOP = 2, FEAT_ALL = 3 };
template <int M, int MODE, int TT, class Exec, class Th, class SH, int FEAT = FEAT_ALL>
PK_HD void run_steps(Exec& ex, const Params& p, const SH& sh, const StepIO& io) {
  // counters may have been written by another SM earlier in this launch (previous step block)
  int failed = PK_LD_COHERENT(io.fail_step + io.r);
  int its_last = io.n0 > 0 ? PK_LD_COHERENT(io.its_last + io.r) : 0;   // later step blocks of a launch continue
  int its_total = PK_LD_COHERENT(io.its_total + io.r), reason = 0;
  DrawIO dio;
  dio.dW_in = io.dW_in; dio.dW_out = io.dW_out; dio.R = io.R; dio.r = io.r; dio.step0 = io.step0;
  dio.nsteps = io.nsteps; dio.rid = io.rid; dio.seed = io.seed; dio.n0 = io.n0;
  int expect = 2;     // iteration at which convergence is expected: there only the residual is evaluated first
  int cert_skip = 0;  // steps during which the convergence certificate is not attempted
  ex.each_tid([&](Th& th, int tid, int nth, bool, int, int) { th.dw_slot = draw_slot(p, tid); if (dio.per == 0) draw_begin<(FEAT & FEAT_COOP) != 0>(p, dio, nth); });
  for (int n = io.n0; n < io.n1; ++n) {
#ifndef PK_EXP_NO_KILL_ALL
    ex.each([&](Th& th) { th.kill_all(); });
#endif
    // Brownian increments are produced (and exported) every step, failed or not (simulation.py:96)
    draw_advance(dio, n);
    ex.each_tid([&](Th& th, int tid, int nth, bool leader, int, int) {
      draw_step<(FEAT & FEAT_COOP) != 0>(p, sh, dio, n, tid, nth, leader, th.dw_hold, th.dw_slot);
    });
    ex.sync_local();        // the increments live in CTA-local shared memory (every CTA of a cluster draws its own copy)
    ex.tick(0);
    if (failed >= 0) {
      if constexpr ((FEAT & FEAT_DUMP) != 0) {
        if (io.diag_every > 0 && (n + 1) % io.diag_every == 0)
          diag_dump<M, MODE, TT, Exec, Th, SH>(ex, p, sh, io, (n + 1) / io.diag_every - 1, true);
        if (io.snap_every > 0 && (n + 1) % io.snap_every == 0)
          snapshot<M, TT, Exec, Th, SH>(ex, p, sh, io, (n + 1) / io.snap_every - 1, failed);
      }
      continue;
    }
    // Barriers between two steps: the neighbours' halo reads of the next assembly need ONE (cluster-wide) barrier after
    // accept() and after the noise arrays are filled.  Scalar noise lives in a register: the barrier that closes the
    // previous step does it; noise arrays: the barrier below does, and the closing one is dropped (see the end of the step).
    if (PK_NOISE_OF(MODE) == NOISE_SCALAR) ex.each([&](Th& th) { th.set_scalar_noise(p, sh); });
    else { ex.each([&](Th& th) { th.fill_noise(p, sh, io.xi, io.xi_x, io.xi_xx); }); ex.sync(); }
    ex.tick(1);

    if (cert_skip > 0) --cert_skip;
    // ---- SNES newtonls, basic line search ----
    int it = 0, status = 0;
    double fnorm0 = 0.0, nrm[3];
    while (true) {
      if (it > 0 && it >= expect) {
        // Convergence is expected here.  The residual is a QUADRATIC map of w (N1 and f are quadratic forms of uh, the
        // rest is linear) and the Jacobian is exact, so R(w + d) = R(w) + J d + Q(d/2) identically, and the Newton
        // step just taken makes the first two terms cancel: |R(w_new)|_2 = |Q(e)|_2, e = du/2.  Cell by cell,
        //   |N1 part| <= Dt |a1-a0| |2a0+a1| / 6 <= 0.6 Dt (a0^2+a1^2),
        //   |f part|  <= (Dt h/3 + a2 Dt/(2h)) (a0^2+a1^2),
        // every cell feeds two nodes and sum_cells (a0^2+a1^2) = 2 |e|^2, so with |.|_2 <= |.|_1
        //   |R(w_new)|_2 <= Dt (2.4 + 4h/3 + 2 a2/h) |e|_2^2 <= cert_k |d|_2^2,   cert_k = Dt (2.4 + 4h/3 + 2 a2/h) / 4.
        // If that bound is four times below the rtol threshold -- and the threshold itself ten times above a generous
        // estimate of the rounding noise of an evaluation of R -- SNES's test `|R| <= rtol |R0|` is decided without
        // evaluating R: same iterates, same iteration count (checked against the oracle in every parity test).
        // Otherwise fall through.
        if (p.cert_k > 0.0 && cert_skip == 0) {
          double n2[3];
          ex.allreduce3_nb([&](Th& th, double* v) { v[0] = 0.0; v[1] = th.sn2; v[2] = th.xn2; }, n2);
          // + an allowance for the linear solve: relative residual 1e-6 of the system just solved (measured: < 1e-9)
          const double B = p.cert_k * n2[1] + 1e-6 * sqrt(nrm[0]), thr2 = p.rtol2 * fnorm0;
          if (B * B <= 0.0625 * thr2 && thr2 >= 100.0 * p.cert_floor2 * ((double)p.ne + n2[2])) {
#ifdef PK_CERT_CHECK
            // checking build (profiles/cert_check.py): evaluate the residual the certificate claims to bound and FAIL the
            // realisation (reason 9) if SNES's own test would not have accepted it -- the certificate must be iterate-neutral
            ex.each([&](Th& th) { th.assemble_begin(p, sh, true); });
            if constexpr (Th::ROLL_LITE) ex.each([&](Th& th) { th.residual_finish_rolled(p, sh); });
            else ex.each([&](Th& th) { th.assemble_finish(p, sh, true); });
            double chk[3];
            ex.allreduce3([&](Th& th, double* v) { v[0] = th.fn2; v[1] = th.sn2; v[2] = th.xn2; }, chk);
            if (!(chk[0] <= thr2)) { status = 2; reason = 9; break; }
#ifdef PK_CERT_CHECK_BOUND      // also: the evaluated |R| (which carries evaluation noise ~ sqrt(cert_floor2 (ne + |w|^2))) within the bound
            { const double fl = sqrt(p.cert_floor2 * ((double)p.ne + n2[2])); if (!(sqrt(chk[0]) <= B + 10.0 * fl)) { status = 2; reason = 8; break; } }
#endif
#endif
            status = 1; break;
          }
          cert_skip = 8;      // not provable here (coarse bound / large noise): do not pay the reduction again for a few steps
          // the residual evaluation below reads the neighbours' new iterate: the barrier that was left to the all-reduce
          if (ex.barrier_free_gather()) ex.sync();
        }
        // evaluate the residual only; fall through to the full
        // assembly (which recomputes it) in the rare case it is not
        ex.each([&](Th& th) { th.assemble_begin(p, sh, true); });
        if constexpr (Th::ROLL_LITE) ex.each([&](Th& th) { th.residual_finish_rolled(p, sh); });
        else ex.each([&](Th& th) { th.assemble_finish(p, sh, true); });
        ex.allreduce3([&](Th& th, double* v) { v[0] = th.fn2; v[1] = th.sn2; v[2] = th.xn2; }, nrm);
        status = snes_test(p, it, nrm, fnorm0, reason);
        ex.tick(2);
        if (status) break;
      }
      ex.each([&](Th& th) { th.assemble_begin(p, sh, false); });
      ex.sync();
      ex.tick(3);
      if constexpr (Th::FS || Th::ROLL) ex.each([&](Th& th) { th.assemble_finish_fs(p, sh, false); });
      else ex.each([&](Th& th) { th.assemble_finish(p, sh, false); });
      ex.tick(4);
      ex.allreduce3([&](Th& th, double* v) { v[0] = th.fn2; v[1] = th.sn2; v[2] = th.xn2; }, nrm);
      status = snes_test(p, it, nrm, fnorm0, reason);
      ex.tick(5);
      if (status) break;
      // ---- linear solve + update ----
      ex.each([&](Th& th) { th.reduce_build(p, sh); });
      // the publish below reuses the slots the neighbours read their tips from: a barrier first, unless there is a second
      // row buffer.  (Single buffer + partitioned solve: the edge threads exchange the tips through a separate area, see
      // SharedClusterT, so the readers of these slots are all in the own CTA.)
      if (!ex.pingpong()) { if (ex.partitioned()) ex.sync_local(); else ex.sync(); }
      const int Tn = (TT > 0 ? TT : p.T);
      if (ex.hybrid(Tn)) {
        // cyclic reduction (compacted onto the leading threads) -> 32 rows -> PCR inside warp 0 -> back-substitution
        ex.each([&](Th& th) { th.hyb_publish(p, sh); });
        ex.sync();
        ex.tick(6);
        int S = 1;
        for (; Tn / (2 * S) >= 32; S *= 2) {
          ex.each_first(Tn / (2 * S), [&](Th& th) { th.hyb_down(p, sh, S); });
          ex.sync();
        }
        ex.warp0_steps(10, [&](Th& th, int step) { th.hyb_warp_step(p, sh, S, step); });
        ex.sync();
        ex.tick(7);
        for (int s2 = S / 2; s2 >= 1; s2 /= 2) {
          ex.each_first(Tn / (2 * s2), [&](Th& th) { th.hyb_up(p, sh, s2); });
          ex.sync();
        }
        ex.each([&](Th& th) { th.hyb_load_xs(p, sh); });
        ex.tick(8);
      } else {
        int base = 0;
        if (ex.pingpong()) {
          // second row buffer: the first publish does not overwrite the tips the neighbours are still reading, so it
          // needs no barrier after reduce_build (whose barrier below is reused)
          base = XB_PP_BASE;
          ex.each([&](Th& th) { th.reduce_publish(p, sh, XB_PP_BASE); });
        } else {
          ex.each([&](Th& th) { th.reduce_publish(p, sh); });
        }
        // single row buffer + partitioned solve: the barrier after reduce_build was cluster-wide (the neighbours have read
        // the tips that the publish overwrote); the first PCR steps read rows of the own CTA only
        if (ex.partitioned() && !ex.pingpong()) ex.sync_local(); else ex.sync();
        ex.tick(6);
        if (ex.partitioned()) {
          // segment-local PCR (block-local barriers and loads), then PCR among the C separators
          const int Tc = ex.tc(), Cn = Tn / Tc;
          // separators solved redundantly in every CTA when the idle row buffer (or, with a single row buffer, the
          // scratch area behind the exchange slots) can hold the scratch rows
          const bool red = ex.redundant() && (!ex.pingpong() || 14 * Tc >= Th::sep_scratch_doubles(Cn));
          for (int d = 1; d <= Tc - 2; d *= 2) {
            if (ex.pingpong()) {
              ex.each([&](Th& th) { th.pcr_step_local(p, sh, d, base, Tc); th.reduce_publish(p, sh, XB_PP_BASE - base, red && (th.t & (Tc - 1)) == 0); });
              base = XB_PP_BASE - base;
              if (2 * d <= Tc - 2) ex.sync_local();
            } else {
              // one row buffer (18 exchange slots instead of 30: room for a second CTA per SM): read, barrier, publish
              ex.each([&](Th& th) { th.pcr_step_local(p, sh, d, 0, Tc); });
              ex.sync_local();
              ex.each([&](Th& th) { th.reduce_publish(p, sh, 0, red && (th.t & (Tc - 1)) == 0); });
              if (2 * d <= Tc - 2) ex.sync_local();
            }
          }
          ex.sync();
          ex.tick(13);
          if (red) {
            // scratch of the separator solve: the idle row buffer, or the area behind the XB_SLOTS exchange slots
            const int scr = ex.pingpong() ? (XB_PP_BASE - base) * Tc : XB_SLOTS * Tc;
            ex.each_cta_first(Cn, [&](Th& th) { th.sep_build(p, sh, base, Tc, Cn, scr); });
            ex.sync_local();
            int half = 0;
            for (int d = 1; d < Cn / 2; d *= 2) {
              ex.each_cta_first(Cn, [&](Th& th) { th.sep_step(sh, d, half, Tc, Cn, scr); });
              half ^= 1;
              ex.sync_local();
            }
            ex.tick(7);
            ex.each_cta_first(Cn, [&](Th& th) { th.sep_final(sh, half, Tc, Cn, scr); });
            ex.sync_local();
            ex.tick(14);
            ex.each([&](Th& th) { th.part_back_local(p, sh, Tc, Cn, scr, base); });
          } else {
          ex.each([&](Th& th) { if ((th.t & (Tc - 1)) == 0) { th.pcr_step(p, sh, 1, base); th.reduce_publish(p, sh, XB_PP_BASE - base); } });
          base = XB_PP_BASE - base;
          ex.sync();
          for (int d = Tc; d < Tn / 2; d *= 2) {
            ex.each([&](Th& th) { if ((th.t & (Tc - 1)) == 0) { th.pcr_step(p, sh, d, base); th.reduce_publish(p, sh, XB_PP_BASE - base); } });
            base = XB_PP_BASE - base;
            ex.sync();
          }
          ex.tick(7);
          ex.each([&](Th& th) { if ((th.t & (Tc - 1)) == 0) th.pcr_final(p, sh, base); });
          ex.sync();
          ex.tick(14);
          ex.each([&](Th& th) { if ((th.t & (Tc - 1)) != 0) th.part_back(p, sh, Tc); });
          }
        } else if (ex.pingpong()) {
          // two row buffers: a step reads one and publishes into the other -> ONE barrier per PCR step
          for (int d = 1; d < Tn / 2; d *= 2) {
            ex.each([&](Th& th) { th.pcr_step(p, sh, d, base); th.reduce_publish(p, sh, XB_PP_BASE - base); });
            ex.sync();
            base = XB_PP_BASE - base;
          }
        } else {
          for (int d = 1; d < Tn / 2; d *= 2) {
            ex.each([&](Th& th) { th.pcr_step(p, sh, d); });
            ex.sync();
            ex.each([&](Th& th) { th.reduce_publish(p, sh); });
            ex.sync();
          }
        }
        if (!ex.partitioned()) {
          ex.tick(7);
          ex.each([&](Th& th) { th.pcr_final(p, sh, base); });
        }
        // the separator solution of the right neighbour feeds the interior solve: a thread of the own CTA, or -- edge
        // thread of a single-buffer cluster kernel -- the next segment's separator in the CTA's own scratch rows
        if (ex.partitioned() && !ex.pingpong()) ex.sync_local(); else ex.sync();
        ex.tick(8);
      }
      if constexpr (Th::FS) ex.each([&](Th& th) { th.interior_solve_fs(p, sh); });
      else ex.each([&](Th& th) { th.interior_solve(p, sh); });
      ex.sync();
      ex.tick(9);
      if constexpr (PK_IS_TEST(MODE)) {
        ex.each([&](Th& th) { th.test_update(p, sh); });
        ex.sync();
        ex.tick(11);
        ++it;
        continue;
      }
      // forward cyclic scan (causal filter); its carries feed the anticausal local pass, whose maps
      // are scanned in reverse; the reverse carries finish the update.
      ex.scan_fwd([&](Th& th, double& A, double& c0, double& c1) { th.helm_causal_local(p, sh, A, c0, c1); },
                  [&](Th& th, double in0, double in1, double& A, double& c0, double& c1) {
                    th.helm_anticausal_local(p, in0, in1, A, c0, c1); });
      ex.tick(10);
      ex.scan_rev([&](Th& th, double in0, double in1) { th.helm_finish_update(p, sh, in0, in1); });
      // the new iterate must be visible to the neighbours before the next assembly; when the convergence certificate
      // is attempted next, the barrier inside its all-reduce comes first and does it
      if (!((it + 1 >= expect) && p.cert_k > 0.0 && cert_skip == 0)) ex.sync();
      ex.tick(11);
      ++it;
    }
    its_last = it; its_total += it;
    if (status == 1) {
      ex.each([&](Th& th) { th.accept(p, sh); });
      expect = it > 1 ? it : 1;
    } else {
      failed = io.step0 + n;
    }
    // accept() wrote U0: a barrier before anybody reads a neighbour's row -- the one after fill_noise of the next step,
    // unless the noise is a scalar (no barrier there) or a dump / snapshot of this state follows; the kernels place
    // their own barrier between the last step and the write-back
    {
      bool need = (PK_NOISE_OF(MODE) == NOISE_SCALAR);
      if constexpr ((FEAT & FEAT_DUMP) != 0)
        need = need || (io.diag_every > 0 && (n + 1) % io.diag_every == 0) || (io.snap_every > 0 && (n + 1) % io.snap_every == 0);
      if (need) ex.sync();
    }
    ex.tick(12);
    if constexpr ((FEAT & FEAT_DUMP) != 0) {
      if (io.diag_every > 0 && (n + 1) % io.diag_every == 0)
        diag_dump<M, MODE, TT, Exec, Th, SH>(ex, p, sh, io, (n + 1) / io.diag_every - 1, failed >= 0);
      if (io.snap_every > 0 && (n + 1) % io.snap_every == 0)
        snapshot<M, TT, Exec, Th, SH>(ex, p, sh, io, (n + 1) / io.snap_every - 1, failed);
      ex.tick(15);
    }
  }
  ex.one([&]() {
    io.fail_step[io.r] = failed; io.its_last[io.r] = its_last; io.its_total[io.r] = its_total;
    if (io.fail_reason) io.fail_reason[io.r] = reason;
  });
}

}  // namespace pk
