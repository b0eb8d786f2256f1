// pk_emu.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Executes the *device* code of peakondrake_b200/csrc/pk_core.h on the CPU by running the threads
// of one CTA one after another between barriers.  The authoring container has no GPU; this lets
// the not-gpu test-suite check the kernel's algorithm (assembly closed forms, two-level
// pentadiagonal solve, Helmholtz scans, SNES logic) against the NumPy/SciPy oracle.  Threads are
// run in ascending or descending order (`order`): any intra-phase race would make the two
// orders disagree bit-wise, which tests/test_emulation.py checks.
#include <cstdio>
#include <cstring>
#include <vector>
#include <cmath>
#include "../../peakondrake_b200/csrc/pk_core.h"
#include "../../peakondrake_b200/csrc/pk_upwind.h"
#include "../../peakondrake_b200/csrc/pk_mu.h"

using namespace pk;

static int g_emu_pp = 1;          // cluster emulation: two PCR row buffers (1) or the single-buffer layout of the two-CTAs-per-SM kernels (0)
static int g_emu_part = 1;        // cluster emulation: Schur solve partitioned by CTA (1, like the device) or flat PCR (0)
static int g_emu_cert = 1;        // 0: always evaluate the last residual (no convergence certificate)
static int g_emu_hybrid = 0;      // 1: compacted Schur solve (hyb_* in pk_core.h) where T >= 32; 0: plain PCR
template <class ThT> struct EmuExecT;
template <int M, int MODE> using EmuExec = EmuExecT<Thread<M, MODE> >;

// Emulated thread-block cluster: the flat index space j*T + t is split over C "CTAs" of Tc threads
// exactly like SharedCluster in pk_lib.cu (same decode); `cur` is the CTA of the running thread and
// every store is checked to be local.
struct EmuSharedCluster {
  double *U0, *U1, *F, *G, *X, *Xx, *Xxx, *xb, *dW;    // array ids: offsets into every CTA's private buffer
  std::vector<std::vector<double> >* bufs;
  int Tc, lgTc, lgT, nxi_;
  int* cur;
  int* bad;
  double* at(const double* a, int i, bool store) const {
    const int tg = i & ((1 << lgT) - 1), j = i >> lgT;
    const int owner = tg >> lgTc;
    if (store && owner != *cur) ++*bad;
    return &(*bufs)[owner][(a - (const double*)0) + (size_t)(j << lgTc) + (tg & (Tc - 1))];
  }
  double ld(const double* a, int i) const { return *at(a, i, false); }
  void st(double* a, int i, double v) const { *at(a, i, true) = v; }
  // slot c of the ring neighbour q of the calling thread (the device reads these through per-thread pointers)
  double ldL(const double* a, int c, int T, int q) const { return *at(a, c * T + q, false); }
  double ldR(const double* a, int c, int T, int q) const { return *at(a, c * T + q, false); }
  // tips / separator solution of a ring neighbour: plain slots here (the single-buffer device kernels route their edge
  // threads through CTA-local copies of the same values)
  void stTip(int c, int T, int t, double v) const { *at(xb, c * T + t, true) = v; }
  double ldTip(int c, int T, int q) const { return *at(xb, c * T + q, false); }
  double ldXsR(int c, int T, int q) const { return *at(xb, (14 + c) * T + q, false); }
  // pairs of doubles (PCR rows): pair slot q of thread tt at doubles 2 (q Tc + tt_local), +1 in tt's CTA
  void ld2(const double* a, int q, int T, int tt, double& v0, double& v1) const {
    const int owner = tt >> lgTc;
    const size_t o = (a - (const double*)0) + 2 * ((size_t)(q << lgTc) + (tt & (Tc - 1)));
    v0 = (*bufs)[owner][o]; v1 = (*bufs)[owner][o + 1];
  }
  void ld2l(const double* a, int q, int T, int tt, double& v0, double& v1) const {     // must be block-local
    if ((tt >> lgTc) != *cur) ++*bad;
    ld2(a, q, T, tt, v0, v1);
  }
  void st2o(double* a, int q, int T, int t, double v0, double v1) const {
    const int owner = (t & ((1 << lgT) - 1)) >> lgTc;
    if (owner != *cur) ++*bad;
    const size_t o = (a - (double*)0) + 2 * ((size_t)(q << lgTc) + (t & (Tc - 1)));
    (*bufs)[owner][o] = v0; (*bufs)[owner][o + 1] = v1;
  }
  // own-slot accessors: must be block-local (checked like the device's sto/ldo, which do not decode)
  double ldc(const double* a, int j, int T, int t) const { if (((t & ((1 << lgT) - 1)) >> lgTc) != *cur) ++*bad; return *at(a, j * T + t, false); }
  double ldo(const double* a, int j, int T, int t) const { if (((t & ((1 << lgT) - 1)) >> lgTc) != *cur) ++*bad; return *at(a, j * T + t, false); }
  void sto(double* a, int j, int T, int t, double v) const { *at(a, j * T + t, true) = v; }
  double ldl(const double* a, int off) const { return (*bufs)[*cur][(a - (const double*)0) + off]; }
  void stl(double* a, int off, double v) const { (*bufs)[*cur][(a - (double*)0) + off] = v; }
  double dw(int k) const { return (*bufs)[*cur][(dW - (double*)0) + k]; }
  void set_dw(int k, double v) const { (*bufs)[*cur][(dW - (double*)0) + k] = v; }
};

template <class ThT>
struct EmuExecT {
  typedef ThT Th;
  std::vector<Th> th;
  int T, order;
  std::vector<double> A, c0, c1;
  EmuExecT(const Params& p, int order_, int C_ = 1) : th(p.T), T(p.T), order(order_), A(p.T), c0(p.T), c1(p.T), ne(p.ne), C(C_), Tc(p.T / C_), cur(0), pp(p) {
    for (int t = 0; t < T; ++t) th[t].init(p, t);
  }
  template <class F> void each(F f) {
    for (int k = 0; k < T; ++k) { const int t = order == 0 ? k : T - 1 - k; cur = t / Tc; f(th[t]); }
  }
  int ne, C, Tc; int cur; Params pp;
  template <class F> void each_tid(F f) {
    for (int k = 0; k < T; ++k) {
      const int t = order == 0 ? k : T - 1 - k;
      const int c = t / Tc; cur = c;
      const int g0 = chunk_start(pp, c * Tc), g1 = (c == C - 1) ? ne : chunk_start(pp, (c + 1) * Tc);
      f(th[t], t % Tc, Tc, c == 0, g0, g1);
    }
  }
  void sync() {}
  void tick(int) {}
  template <class F> void one(F f) { f(); }
  bool any_flag(int* flag) const { return *flag != 0; }
  bool hybrid(int Tn) const { return g_emu_hybrid && C == 1 && Tn >= 32; }
  bool pingpong() const { return C > 1 && g_emu_pp; }      // like the cluster kernels: XB_SLOTS_PP slots, or one row buffer
  bool partitioned() const { return C > 1 && g_emu_part; }
  bool redundant() const { return g_emu_part == 1 || !g_emu_pp; }
  template <class F> void each_cta_first(int n, F f) {
    for (int c = 0; c < C; ++c)
      for (int k = 0; k < n && k < Tc; ++k) { const int kk = order == 0 ? k : n - 1 - k; cur = c; f(th[c * Tc + kk]); }
  }
  int tc() const { return Tc; }
  void sync_local() {}
  template <class F> void each_first(int n, F f) {
    for (int k = 0; k < n; ++k) { const int t = order == 0 ? k : n - 1 - k; cur = t / Tc; f(th[t]); }
  }
  template <class F> void warp0_steps(int n, F f) {
    for (int s = 0; s < n; ++s)
      for (int k = 0; k < 32; ++k) { const int t = order == 0 ? k : 31 - k; f(th[t], s); }
  }
  template <class F> void argmax4sum3(F f, double* outv, int* outi, double* outs, double*) {
    for (int c = 0; c < 4; ++c) { outv[c] = -INFINITY; outi[c] = 0x7fffffff; }
    std::vector<double> part(3 * T);
    for (int k = 0; k < T; ++k) {
      const int t = order == 0 ? k : T - 1 - k; cur = t / Tc;
      double v[4], s3[3] = {0.0, 0.0, 0.0}; int ix[4];
      for (int c = 0; c < 4; ++c) { v[c] = -INFINITY; ix[c] = 0x7fffffff; }
      f(th[t], v, ix, s3);
      for (int c = 0; c < 4; ++c) if (v[c] > outv[c] || (v[c] == outv[c] && ix[c] < outi[c])) { outv[c] = v[c]; outi[c] = ix[c]; }
      for (int c = 0; c < 3; ++c) part[3 * t + c] = s3[c];
    }
    for (int c = 0; c < 3; ++c) { outs[c] = 0.0; for (int t = 0; t < T; ++t) outs[c] += part[3 * t + c]; }   // thread order, whatever `order`
  }
  template <class F> void minint2(F f, int* out, double*) {
    out[0] = out[1] = 0x7fffffff;
    for (int k = 0; k < T; ++k) {
      const int t = order == 0 ? k : T - 1 - k; cur = t / Tc;
      int kk[2] = {0x7fffffff, 0x7fffffff};
      f(th[t], kk);
      if (kk[0] < out[0]) out[0] = kk[0];
      if (kk[1] < out[1]) out[1] = kk[1];
    }
  }
  template <class F> void allreduce3_nb(F f, double* out) { allreduce3(f, out); }
  bool barrier_free_gather() const { return false; }
  template <class F> void allreduce3(F f, double* out) {
    out[0] = out[1] = out[2] = 0.0;
    for (int t = 0; t < T; ++t) { double v[3]; f(th[t], v); out[0] += v[0]; out[1] += v[1]; out[2] += v[2]; }
  }
  // cyclic exclusive scan in thread order (dir=+1: 0..T-1, dir=-1: T-1..0)
  void cyc_scan(int dir, std::vector<double>& in0, std::vector<double>& in1) {
    double At = 1.0, t0 = 0.0, t1 = 0.0;
    for (int k = 0; k < T; ++k) {
      const int t = dir > 0 ? k : T - 1 - k;
      t0 = A[t] * t0 + c0[t]; t1 = A[t] * t1 + c1[t]; At *= A[t];
    }
    double x0 = t0 / (1.0 - At), x1 = t1 / (1.0 - At);
    for (int k = 0; k < T; ++k) {
      const int t = dir > 0 ? k : T - 1 - k;
      in0[t] = x0; in1[t] = x1;
      x0 = A[t] * x0 + c0[t]; x1 = A[t] * x1 + c1[t];
    }
  }
  template <class G, class P> void scan_fwd(G get, P put) {
    each([&](Th& h) { get(h, A[h.t], c0[h.t], c1[h.t]); });
    std::vector<double> in0(T), in1(T);
    cyc_scan(+1, in0, in1);
    each([&](Th& h) { double a, b, c; put(h, in0[h.t], in1[h.t], a, b, c); A[h.t] = a; c0[h.t] = b; c1[h.t] = c; });
  }
  template <class F> void scan_rev(F fin) {
    std::vector<double> in0(T), in1(T);
    cyc_scan(-1, in0, in1);
    each([&](Th& h) { fin(h, in0[h.t], in1[h.t]); });
  }
};

template <int M, int MODE>
static int run_one_cluster(const Params& p, StepIO io, double* u, double* u1, double* dF, double* dG, int order, int C) {
  const int Tc = p.T / C, n = Tc * p.M;
  const int nxb = std::max(cluster_xb_doubles(Tc, false), cluster_xb_doubles(Tc, true));
  std::vector<std::vector<double> > bufs(C, std::vector<double>(7 * n + nxb + (p.nxi > 0 ? p.nxi : 1), 0.0));
  EmuExec<M, MODE> ex(p, order, C);
  int bad = 0;
  EmuSharedCluster sh;
  double* z = (double*)0;
  sh.U0 = z; sh.U1 = z + n; sh.F = z + 2 * n; sh.G = z + 3 * n; sh.X = z + 4 * n; sh.Xx = z + 5 * n; sh.Xxx = z + 6 * n;
  sh.xb = z + 7 * n; sh.dW = z + 7 * n + nxb;
  sh.bufs = &bufs; sh.Tc = Tc; sh.nxi_ = p.nxi; sh.cur = &ex.cur; sh.bad = &bad;
  sh.lgTc = 0; while ((1 << sh.lgTc) < Tc) ++sh.lgTc;
  sh.lgT = 0; while ((1 << sh.lgT) < p.T) ++sh.lgT;
  const long long off = (long long)io.r * p.ne;
  for (int i = 0; i < p.ne; ++i) {
    const int sl = row_slot(p, i);
    ex.cur = (sl & (p.T - 1)) / Tc;
    sh.st(sh.U0, sl, u[off + i]); sh.st(sh.U1, sl, u1[off + i]); sh.st(sh.F, sl, dF[off + i]);
    sh.st(sh.G, sl, (MODE == NOISE_FULL) ? dG[off + i] : 0.0);
  }
  run_steps<M, MODE, 0, EmuExec<M, MODE>, Thread<M, MODE> >(ex, p, sh, io);
  for (int i = 0; i < p.ne; ++i) {
    const int sl = row_slot(p, i);
    u[off + i] = sh.ld(sh.U0, sl); u1[off + i] = sh.ld(sh.U1, sl); dF[off + i] = sh.ld(sh.F, sl);
    if (MODE == NOISE_FULL) dG[off + i] = sh.ld(sh.G, sl);
  }
  return bad;
}

static int g_emu_nblk = 0, g_emu_fs = 0;
template <int M, int MODE, int FS>
static void run_one_t(const Params& p, StepIO io, double* u, double* u1, double* dF, double* dG, int order) {
  const int n = p.T * p.M;
  std::vector<double> buf(7 * n + XB_SLOTS * p.T + (p.nxi > 0 ? p.nxi : 1) + 6 * (p.M - 2) * p.T, 0.0);
  Shared sh;
  sh.U0 = &buf[0]; sh.U1 = &buf[n]; sh.F = &buf[2 * n]; sh.G = &buf[3 * n];
  sh.X = &buf[4 * n]; sh.Xx = &buf[5 * n]; sh.Xxx = &buf[6 * n];
  sh.xb = &buf[7 * n]; sh.dW = &buf[7 * n + XB_SLOTS * p.T];
  sh.fac = FS == 1 ? &buf[7 * n + XB_SLOTS * p.T + (p.nxi > 0 ? p.nxi : 1)] : nullptr;
  const long long off = (long long)io.r * p.ne;
  for (int i = 0; i < p.ne; ++i) {
    const int sl = row_slot(p, i);
    sh.U0[sl] = u[off + i]; sh.U1[sl] = u1[off + i]; sh.F[sl] = dF[off + i];
    sh.G[sl] = (MODE == NOISE_FULL) ? dG[off + i] : 0.0;
  }
  // g_emu_nblk > 0: advance in work items of g_emu_nblk steps, each with fresh per-thread state (what a
  // different persistent CTA sees when it pulls the next block of the realisation from the queue)
  typedef Thread<M, MODE, 0, FS> Th;
  const int nblk = g_emu_nblk > 0 ? g_emu_nblk : io.nsteps;
  for (int b0 = 0; b0 < io.nsteps || b0 == 0; b0 += nblk) {
    EmuExecT<Th> ex(p, order);
    io.n0 = b0; io.n1 = b0 + nblk < io.nsteps ? b0 + nblk : io.nsteps;
    run_steps<M, MODE, 0, EmuExecT<Th>, Th>(ex, p, sh, io);
    if (nblk <= 0) break;
  }
  for (int i = 0; i < p.ne; ++i) {
    const int sl = row_slot(p, i);
    u[off + i] = sh.U0[sl]; u1[off + i] = sh.U1[sl]; dF[off + i] = sh.F[sl];
    if (MODE == NOISE_FULL) dG[off + i] = sh.G[sl];
  }
}
template <int M, int MODE>
static void run_one(const Params& p, StepIO io, double* u, double* u1, double* dF, double* dG, int order) {
  if (g_emu_fs == 1) run_one_t<M, MODE, 1>(p, io, u, u1, dF, dG, order);
  else if (g_emu_fs == 2) run_one_t<M, MODE, 2>(p, io, u, u1, dF, dG, order);
  else if (g_emu_fs == 4) run_one_t<M, MODE, 4>(p, io, u, u1, dF, dG, order);
  else run_one_t<M, MODE, 0>(p, io, u, u1, dF, dG, order);
}

static int g_emuC = 1, g_bad = 0;
template <int M>
static void run_mode(const Params& p, StepIO io, double* u, double* u1, double* dF, double* dG, int order) {
  if (g_emuC > 1) {
    switch (p.mode) {
      case NOISE_SCALAR: g_bad += run_one_cluster<M, NOISE_SCALAR>(p, io, u, u1, dF, dG, order, g_emuC); break;
      case NOISE_X: g_bad += run_one_cluster<M, NOISE_X>(p, io, u, u1, dF, dG, order, g_emuC); break;
      case TEST_SCALAR: g_bad += run_one_cluster<M, TEST_SCALAR>(p, io, u, u1, dF, dG, order, g_emuC); break;
      case TEST_X: g_bad += run_one_cluster<M, TEST_X>(p, io, u, u1, dF, dG, order, g_emuC); break;
      default: g_bad += run_one_cluster<M, NOISE_FULL>(p, io, u, u1, dF, dG, order, g_emuC); break;
    }
    return;
  }
  switch (p.mode) {
    case NOISE_SCALAR: run_one<M, NOISE_SCALAR>(p, io, u, u1, dF, dG, order); break;
    case NOISE_X: run_one<M, NOISE_X>(p, io, u, u1, dF, dG, order); break;
    case TEST_SCALAR: run_one<M, TEST_SCALAR>(p, io, u, u1, dF, dG, order); break;
    case TEST_X: run_one<M, TEST_X>(p, io, u, u1, dF, dG, order); break;
    default: run_one<M, NOISE_FULL>(p, io, u, u1, dF, dG, order); break;
  }
}
// emulate a thread-block cluster of C CTAs per realisation (1 = off); returns the number of
// non-local stores seen so far (must stay 0)
extern "C" int pk_emu_set_cluster(int C) { g_emuC = C; return g_bad; }

// per-thread code variant (Thread<..., VAR>): 0 default, 1 FS (factors in shared memory, rolled row loops),
// 2 rolled assembly with register-resident factors, 4 everything unrolled
extern "C" int pk_emu_set_fs(int on) { g_emu_fs = on; return 0; }

extern "C" int pk_emu_set_hybrid(int on) { g_emu_hybrid = on; return 0; }
extern "C" int pk_emu_set_certificate(int on) { g_emu_cert = on; return 0; }
extern "C" int pk_emu_set_partitioned(int on) { g_emu_part = on; return 0; }
extern "C" int pk_emu_set_pingpong(int on) { g_emu_pp = on; return 0; }

// fused dumps (StepIO::diag_every etc.) for the next pk_emu_step calls; every = 0 switches them off
static int g_dump_every = 0; static double* g_dump_out = nullptr; static int* g_dump_idx = nullptr;
static double* g_dump_sv = nullptr; static int* g_dump_si = nullptr;
extern "C" int pk_emu_set_dump(int every, double* out, int* idx, double* suite_vals, int* suite_ids) {
  g_dump_every = every; g_dump_out = out; g_dump_idx = idx; g_dump_sv = suite_vals; g_dump_si = suite_ids; return 0;
}

static int g_snap_every = 0; static double* g_snap_u = nullptr; static int* g_snap_flag = nullptr;
extern "C" int pk_emu_set_snap(int every, double* u, int* flag) { g_snap_every = every; g_snap_u = u; g_snap_flag = flag; return 0; }

// advance in step blocks of nblk steps (0 = whole launch at once), see run_one
extern "C" int pk_emu_set_block(int nblk) { g_emu_nblk = nblk; return 0; }

extern "C" int pk_emu_layout(int ne, int* T, int* M, int* rem) { return choose_layout(ne, T, M, rem) ? 0 : -1; }
extern "C" int pk_emu_layout_wide(int ne, int* T, int* M, int* rem) { return choose_layout_wide(ne, T, M, rem) ? 0 : -1; }

extern "C" double pk_emu_philox_normal(unsigned long long seed, unsigned rid, unsigned step, unsigned mode, unsigned sub) {
  return philox_normal(seed, rid, step, mode, sub);
}

// Same argument meaning as pk_step() in include/peakon_b200.h, host pointers.
extern "C" int pk_emu_step(int ne, double Ld, double dt, double alphasq, int mode, int nxi, int nupd,
                           double xi_const, int R, int nsteps, int step0,
                           double* u, double* u1, double* dF, double* dG,
                           const double* xi, const double* xi_x, const double* xi_xx,
                           const double* dW_in, double* dW_out, unsigned long long seed, unsigned rid0,
                           int* fail_step, int* its_last, int* its_total, int* fail_reason,
                           int order, int forceT, int forceM) {
  Params p;
  memset(&p, 0, sizeof(p));
  p.ne = ne; p.h = Ld / ne; p.dt = dt; p.alphasq = alphasq; p.mode = mode; p.nxi = nxi; p.nupd = nupd;
  p.max_it = 50; p.rtol = 1e-8; p.atol = 1e-50; p.stol = 1e-8; p.divtol = 1e4; p.xi_const = xi_const;
  if (forceT > 0) { p.T = forceT; p.M = forceM; p.rem = ne - forceT * (forceM - 1); if (p.rem < 0 || p.rem > p.T) return -2; }
  else if (!choose_layout(ne, &p.T, &p.M, &p.rem)) return -1;
  fill_constants(p);
  if (!g_emu_cert) p.cert_k = 0.0;
  std::vector<double> pt[3];
  const double* src[3] = {xi, xi_x, xi_xx};
  const double* tab[3] = {nullptr, nullptr, nullptr};
  for (int q = 0; q < 3; ++q) if (src[q] && nxi > 0) { pt[q].resize((size_t)nxi * p.T * p.M); permute_table_host(p, src[q], pt[q].data()); tab[q] = pt[q].data(); }
  for (int r = 0; r < R; ++r) {
    StepIO io;
    io.dW_in = dW_in; io.dW_out = dW_out; io.xi = tab[0]; io.xi_x = tab[1]; io.xi_xx = tab[2];
    io.R = R; io.r = r; io.rid = rid0 + r; io.seed = seed; io.step0 = step0; io.nsteps = nsteps;
    io.n0 = 0; io.n1 = nsteps;
    io.fail_step = fail_step; io.its_last = its_last; io.its_total = its_total; io.fail_reason = fail_reason;
    io.diag_every = g_dump_every; io.diag_out = g_dump_out; io.diag_idx = g_dump_idx; io.suite_vals = g_dump_sv; io.suite_ids = g_dump_si;
    io.snap_every = g_snap_every; io.snap_u = g_snap_u; io.snap_flag = g_snap_flag;
    switch (p.M) {
      case 4: run_mode<4>(p, io, u, u1, dF, dG, order); break;
      case 5: run_mode<5>(p, io, u, u1, dF, dG, order); break;
      case 6: run_mode<6>(p, io, u, u1, dF, dG, order); break;
      case 7: run_mode<7>(p, io, u, u1, dF, dG, order); break;
      case 8: run_mode<8>(p, io, u, u1, dF, dG, order); break;
      case 9: run_mode<9>(p, io, u, u1, dF, dG, order); break;
      case 10: run_mode<10>(p, io, u, u1, dF, dG, order); break;
      default: return -3;
    }
  }
  return 0;
}

// ---- upwind / ssprk3 ---------------------------------------------------------------------------
template <int M, int XARR>
static void run_upwind_one(const Params& p, UpwindIO io, int order) {
  std::vector<double> buf(UW_SLOTS * p.T + (p.nxi > 0 ? p.nxi : 1), 0.0);
  Shared sh;
  sh.U0 = sh.U1 = sh.F = sh.G = sh.X = sh.Xx = sh.Xxx = nullptr;
  sh.xb = &buf[0]; sh.dW = &buf[UW_SLOTS * p.T];
  typedef UpwindThread<M, XARR, 0> Th;
  EmuExecT<Th> ex(p, order);
  int flag = 0;
  run_upwind_steps<M, XARR, 0, EmuExecT<Th>, Th>(ex, p, sh, io, &flag);
}

// thread-block-cluster emulation of the upwind kernel: exchange slots distributed over C CTAs (EmuSharedCluster decode)
template <int M, int XARR>
static int run_upwind_one_cluster(const Params& p, UpwindIO io, int order, int C) {
  const int Tc = p.T / C;
  std::vector<std::vector<double> > bufs(C, std::vector<double>(UW_SLOTS * Tc + (p.nxi > 0 ? p.nxi : 1), 0.0));
  typedef UpwindThread<M, XARR, 0> Th;
  EmuExecT<Th> ex(p, order, C);
  int bad = 0;
  EmuSharedCluster sh;
  double* z = (double*)0;
  sh.U0 = sh.U1 = sh.F = sh.G = sh.X = sh.Xx = sh.Xxx = z;
  sh.xb = z; sh.dW = z + UW_SLOTS * Tc;
  sh.bufs = &bufs; sh.Tc = Tc; sh.nxi_ = p.nxi; sh.cur = &ex.cur; sh.bad = &bad;
  sh.lgTc = 0; while ((1 << sh.lgTc) < Tc) ++sh.lgTc;
  sh.lgT = 0; while ((1 << sh.lgT) < p.T) ++sh.lgT;
  int flag = 0;
  run_upwind_steps<M, XARR, 0, EmuExecT<Th>, Th>(ex, p, sh, io, &flag);
  return bad;
}

extern "C" int pk_emu_step_upwind(int ne, double Ld, double dt, double alphasq, int xarr, int nxi, int nupd,
                                  double xi_const, int R, int nsteps, int step0, double* m, double* u,
                                  const double* xi, const double* dW_in, double* dW_out,
                                  unsigned long long seed, unsigned rid0, int* fail_step, int order) {
  Params p;
  memset(&p, 0, sizeof(p));
  p.ne = ne; p.h = Ld / ne; p.dt = dt; p.alphasq = alphasq; p.mode = xarr ? NOISE_X : NOISE_SCALAR; p.nxi = nxi; p.nupd = nupd;
  p.max_it = 50; p.rtol = 1e-8; p.atol = 1e-50; p.stol = 1e-8; p.divtol = 1e4; p.xi_const = xi_const;
  if (!choose_layout(ne, &p.T, &p.M, &p.rem)) return -1;
  fill_constants(p);
  std::vector<double> pt;
  if (xi && nxi > 0) { pt.resize((size_t)nxi * p.T * p.M); permute_table_host(p, xi, pt.data()); }
  for (int r = 0; r < R; ++r) {
    UpwindIO io;
    io.m = m; io.u = u; io.dW_in = dW_in; io.dW_out = dW_out; io.xi = pt.empty() ? nullptr : pt.data(); io.R = R; io.r = r; io.rid = rid0 + r;
    io.seed = seed; io.step0 = step0; io.nsteps = nsteps; io.fail_step = fail_step;
#define UW_CASE(MM) case MM: if (g_emuC > 1) { g_bad += xarr ? run_upwind_one_cluster<MM, 1>(p, io, order, g_emuC) : run_upwind_one_cluster<MM, 0>(p, io, order, g_emuC); } \
                             else if (xarr) run_upwind_one<MM, 1>(p, io, order); else run_upwind_one<MM, 0>(p, io, order); break;
    switch (p.M) { UW_CASE(5) UW_CASE(6) UW_CASE(7) UW_CASE(8) default: return -3; }
#undef UW_CASE
  }
  return 0;
}

// ---- mu diagnostic (pk_mu.h): the device routine on host arrays ------------------------------------
extern "C" int pk_emu_mu(const double* u, const double* du, const double* x, int n, double Ld, double* old4, double* alt4) {
  return mu_serial(u, du, x, n, Ld, old4, alt4);
}
extern "C" double pk_emu_pairwise_sum(const double* a, int n) { return np_pairwise_sum(a, n); }

// ---- closed-form element integrals of pk_core.h (CellC) for one cell, for the symbolic check (tests/test_symbolic.py) ----
// A7, B7 = NodeS (a, v, w, F, G, x, y) of the left / right node; out[18] = N1L N1R fL fR gL gR | dN1L0 dN1L1 dN1R0 dN1R1 |
// dfL0 dfL1 dfR0 dfR1 | dgL0 dgL1 dgR0 dgR1, as the kernel holds them (f, g: HALF the cell contributions; derivatives
// pre-scaled: N1'/2, f'/4, g'/2).  full != 0: compute_full (values from the Jacobian entries), else compute_res + compute_jac.
extern "C" void pk_emu_cell(double h, double dt, double alphasq, const double* A7, const double* B7, int full, double* out) {
  Params p;
  memset(&p, 0, sizeof(p));
  p.ne = 8; p.h = h; p.dt = dt; p.alphasq = alphasq; p.mode = NOISE_FULL; p.nupd = 1;
  p.rtol = 1e-8; p.atol = 1e-50; p.stol = 1e-8; p.divtol = 1e4;
  fill_constants(p);
  NodeS A, B;
  A.a = A7[0]; A.v = A7[1]; A.w = A7[2]; A.F = A7[3]; A.G = A7[4]; A.x = A7[5]; A.y = A7[6];
  B.a = B7[0]; B.v = B7[1]; B.w = B7[2]; B.F = B7[3]; B.G = B7[4]; B.x = B7[5]; B.y = B7[6];
  CellC<NOISE_FULL> c;
  if (full) c.compute_full(p, A, B);
  else { c.compute_jac(p, A, B); c.compute_res(p, A, B); }
  const double v[18] = {c.N1L, c.N1R, c.fL, c.fR, c.gL, c.gR, c.dN1L0, c.dN1L1, c.dN1R0, c.dN1R1,
                        c.dfL0, c.dfL1, c.dfR0, c.dfR1, c.dgL0, c.dgL1, c.dgR0, c.dgR1};
  for (int i = 0; i < 18; ++i) out[i] = v[i];
}
