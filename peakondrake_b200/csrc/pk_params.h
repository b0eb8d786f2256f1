// pk_params.h -- POD parameter block shared by the CUDA kernels, the C-ABI host code and the
// CPU emulation harness used by the not-gpu tests.  Plain C++, no CUDA types.
//
// Discretisation constants follow SURVEY.md App. A (CG1 on PeriodicIntervalMesh(ne, Ld),
// reference experiment.py:171,178,208):  M_e = h/6 [[2,1],[1,2]], K_e = 1/h [[1,-1],[-1,1]],
// D_e = 1/2 [[-1,-1],[1,1]],  H = M + alphasq K.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define PK_HD __host__ __device__ __forceinline__
#define PK_UNROLL _Pragma("unroll")
#else
#define PK_HD inline __attribute__((always_inline))
#define PK_UNROLL
#endif

namespace pk {

enum NoiseMode {
  NOISE_SCALAR = 0,   // dXi is one number per step (num_Xis==0, or Xi_family 'constant'); dXi_x=dXi_xx=0 => dG == 0
  NOISE_X = 1,        // dXi varies in space, dXi_x=dXi_xx=0 ('no_gradient', stochastic_functions.py:121,189,232) => dG == 0
  NOISE_FULL = 2,     // 'hydrodynamic' with tables xi, xi_x, xi_xx (stochastic_functions.py:128-154,189-201)
  // scheme 'test' (equations.py:132-145): one CG1 unknown u1, residual M(u1-u) - int psi uh (dv)_x, noise as in
  // NOISE_SCALAR / NOISE_X.  Encoded in the mode so that the same kernels, executors and dispatch serve it.
  TEST_SCALAR = 3,
  TEST_X = 4
};
#define PK_IS_TEST(MODE) ((MODE) >= 3)
#define PK_NOISE_OF(MODE) ((MODE) >= 3 ? (MODE) - 3 : (MODE))

struct Params {
  int ne;        // cells == CG1 dofs
  int T;         // threads (row chunks) per realisation, power of two, >= 2
  int M;         // max rows per chunk; chunks have M or M-1 rows
  int rem;       // chunks t < rem have M rows, the others M-1:  ne = T*(M-1) + rem
  int nxi;       // number of noise modes (num_Xis)
  int nupd;      // nXi_updates (stochastic_functions.py:213-217)
  int mode;      // NoiseMode
  int max_it;    // SNES max_it (50)
  double h, dt, alphasq;
  double mb, ma;          // M: off-diagonal h/6, diagonal 2h/3
  double hb, ha;          // H: off-diagonal h/6 - a2/h, diagonal 2h/3 + 2 a2/h
  double hm0, hm1, hm2;   // H*M circulant, offsets 0, +-1, +-2
  double c1, c2;          // Dt*h/12, alphasq*Dt/(4h)
  double h6, a26;         // h/6, alphasq/6
  double c1q, c1h, c2h, h12, a212, dt12;   // c1/4, c1/2, c2/2, h/12, alphasq/12, dt/12: the Jacobian enters the Newton
                                           // matrix as N1'/2, f'/4 (+-), g'/2, assembled pre-scaled (exact binary scalings)
  double rho, ikappa;     // H = kappa (1 - rho S)(1 - rho S^-1): |rho|<1
  double rpow[12];        // rho^0 .. rho^11 (chunk-local carry weights; a chunk has at most 10 rows)
  double irho;            // 1 / rho
  double inv1mrn;         // 1 / (1 - rho^ne): cyclic closure of the Helmholtz scans
  double cert_k, cert_floor2;   // convergence certificate of the Newton iteration (run_steps); cert_k = 0 disables it
  double ih2;             // 2 / h   (inverse DG1 element mass matrix M_e^-1 = (2/h) [[2,-1],[-1,2]])
  double ih;              // RN(1 / h): reciprocal for the exact quotient x / h of the fused diagnostics (div_h)
  double c1k, c2k, h6k, a26k;   // c1, c2, h6, a26 times ikappa (Helmholtz right-hand sides pre-divided by kappa)
  double rtol, atol, stol, divtol;  // PETSc SNES defaults 1e-8, 1e-50, 1e-8, 1e4
  double rtol2, atol2, stol2, divtol2;  // their squares (fill_constants): the tests compare squared norms
  double sdt;             // sqrt(dt / nupd)
  double xi_const;        // NOISE_SCALAR: the (spatially constant) value of xi_0, 0 if nxi==0
};

// reduction / scan scratch of the thread-block-cluster executors: two halves (ping-pong) of 3 values per warp, the CTA's
// aggregate (3) and the cluster-wide result broadcast inside the CTA (up to 3)
PK_HD int cluster_red_half(int Tc) { return 3 * (Tc >> 5) + 6; }
PK_HD int cluster_red_doubles(int Tc) { return 2 * cluster_red_half(Tc); }
// exchange area of the midpoint cluster kernels: two PCR row buffers (XB_SLOTS_PP = 30 slots per thread), or -- two CTAs per
// SM -- one row buffer (XB_SLOTS = 18 slots) + the scratch rows of the redundant separator solve (30 doubles per CTA of the
// cluster, at most 16 CTAs) + 16 doubles for the copy of the last thread's tips
PK_HD int cluster_xb_doubles(int Tc, bool compact) { return compact ? 18 * Tc + 30 * 16 + 16 : 30 * Tc; }
// communication area of the push exchanges (single-buffer kernels): all-gather inboxes + mbarriers (pk_kernels.cuh, pkx)
PK_HD int cluster_comm_doubles() { return 160; }

// chunk geometry
PK_HD int chunk_rows(const Params& p, int t) { return t < p.rem ? p.M : p.M - 1; }
PK_HD int chunk_start(const Params& p, int t) { return t * (p.M - 1) + (t < p.rem ? t : p.rem); }

// Choose (T, M, rem) for a given ne: T = smallest power of two with 8T >= ne, chunks of
// ceil/floor(ne/T) rows, which is then in [4, 8].  Returns false if ne < 8.
static inline bool choose_layout(int ne, int* T, int* M, int* rem) {
  if (ne < 8) return false;
  int t = 2;
  while (8 * t < ne) t *= 2;
  int lo = ne / t;             // floor
  int r = ne - lo * t;         // chunks with lo+1 rows
  if (r == 0) { *M = lo; *rem = t; if (lo < 5) { *M = lo + 1; *rem = 0; } }
  else { *M = lo + 1; *rem = r; }
  *T = t;
  return (*M >= 5 && *M <= 8);
}

// Wide chunks for the thread-block-cluster kernels: when the standard layout ends at 5 rows per thread (ne just above
// a power of two times 4: ne = 20 000 -> 4096 x 5), half the threads with 9 or 10 rows each do the same assembly work
// with half the separator rows, half the CTAs per cluster and one PCR level less.
static inline bool choose_layout_wide(int ne, int* T, int* M, int* rem) {
  int t, m, r;
  if (!choose_layout(ne, &t, &m, &r) || m != 5 || t < 4) return false;
  t /= 2;
  const int lo = ne / t;
  r = ne - lo * t;
  if (r == 0) { m = lo; r = t; } else m = lo + 1;
  if (m < 9 || m > 10) return false;
  *T = t; *M = m; *rem = r;
  return true;
}

static inline void fill_constants(Params& p) {
  const double h = p.h, a2 = p.alphasq;
  p.mb = h / 6.0; p.ma = 2.0 * h / 3.0;
  p.hb = h / 6.0 - a2 / h; p.ha = 2.0 * h / 3.0 + 2.0 * a2 / h;
  p.hm0 = 2.0 * p.hb * p.mb + p.ha * p.ma;
  p.hm1 = p.hb * p.ma + p.ha * p.mb;
  p.hm2 = p.hb * p.mb;
  if (PK_IS_TEST(p.mode)) { p.hm0 = p.ma; p.hm1 = p.mb; p.hm2 = 0.0; }   // 'test': the Newton matrix is M - T'/2 (tridiagonal)
  p.c1 = p.dt * h / 12.0; p.c2 = a2 * p.dt / (4.0 * h);
  p.h6 = h / 6.0; p.a26 = a2 / 6.0;
  p.c1q = 0.25 * p.c1; p.c1h = 0.5 * p.c1; p.c2h = 0.5 * p.c2; p.h12 = 0.5 * p.h6; p.a212 = 0.5 * p.a26; p.dt12 = p.dt * (1.0 / 12.0);
  const double disc = std::sqrt(p.ha * p.ha - 4.0 * p.hb * p.hb);
  const double kappa = 0.5 * (p.ha + disc);
  p.rho = -p.hb / kappa;
  p.ikappa = 1.0 / kappa;
  p.rpow[0] = 1.0;
  for (int i = 1; i < 12; ++i) p.rpow[i] = p.rpow[i - 1] * p.rho;
  p.irho = 1.0 / p.rho;
  p.inv1mrn = 1.0 / (1.0 - std::pow(p.rho, (double)p.ne));
  // |R(x + d)|_2 <= cert_k |d|_2^2 once R(x) + J d = 0 (see run_steps); evaluation noise of R: ~1.5e-14 h sqrt(ne + |w|^2)
  p.cert_k = p.dt * (2.4 + 4.0 * h / 3.0 + 2.0 * a2 / h) * 0.25;
  p.cert_floor2 = (1.5e-14 * h) * (1.5e-14 * h);
  p.ih2 = 2.0 / h; p.ih = 1.0 / h;
  p.c1k = p.c1 * p.ikappa; p.c2k = p.c2 * p.ikappa; p.h6k = p.h6 * p.ikappa; p.a26k = p.a26 * p.ikappa;
  p.sdt = std::sqrt(p.dt / (p.nupd > 0 ? p.nupd : 1));
  p.rtol2 = p.rtol * p.rtol; p.atol2 = p.atol * p.atol; p.stol2 = p.stol * p.stol; p.divtol2 = p.divtol * p.divtol;
}

}  // namespace pk
