// pk_aux_lib.cu -- the small per-realisation companions of the step kernels:
//   pk_peakon_sde_step   PeakonEquations.update (peakon_equations.py:34-65) for every realisation, K steps per launch
//   pk_ensemble_scalars  per-diagnostic ensemble sums / extrema / fail count of one dump
//   pk_du_smooth_loc     du_smooth (CG5 L2 projection of the DG0 slope, diagnostic_equations.py:115-119) and the
//                        locations of its extrema (outputting.py:135-138)
#include <cuda_runtime.h>
#include <cstdio>
#include <cmath>
#include "pk_core.h"
#include "pk_ctx.h"

using namespace pk;

// ------------------------------------------------------------------------------------------------
// Peakon SDE.  One thread per realisation; the arithmetic follows the reference statement by statement with
// explicitly rounded operations (no FMA contraction), so that (p, q) are the reference's own numbers when the
// increments are the reference's:
//   dq = (p/2)*(1+E)/(1-E) * dt                                              peakon_equations.py:37
//   ito_euler: dp += p/2*(xi_x*xi_x - xi*xi_xx)*dt - p*xi_x*dW ;  dq += 0.5*xi*xi_x*dt + xi*dW     :45-46
//   milstein : dt -> dW**2                                                                       :55-56
//   q wraps into [0, Ld]                                                                          :62-65
// xi, xi_x, xi_xx are the CG1 fields pure_xi* evaluated at q (`.at`): s = q/h, i = floor(s) mod n,
// th = s - floor(s), value = (1-th) f_i + th f_{i+1}.
// Increments: dW_in [nsteps, R, nxi], or the in-kernel Philox stream with the PDE kernels' counters
// (global realisation id, global step, mode, sub-update) -- the SAME numbers pk_step draws for that step.
// ------------------------------------------------------------------------------------------------
struct SdeArgs {
  double *p, *q;
  const double* dW_in;
  const double *xi, *xi_x, *xi_xx;      // [nxi, ne], natural node order
  double* traj;                         // [nsteps / stride, R, 2] or null: (p, q) after every `stride` steps
  int R, ne, nxi, nupd, nsteps, method, stride;
  long long step0;
  unsigned rid0;
  unsigned long long seed;
  double dt, Ld, h, E, sdt;
};

__device__ __forceinline__ double cg1_at(const double* __restrict__ f, int n, double h, double x) {
  const double s = __ddiv_rn(x, h);
  const double fl = floor(s);
  long long i = (long long)fl % n;
  if (i < 0) i += n;                            // Python's modulo
  const int i1 = (int)((i + 1) % n);
  const double th = __dsub_rn(s, fl);
  return __dadd_rn(__dmul_rn(__dsub_rn(1.0, th), f[i]), __dmul_rn(th, f[i1]));
}

__global__ void __launch_bounds__(128) pk_peakon_sde_kernel(const __grid_constant__ SdeArgs a) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.R) return;
  double p = a.p[r], q = a.q[r];
  const double c1 = __dadd_rn(1.0, a.E), c2 = __dsub_rn(1.0, a.E);
  for (int n = 0; n < a.nsteps; ++n) {
    double dp = 0.0;
    double dq = __dmul_rn(__ddiv_rn(__dmul_rn(__ddiv_rn(p, 2.0), c1), c2), a.dt);
    for (int k = 0; k < a.nxi; ++k) {
      double w;
      if (a.dW_in) w = a.dW_in[((long long)n * a.R + r) * a.nxi + k];
      else w = philox_increment(a.seed, a.rid0 + (unsigned)r, (unsigned)(a.step0 + n), (unsigned)k, a.nupd, a.sdt);
      const double xi = cg1_at(a.xi + (size_t)k * a.ne, a.ne, a.h, q);
      const double xi_x = cg1_at(a.xi_x + (size_t)k * a.ne, a.ne, a.h, q);
      const double xi_xx = cg1_at(a.xi_xx + (size_t)k * a.ne, a.ne, a.h, q);
      const double scale = (a.method == 0) ? a.dt : __dmul_rn(w, w);
      const double br = __dsub_rn(__dmul_rn(xi_x, xi_x), __dmul_rn(xi, xi_xx));
      const double t1 = __dmul_rn(__dmul_rn(__ddiv_rn(p, 2.0), br), scale);
      const double t2 = __dmul_rn(__dmul_rn(p, xi_x), w);
      dp = __dadd_rn(dp, __dsub_rn(t1, t2));
      const double t3 = __dmul_rn(__dmul_rn(__dmul_rn(0.5, xi), xi_x), scale);
      dq = __dadd_rn(dq, __dadd_rn(t3, __dmul_rn(xi, w)));
    }
    p = __dadd_rn(p, dp);
    q = __dadd_rn(q, dq);
    if (q < 0.0) q = __dadd_rn(q, a.Ld);
    if (q > a.Ld) q = __dsub_rn(q, a.Ld);
    if (a.traj && (n + 1) % a.stride == 0) {
      const long long o = ((long long)((n + 1) / a.stride - 1) * a.R + r) * 2;
      a.traj[o] = p; a.traj[o + 1] = q;
    }
  }
  a.p[r] = p; a.q[r] = q;
}

extern "C" int pk_peakon_sde_step(pk_ctx* c, int32_t nsteps, double* p, double* q, const double* dW_in, int64_t step0,
                                  const double* xi, const double* xi_x, const double* xi_xx, int32_t method,
                                  double E, double* traj, int32_t stride, void* stream) {
  if (!c) return -1;
  if (!p || !q) return fail(c, "pk_peakon_sde_step: null state pointer");
  if (nsteps < 0) return fail(c, "pk_peakon_sde_step: nsteps < 0");
  if (method != 0 && method != 1) return fail(c, "pk_peakon_sde_step: method must be 0 (ito_euler) or 1 (milstein)");
  if (c->p.nxi > 0 && (!xi || !xi_x || !xi_xx)) return fail(c, "pk_peakon_sde_step: the pure_xi tables are required");
  if (traj && stride < 1) return fail(c, "pk_peakon_sde_step: stride must be >= 1");
  if (nsteps == 0) return 0;
  CK(c, cudaSetDevice(c->cfg.device));
  SdeArgs a;
  a.p = p; a.q = q; a.dW_in = dW_in; a.xi = xi; a.xi_x = xi_x; a.xi_xx = xi_xx; a.traj = traj;
  a.R = c->cfg.R; a.ne = c->p.ne; a.nxi = c->p.nxi; a.nupd = c->p.nupd; a.nsteps = nsteps; a.method = method;
  a.stride = traj ? stride : 1;
  a.step0 = step0; a.rid0 = c->cfg.rid_offset; a.seed = c->cfg.seed;
  a.dt = c->p.dt; a.Ld = c->cfg.Ld; a.h = c->cfg.Ld / c->p.ne; a.E = E; a.sdt = c->p.sdt;
  pk_peakon_sde_kernel<<<(a.R + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
  ++c->nlaunch;
  CK(c, cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Ensemble scalars of one dump: over the healthy realisations, per diagnostic column of pk_diagnostics,
//   sums [2*PK_NDIAG + 2] = sum_r v | sum_r v^2 | number healthy | number failed       (all-reduce: SUM)
//   ext  [2*PK_NDIAG]     = min_r v | -max_r v                                          (all-reduce: MIN)
// One block, fixed summation order (deterministic); NaN entries (columns that were not requested) are skipped.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pk_ensemble_scalars_kernel(const double* __restrict__ diag, const int* __restrict__ fail_step,
                                                                  int R, double* __restrict__ sums, double* __restrict__ ext) {
  __shared__ double s1[256], s2[256], mn[256], mx[256];
  __shared__ int cnt[2][256];
  const int tid = threadIdx.x;
  int nh = 0, nf = 0;
  for (int r = tid; r < R; r += 256) { if (fail_step[r] >= 0) ++nf; else ++nh; }
  cnt[0][tid] = nh; cnt[1][tid] = nf;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { cnt[0][tid] += cnt[0][tid + o]; cnt[1][tid] += cnt[1][tid + o]; }
    __syncthreads();
  }
  if (tid == 0) { sums[2 * PK_NDIAG] = (double)cnt[0][0]; sums[2 * PK_NDIAG + 1] = (double)cnt[1][0]; }
  for (int k = 0; k < PK_NDIAG; ++k) {
    double a1 = 0.0, a2 = 0.0, lo = INFINITY, hi = -INFINITY;
    for (int r = tid; r < R; r += 256) {
      if (fail_step[r] >= 0) continue;
      const double v = diag[(size_t)r * PK_NDIAG + k];
      if (v != v) continue;
      a1 += v; a2 += v * v; lo = fmin(lo, v); hi = fmax(hi, v);
    }
    __syncthreads();
    s1[tid] = a1; s2[tid] = a2; mn[tid] = lo; mx[tid] = hi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (tid < o) {
        s1[tid] += s1[tid + o]; s2[tid] += s2[tid + o];
        mn[tid] = fmin(mn[tid], mn[tid + o]); mx[tid] = fmax(mx[tid], mx[tid + o]);
      }
      __syncthreads();
    }
    if (tid == 0) { sums[k] = s1[0]; sums[PK_NDIAG + k] = s2[0]; ext[k] = mn[0]; ext[PK_NDIAG + k] = -mx[0]; }
  }
}

extern "C" int pk_ensemble_scalars(pk_ctx* c, const double* diag, const int32_t* fail_step, double* sums, double* ext,
                                   void* stream) {
  if (!c) return -1;
  if (!diag || !fail_step || !sums || !ext) return fail(c, "pk_ensemble_scalars: null pointer");
  CK(c, cudaSetDevice(c->cfg.device));
  pk_ensemble_scalars_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(diag, fail_step, c->cfg.R, sums, ext);
  ++c->nlaunch;
  CK(c, cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// du_smooth: L2 projection of the DG0 slope du onto periodic CG5 (Projector(du, du_smooth),
// diagnostic_equations.py:115-119), and the locations of its maximum / minimum in the CG5 coordinate field
// (find_max / find_min on du_smooth with smooth_coords, outputting.py:135-138).
//
// On the uniform periodic mesh the CG5 mass matrix is, per cell, h * Mref (6x6, equispaced Lagrange nodes: the
// oracle's convention for Firedrake's CG5).  Static condensation of the 4 interior dofs of
// every cell leaves a constant-coefficient cyclic tridiagonal system for the vertex values,
//     sb (v_{i-1} + v_{i+1}) + sa v_i = cL du_{i-1} + cR du_i,
// solved with the explicit (exponentially decaying) circulant inverse like the m diagnostic; the interior dofs
// follow as  w_k = e_k du_i + l_k v_i + r_k v_{i+1}.  The host passes the nine condensed constants.
// dof order (the oracle's CG5 numbering): vertex i -> 5 i, interior k = 1..4 of cell i -> 5 i + k; coordinate of
// dof 5 i + k is (i + k/5) h.  Ties -> lowest dof index (NumPy argmax / argmin).
// ------------------------------------------------------------------------------------------------
struct SmoothArgs {
  const double* u; const int* fail_step;
  double* out;          // [R, 4]: max value, x of max, min value, x of min
  int* idx;             // [R, 2]: dof index of max, of min
  double* scratch;      // [R, 2, ne]: vertex rhs, vertex values
  double* field;        // [R, 5 ne] or null: du_smooth itself
  int ne, K;
  double h, rho, rn, gscale, cL, cR;
  double ek[4], lk[4], rk[4];
};

struct ArgS { double v; int i; };
__device__ __forceinline__ void smax(ArgS& a, double v, int i) { if (v > a.v || (v == a.v && i < a.i)) { a.v = v; a.i = i; } }
__device__ __forceinline__ void smin(ArgS& a, double v, int i) { if (v < a.v || (v == a.v && i < a.i)) { a.v = v; a.i = i; } }

__global__ void __launch_bounds__(256) pk_du_smooth_kernel(const __grid_constant__ SmoothArgs a) {
  __shared__ ArgS smx[8], smn[8];
  const int r = blockIdx.x, ne = a.ne;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  double* out = a.out + (size_t)r * 4;
  int* idx = a.idx + (size_t)r * 2;
  if (a.fail_step && a.fail_step[r] >= 0) {
    if (threadIdx.x < 4) out[threadIdx.x] = qnan;
    if (threadIdx.x < 2) idx[threadIdx.x] = -1;
    return;
  }
  const double* u = a.u + (size_t)r * ne;
  double* b = a.scratch + (size_t)r * 2 * ne;
  double* v = b + ne;
  auto du = [&](int i) { const int ip = (i + 1 == ne) ? 0 : i + 1; return (u[ip] - u[i]) / a.h; };
  for (int i = threadIdx.x; i < ne; i += blockDim.x) {
    const int im = (i == 0) ? ne - 1 : i - 1;
    b[i] = a.cL * du(im) + a.cR * du(i);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ne; i += blockDim.x) {
    double acc;
    if (a.K < 0) {                // small meshes: the full cyclic sum  g_k = (rho^k + rho^(n-k)) * gscale
      acc = 0.0;
      double pk_ = 1.0, pnk = a.rn;
      const double irho = 1.0 / a.rho;
      for (int k = 0; k < ne; ++k) {
        int j = i + k; if (j >= ne) j -= ne;
        acc += (pk_ + pnk) * b[j];
        pk_ *= a.rho; pnk *= irho;
      }
    } else {
      acc = b[i];
      double pw = 1.0;
      for (int k = 1; k <= a.K; ++k) {
        pw *= a.rho;
        int jp = i + k; if (jp >= ne) jp -= ne;
        int jm = i - k; if (jm < 0) jm += ne;
        acc += pw * (b[jp] + b[jm]);
      }
    }
    v[i] = acc * a.gscale;
  }
  __syncthreads();
  ArgS mx = {-INFINITY, 0x7fffffff}, mn = {INFINITY, 0x7fffffff};
  for (int i = threadIdx.x; i < ne; i += blockDim.x) {
    const int ip = (i + 1 == ne) ? 0 : i + 1;
    const double d = du(i), v0 = v[i], v1 = v[ip];
    smax(mx, v0, 5 * i); smin(mn, v0, 5 * i);
    if (a.field) a.field[(size_t)r * 5 * ne + 5 * i] = v0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double w = a.ek[k] * d + a.lk[k] * v0 + a.rk[k] * v1;
      smax(mx, w, 5 * i + k + 1); smin(mn, w, 5 * i + k + 1);
      if (a.field) a.field[(size_t)r * 5 * ne + 5 * i + k + 1] = w;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgS t;
    t.v = __shfl_xor_sync(0xffffffffu, mx.v, o); t.i = __shfl_xor_sync(0xffffffffu, mx.i, o); smax(mx, t.v, t.i);
    t.v = __shfl_xor_sync(0xffffffffu, mn.v, o); t.i = __shfl_xor_sync(0xffffffffu, mn.i, o); smin(mn, t.v, t.i);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { smx[w] = mx; smn[w] = mn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) { smax(mx, smx[k].v, smx[k].i); smin(mn, smn[k].v, smn[k].i); }
    out[0] = mx.v; out[1] = ((mx.i / 5) + (mx.i % 5) * 0.2) * a.h;
    out[2] = mn.v; out[3] = ((mn.i / 5) + (mn.i % 5) * 0.2) * a.h;
    idx[0] = mx.i; idx[1] = mn.i;
  }
}

extern "C" int pk_du_smooth_loc(pk_ctx* c, const double* u, const int32_t* fail_step, const double* consts,
                                double* out, int32_t* idx, double* field, void* stream) {
  if (!c) return -1;
  if (!u || !consts || !out || !idx) return fail(c, "pk_du_smooth_loc: null pointer");
  CK(c, cudaSetDevice(c->cfg.device));
  const int ne = c->p.ne, R = c->cfg.R;
  if (pk_ensure_scratch(c, 2 * (size_t)R * ne)) return -1;
  SmoothArgs a;
  a.u = u; a.fail_step = fail_step; a.out = out; a.idx = idx; a.scratch = c->d_scratch; a.field = field;
  a.ne = ne; a.h = c->p.h;
  // consts (host, 16 doubles): sa, sb (Schur tridiagonal, per unit h), cL, cR, e[4], l[4], r[4]
  const double sa = consts[0] * a.h, sb = consts[1] * a.h;
  a.cL = consts[2] * a.h; a.cR = consts[3] * a.h;
  for (int k = 0; k < 4; ++k) { a.ek[k] = consts[4 + k]; a.lk[k] = consts[8 + k]; a.rk[k] = consts[12 + k]; }
  // sa v_i + sb (v_{i-1} + v_{i+1}) = kappa (1 - rho S)(1 - rho S^-1) v
  const double kappa = 0.5 * (sa + sqrt(sa * sa - 4.0 * sb * sb));
  a.rho = -sb / kappa;
  int K = 1;
  { double pw = fabs(a.rho); while (pw > 1e-40 && K < ne) { pw *= fabs(a.rho); ++K; } }
  a.rn = pow(a.rho, (double)ne);
  if (2 * K + 1 <= ne) { a.K = K; a.gscale = 1.0 / (kappa * (1.0 - a.rho * a.rho)); }      // |rho|^K < 1e-40
  else { a.K = -1; a.gscale = 1.0 / (kappa * (1.0 - a.rho * a.rho) * (1.0 - a.rn)); }
  pk_du_smooth_kernel<<<R, 256, 0, (cudaStream_t)stream>>>(a);
  ++c->nlaunch;
  CK(c, cudaGetLastError());
  return 0;
}
