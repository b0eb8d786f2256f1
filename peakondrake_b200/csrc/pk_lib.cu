// pk_lib.cu -- CUDA kernels (sm_100a) and the C ABI declared in include/peakon_b200.h.
//
// Build (see __graft_entry__.build):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared
//        -o peakondrake_b200/libpeakon_b200.so peakondrake_b200/csrc/pk_lib.cu
//
// One CTA integrates one realisation for `nsteps` timesteps with its whole state resident in
// shared memory (pk_core.h holds the thread-level phases, shared with the CPU emulation harness).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <string>
#include <cmath>
#include "pk_core.h"
#include "../../include/peakon_b200.h"

using namespace pk;

#include "pk_args.h"
#include "pk_dispatch.h"
#include "pk_upwind.h"
const void* pk_upwind_cluster_kernel_ptr(const pk::Params& p, int Tc);     // pk_upwind_lib.cu

// ------------------------------------------------------------------------------------------------
// diagnostics kernel: one CTA per realisation, natural order, coalesced
// ------------------------------------------------------------------------------------------------
struct DiagArgs {
  const double* u;
  const double* pq;        // [R,2] (p_true, q) of the peakon SDE or null
  double* out;             // [R, PK_NDIAG]
  int* idx;                // [R, PK_NIDX]
  double* du_out;          // [R, ne] or null
  double* scratch;         // [R, 2, ne] (b = H u, m) or null -> skip the m diagnostics
  const int* fail_step;
  int ne, periodic;
  double h, alphasq, Ld;
  double rhoM, gscale;     // mass-matrix inverse: m_i = gscale * sum_k g_k b_{i+k}
  int K;                   // truncation radius of the (exponentially decaying) inverse stencil, or -1: full cyclic sum
};

struct ArgV { double v; int i; };
__device__ __forceinline__ void amax(ArgV& a, double v, int i) { if (v > a.v || (v == a.v && i < a.i)) { a.v = v; a.i = i; } }
__device__ __forceinline__ void amin(ArgV& a, double v, int i) { if (v < a.v || (v == a.v && i < a.i)) { a.v = v; a.i = i; } }

template <bool IS_MAX>
__device__ __forceinline__ ArgV block_arg(ArgV a, ArgV* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgV b; b.v = __shfl_xor_sync(0xffffffffu, a.v, o); b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
    if (IS_MAX) amax(a, b.v, b.i); else amin(a, b.v, b.i);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) sm[w] = a;
  __syncthreads();
  ArgV r = sm[0];
  for (int k = 1; k < nw; ++k) { if (IS_MAX) amax(r, sm[k].v, sm[k].i); else amin(r, sm[k].v, sm[k].i); }
  __syncthreads();
  return r;
}

__device__ __forceinline__ double block_sum(double v, double* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) sm[w] = v;
  __syncthreads();
  double s = 0.0;
  for (int k = 0; k < nw; ++k) s += sm[k];
  __syncthreads();
  return s;
}

// u_sde nodal value, diagnostic_equations.py:254-279 with the amplitude of :377 (third periodic
// branch reproduces the reference's parenthesis at :270).
__device__ __forceinline__ double usde_at(double x, double p_true, double q, double Ld, double alphasq, int periodic) {
  const double al = sqrt(alphasq);
  const double E = exp(-Ld / al);
  const double p = p_true * 0.5 * (1.0 + E) / (1.0 - E);
  if (periodic) {
    if (x < q - 0.5 * Ld) return p * ((exp(-(x - q + Ld) / al) + E * exp((x - q + Ld) / al)) / (1.0 - E));
    if (x < q + 0.5 * Ld) { const double d = sqrt((q - x) * (q - x) / alphasq); return p * ((exp(-d) + E * exp(d)) / (1.0 - E)); }
    return p * ((exp(-(q + Ld - x) / al) + exp(-Ld / al * exp((q + Ld - x) / al))) / (1.0 - E));
  }
  if (x < q - 0.5 * Ld) return p * exp(-(x - q + Ld) / al);
  if (x < q + 0.5 * Ld) return p * exp(-sqrt((q - x) * (q - x) / alphasq));
  return p * exp(-(q + Ld - x) / al);
}

// x / h, correctly rounded, from ih = RN(1/h): q0 = RN(x ih) is within ~1.5 ulp, one FMA correction with the exact
// remainder makes it faithful, a second one (Markstein's theorem: correctly rounded reciprocal + faithful quotient + exact
// remainder) makes it RN(x / h).  Five instructions instead of the ~40 of a division.
static __device__ __forceinline__ double div_h(double x, double h, double ih) {
  const double q0 = x * ih;
  const double q = fma(fma(-q0, h, x), ih, q0);
  const double res = fma(fma(-q, h, x), ih, q);
  return res == 0.0 ? q0 : res;            // (-0) / h = -0: the corrections lose the sign of a zero
}

// EXTRA = false: the plain dump (no du field, no m columns, no true-peakon error) -- what the per-step dumps of BASELINE
// config 4 ask for: a light loop body (4 CTAs per SM instead of 2 at 95 registers) unrolled four times so that twelve
// loads per thread are in flight (the rolled loop sat at 13 % of the HBM bandwidth with `long_scoreboard` 7.7 per issue,
// profiles/r2zp).  The per-thread accumulation order is unchanged: same bits.
template <bool EXTRA>
__global__ void __launch_bounds__(256, EXTRA ? 2 : 4) pk_diag_kernel(const __grid_constant__ DiagArgs a) {
  __shared__ double ssum[32];
  __shared__ ArgV sarg[32];
  const int r = blockIdx.x, ne = a.ne;
  const double* u = a.u + (long long)r * ne;
  double* out = a.out + (long long)r * PK_NDIAG;
  int* idx = a.idx + (long long)r * PK_NIDX;
  if (a.fail_step && a.fail_step[r] >= 0) {   // outputting.py:107-108
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int k = threadIdx.x; k < PK_NDIAG; k += blockDim.x) out[k] = qnan;
    for (int k = threadIdx.x; k < PK_NIDX; k += blockDim.x) idx[k] = -1;
    return;
  }
  const double h = a.h, ih = 1.0 / a.h, a2 = a.alphasq;
  double sM = 0.0, sK = 0.0, sU = 0.0, sE = 0.0;
  ArgV mxdu = {-INFINITY, 0x7fffffff}, mndu = {INFINITY, 0x7fffffff}, mxu = {-INFINITY, 0x7fffffff}, mxj = {-INFINITY, 0x7fffffff};
  double* b = (EXTRA && a.scratch) ? a.scratch + (long long)r * 2 * ne : nullptr;
  double* mm = b ? b + ne : nullptr;
  const double pt = (EXTRA && a.pq) ? a.pq[2 * r] : 0.0, qq = (EXTRA && a.pq) ? a.pq[2 * r + 1] : 0.0;
#pragma unroll 4
  for (int i = threadIdx.x; i < ne; i += 256) {
    const int im = (i == 0) ? ne - 1 : i - 1, ip = (i == ne - 1) ? 0 : i + 1;
    const double um = u[im], u0 = u[i], up = u[ip];
    // x / h, correctly rounded, without a division (Markstein: q = x RN(1/h), the exact remainder by FMA, one fused
    // correction -- 3 instructions instead of ~40; the two divisions were most of this kernel's instructions and held it
    // at 13 % of the HBM bandwidth).  Bit-identical to `/ h`: tests/test_gpu_parity.py compares du with NumPy's division.
    const double du = div_h((up - u0), h, ih);           // DG0 cell i          (diagnostic_equations.py:94-101)
    const double dum = div_h((u0 - um), h, ih);          // DG0 cell i-1
    const double jmp = fabs(du - dum);         // |jump(du)| at vertex i (:103-113)
    sM += (h * (1.0 / 3.0)) * (u0 * u0 + u0 * up + up * up);   // cell integral of u^2
    sK += (up - u0) * (up - u0) * ih;                          // cell integral of u_x^2
    sU += h * u0;
    amax(mxdu, du, i); amin(mndu, du, i); amax(mxu, u0, i); amax(mxj, jmp, i);
    if (EXTRA && a.du_out) a.du_out[(long long)r * ne + i] = du;
    if (EXTRA && b) b[i] = a.h / 6.0 * (um + 4.0 * u0 + up) + a2 * ih * (2.0 * u0 - um - up);
    if (EXTRA && a.pq) {
      const double e0 = u0 - usde_at(i * h, pt, qq, a.Ld, a2, a.periodic);
      const double e1 = up - usde_at(ip * h, pt, qq, a.Ld, a2, a.periodic);
      sE += (h * (1.0 / 3.0)) * (e0 * e0 + e0 * e1 + e1 * e1);
    }
  }
  sM = block_sum(sM, ssum); sK = block_sum(sK, ssum); sU = block_sum(sU, ssum);
  if (EXTRA && a.pq) sE = block_sum(sE, ssum);
  mxdu = block_arg<true>(mxdu, sarg); mndu = block_arg<false>(mndu, sarg);
  mxu = block_arg<true>(mxu, sarg); mxj = block_arg<true>(mxj, sarg);
  double sMm = 0.0, sKm = 0.0, sm1 = 0.0, sUm = 0.0, mmax = -INFINITY;
  if (b) {
    __syncthreads();
    // m = M^-1 b by the explicit (exponentially decaying) circulant inverse
    for (int i = threadIdx.x; i < ne; i += blockDim.x) {
      double acc = 0.0;
      if (a.K < 0) {
        const double rn = pow(a.rhoM, (double)ne);
        double pk_ = 1.0, pnk = rn;     // rho^k, rho^(n-k)
        const double irho = 1.0 / a.rhoM;
        for (int k = 0; k < ne; ++k) {
          int j = i + k; if (j >= ne) j -= ne;
          acc += (pk_ + pnk) * b[j];
          pk_ *= a.rhoM; pnk *= irho;
        }
      } else {
        double pk_ = 1.0;
        acc = b[i];
        for (int k = 1; k <= a.K; ++k) {
          pk_ *= a.rhoM;
          int jp = i + k; if (jp >= ne) jp -= ne;
          int jm = i - k; if (jm < 0) jm += ne;
          acc += pk_ * (b[jp] + b[jm]);
        }
      }
      mm[i] = acc * a.gscale;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ne; i += blockDim.x) {
      const int ip = (i == ne - 1) ? 0 : i + 1;
      const double m0 = mm[i], mp = mm[ip], u0 = u[i], up = u[ip];
      sMm += (h * (1.0 / 3.0)) * (m0 * m0 + m0 * mp + mp * mp);
      sKm += (mp - m0) * (mp - m0) * ih;
      sm1 += h * m0;
      sUm += (h * (1.0 / 6.0)) * (2.0 * u0 * m0 + u0 * mp + up * m0 + 2.0 * up * mp);
      mmax = fmax(mmax, m0);
    }
    sMm = block_sum(sMm, ssum); sKm = block_sum(sKm, ssum); sm1 = block_sum(sm1, ssum); sUm = block_sum(sUm, ssum);
    ArgV t = {mmax, 0}; t = block_arg<true>(t, sarg); mmax = t.v;
  }
  if (threadIdx.x == 0) {
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    out[PK_D_ENERGY] = sM + a2 * sK;
    out[PK_D_L2U2] = sM; out[PK_D_H1SEMI2] = sK; out[PK_D_MASS_U] = sU;
    out[PK_D_MAX_DU] = mxdu.v; out[PK_D_MIN_DU] = mndu.v; out[PK_D_MAX_JUMP_LOCAL] = mxj.v; out[PK_D_MAX_U] = mxu.v;
    out[PK_D_L2M2] = b ? sMm : qnan; out[PK_D_H1SEMI_M2] = b ? sKm : qnan; out[PK_D_MASS_M] = b ? sm1 : qnan;
    out[PK_D_M_MAX] = b ? mmax : qnan; out[PK_D_UMU] = b ? sUm : qnan;
    out[PK_D_ERR_SDE2] = (EXTRA && a.pq) ? sE : qnan;
    idx[PK_I_ARGMAX_DU] = mxdu.i; idx[PK_I_ARGMIN_DU] = mndu.i; idx[PK_I_ARGMAX_U] = mxu.i; idx[PK_I_ARGMAX_JUMP] = mxj.i;
  }
}



// ------------------------------------------------------------------------------------------------
// peakon suite (outputting.py:373-450): the two sequential scans around the highest peak as
// block-parallel "first index satisfying" searches. One CTA per realisation.
//   right scan: first k in [0,S): D=(pc+k) mod n with du[D] > du[D-1]  -> min_du = du[D-1], index D-1
//   left  scan: first k in [0,S): D=(pc-1-k) mod n with du[D] < du[D+1] -> max_du = du[D+1], index D-1
// (the "D-1" of the left scan is the reference's own indexing, :430, reproduced)
// outputs per realisation: vals[2] = (min_du, max_du) (NaN if not found), ids[3] = (argmax u,
// min_du_idx, max_du_idx) as the NON-NEGATIVE array positions Python's negative indexing resolves to
// (-1 if not found).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pk_peakon_suite_kernel(const double* __restrict__ uall, const int* __restrict__ didx,
                                                              const int* __restrict__ fail_step, int ne, double h,
                                                              double* __restrict__ vals, int* __restrict__ ids) {
  __shared__ int found[2];
  const int r = blockIdx.x;
  const double* u = uall + (long long)r * ne;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  if (fail_step && fail_step[r] >= 0) {
    if (threadIdx.x == 0) { vals[2 * r] = qnan; vals[2 * r + 1] = qnan; ids[3 * r] = -1; ids[3 * r + 1] = -1; ids[3 * r + 2] = -1; }
    return;
  }
  const int pc = didx[r * PK_NIDX + PK_I_ARGMAX_U];
  const int S = ne / 2;
  auto du = [&](int i) { i %= ne; if (i < 0) i += ne; const int ip = (i + 1 == ne) ? 0 : i + 1; return (u[ip] - u[i]) / h; };
  if (threadIdx.x == 0) { found[0] = 0x7fffffff; found[1] = 0x7fffffff; }
  __syncthreads();
  for (int base = 0; base < S; base += blockDim.x) {
    const int k = base + threadIdx.x;
    if (k < S) {
      if (found[0] == 0x7fffffff) { const int D = pc + k; if (du(D) > du(D - 1)) atomicMin(&found[0], k); }
      if (found[1] == 0x7fffffff) { const int D = pc - 1 - k; if (du(D) < du(D + 1)) atomicMin(&found[1], k); }
    }
    __syncthreads();
    if (found[0] != 0x7fffffff && found[1] != 0x7fffffff) break;
  }
  if (threadIdx.x == 0) {
    ids[3 * r] = pc;
    if (found[0] != 0x7fffffff) {
      int D = (pc + found[0]) % ne; int j = D - 1; if (j < 0) j += ne;
      vals[2 * r] = du(j); ids[3 * r + 1] = j;
    } else { vals[2 * r] = qnan; ids[3 * r + 1] = -1; }
    if (found[1] != 0x7fffffff) {
      int D = (pc - 1 - found[1]) % ne; if (D < 0) D += ne;
      int j = D - 1; if (j < 0) j += ne;
      vals[2 * r + 1] = du(D + 1); ids[3 * r + 2] = j;
    } else { vals[2 * r + 1] = qnan; ids[3 * r + 2] = -1; }
  }
}

// ------------------------------------------------------------------------------------------------
// ensemble statistics: per-x moments of u over the healthy realisations + histograms of two scalar
// diagnostics, packed into ONE fp64 buffer so that a multi-GPU run needs a single all-reduce
// (SURVEY.md 8e). Deterministic: fixed-order partial sums, integer histogram counts.
//   packed = [ sum_r u (ne) | sum_r u^2 (ne) | n_healthy (1) | hist q_pde (nbins) | hist min_du (nbins) ]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pk_moments_partial(const double* __restrict__ u, const int* __restrict__ fail_step,
                                                          int R, int ne, int rows_per_tile, double* __restrict__ scratch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int tile = blockIdx.y;
  const int r0 = tile * rows_per_tile, r1 = min(R, r0 + rows_per_tile);
  double s1 = 0.0, s2 = 0.0;
  if (i < ne) {
    for (int r = r0; r < r1; ++r) {
      if (fail_step[r] >= 0) continue;
      const double v = u[(long long)r * ne + i];
      s1 += v; s2 += v * v;
    }
    scratch[(long long)tile * 2 * ne + i] = s1;
    scratch[(long long)tile * 2 * ne + ne + i] = s2;
  }
}

__global__ void __launch_bounds__(256) pk_moments_final(const double* __restrict__ scratch, int ntile, int ne,
                                                        const int* __restrict__ fail_step, const double* __restrict__ diag,
                                                        const int* __restrict__ idx, int R, int nbins, double h,
                                                        double q_lo, double q_hi, double d_lo, double d_hi,
                                                        double* __restrict__ packed) {
  extern __shared__ int hist[];   // [2*nbins + 1]
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = tid; i < 2 * ne; i += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int t = 0; t < ntile; ++t) s += scratch[(long long)t * 2 * ne + i];
    packed[i] = s;
  }
  if (blockIdx.x == 0) {
    for (int k = threadIdx.x; k < 2 * nbins + 1; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
      if (fail_step[r] >= 0) continue;
      atomicAdd(&hist[2 * nbins], 1);
      const double q = (idx[r * PK_NIDX + PK_I_ARGMAX_U] + 0.5) * h;
      int b = (int)floor((q - q_lo) / (q_hi - q_lo) * nbins);
      b = b < 0 ? 0 : (b >= nbins ? nbins - 1 : b);
      atomicAdd(&hist[b], 1);
      const double d = diag[r * PK_NDIAG + PK_D_MIN_DU];
      int c = (int)floor((d - d_lo) / (d_hi - d_lo) * nbins);
      c = c < 0 ? 0 : (c >= nbins ? nbins - 1 : c);
      atomicAdd(&hist[nbins + c], 1);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 2 * nbins + 1; k += blockDim.x) {
      const int dst = (k == 2 * nbins) ? 2 * ne : 2 * ne + 1 + k;
      packed[dst] = (double)hist[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
#include "pk_ctx.h"
#include "pk_mu.h"

std::string g_create_err;
int pk_fail(pk_ctx* c, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  if (c) c->err = buf; else g_create_err = buf;
  return -1;
}

int pk_ensure_scratch(pk_ctx* c, size_t doubles) {
  if (c->scratch_cap >= doubles) return 0;
  if (c->d_scratch) { cudaStreamSynchronize(0); cudaDeviceSynchronize(); cudaFree(c->d_scratch); c->d_scratch = nullptr; c->scratch_cap = 0; }
  if (cudaMalloc(&c->d_scratch, sizeof(double) * doubles) != cudaSuccess) return fail(c, "out of device memory for the diagnostics scratch (%zu doubles)", doubles);
  c->scratch_cap = doubles;
  return 0;
}

static kern_t pick_kernel(const Params& p, int block, int fs, int feat = 1) {
  // fs = 1: FS variant (factors in shared memory, rolled row loops, 2 CTAs/SM); 2: rolled assembly with register
  // factors; 0: fully unrolled register variant
  const int v = fs == 1 ? 2 : (fs == 2 ? 4 : (fs == 4 ? 5 : (fs == 6 ? 6 : (fs == 7 ? 7 : 3))));
  switch (p.M) {
    case 5: return pk_pick_kernel_m5(p.mode, block, p.T, v, feat);
    case 6: return pk_pick_kernel_m6(p.mode, block, p.T, v, feat);
    case 7: return pk_pick_kernel_m7(p.mode, block, p.T, v, feat);
    case 8: return pk_pick_kernel_m8(p.mode, block, p.T, v, feat);
  }
  return nullptr;
}
// Split K steps of R realisations into work items for S resident CTAs: nb blocks of nblk steps per
// realisation.  A round (all slots advance one item) costs nblk steps plus the state reload / noise
// window redraw (~0.15 step, measured); blocks of one realisation are sequential, so rounds >= nb.
static void choose_blocks(int R, int K, int S, int* nb_out, int* nblk_out) {
  double best = 1e300; int bnb = 1, bblk = K;
  for (int nblk = 1; nblk <= K; ++nblk) {
    const int nb = (K + nblk - 1) / nblk;
    long long rounds = ((long long)R * nb + S - 1) / S;
    if (rounds < nb) rounds = nb;
    const double cost = (double)rounds * (nblk + 0.15);
    if (cost < best - 1e-9) { best = cost; bnb = nb; bblk = nblk; }
  }
  *nb_out = bnb; *nblk_out = bblk;
}
static ckern_t pick_cluster_kernel(const Params& p, int Tc, int minb) {
  switch (p.M) {
    case 5: return pk_pick_cluster_kernel_m5(p.mode, Tc, minb);
    case 6: return pk_pick_cluster_kernel_m6(p.mode, Tc, minb);
    case 7: return pk_pick_cluster_kernel_m7(p.mode, Tc, minb);
    case 8: return pk_pick_cluster_kernel_m8(p.mode, Tc, minb);
    case 9: return pk_pick_cluster_kernel_m9(p.mode, Tc, minb);
    case 10: return pk_pick_cluster_kernel_m10(p.mode, Tc, minb);
  }
  return nullptr;
}

extern "C" const char* pk_last_error(const pk_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

extern "C" int pk_create(const pk_config* cfg, pk_ctx** out) {
  if (!cfg || !out) return fail(nullptr, "pk_create: null argument");
  *out = nullptr;
  if (!(cfg->alphasq > 0.0)) return fail(nullptr, "Schemes for alphasq <= 0 are not implemented (reference equations.py:166-167)");
  if (cfg->ne < 8) return fail(nullptr, "pk_create: resolution ne=%d too small (need >= 8)", cfg->ne);
  if (cfg->R < 1) return fail(nullptr, "pk_create: R must be >= 1");
  if (cfg->noise_mode < 0 || cfg->noise_mode > 2) return fail(nullptr, "pk_create: bad noise_mode %d", cfg->noise_mode);
  if (cfg->num_xis < 0) return fail(nullptr, "pk_create: num_xis < 0");
  if (cfg->scheme < 0 || cfg->scheme > 2) return fail(nullptr, "pk_create: bad scheme %d", cfg->scheme);
  if (cfg->scheme == 2 && cfg->noise_mode == PK_NOISE_FULL)
    return fail(nullptr, "pk_create: the 'test' scheme takes dXi only (noise_mode SCALAR or X)");
  if (cfg->scheme == 1 && cfg->noise_mode == PK_NOISE_FULL)
    return fail(nullptr, "pk_create: the upwind scheme takes dXi only (noise_mode SCALAR or X)");
  if (cfg->noise_mode == PK_NOISE_SCALAR && cfg->num_xis > 1)
    return fail(nullptr, "pk_create: scalar noise supports one Xi only (reference stochastic_functions.py:111-112)");
  pk_ctx* c = new pk_ctx();
  c->cfg = *cfg;
  c->slots = 0; c->d_queue = nullptr; c->last_nb = c->last_nblk = c->last_grid = 0;
  c->sms = 148;
  cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, cfg->device);
  Params& p = c->p;
  memset(&p, 0, sizeof(p));
  p.ne = cfg->ne; p.h = cfg->Ld / cfg->ne; p.dt = cfg->dt; p.alphasq = cfg->alphasq;
  p.mode = cfg->noise_mode + (cfg->scheme == 2 ? 3 : 0); p.nxi = cfg->num_xis; p.nupd = cfg->nxi_updates > 0 ? cfg->nxi_updates : 1;
  p.max_it = cfg->max_it > 0 ? cfg->max_it : 50;
  p.rtol = cfg->rtol > 0 ? cfg->rtol : 1e-8; p.atol = cfg->atol > 0 ? cfg->atol : 1e-50;
  p.stol = cfg->stol > 0 ? cfg->stol : 1e-8; p.divtol = cfg->divtol > 0 ? cfg->divtol : 1e4;
  p.xi_const = cfg->xi_const;
  if (!choose_layout(p.ne, &p.T, &p.M, &p.rem)) { delete c; return fail(nullptr, "pk_create: no thread layout for ne=%d", cfg->ne); }
  // cluster meshes that would end at 5 rows per thread: 9 or 10 rows on half the threads (choose_layout_wide)
  if (cfg->scheme == 0 && p.T > 512) {
    const char* e = getenv("PK_NO_WIDE_CHUNKS");
    if (!(e && atoi(e) == 1)) choose_layout_wide(p.ne, &p.T, &p.M, &p.rem);
  }
  { // experiments: PK_FORCE_TM="T,M" overrides the layout (chunks of M or M-1 rows, M in 5..8)
    const char* e = getenv("PK_FORCE_TM"); int ft = 0, fm = 0;
    if (e && sscanf(e, "%d,%d", &ft, &fm) == 2) {
      const int rem = p.ne - ft * (fm - 1);
      if (rem < 0 || rem > ft || fm < 5 || fm > 8) { delete c; return fail(nullptr, "PK_FORCE_TM=%s does not fit ne=%d", e, cfg->ne); }
      p.T = ft; p.M = fm; p.rem = rem;
    }
  }
  fill_constants(p);
  { const char* e = getenv("PK_NO_CERTIFICATE"); if (e && atoi(e) == 1) p.cert_k = 0.0; }     // always evaluate the last residual
  const int narr = p.mode == NOISE_FULL ? 7 : (PK_NOISE_OF(p.mode) == NOISE_X ? 4 : 3);
  c->block = p.T < 32 ? 32 : p.T;
  c->cluster = 1; c->lgTc = 0;
  c->smem = sizeof(double) * ((size_t)narr * p.T * p.M + (size_t)XB_SLOTS * p.T + 6 * (c->block / 32) + (p.nxi > 0 ? p.nxi : 1));
  c->fs = 0; c->cminb = 1;
  { const char* e = getenv("PK_FS"); const int want = e ? atoi(e) : 0;
    if (want && cfg->scheme != 1 && p.T == 128) {      // FS variant: + level-1 factors in shared memory
      c->fs = (want == 1 || want == 2 || want == 4 || want == 6 || want == 7) ? want : 1;     // needs a -DPK_EXPERIMENTAL_VARIANTS build
      if (c->fs == 1) c->smem += sizeof(double) * ((size_t)6 * (p.M - 2) * p.T + 1);
    } }
  if (cfg->scheme == 1) {
    // upwind/ssprk3: state in registers, only exchange slots in shared memory (pk_upwind_lib.cu)
    c->smem = sizeof(double) * ((size_t)UW_SLOTS * p.T + 6 * (c->block / 32) + (p.nxi > 0 ? p.nxi : 1) + 2);
  }
  int force = 0;
  { const char* e = getenv("PK_FORCE_CLUSTER"); force = e ? atoi(e) : 0; }   // testing: split small meshes too
  if (force > 1 || c->block > 512 || c->smem > 227 * 1024) {
    // one realisation per thread-block cluster: C CTAs of Tc threads
    int Tc = 256, C = p.T / Tc;
    if (force > 1) { C = force; Tc = p.T / C; }
    while (C > 16 && Tc < 512) { Tc *= 2; C /= 2; }
    int lg = 0; while ((1 << lg) < Tc) ++lg;
    const size_t sm = cfg->scheme == 1
        ? sizeof(double) * ((size_t)UW_SLOTS * Tc + cluster_red_doubles(Tc) + (p.nxi > 0 ? p.nxi : 1) + 2)
        : sizeof(double) * ((size_t)narr * Tc * p.M + (size_t)cluster_xb_doubles(Tc, false) + cluster_red_doubles(Tc) + (p.nxi > 0 ? p.nxi : 1));
    // two CTAs per SM: one PCR row buffer + separator scratch instead of two row buffers
    const size_t sm2 = sizeof(double) * ((size_t)narr * Tc * p.M + (size_t)cluster_xb_doubles(Tc, true) + cluster_red_doubles(Tc) + (p.nxi > 0 ? p.nxi : 1) + 1 + cluster_comm_doubles());
    if (C < 2 || C > 16 || (C & (C - 1)) || Tc < 32 || C * Tc != p.T || sm > 227 * 1024) {
      const int ne = p.ne; delete c;
      return fail(nullptr, "pk_create: no cluster layout for ne=%d (threads %d)", ne, p.T);
    }
    c->cluster = C; c->block = Tc; c->lgTc = lg; c->smem = sm;
    // two CTAs per SM (two clusters interleaved on the same SMs: one computes while the other waits at a cluster
    // barrier) when the shared memory allows it -- scalar noise.  These instantiations keep ONE PCR row buffer (18 exchange
    // slots + separator scratch instead of 30 slots: -22 KB per CTA at 256 threads, so that the ne = 20 000 mesh fits as
    // well and the spills of the 128-register code find room in L1).  With the compile-time shared-memory offsets two
    // clusters per SM set are 14 % faster than one with 255 registers (C4: 5.42e5 vs 4.76e5 realisation-steps/s,
    // profiles/r2w_cluster_ab.log; before that change the two ran at the same rate).  PK_CLUSTER_MINB=1 keeps one.
    c->cminb = 1;
    const char* e2 = getenv("PK_CLUSTER_MINB");
    // (the older Schur-solve variants, kept for A/B runs, need the two row buffers)
    const char *ef = getenv("PK_CLUSTER_FLAT_PCR"), *el = getenv("PK_CLUSTER_LEADER_SEP");
    const bool older = (ef && atoi(ef) == 1) || (el && atoi(el) == 1);
    if (!(e2 && atoi(e2) == 1) && !older && cfg->scheme != 1 && Tc <= 256 && Tc >= 32 && p.mode == NOISE_SCALAR && 2 * (sm2 + 1024) <= 227 * 1024) {
      // does the 128-register variant really get two CTAs per SM?
      cudaSetDevice(cfg->device);
      const void* k2 = (const void*)pick_cluster_kernel(p, Tc, 2);
      int per_sm = 0;
      if (k2 && cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2) == cudaSuccess &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k2, Tc, sm2) == cudaSuccess && per_sm >= 2) {
        c->cminb = 2; c->smem = sm2;
      }
      cudaGetLastError();
    }
  }
  c->xi = c->xi_x = c->xi_xx = nullptr;
  c->w1_resident = false;
  c->step = 0; c->d_state = nullptr; c->d_dWin = c->d_dWout = nullptr; c->d_cnt = nullptr; c->dW_cap = 0; c->d_scratch = nullptr; c->scratch_cap = 0; c->d_mom = nullptr; c->d_prof = nullptr; c->d_tab[0] = c->d_tab[1] = c->d_tab[2] = nullptr;
  c->timed = false; c->nlaunch = 0; c->h_stage = nullptr;
  cudaError_t e = cudaSetDevice(cfg->device);
  if (e != cudaSuccess) { std::string m = cudaGetErrorString(e); delete c; return fail(nullptr, "cudaSetDevice(%d): %s", cfg->device, m.c_str()); }
  const void* k = cfg->scheme == 1 ? (c->cluster > 1 ? pk_upwind_cluster_kernel_ptr(p, c->block) : (const void*)&g_create_err /* own launch path */)
                  : (c->cluster > 1 ? (const void*)pick_cluster_kernel(p, c->block, c->cminb) : (const void*)pick_kernel(p, c->block, c->fs));
  if (!k) { delete c; return fail(nullptr, "pk_create: no kernel for this layout"); }
  if (c->cluster > 8) cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  e = (cfg->scheme == 1 && c->cluster == 1) ? cudaSuccess : cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem);
  if (e != cudaSuccess) { std::string m = cudaGetErrorString(e); delete c; return fail(nullptr, "cudaFuncSetAttribute(smem=%zu): %s", c->smem, m.c_str()); }
  if (cfg->scheme != 1 && c->cluster == 1) {
    // the lean instantiation of the same kernel (no fused dumps / shared sub-updates in the time loop) needs the attribute too
    const void* k0 = (const void*)pick_kernel(p, c->block, c->fs, 0);
    if (k0 && k0 != k) cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem);
    // persistent CTAs: as many as are resident at once
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, c->block, c->smem);
    if (e != cudaSuccess || per_sm < 1) { std::string m = cudaGetErrorString(e); delete c; return fail(nullptr, "occupancy query failed: %s", m.c_str()); }
    c->slots = per_sm * c->sms;
  }
  for (int q = 0; q < 64; ++q) { cudaEventCreate(&c->evr[0][q]); cudaEventCreate(&c->evr[1][q]); }
  c->ev0 = c->evr[0][0]; c->ev1 = c->evr[1][0]; c->evn = c->evq = 0;
  for (int q = 0; q < 4; ++q) cudaStreamCreateWithFlags(&c->hs[q], cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&c->hev, cudaEventDisableTiming);
  *out = c;
  return 0;
}

extern "C" void pk_destroy(pk_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  if (c->d_state) cudaFree(c->d_state);
  if (c->d_dWin) cudaFree(c->d_dWin);
  if (c->d_dWout) cudaFree(c->d_dWout);
  if (c->d_cnt) cudaFree(c->d_cnt);
  if (c->d_queue) cudaFree(c->d_queue);
  if (c->d_scratch) cudaFree(c->d_scratch);
  if (c->d_mom) cudaFree(c->d_mom);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  for (int q = 0; q < 3; ++q) if (c->d_tab[q]) cudaFree(c->d_tab[q]);
  for (int q = 0; q < 64; ++q) { cudaEventDestroy(c->evr[0][q]); cudaEventDestroy(c->evr[1][q]); }
  cudaEventDestroy(c->hev);
  for (int q = 0; q < 4; ++q) cudaStreamDestroy(c->hs[q]);
  delete c;
}

// chunk-interleaved device copy of a [nxi, ne] table (see Thread::fill_noise)
__global__ void pk_permute_table_kernel(const __grid_constant__ Params p, const double* __restrict__ src, double* __restrict__ dst) {
  const int TM = p.T * p.M;
  const int k = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.ne; i += gridDim.x * blockDim.x)
    dst[(size_t)k * TM + row_slot(p, i)] = src[(size_t)k * p.ne + i];
}

extern "C" int pk_set_tables(pk_ctx* c, const double* xi, const double* xi_x, const double* xi_xx) {
  if (!c) return -1;
  if (PK_NOISE_OF(c->p.mode) != NOISE_SCALAR && c->p.nxi > 0 && !xi) return fail(c, "pk_set_tables: xi is required");
  if (c->p.mode == NOISE_FULL && c->p.nxi > 0 && (!xi_x || !xi_xx)) return fail(c, "pk_set_tables: xi_x and xi_xx are required for PK_NOISE_FULL");
  c->xi = c->xi_x = c->xi_xx = nullptr;
  if (c->p.nxi == 0 || PK_NOISE_OF(c->p.mode) == NOISE_SCALAR) return 0;
  CK(c, cudaSetDevice(c->cfg.device));
  const size_t n = (size_t)c->p.nxi * c->p.T * c->p.M;
  const double* src[3] = {xi, xi_x, xi_xx};
  for (int q = 0; q < 3; ++q) {
    if (!src[q] || (q > 0 && c->p.mode != NOISE_FULL)) continue;
    if (!c->d_tab[q]) CK(c, cudaMalloc(&c->d_tab[q], sizeof(double) * n));
    CK(c, cudaMemset(c->d_tab[q], 0, sizeof(double) * n));
    dim3 g((c->p.ne + 255) / 256, c->p.nxi);
    pk_permute_table_kernel<<<g, 256>>>(c->p, src[q], c->d_tab[q]);
    ++c->nlaunch;
    CK(c, cudaGetLastError());
  }
  CK(c, cudaDeviceSynchronize());
  c->xi = c->d_tab[0]; c->xi_x = c->d_tab[1]; c->xi_xx = c->d_tab[2];
  return 0;
}

extern "C" int64_t pk_launch_count(const pk_ctx* c) { return c ? c->nlaunch : -1; }

extern "C" int pk_set_step_counter(pk_ctx* c, int64_t s) { if (!c) return -1; c->step = s; return 0; }
extern "C" int64_t pk_get_step_counter(const pk_ctx* c) { return c ? c->step : -1; }

extern "C" int pk_get_layout(const pk_ctx* c, int32_t* T, int32_t* M, int32_t* rem, int32_t* smem) {
  if (!c) return -1;
  if (T) *T = c->p.T; if (M) *M = c->p.M; if (rem) *rem = c->p.rem; if (smem) *smem = (int32_t)c->smem;
  return c->cluster;    // >= 1: CTAs per realisation
}

extern "C" int pk_get_work_split(const pk_ctx* c, int32_t* group, int32_t* nb, int32_t* nblk, int32_t* grid) {
  if (!c) return -1;
  if (group) *group = c->slots; if (nb) *nb = c->last_nb; if (nblk) *nblk = c->last_nblk; if (grid) *grid = c->last_grid;
  return 0;
}

// Launch the step kernel for realisations [r0, r0 + nr) of the context on stream st.  `slot` selects one of the
// PK_HOST_CHUNKS queue areas (launches that may run concurrently need their own).  Does not advance c->step.
enum { PK_HOST_CHUNKS = 4 };
static int launch_range(pk_ctx* c, StepArgs a, int nsteps, int r0, int nr, int slot, cudaStream_t st, bool timed) {
  a.r0 = r0; a.Rq = nr;
  if (timed) { const int k = (int)(c->evn % 64); c->ev0 = c->evr[0][k]; c->ev1 = c->evr[1][k]; ++c->evn; }
  if (c->cluster > 1) {
    if (r0 != 0 || nr != c->cfg.R) return fail(c, "internal: the cluster kernel advances the whole ensemble");
    if (timed) CK(c, cudaEventRecord(c->ev0, st));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)(c->cfg.R * c->cluster)); lc.blockDim = dim3((unsigned)c->block);
    lc.dynamicSmemBytes = c->smem; lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)c->cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    CK(c, cudaLaunchKernelEx(&lc, pick_cluster_kernel(c->p, c->block, c->cminb), c->p, a, c->lgTc));
    ++c->nlaunch;
  } else {
    const int R = c->cfg.R;
    if (!c->d_queue) CK(c, cudaMalloc(&c->d_queue, sizeof(int) * (size_t)(R + 1) * PK_HOST_CHUNKS));
    int* queue = c->d_queue + (size_t)slot * (R + 1);
    const int grid = nr < c->slots ? nr : c->slots;
    int nb = 1, nblk = nsteps;
    choose_blocks(nr, nsteps, grid, &nb, &nblk);
    { const char* e = getenv("PK_NBLK"); const int v = e ? atoi(e) : 0; if (v > 0) { nblk = v < nsteps ? v : nsteps; nb = (nsteps + nblk - 1) / nblk; } }
    a.queue = queue; a.nblk = nblk; a.nb = nb;
    c->last_nb = nb; c->last_nblk = nblk; c->last_grid = grid;
    CK(c, cudaMemsetAsync(queue, 0, sizeof(int) * (size_t)(nr + 1), st));
    if (timed) CK(c, cudaEventRecord(c->ev0, st));      // time the kernel alone
    // lean time loop unless this launch dumps inside the kernel or shares Philox sub-updates within warps
    const int feat = (a.diag_every > 0 || a.snap_every > 0 || (!a.dW_in && c->p.nupd >= PK_COOP_NUPD)) ? 1 : 0;
    kern_t k = pick_kernel(c->p, c->block, c->fs, feat);
    k<<<grid, c->block, c->smem, st>>>(c->p, a);
    ++c->nlaunch;
  }
  CK(c, cudaGetLastError());
  if (timed) { CK(c, cudaEventRecord(c->ev1, st)); c->timed = true; }
  return 0;
}

static int step_args(pk_ctx* c, int32_t nsteps, double* u, double* u1, double* dF, double* dG, const double* dW_in,
                     double* dW_out, int32_t* fail_step, int32_t* its_last, int32_t* its_total, StepArgs* a) {
  if (nsteps < 0) return fail(c, "pk_step: nsteps < 0");
  if (c->cfg.scheme == 1) return fail(c, "pk_step: context was created for the upwind scheme; use pk_step_upwind");
  if (!u || !u1 || !dF || !fail_step || !its_last || !its_total) return fail(c, "pk_step: null state/counter pointer");
  if (c->p.mode == NOISE_FULL && !dG) return fail(c, "pk_step: dG is required for PK_NOISE_FULL");
  if (PK_NOISE_OF(c->p.mode) != NOISE_SCALAR && c->p.nxi > 0 && !c->xi) return fail(c, "pk_step: noise tables not set (pk_set_tables)");
  a->u = u; a->u1 = u1; a->dF = dF; a->dG = dG; a->xi = c->xi; a->xi_x = c->xi_x; a->xi_xx = c->xi_xx;
  a->dW_in = dW_in; a->dW_out = dW_out; a->fail_step = fail_step; a->its_last = its_last; a->its_total = its_total;
  a->prof = c->d_prof;
  a->queue = nullptr; a->nblk = nsteps; a->nb = 1; a->r0 = 0; a->Rq = c->cfg.R;
  { const char* e = getenv("PK_CLUSTER_FLAT_PCR"); a->part = (e && atoi(e) == 1) ? 0 : 1;
    e = getenv("PK_CLUSTER_LEADER_SEP"); if (a->part && e && atoi(e) == 1) a->part = 2;
    if (c->cminb == 2) a->part = 1; }      // single row buffer: separators solved redundantly per CTA
  a->R = c->cfg.R; a->step0 = (int)c->step; a->nsteps = nsteps; a->rid0 = c->cfg.rid_offset; a->seed = c->cfg.seed;
  a->diag_every = 0; a->diag_out = nullptr; a->diag_idx = nullptr; a->suite_vals = nullptr; a->suite_ids = nullptr;
  a->snap_every = 0; a->snap_u = nullptr; a->snap_flag = nullptr;
  return 0;
}

extern "C" int pk_step(pk_ctx* c, int32_t nsteps, double* u, double* u1, double* dF, double* dG,
                       const double* dW_in, double* dW_out, int32_t* fail_step, int32_t* its_last,
                       int32_t* its_total, void* stream) {
  if (!c) return -1;
  StepArgs a;
  int rc = step_args(c, nsteps, u, u1, dF, dG, dW_in, dW_out, fail_step, its_last, its_total, &a);
  if (rc) return rc;
  if (nsteps == 0) return 0;
  CK(c, cudaSetDevice(c->cfg.device));
  rc = launch_range(c, a, nsteps, 0, c->cfg.R, 0, (cudaStream_t)stream, true);
  if (rc) return rc;
  c->step += nsteps;
  return 0;
}

static_assert((int)DIAG_NCOL == (int)PK_NDIAG && (int)DIAG_NIDX == (int)PK_NIDX && (int)DG_ENERGY == (int)PK_D_ENERGY && (int)DG_L2U2 == (int)PK_D_L2U2 &&
              (int)DG_H1SEMI2 == (int)PK_D_H1SEMI2 && (int)DG_MASS_U == (int)PK_D_MASS_U && (int)DG_MAX_DU == (int)PK_D_MAX_DU && (int)DG_MIN_DU == (int)PK_D_MIN_DU &&
              (int)DG_MAX_JUMP == (int)PK_D_MAX_JUMP_LOCAL && (int)DG_MAX_U == (int)PK_D_MAX_U && (int)DGI_ARGMAX_DU == (int)PK_I_ARGMAX_DU &&
              (int)DGI_ARGMIN_DU == (int)PK_I_ARGMIN_DU && (int)DGI_ARGMAX_U == (int)PK_I_ARGMAX_U && (int)DGI_ARGMAX_JUMP == (int)PK_I_ARGMAX_JUMP,
              "fused dump layout (pk_core.h) must match include/peakon_b200.h");

extern "C" int pk_step_dump(pk_ctx* c, int32_t nsteps, double* u, double* u1, double* dF, double* dG,
                            const double* dW_in, double* dW_out, int32_t* fail_step, int32_t* its_last,
                            int32_t* its_total, int32_t dump_every, double* diag_out, int32_t* diag_idx,
                            double* suite_vals, int32_t* suite_ids, void* stream) {
  if (!c) return -1;
  if (dump_every < 1 || !diag_out || !diag_idx) return fail(c, "pk_step_dump: dump_every >= 1 and the dump arrays are required");
  if ((suite_vals == nullptr) != (suite_ids == nullptr)) return fail(c, "pk_step_dump: suite_vals and suite_ids go together");
  if (c->step % dump_every != 0) return fail(c, "pk_step_dump: the step counter (%lld) must be a multiple of dump_every", c->step);
  StepArgs a;
  int rc = step_args(c, nsteps, u, u1, dF, dG, dW_in, dW_out, fail_step, its_last, its_total, &a);
  if (rc) return rc;
  if (nsteps == 0) return 0;
  a.diag_every = dump_every; a.diag_out = diag_out; a.diag_idx = diag_idx; a.suite_vals = suite_vals; a.suite_ids = suite_ids;
  CK(c, cudaSetDevice(c->cfg.device));
  rc = launch_range(c, a, nsteps, 0, c->cfg.R, 0, (cudaStream_t)stream, true);
  if (rc) return rc;
  c->step += nsteps;
  return 0;
}

extern "C" int pk_step_snap(pk_ctx* c, int32_t nsteps, double* u, double* u1, double* dF, double* dG,
                            const double* dW_in, double* dW_out, int32_t* fail_step, int32_t* its_last,
                            int32_t* its_total, int32_t snap_every, double* snap_u, int32_t* snap_flag, void* stream) {
  if (!c) return -1;
  if (snap_every < 1 || !snap_u || !snap_flag) return fail(c, "pk_step_snap: snap_every >= 1 and the snapshot arrays are required");
  if (c->step % snap_every != 0) return fail(c, "pk_step_snap: the step counter (%lld) must be a multiple of snap_every", c->step);
  StepArgs a;
  int rc = step_args(c, nsteps, u, u1, dF, dG, dW_in, dW_out, fail_step, its_last, its_total, &a);
  if (rc) return rc;
  if (nsteps == 0) return 0;
  a.snap_every = snap_every; a.snap_u = snap_u; a.snap_flag = snap_flag;
  CK(c, cudaSetDevice(c->cfg.device));
  rc = launch_range(c, a, nsteps, 0, c->cfg.R, 0, (cudaStream_t)stream, true);
  if (rc) return rc;
  c->step += nsteps;
  return 0;
}

extern "C" int pk_reset_warm_start(pk_ctx* c) {
  if (!c) return -1;
  c->w1_resident = false;
  return 0;
}

extern "C" double pk_step_kernel_ms_since(pk_ctx* c, int32_t* count) {
  if (count) *count = 0;
  if (!c || !c->timed) return -1.0;
  if (cudaEventSynchronize(c->ev1) != cudaSuccess) return -1.0;
  long long first = c->evq;
  if (c->evn - first > 64) first = c->evn - 64;          // older pairs have been overwritten
  double tot = 0.0;
  for (long long k = first; k < c->evn; ++k) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->evr[0][k % 64], c->evr[1][k % 64]) != cudaSuccess) return -1.0;
    tot += ms;
  }
  if (count) *count = (int32_t)(c->evn - first);
  c->evq = c->evn;
  return tot;
}

extern "C" double pk_last_step_kernel_ms(pk_ctx* c) {
  if (!c || !c->timed) return -1.0;
  if (cudaEventSynchronize(c->ev1) != cudaSuccess) return -1.0;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return -1.0;
  return (double)ms;
}

extern "C" int pk_step_host(pk_ctx* c, int32_t nsteps, double* u, double* u1, double* dF, double* dG,
                            const double* dW_in, double* dW_out, int32_t* fail_step, int32_t* its_last,
                            int32_t* its_total) {
  if (!c) return -1;
  const bool resident = (!u1 && !dF && !dG);       // w1 = (u1, dF, dG) stays in the context (see the header)
  if (!u || !fail_step || !its_last || !its_total || (!resident && (!u1 || !dF)))
    return fail(c, "pk_step_host: null state/counter pointer");
  CK(c, cudaSetDevice(c->cfg.device));
  const size_t R = c->cfg.R, ne = c->p.ne, fld = R * ne;
  const bool hasG = (c->p.mode == NOISE_FULL);
  if (hasG && !dG && !resident) return fail(c, "pk_step_host: dG is required for PK_NOISE_FULL");
  if (!c->d_state) {
    CK(c, cudaMalloc(&c->d_state, sizeof(double) * 4 * fld));
    CK(c, cudaMalloc(&c->d_cnt, sizeof(int) * 3 * R));
    c->w1_resident = false;
  }
  if (resident && !c->w1_resident) {               // first call in this mode: w1 = 0 (equations.py:71)
    CK(c, cudaMemset(c->d_state + fld, 0, sizeof(double) * 3 * fld));
    c->w1_resident = true;
  }
  if (!resident) c->w1_resident = false;
  const size_t ndw = (size_t)nsteps * R * (size_t)c->p.nxi;
  if (ndw > c->dW_cap) {
    if (c->d_dWin) cudaFree(c->d_dWin);
    if (c->d_dWout) cudaFree(c->d_dWout);
    CK(c, cudaMalloc(&c->d_dWin, sizeof(double) * ndw));
    CK(c, cudaMalloc(&c->d_dWout, sizeof(double) * ndw));
    c->dW_cap = ndw;
  }
  double *du = c->d_state, *du1 = du + fld, *ddF = du + 2 * fld, *ddG = du + 3 * fld;
  if (!c->h_stage) CK(c, cudaMallocHost(&c->h_stage, sizeof(int) * 3 * R));
  StepArgs a;
  int rc = step_args(c, nsteps, du, du1, ddF, hasG ? ddG : nullptr, (dW_in && ndw) ? c->d_dWin : nullptr,
                     (dW_out && ndw) ? c->d_dWout : nullptr, c->d_cnt, c->d_cnt, c->d_cnt, &a);
  if (rc) return rc;
  if (nsteps == 0) return 0;
  // Chunk pipeline: the realisations are split into NC contiguous chunks, each with its own stream doing
  // H2D -> step kernel -> D2H, so the copies of one chunk overlap the kernel of another (the persistent kernels of
  // two chunks share the SMs work-conservingly).  Host buffers should be pinned for the copies to be asynchronous.
  // The three counters travel through ONE pinned staging block per chunk, laid out chunk by chunk as
  // [fail_step | its_total | its_last] (device and staging alike): one H2D of the first two and one D2H of all three
  // per chunk instead of five small copies (each small copy costs a launch latency, as much as the state copies).
  int NC = (c->cluster > 1 || R < 256) ? 1 : 2;
  { const char* e = getenv("PK_HOST_NCHUNK"); const int v = e ? atoi(e) : 0; if (v >= 1 && v <= PK_HOST_CHUNKS && c->cluster == 1) NC = v; }
  if (dW_in && ndw) {       // [nsteps, R, nxi]: not contiguous per chunk -> one copy up front, every chunk waits for it
    CK(c, cudaMemcpyAsync(c->d_dWin, dW_in, sizeof(double) * ndw, cudaMemcpyHostToDevice, c->hs[0]));
    CK(c, cudaEventRecord(c->hev, c->hs[0]));
  }
  for (int k = 0; k < NC; ++k) {
    const size_t r0 = R * k / NC, r1 = R * (k + 1) / NC, nr = r1 - r0, o = r0 * ne, nb = sizeof(double) * nr * ne;
    cudaStream_t st = c->hs[k];
    if (dW_in && ndw && k > 0) CK(c, cudaStreamWaitEvent(st, c->hev, 0));
    int* hs = c->h_stage + 3 * r0;
    memcpy(hs, fail_step + r0, sizeof(int) * nr);
    memcpy(hs + nr, its_total + r0, sizeof(int) * nr);
    int* dc = c->d_cnt + 3 * r0;
    CK(c, cudaMemcpyAsync(du + o, u + o, nb, cudaMemcpyHostToDevice, st));
    if (!resident) {
      CK(c, cudaMemcpyAsync(du1 + o, u1 + o, nb, cudaMemcpyHostToDevice, st));
      CK(c, cudaMemcpyAsync(ddF + o, dF + o, nb, cudaMemcpyHostToDevice, st));
      if (hasG) CK(c, cudaMemcpyAsync(ddG + o, dG + o, nb, cudaMemcpyHostToDevice, st));
    }
    CK(c, cudaMemcpyAsync(dc, hs, sizeof(int) * 2 * nr, cudaMemcpyHostToDevice, st));
    // the kernels index the counters by the realisation number r in [r0, r1)
    a.fail_step = dc - r0; a.its_total = dc + nr - r0; a.its_last = dc + 2 * nr - r0;
    rc = launch_range(c, a, nsteps, (int)r0, (int)nr, k, st, false);
    if (rc) return rc;
    CK(c, cudaMemcpyAsync(u + o, du + o, nb, cudaMemcpyDeviceToHost, st));
    if (!resident) {
      CK(c, cudaMemcpyAsync(u1 + o, du1 + o, nb, cudaMemcpyDeviceToHost, st));
      CK(c, cudaMemcpyAsync(dF + o, ddF + o, nb, cudaMemcpyDeviceToHost, st));
      if (hasG) CK(c, cudaMemcpyAsync(dG + o, ddG + o, nb, cudaMemcpyDeviceToHost, st));
    }
    CK(c, cudaMemcpyAsync(hs, dc, sizeof(int) * 3 * nr, cudaMemcpyDeviceToHost, st));
  }
  for (int k = 0; k < NC; ++k) {
    CK(c, cudaStreamSynchronize(c->hs[k]));
    const size_t r0 = R * k / NC, r1 = R * (k + 1) / NC, nr = r1 - r0;
    const int* hs = c->h_stage + 3 * r0;
    memcpy(fail_step + r0, hs, sizeof(int) * nr);
    memcpy(its_total + r0, hs + nr, sizeof(int) * nr);
    memcpy(its_last + r0, hs + 2 * nr, sizeof(int) * nr);
  }
  if (dW_out && ndw) CK(c, cudaMemcpy(dW_out, c->d_dWout, sizeof(double) * ndw, cudaMemcpyDeviceToHost));
  c->timed = false;
  c->step += nsteps;
  return 0;
}

extern "C" int pk_diagnostics_rows(pk_ctx* c, const double* u, int32_t nrows, const double* pq, int32_t periodic, int32_t want_m,
                                   double* out, int32_t* idx, double* du_out, const int32_t* row_flag, void* stream) {
  if (!c) return -1;
  if (!u || !out || !idx || nrows < 1) return fail(c, "pk_diagnostics: null pointer / no rows");
  CK(c, cudaSetDevice(c->cfg.device));
  DiagArgs a;
  a.u = u; a.pq = pq; a.out = out; a.idx = idx; a.du_out = du_out; a.fail_step = row_flag;
  a.ne = c->p.ne; a.periodic = periodic; a.h = c->p.h; a.alphasq = c->p.alphasq; a.Ld = c->cfg.Ld;
  a.scratch = nullptr;
  if (want_m) {
    if (pk_ensure_scratch(c, 2 * (size_t)nrows * c->p.ne)) return -1;
    a.scratch = c->d_scratch;
  }
  // M = kappa (1 - rho S)(1 - rho S^-1), rho = -(2 - sqrt 3); circulant inverse entries
  // g_k = (rho^k + rho^(n-k)) / (kappa (1-rho^2)(1-rho^n))
  const double ma = c->p.ma, mb = c->p.mb;
  const double kappa = 0.5 * (ma + sqrt(ma * ma - 4.0 * mb * mb));
  const double rho = -mb / kappa;
  a.rhoM = rho;
  const double rn = pow(rho, (double)c->p.ne);
  if (c->p.ne >= 160) { a.K = 64; a.gscale = 1.0 / (kappa * (1.0 - rho * rho)); }   // |rho|^64 ~ 2e-37
  else { a.K = -1; a.gscale = 1.0 / (kappa * (1.0 - rho * rho) * (1.0 - rn)); }
  if (a.du_out || a.scratch || a.pq) pk_diag_kernel<true><<<nrows, 256, 0, (cudaStream_t)stream>>>(a);
  else pk_diag_kernel<false><<<nrows, 256, 0, (cudaStream_t)stream>>>(a);
  ++c->nlaunch;
  CK(c, cudaGetLastError());
  return 0;
}

extern "C" int pk_diagnostics(pk_ctx* c, const double* u, const double* pq, int32_t periodic, int32_t want_m,
                              double* out, int32_t* idx, double* du_out, const int32_t* fail_step, void* stream) {
  if (!c) return -1;
  return pk_diagnostics_rows(c, u, c->cfg.R, pq, periodic, want_m, out, idx, du_out, fail_step, stream);
}

extern "C" int pk_ensemble_stats(pk_ctx* c, const double* u, const int32_t* fail_step, const double* diag,
                                 const int32_t* idx, int32_t nbins, double q_lo, double q_hi, double d_lo,
                                 double d_hi, double* packed, void* stream) {
  if (!c) return -1;
  if (!u || !fail_step || !diag || !idx || !packed || nbins < 1) return fail(c, "pk_ensemble_stats: bad argument");
  CK(c, cudaSetDevice(c->cfg.device));
  const int R = c->cfg.R, ne = c->p.ne;
  const int ntile = R < 32 ? 1 : 32;
  const int rows = (R + ntile - 1) / ntile;
  if (!c->d_mom) CK(c, cudaMalloc(&c->d_mom, sizeof(double) * 2 * (size_t)ne * 32));
  cudaStream_t st = (cudaStream_t)stream;
  dim3 g((ne + 127) / 128, ntile);
  pk_moments_partial<<<g, 128, 0, st>>>(u, fail_step, R, ne, rows, c->d_mom);
  ++c->nlaunch;
  CK(c, cudaGetLastError());
  int nb = (2 * ne + 255) / 256; if (nb > 64) nb = 64;
  pk_moments_final<<<nb, 256, sizeof(int) * (2 * nbins + 1), st>>>(c->d_mom, ntile, ne, fail_step, diag, idx, R, nbins,
                                                                c->p.h, q_lo, q_hi, d_lo, d_hi, packed);
  ++c->nlaunch;
  CK(c, cudaGetLastError());
  return 0;
}

extern "C" int pk_peakon_suite_rows(pk_ctx* c, const double* u, int32_t nrows, const int32_t* idx, const int32_t* row_flag,
                                    double* vals, int32_t* ids, void* stream) {
  if (!c) return -1;
  if (!u || !idx || !vals || !ids || nrows < 1) return fail(c, "pk_peakon_suite: null pointer / no rows");
  CK(c, cudaSetDevice(c->cfg.device));
  pk_peakon_suite_kernel<<<nrows, 128, 0, (cudaStream_t)stream>>>(u, idx, row_flag, c->p.ne, c->p.h, vals, ids);
  ++c->nlaunch;
  CK(c, cudaGetLastError());
  return 0;
}

extern "C" int pk_peakon_suite(pk_ctx* c, const double* u, const int32_t* idx, const int32_t* fail_step,
                               double* vals, int32_t* ids, void* stream) {
  if (!c) return -1;
  return pk_peakon_suite_rows(c, u, c->cfg.R, idx, fail_step, vals, ids, stream);
}

// ---- mu diagnostic (find_mus, outputting.py:260-370): one realisation per block, the serial reference logic run by
// thread 0 on shared-memory copies of u and du (global memory when the mesh does not fit) -- pk_mu.h
__global__ void pk_mu_kernel(const double* __restrict__ u, const double* __restrict__ x, const int32_t* __restrict__ fail_step,
                             int ne, double h, double Ld, int use_smem, double* __restrict__ du_scratch,
                             double* __restrict__ out, int32_t* __restrict__ status) {
  extern __shared__ __align__(16) double mu_smem[];
  const int r = blockIdx.x;
  const double* ur = u + (size_t)r * ne;
  double* su = use_smem ? mu_smem : nullptr;
  double* sdu = use_smem ? mu_smem + ne : du_scratch + (size_t)r * ne;
  for (int i = threadIdx.x; i < ne; i += blockDim.x) {
    const int ip = (i + 1 == ne) ? 0 : i + 1;
    const double u0 = ur[i];
    if (su) su[i] = u0;
    sdu[i] = (ur[ip] - u0) / h;              // DG0 slope, the expression of pk_diag_kernel
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double old4[4], alt4[4];
    int rc = -3;
    if (fail_step && fail_step[r] >= 0) {
      for (int j = 0; j < 4; ++j) { old4[j] = pk::nan_value(); alt4[j] = pk::nan_value(); }
    } else {
      rc = pk::mu_serial(su ? su : ur, sdu, x, ne, Ld, old4, alt4);
    }
    for (int j = 0; j < 4; ++j) { out[(size_t)r * 8 + j] = old4[j]; out[(size_t)r * 8 + 4 + j] = alt4[j]; }
    status[r] = rc;
  }
}

extern "C" int pk_mu(pk_ctx* c, const double* u, const double* x_cells, const int32_t* fail_step,
                     double* out, int32_t* status, void* stream) {
  if (!c) return -1;
  if (!u || !x_cells || !out || !status) return fail(c, "pk_mu: null pointer");
  CK(c, cudaSetDevice(c->cfg.device));
  const int ne = c->p.ne, R = c->cfg.R;
  const size_t sm = sizeof(double) * 2 * (size_t)ne;
  const int use_smem = sm <= 200 * 1024;
  if (!use_smem && pk_ensure_scratch(c, 2 * (size_t)R * ne)) return -1;
  if (use_smem && sm > 48 * 1024) CK(c, cudaFuncSetAttribute(pk_mu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  pk_mu_kernel<<<R, 128, use_smem ? sm : 0, (cudaStream_t)stream>>>(u, x_cells, fail_step, ne, c->p.h, c->cfg.Ld, use_smem,
                                                                    c->d_scratch, out, status);
  ++c->nlaunch;
  CK(c, cudaGetLastError());
  return 0;
}

// Profiling builds (-DPK_PROFILE) only: per-phase cycle buckets of the last launches, [R,16] host copy.
extern "C" int pk_profile_read(pk_ctx* c, long long* host_out) {
  if (!c) return -1;
#ifdef PK_PROFILE
  if (!c->d_prof) return fail(c, "no profile yet");
  CK(c, cudaMemcpy(host_out, c->d_prof, sizeof(long long) * 16 * c->cfg.R, cudaMemcpyDeviceToHost));
  return 0;
#else
  (void)host_out;
  return fail(c, "library built without PK_PROFILE");
#endif
}
extern "C" int pk_profile_enable(pk_ctx* c) {
  if (!c) return -1;
#ifdef PK_PROFILE
  if (!c->d_prof) { CK(c, cudaMalloc(&c->d_prof, sizeof(long long) * 16 * c->cfg.R)); }
  CK(c, cudaMemset(c->d_prof, 0, sizeof(long long) * 16 * c->cfg.R));
  return 0;
#else
  return fail(c, "library built without PK_PROFILE");
#endif
}

#ifdef PK_PROFILE_SINGLE_M
// profiling build: only the M=8 kernels are compiled (profiles/phase_cycles.py)
kern_t pk_pick_kernel_m5(int, int, int, int, int) { return nullptr; }
kern_t pk_pick_kernel_m6(int, int, int, int, int) { return nullptr; }
kern_t pk_pick_kernel_m7(int, int, int, int, int) { return nullptr; }
ckern_t pk_pick_cluster_kernel_m5(int, int, int) { return nullptr; }
ckern_t pk_pick_cluster_kernel_m6(int, int, int) { return nullptr; }
ckern_t pk_pick_cluster_kernel_m7(int, int, int) { return nullptr; }
ckern_t pk_pick_cluster_kernel_m9(int, int, int) { return nullptr; }
#ifndef PK_VARIANT_WITH_M10
ckern_t pk_pick_cluster_kernel_m10(int, int, int) { return nullptr; }
#endif
#endif
