// pk_upwind_lib.cu -- CUDA kernels (sm_100a) + C ABI entry points of the 'upwind'/'ssprk3' scheme and
// the diagnostics of its prognostic DG1 field m.  Thread-level phases: pk_upwind.h.
#include <cuda_runtime.h>
#include <cstdlib>
#include "pk_kernels.cuh"      // SharedCluster, GpuExecClusterT (the thread-block-cluster executor of the midpoint kernels)
#include "pk_upwind.h"
#include "pk_ctx.h"

// the block executor of pk_lib.cu, restated for this translation unit (same collectives)
template <class ThT, bool AA, int TBC = 0>
struct UwExec {
  typedef ThT Th;
  static __device__ __forceinline__ int nwarps() { return TBC > 0 ? TBC / 32 : (int)(blockDim.x >> 5); }
  Th& th;
  double* red;
  int flip, ne_;
  double rA, rc0, rc1;
  double inv;           // 1 / (1 - rho^ne): the cyclic closure of both scans (the product of all maps' A is rho^ne)
  __device__ __forceinline__ UwExec(Th& t, double* scratch) : th(t), red(scratch), flip(0) {}
  template <class F> __device__ __forceinline__ void each(F f) { if (AA || th.active) f(th); }
  template <class F> __device__ __forceinline__ void each_tid(F f) { f(th, (int)threadIdx.x, (int)blockDim.x, true, 0, ne_); }
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ void tick(int) {}
  template <class F> __device__ __forceinline__ void one(F f) { if (threadIdx.x == 0) f(); }
  __device__ __forceinline__ bool any_flag(int* flag) const { return *flag != 0; }
  __device__ __forceinline__ double* area() {
    const int nw = nwarps();
    double* a = red + flip * 3 * nw;
    flip ^= 1;
    return a;
  }
  template <int DIR>
  __device__ __forceinline__ void cyc_scan(double A, double c0, double& in0) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = nwarps();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double Ao, c0o;
      if (DIR > 0) { Ao = __shfl_up_sync(0xffffffffu, A, o); c0o = __shfl_up_sync(0xffffffffu, c0, o); }
      else { Ao = __shfl_down_sync(0xffffffffu, A, o); c0o = __shfl_down_sync(0xffffffffu, c0, o); }
      const bool ok = (DIR > 0) ? (lane >= o) : (lane + o < 32);
      if (ok) { c0 = A * c0o + c0; A = A * Ao; }
    }
    const int last = (DIR > 0) ? 31 : 0;
    double* sc = area();
    if (lane == last) { sc[w] = A; sc[nw + w] = c0; }
    __syncthreads();
    // compose the warp aggregates in scan order: first the `pos` warps in front of mine (their composition enters my
    // warp), then the rest (warp-uniform branches, no selects; unrolled exactly when the warp count is a constant)
    double tA = 1.0, t0 = 0.0;
    const int pos = (DIR > 0) ? w : nw - 1 - w;
#pragma unroll
    for (int k = 0; k < (TBC > 0 ? TBC / 32 : nw); ++k) {
      if (k < pos) {
        const int ww = (DIR > 0) ? k : nw - 1 - k;
        const double a = sc[ww];
        t0 = a * t0 + sc[nw + ww]; tA = a * tA;
      }
    }
    const double eA = tA, e0 = t0;
#pragma unroll
    for (int k = 0; k < (TBC > 0 ? TBC / 32 : nw); ++k) {
      if (k >= pos) {
        const int ww = (DIR > 0) ? k : nw - 1 - k;
        t0 = sc[ww] * t0 + sc[nw + ww];
      }
    }
    const double x0 = t0 * inv;
    const double w0 = eA * x0 + e0;
    const double o0 = A * w0 + c0;
    const double p0 = (DIR > 0) ? __shfl_up_sync(0xffffffffu, o0, 1) : __shfl_down_sync(0xffffffffu, o0, 1);
    const bool first = (DIR > 0) ? (lane == 0) : (lane == 31);
    in0 = first ? w0 : p0;
  }
  template <class G, class P> __device__ __forceinline__ void scan_fwd(G get, P put) {
    double A = 1.0, c0 = 0.0, c1 = 0.0, in0;
    if (AA || th.active) get(th, A, c0, c1);
    cyc_scan<+1>(A, c0, in0);
    rA = 1.0; rc0 = 0.0; rc1 = 0.0;
    if (AA || th.active) put(th, in0, 0.0, rA, rc0, rc1);
  }
  template <class F> __device__ __forceinline__ void scan_rev(F fin) {
    double in0;
    cyc_scan<-1>(rA, rc0, in0);
    if (AA || th.active) fin(th, in0, 0.0);
  }
};

struct UpwindArgs {
  double *m, *u;
  const double* xi;
  const double* dW_in;
  double* dW_out;
  int* fail_step;
  int R, step0, nsteps;
  unsigned rid0;
  unsigned long long seed;
};

template <int M, int XARR, int TB, int MINB, int TT, bool COOP = false>
__global__ void __launch_bounds__(TB, MINB) pk_upwind_kernel(const __grid_constant__ Params p, const __grid_constant__ UpwindArgs a) {
  extern __shared__ __align__(16) double smem[];
  const int Tn = TT > 0 ? TT : p.T;
  Shared sh;
  sh.U0 = sh.U1 = sh.F = sh.G = sh.X = sh.Xx = sh.Xxx = nullptr;
  sh.xb = smem;
  double* red = smem + UW_SLOTS * Tn;
  sh.dW = red + 6 * (blockDim.x >> 5);
  int* flag = (int*)(sh.dW + (p.nxi > 0 ? p.nxi : 1));
  typedef UpwindThread<M, XARR, TT> Th;
  Th th;
  th.init(p, threadIdx.x);
  UwExec<Th, (TB > 32), TT> ex(th, red);
  ex.ne_ = p.ne; ex.inv = p.inv1mrn;
  UpwindIO io;
  io.m = a.m; io.u = a.u; io.dW_in = a.dW_in; io.dW_out = a.dW_out; io.xi = a.xi;
  io.R = a.R; io.r = blockIdx.x; io.rid = a.rid0 + blockIdx.x; io.seed = a.seed; io.step0 = a.step0; io.nsteps = a.nsteps;
  io.fail_step = a.fail_step;
  run_upwind_steps<M, XARR, TT, UwExec<Th, (TB > 32), TT>, Th, Shared, COOP>(ex, p, sh, io, flag);
}

// Thread-block-cluster variant (meshes beyond 512 threads: ne > 4096): one realisation per cluster of C CTAs.  The state
// still lives in the registers of its threads; the chunk-edge exchange slots are distributed over the CTAs (foreign
// slots read through DSMEM, all stores local) and the two Helmholtz scans get a cross-CTA level -- the executor and the
// shared-memory decode of the midpoint cluster kernels.
template <int M, int XARR, int TB>
__global__ void __launch_bounds__(TB, 1) pk_upwind_cluster_kernel(const __grid_constant__ Params p, const __grid_constant__ UpwindArgs a, int lgTc) {
  extern __shared__ __align__(16) double smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank(), C = cluster.num_blocks();
  const int Tc = blockDim.x;
  SharedCluster sh;
  sh.Tc_ = Tc; sh.lgTc_ = lgTc; sh.rank = rank; sh.base = sh.nbL = sh.nbR = nullptr; sh.edge = nullptr; sh.edgeL = sh.sepR = nullptr;
  { int lg = 0; while ((1 << lg) < p.T) ++lg; sh.lgT = lg; }
  sh.U0 = sh.U1 = sh.F = sh.G = sh.X = sh.Xx = sh.Xxx = nullptr;
  sh.xb = smem;
  double* red = smem + UW_SLOTS * Tc;
  sh.dW = red + cluster_red_doubles(Tc);
  int* flag = (int*)(sh.dW + (p.nxi > 0 ? p.nxi : 1));
  if (threadIdx.x == 0) *flag = 0;
  typedef UpwindThread<M, XARR, 0> Th;
  Th th;
  th.init(p, (int)rank * Tc + (int)threadIdx.x);
  GpuExecClusterT<Th> ex(th, red);
  const int r = blockIdx.x / C;
  ex.rank = rank; ex.C = C; ex.part = 0;
  ex.g0 = chunk_start(p, (int)rank * Tc); ex.g1 = (rank == C - 1) ? p.ne : chunk_start(p, ((int)rank + 1) * Tc);
  UpwindIO io;
  io.m = a.m; io.u = a.u; io.dW_in = a.dW_in; io.dW_out = a.dW_out; io.xi = a.xi;
  io.R = a.R; io.r = r; io.rid = a.rid0 + r; io.seed = a.seed; io.step0 = a.step0; io.nsteps = a.nsteps;
  io.fail_step = a.fail_step;
  cluster.sync();
  run_upwind_steps<M, XARR, 0, GpuExecClusterT<Th>, Th, SharedCluster, true>(ex, p, sh, io, flag);
  cluster.sync();     // nobody leaves while a peer may still read its shared memory
}

typedef void (*ukern_t)(const Params, const UpwindArgs);
typedef void (*uckern_t)(const Params, const UpwindArgs, int);
template <int M> static uckern_t pick_uck(bool x, int Tc) {
  if (Tc <= 256) return x ? pk_upwind_cluster_kernel<M, 1, 256> : pk_upwind_cluster_kernel<M, 0, 256>;
  return x ? pk_upwind_cluster_kernel<M, 1, 512> : pk_upwind_cluster_kernel<M, 0, 512>;
}
static uckern_t pick_ucluster_kernel(const Params& p, int Tc) {
  const bool x = (p.mode != NOISE_SCALAR);
  switch (p.M) {
    case 5: return pick_uck<5>(x, Tc);
    case 6: return pick_uck<6>(x, Tc);
    case 7: return pick_uck<7>(x, Tc);
    case 8: return pick_uck<8>(x, Tc);
  }
  return nullptr;
}
const void* pk_upwind_cluster_kernel_ptr(const Params& p, int Tc) { return (const void*)pick_ucluster_kernel(p, Tc); }

template <int M, int XARR, bool COOP> static ukern_t pick_utb(int block, int T) {
  if (T < 32) return pk_upwind_kernel<M, XARR, 32, 1, 0, COOP>;
  if (T == 128) return pk_upwind_kernel<M, XARR, 128, 4, 128, COOP>;
  if (block <= 128) return pk_upwind_kernel<M, XARR, 128, 4, 0, COOP>;
  if (block <= 256) return pk_upwind_kernel<M, XARR, 256, 2, 256, COOP>;
  return pk_upwind_kernel<M, XARR, 512, 1, 512, COOP>;
}
template <bool COOP> static ukern_t pick_ukernel_t(const Params& p, int block) {
  const bool x = (p.mode != NOISE_SCALAR);
  switch (p.M) {
    case 5: return x ? pick_utb<5, 1, COOP>(block, p.T) : pick_utb<5, 0, COOP>(block, p.T);
    case 6: return x ? pick_utb<6, 1, COOP>(block, p.T) : pick_utb<6, 0, COOP>(block, p.T);
    case 7: return x ? pick_utb<7, 1, COOP>(block, p.T) : pick_utb<7, 0, COOP>(block, p.T);
    case 8: return x ? pick_utb<8, 1, COOP>(block, p.T) : pick_utb<8, 0, COOP>(block, p.T);
  }
  return nullptr;
}
// coop: this launch draws Philox increments with many sub-updates (shared by the lanes of a warp)
static ukern_t pick_ukernel(const Params& p, int block, bool coop) {
  return coop ? pick_ukernel_t<true>(p, block) : pick_ukernel_t<false>(p, block);
}

extern "C" int pk_step_upwind(pk_ctx* c, int32_t nsteps, double* m, double* u, const double* dW_in,
                              double* dW_out, int32_t* fail_step, void* stream) {
  if (!c) return -1;
  if (c->cfg.scheme != 1) return fail(c, "pk_step_upwind: context was not created for the upwind scheme");
  if (nsteps < 0) return fail(c, "pk_step_upwind: nsteps < 0");
  if (!m || !u || !fail_step) return fail(c, "pk_step_upwind: null state pointer");
  if (c->p.mode != NOISE_SCALAR && c->p.nxi > 0 && !c->xi) return fail(c, "pk_step_upwind: noise table not set (pk_set_tables)");
  if (nsteps == 0) return 0;
  CK(c, cudaSetDevice(c->cfg.device));
  UpwindArgs a;
  a.m = m; a.u = u; a.xi = c->xi; a.dW_in = dW_in; a.dW_out = dW_out; a.fail_step = fail_step;
  a.R = c->cfg.R; a.step0 = (int)c->step; a.nsteps = nsteps; a.rid0 = c->cfg.rid_offset; a.seed = c->cfg.seed;
  cudaStream_t st = (cudaStream_t)stream;
  { const int k = (int)(c->evn % 64); c->ev0 = c->evr[0][k]; c->ev1 = c->evr[1][k]; ++c->evn; }     // event ring (pk_step_kernel_ms_since)
  CK(c, cudaEventRecord(c->ev0, st));
  if (c->cluster > 1) {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)(c->cfg.R * c->cluster)); lc.blockDim = dim3((unsigned)c->block);
    lc.dynamicSmemBytes = c->smem; lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)c->cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    CK(c, cudaLaunchKernelEx(&lc, pick_ucluster_kernel(c->p, c->block), c->p, a, c->lgTc));
  } else {
    ukern_t k = pick_ukernel(c->p, c->block, !dW_in && c->p.nupd >= PK_COOP_NUPD);
    k<<<c->cfg.R, c->block, c->smem, st>>>(c->p, a);
  }
  ++c->nlaunch;
  CK(c, cudaGetLastError());
  CK(c, cudaEventRecord(c->ev1, st));
  c->timed = true;
  c->step += nsteps;
  return 0;
}

// ---- diagnostics of the prognostic DG1 field m (outputting.py:112,118,120,164,174 with m prognostic) ----
// writes the m-columns of the pk_diagnostics table: PK_D_L2M2, PK_D_H1SEMI_M2, PK_D_MASS_M, PK_D_M_MAX, PK_D_UMU
__global__ void __launch_bounds__(256) pk_diag_dg1_kernel(const double* __restrict__ mall, const double* __restrict__ uall,
                                                          const int* __restrict__ fail_step, int ne, double h,
                                                          double* __restrict__ out) {
  __shared__ double sm[5][8];
  const int r = blockIdx.x;
  if (fail_step && fail_step[r] >= 0) return;      // pk_diagnostics already wrote NaN
  const double* m = mall + 2LL * r * ne;
  const double* u = uall + (long long)r * ne;
  double l2 = 0.0, semi = 0.0, mass = 0.0, mx = -INFINITY, um = 0.0;
  for (int i = threadIdx.x; i < ne; i += blockDim.x) {
    const double a = m[2 * i], b = m[2 * i + 1];
    const double u0 = u[i], u1 = u[i + 1 == ne ? 0 : i + 1];
    l2 += (h * (1.0 / 3.0)) * (a * a + a * b + b * b);
    semi += (b - a) * (b - a) / h;
    mass += 0.5 * h * (a + b);
    mx = fmax(mx, fmax(a, b));
    um += (h * (1.0 / 6.0)) * (2.0 * u0 * a + u0 * b + u1 * a + 2.0 * u1 * b);
  }
  double v[5] = {l2, semi, mass, mx, um};
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const double w = __shfl_xor_sync(0xffffffffu, v[k], o);
      v[k] = (k == 3) ? fmax(v[k], w) : v[k] + w;
    }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) for (int k = 0; k < 5; ++k) sm[k][w] = v[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    double s[5] = {0.0, 0.0, 0.0, -INFINITY, 0.0};
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i)
      for (int k = 0; k < 5; ++k) s[k] = (k == 3) ? fmax(s[k], sm[k][i]) : s[k] + sm[k][i];
    double* o = out + (long long)r * PK_NDIAG;
    o[PK_D_L2M2] = s[0]; o[PK_D_H1SEMI_M2] = s[1]; o[PK_D_MASS_M] = s[2]; o[PK_D_M_MAX] = s[3]; o[PK_D_UMU] = s[4];
  }
}

extern "C" int pk_diagnostics_dg1(pk_ctx* c, const double* m, const double* u, const int32_t* fail_step,
                                  double* out, void* stream) {
  if (!c) return -1;
  if (!m || !u || !out) return fail(c, "pk_diagnostics_dg1: null pointer");
  CK(c, cudaSetDevice(c->cfg.device));
  pk_diag_dg1_kernel<<<c->cfg.R, 256, 0, (cudaStream_t)stream>>>(m, u, fail_step, c->p.ne, c->p.h, out);
  ++c->nlaunch;
  CK(c, cudaGetLastError());
  return 0;
}
