// pk_kernels.cuh -- device executors and the step kernels (single-CTA and thread-block-cluster
// variants).  Included by pk_kernels_m.cu, which is compiled once per chunk size M (parallel build).
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdlib>
#include "pk_core.h"

using namespace pk;
namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------------------
// device-side executor: one CUDA thread per chunk
// ------------------------------------------------------------------------------------------------
// AA ("all active") is true for every launch with T >= 32 chunks: then no thread of the block is idle and
// the phase calls are unconditional, which lets the compiler see the per-thread arrays die between phases.
// TBC > 0: block size known at compile time (warp counts fold, the small per-warp loops unroll exactly).
template <class ThT, bool AA, int TBC = 0>
struct GpuExecT {
  typedef ThT Th;
  static __device__ __forceinline__ int nwarps() { return TBC > 0 ? TBC / 32 : (int)(blockDim.x >> 5); }
  Th& th;
  double* red;          // [2][3*nw] reduction / scan scratch in shared memory, used alternately (ping-pong)
  int flip;             // so that a collective needs one barrier only: the area written by collective k
                        // is next written by collective k+2, i.e. after the barrier of collective k+1
  double rA, rc0, rc1;  // maps of the pending reverse scan
  __device__ __forceinline__ GpuExecT(Th& t, double* scratch) : th(t), red(scratch), flip(0) {}

  template <class F> __device__ __forceinline__ void each(F f) { if (AA || th.active) f(th); }
  int ne_;
  template <class F> __device__ __forceinline__ void each_tid(F f) { f(th, (int)threadIdx.x, (int)blockDim.x, true, 0, ne_); }
  __device__ __forceinline__ void sync() { __syncthreads(); }
#ifdef PK_PROFILE
  long long tlast; long long* tbuf;   // per-phase cycle buckets (thread 0 of each CTA), profiling builds only
  __device__ __forceinline__ void tick(int id) {
    if (threadIdx.x == 0) { const long long c = clock64(); tbuf[id] += c - tlast; tlast = c; }
  }
#else
  __device__ __forceinline__ void tick(int) {}
#endif
  template <class F> __device__ __forceinline__ void one(F f) { if (threadIdx.x == 0) f(); }
  __device__ __forceinline__ bool pingpong() const { return false; }     // no shared memory to spare at 3 CTAs/SM
  __device__ __forceinline__ bool partitioned() const { return false; }
  __device__ __forceinline__ bool redundant() const { return false; }
  template <class F> __device__ __forceinline__ void each_cta_first(int, F) {}
  __device__ __forceinline__ int tc() const { return (int)blockDim.x; }
  __device__ __forceinline__ void sync_local() { __syncthreads(); }
  // compacted Schur solve (pk_core.h, hyb_*): needs >= 32 threads per realisation, all of one block.  Measured
  // 12 % SLOWER than plain PCR on B200 (the serial single-warp phases stall the whole CTA; profiles/r1d_exp_variants.log),
  // so it is only compiled into -DPK_HYBRID_SCHUR experiment builds.
#ifdef PK_HYBRID_SCHUR
  __device__ __forceinline__ bool hybrid(int T) const { return AA && T >= 32; }
#else
  __device__ __forceinline__ bool hybrid(int) const { return false; }
#endif
  template <class F> __device__ __forceinline__ void each_first(int n, F f) { if ((int)threadIdx.x < n) f(th); }
  template <class F> __device__ __forceinline__ void warp0_steps(int n, F f) {
    if (threadIdx.x < 32) {
      for (int s = 0; s < n; ++s) { f(th, s); __syncwarp(); }
    }
  }
  __device__ __forceinline__ double* area() {
    const int nw = nwarps();
    double* a = red + flip * 3 * nw;
    flip ^= 1;
    return a;
  }


  // ---- collectives of the fused dump diagnostics (diag_dump in pk_core.h) ----
  // argmax4sum3: four (value, index) candidates per thread -> largest value, lowest index among equals; three doubles per
  // thread -> sums; results known to every thread.  minint2: two ints per thread -> minima.  scratch: CTA-local shared
  // memory (the exchange area, idle at the end of a step): doubles [0, 4 nw) values, [4 nw, 6 nw) indices, [6 nw, 9 nw) sums;
  // minint2 uses the ints from double offset 9 nw on.
  template <class F> __device__ __forceinline__ void argmax4sum3(F f, double* outv, int* outi, double* outs, double* scratch) {
    double v[4], s3[3] = {0.0, 0.0, 0.0}; int ix[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { v[c] = -INFINITY; ix[c] = 0x7fffffff; }
    if (AA || th.active) f(th, v, ix, s3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double ov = __shfl_xor_sync(0xffffffffu, v[c], o);
        const int oi = __shfl_xor_sync(0xffffffffu, ix[c], o);
        if (ov > v[c] || (ov == v[c] && oi < ix[c])) { v[c] = ov; ix[c] = oi; }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) s3[c] += __shfl_xor_sync(0xffffffffu, s3[c], o);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = nwarps();
    int* si = reinterpret_cast<int*>(scratch + 4 * nw);
    double* ss = scratch + 6 * nw;
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) { scratch[c * nw + w] = v[c]; si[c * nw + w] = ix[c]; }
#pragma unroll
      for (int c = 0; c < 3; ++c) ss[c * nw + w] = s3[c];
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double rv = scratch[c * nw]; int ri = si[c * nw];
      for (int k = 1; k < nw; ++k) {
        const double ov = scratch[c * nw + k]; const int oi = si[c * nw + k];
        if (ov > rv || (ov == rv && oi < ri)) { rv = ov; ri = oi; }
      }
      outv[c] = rv; outi[c] = ri;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double t = 0.0;
      for (int k = 0; k < nw; ++k) t += ss[c * nw + k];
      outs[c] = t;
    }
  }
  template <class F> __device__ __forceinline__ void minint2(F f, int* out, double* scratch) {
    int k[2] = {0x7fffffff, 0x7fffffff};
    if (AA || th.active) f(th, k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      k[0] = min(k[0], __shfl_xor_sync(0xffffffffu, k[0], o));
      k[1] = min(k[1], __shfl_xor_sync(0xffffffffu, k[1], o));
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = nwarps();
    int* si = reinterpret_cast<int*>(scratch + 9 * nw);
    if (lane == 0) { si[w] = k[0]; si[nw + w] = k[1]; }
    __syncthreads();
    int r0 = si[0], r1 = si[nw];
    for (int q = 1; q < nw; ++q) { r0 = min(r0, si[q]); r1 = min(r1, si[nw + q]); }
    out[0] = r0; out[1] = r1;
  }

  template <class F> __device__ __forceinline__ void allreduce3_nb(F f, double* out) { allreduce3(f, out); }
  __device__ __forceinline__ bool barrier_free_gather() const { return false; }
  template <class F> __device__ __forceinline__ void allreduce3(F f, double* out) {
    double v[3] = {0.0, 0.0, 0.0};
    if (AA || th.active) f(th, v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
      v[1] += __shfl_xor_sync(0xffffffffu, v[1], o);
      v[2] += __shfl_xor_sync(0xffffffffu, v[2], o);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = nwarps();
    double* sc = area();
    if (lane == 0) { sc[w] = v[0]; sc[nw + w] = v[1]; sc[2 * nw + w] = v[2]; }
    __syncthreads();
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int i = 0; i < nw; ++i) { s0 += sc[i]; s1 += sc[nw + i]; s2 += sc[2 * nw + i]; }
    out[0] = s0; out[1] = s1; out[2] = s2;
  }

  // Cyclic exclusive scan of affine maps over the block. DIR=+1: thread order 0..T-1, value
  // entering thread t comes from t-1; DIR=-1: order T-1..0. Inactive threads carry the identity.
  template <int DIR>
  __device__ __forceinline__ void cyc_scan(double A, double c0, double c1, double& in0, double& in1) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = nwarps();
    // inclusive scan inside the warp (Kogge-Stone); "earlier" = lower lane for DIR=+1, higher for -1
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double Ao, c0o, c1o;
      if (DIR > 0) {
        Ao = __shfl_up_sync(0xffffffffu, A, o); c0o = __shfl_up_sync(0xffffffffu, c0, o); c1o = __shfl_up_sync(0xffffffffu, c1, o);
      } else {
        Ao = __shfl_down_sync(0xffffffffu, A, o); c0o = __shfl_down_sync(0xffffffffu, c0, o); c1o = __shfl_down_sync(0xffffffffu, c1, o);
      }
      const bool ok = (DIR > 0) ? (lane >= o) : (lane + o < 32);
      if (ok) { c0 = A * c0o + c0; c1 = A * c1o + c1; A = A * Ao; }
    }
    const int last = (DIR > 0) ? 31 : 0;
    double* sc = area();
    if (lane == last) { sc[w] = A; sc[nw + w] = c0; sc[2 * nw + w] = c1; }
    __syncthreads();
    // compose the warp aggregates in scan order; the prefix in front of my warp is the state after `pos`
    // of them (warp-uniform branches, no selects; unrolled exactly when the warp count is a constant)
    double tA = 1.0, t0 = 0.0, t1 = 0.0;
    const int pos = (DIR > 0) ? w : nw - 1 - w;
#pragma unroll
    for (int k = 0; k < (TBC > 0 ? TBC / 32 : nw); ++k) {
      if (k < pos) {
        const int ww = (DIR > 0) ? k : nw - 1 - k;
        const double a = sc[ww];
        t0 = a * t0 + sc[nw + ww]; t1 = a * t1 + sc[2 * nw + ww]; tA = a * tA;
      }
    }
    const double eA = tA, e0 = t0, e1 = t1;
#pragma unroll
    for (int k = 0; k < (TBC > 0 ? TBC / 32 : nw); ++k) {
      if (k >= pos) {
        const int ww = (DIR > 0) ? k : nw - 1 - k;
        const double a = sc[ww];
        t0 = a * t0 + sc[nw + ww]; t1 = a * t1 + sc[2 * nw + ww]; tA = a * tA;
      }
    }
    const double inv = rcp_t<(TBC > 0)>(1.0 - tA);
    const double x0 = t0 * inv, x1 = t1 * inv;          // value entering the first thread (cyclic closure)
    const double w0 = eA * x0 + e0, w1 = eA * x1 + e1;  // value entering my warp
    const double o0 = A * w0 + c0, o1 = A * w1 + c1;    // value leaving my thread
    double p0, p1;
    if (DIR > 0) { p0 = __shfl_up_sync(0xffffffffu, o0, 1); p1 = __shfl_up_sync(0xffffffffu, o1, 1); }
    else { p0 = __shfl_down_sync(0xffffffffu, o0, 1); p1 = __shfl_down_sync(0xffffffffu, o1, 1); }
    const bool first = (DIR > 0) ? (lane == 0) : (lane == 31);
    in0 = first ? w0 : p0; in1 = first ? w1 : p1;
  }

  template <class G, class P> __device__ __forceinline__ void scan_fwd(G get, P put) {
    double A = 1.0, c0 = 0.0, c1 = 0.0, in0, in1;
    if (AA || th.active) get(th, A, c0, c1);
    cyc_scan<+1>(A, c0, c1, in0, in1);
    rA = 1.0; rc0 = 0.0; rc1 = 0.0;
    if (AA || th.active) put(th, in0, in1, rA, rc0, rc1);
  }
  template <class F> __device__ __forceinline__ void scan_rev(F fin) {
    double in0, in1;
    cyc_scan<-1>(rA, rc0, rc1, in0, in1);
    if (AA || th.active) fin(th, in0, in1);
  }
};

template <int M, int MODE, bool AA, int TT, int VAR = 0> using GpuExec = GpuExecT<Thread<M, MODE, TT, VAR>, AA, TT>;

#include "pk_args.h"


template <int MODE> struct NArr { enum { n = (MODE == NOISE_FULL) ? 7 : (MODE == NOISE_X ? 4 : 3) }; };

template <int M, int MODE, int TB, int MINB, int TT, int VAR = 0, int FEAT = FEAT_ALL>
__global__ void __launch_bounds__(TB, MINB) pk_step_kernel(const __grid_constant__ Params p, const __grid_constant__ StepArgs a) {
  extern __shared__ __align__(16) double smem[];
  const int Tn = TT > 0 ? TT : p.T;
  const int n = Tn * M;
  Shared sh;
  sh.U0 = smem; sh.U1 = smem + n; sh.F = smem + 2 * n;
  double* nxt = smem + 3 * n;
  if (MODE == NOISE_FULL) { sh.G = nxt; sh.X = nxt + n; sh.Xx = nxt + 2 * n; sh.Xxx = nxt + 3 * n; nxt += 4 * n; }
  else if (PK_NOISE_OF(MODE) == NOISE_X) { sh.G = nullptr; sh.X = nxt; sh.Xx = nullptr; sh.Xxx = nullptr; nxt += n; }
  else { sh.G = nullptr; sh.X = nullptr; sh.Xx = nullptr; sh.Xxx = nullptr; }
  sh.xb = nxt; nxt += XB_SLOTS * Tn;
  double* red = nxt; nxt += 6 * (blockDim.x >> 5);
  sh.dW = nxt; nxt += (p.nxi > 0 ? p.nxi : 1);
  sh.fac = (VAR == 1) ? nxt + ((nxt - smem) & 1) : nullptr;     // [6*(M-2)*T], 16-byte aligned like the rest

  Thread<M, MODE, TT, VAR> th;
  th.init(p, threadIdx.x);
  GpuExec<M, MODE, (TB > 32), TT, VAR> ex(th, red);
  ex.ne_ = p.ne;
#ifdef PK_PROFILE
  __shared__ long long tb[16];
  if (threadIdx.x < 16) tb[threadIdx.x] = 0;
  __syncthreads();
  ex.tbuf = tb; ex.tlast = clock64();
#endif
  StepIO io;
  io.dW_in = a.dW_in; io.dW_out = a.dW_out; io.xi = a.xi; io.xi_x = a.xi_x; io.xi_xx = a.xi_xx;
  io.R = a.R; io.seed = a.seed; io.step0 = a.step0; io.nsteps = a.nsteps;
  io.fail_step = a.fail_step; io.its_last = a.its_last; io.its_total = a.its_total; io.fail_reason = nullptr;
  io.diag_every = a.diag_every; io.diag_out = a.diag_out; io.diag_idx = a.diag_idx; io.suite_vals = a.suite_vals; io.suite_ids = a.suite_ids;
  io.snap_every = a.snap_every; io.snap_u = a.snap_u; io.snap_flag = a.snap_flag;

  // Persistent CTAs pull work items from a global queue: item i = `nblk` consecutive time steps
  // (block i / R) of realisation i % R.  Block-major order: the previous block of a realisation was handed
  // out R items earlier and is normally long finished; its completion counter is awaited before the state
  // is loaded.  This keeps every resident CTA busy until the queue is dry, whatever R / (CTAs per wave) is.
  __shared__ int s_item;
  const int nitems = a.nb * a.Rq;
  int* const done = a.queue + 1;
  while (true) {
    if (threadIdx.x == 0) {
      const int item = atomicAdd(a.queue, 1);
      if (item < nitems) {
        const int b = item / a.Rq, rl = item - b * a.Rq;
        while (atomicAdd(done + rl, 0) < b) __nanosleep(200);
        __threadfence();
      }
      s_item = item;
    }
    __syncthreads();
    const int item = s_item;
    if (item >= nitems) break;
    const int b = item / a.Rq, rl = item - b * a.Rq, r = a.r0 + rl;
    const long long off = (long long)r * p.ne;
    // coalesced global reads (L2: the state may have been written by another SM during this launch),
    // scattered into the chunk-interleaved shared layout
    for (int i = threadIdx.x; i < p.ne; i += blockDim.x) {
      const int sl = row_slot_c<M, TT>(p, i);
      sh.U0[sl] = __ldcg(a.u + off + i); sh.U1[sl] = __ldcg(a.u1 + off + i); sh.F[sl] = __ldcg(a.dF + off + i);
      if (MODE == NOISE_FULL) sh.G[sl] = __ldcg(a.dG + off + i);
    }
    __syncthreads();
    io.r = r; io.rid = a.rid0 + (unsigned)r;
    io.n0 = b * a.nblk; io.n1 = min(io.n0 + a.nblk, a.nsteps);
    run_steps<M, MODE, TT, GpuExec<M, MODE, (TB > 32), TT, VAR>, Thread<M, MODE, TT, VAR>, Shared, FEAT>(ex, p, sh, io);
    __syncthreads();
    for (int i = threadIdx.x; i < p.ne; i += blockDim.x) {
      const int sl = row_slot_c<M, TT>(p, i);
      a.u[off + i] = sh.U0[sl]; a.u1[off + i] = sh.U1[sl]; a.dF[off + i] = sh.F[sl];
      if (MODE == NOISE_FULL) a.dG[off + i] = sh.G[sl];
    }
    __threadfence();
    __syncthreads();      // all stores of the block are fenced; s_item may be overwritten
    if (threadIdx.x == 0) { __threadfence(); atomicExch(done + rl, b + 1); }
  }
#ifdef PK_PROFILE
  if (a.prof && threadIdx.x < 16) a.prof[blockIdx.x * 16 + threadIdx.x] = tb[threadIdx.x];
#endif
}


// ------------------------------------------------------------------------------------------------
// Thread-block-cluster variant: one realisation per CLUSTER of C CTAs (meshes whose state does not
// fit one SM: ne = 16 384 / 20 000 of BASELINE configs 4 and 5).  Every CTA keeps the rows of its own
// Tc threads in its shared memory; the flat index space j*T + t (T = C*Tc threads) is decoded to
// (owner CTA, local offset) and foreign rows are read through distributed shared memory.  All writes
// are to own rows.  Barriers are cluster barriers.
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// PUSH exchanges of the single-buffer cluster kernels: the producer writes into the CONSUMER's shared memory with
// st.async; the bytes are counted by an mbarrier in the consumer's CTA, which the consumer waits on with acquire
// semantics at cluster scope before it reads its own shared memory.  No fence, no cluster barrier, no DSMEM load on
// the consumer's critical path: 580 instead of 1 022 cycles per all-gather (profiles/micro/cluster_gather.cu).
// Every mbarrier has ONE pending arrival per phase -- the consumer's own arrive.expect_tx(bytes) -- so a phase cannot
// complete before the consumer has armed it, whatever the arrival order of the producers' bytes.
// ------------------------------------------------------------------------------------------------
namespace pkx {
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(unsigned a, unsigned rank) {
  unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r;
}
__device__ __forceinline__ void mb_init(unsigned mb, int cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mb), "r"(cnt) : "memory"); }
__device__ __forceinline__ void mb_arm(unsigned mb, int tx) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"(tx) : "memory"); }
__device__ __forceinline__ void mb_wait(unsigned mb, int parity) {
  unsigned ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(mb), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void push2(unsigned dst, double a, double b, unsigned mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" :: "r"(dst), "d"(a), "d"(b), "r"(mbar) : "memory");
}
__device__ __forceinline__ void push1(unsigned dst, double a, unsigned mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" :: "r"(dst), "d"(a), "r"(mbar) : "memory");
}
// layout of the communication area (doubles from its start; pk_params.h: cluster_comm_doubles())
enum { GATH_OFF = 0,            // [2 phases][16 source CTAs][4]: the three doubles of an all-gather
       MB_OFF = 128,            // mbarriers (8 bytes each): 0, 1 = all-gather phases
       MB_GATH = 0 };
}  // namespace pkx

// TCC > 0: threads per CTA known at compile time (= blockDim.x): the block-local offsets (slot * Tc + thread) fold into
// immediate offsets of the loads and stores -- the runtime shifts of the generic variant were a third of the
// cluster kernel's instructions (profiles/r2qc source page: IMAD / SHF / LEA / BRA 30 %, fp64 26 %).
template <int TCC> struct Log2C { enum { v = 1 + Log2C<TCC / 2>::v }; };
template <> struct Log2C<1> { enum { v = 0 }; };
template <> struct Log2C<0> { enum { v = 0 }; };
// CMP ("compact", the two-CTAs-per-SM instantiations): one PCR row buffer; behind the 18 exchange slots lie the scratch
// rows of the separator solve and a 14-double EDGE area that holds a copy of the tips of the CTA's last thread.
template <int TCC = 0, bool CMP = false>
struct SharedClusterT {
  double *U0, *U1, *F, *G, *X, *Xx, *Xxx, *xb, *dW;
  // CMP: my copy of the last thread's tips, the left CTA's copy (DSMEM), the next segment's separator solution in the
  // scratch rows of the OWN CTA (every CTA solves the separator system itself)
  double* edge; const double *edgeL, *sepR;
  __device__ __forceinline__ void stTip(int c, int, int, double v) const {
    xb[c * tcn() + (int)threadIdx.x] = v;
    if (CMP && (int)threadIdx.x == tcn() - 1) edge[c] = v;
  }
  __device__ __forceinline__ double ldTip(int c, int T, int q) const {
    if (!CMP) return ldL(xb, c, T, q);
    return threadIdx.x == 0 ? edgeL[c] : xb[c * tcn() + (int)threadIdx.x - 1];
  }
  __device__ __forceinline__ double ldXsR(int c, int T, int q) const {
    if (!CMP) return ldR(xb, 14 + c, T, q);
    return (int)threadIdx.x == tcn() - 1 ? sepR[c] : xb[(14 + c) * tcn() + (int)threadIdx.x + 1];
  }
  int Tc_, lgTc_, lgT;
  unsigned rank;
  // rows of the two NEIGHBOUR threads (t-1, t+1 on the ring): generic pointers to element [0] of the neighbour's column
  // of this CTA's / the adjacent CTA's shared memory (map_shared_rank), so that slot c of array a of that thread is
  // nb[(a - base) + c * Tc] -- one load with a constant offset, no owner decode, no branch
  const double *base, *nbL, *nbR;
  __device__ __forceinline__ int tcn() const { return TCC > 0 ? TCC : Tc_; }
  __device__ __forceinline__ int lgtc() const { return TCC > 0 ? (int)Log2C<TCC>::v : lgTc_; }
  __device__ __forceinline__ void set_neighbours(double* smem0, int tl, int tr) {
    base = smem0;
    cg::cluster_group cl = cg::this_cluster();
    nbL = cl.map_shared_rank(smem0 + (tl & (tcn() - 1)), (unsigned)tl >> lgtc());
    nbR = cl.map_shared_rank(smem0 + (tr & (tcn() - 1)), (unsigned)tr >> lgtc());
  }
  __device__ __forceinline__ double ldL(const double* a, int c, int, int) const { return nbL[(a - base) + c * tcn()]; }
  __device__ __forceinline__ double ldR(const double* a, int c, int, int) const { return nbR[(a - base) + c * tcn()]; }
  __device__ __forceinline__ double ld(const double* a, int i) const {
    const int tg = i & ((1 << lgT) - 1), j = i >> lgT;
    const unsigned owner = (unsigned)tg >> lgtc();
    double* pl = const_cast<double*>(a) + j * tcn() + (tg & (tcn() - 1));
    if (owner == rank) return *pl;
    return *cg::this_cluster().map_shared_rank(pl, owner);
  }
  __device__ __forceinline__ void st(double* a, int i, double v) const {
    const int tg = i & ((1 << lgT) - 1), j = i >> lgT;
    a[j * tcn() + (tg & (tcn() - 1))] = v;
  }
  // pairs of doubles (PCR rows): pair slot q of thread tt, 16-byte accesses; remote rows through DSMEM
  __device__ __forceinline__ void ld2(const double* a, int q, int, int tt, double& v0, double& v1) const {
    const unsigned owner = (unsigned)tt >> lgtc();
    const double2* pl = reinterpret_cast<const double2*>(a) + q * tcn() + (tt & (tcn() - 1));
    const double2 v = (owner == rank) ? *pl : *cg::this_cluster().map_shared_rank(const_cast<double2*>(pl), owner);
    v0 = v.x; v1 = v.y;
  }
  // row of a thread of my own CTA: no owner decode
  __device__ __forceinline__ void ld2l(const double* a, int q, int, int tt, double& v0, double& v1) const {
    const double2 v = reinterpret_cast<const double2*>(a)[q * tcn() + (tt & (tcn() - 1))];
    v0 = v.x; v1 = v.y;
  }
  __device__ __forceinline__ double ldl(const double* a, int off) const { return a[off]; }
  __device__ __forceinline__ void stl(double* a, int off, double v) const { a[off] = v; }
  __device__ __forceinline__ void st2o(double* a, int q, int, int, double v0, double v1) const {
    reinterpret_cast<double2*>(a)[q * tcn() + (int)threadIdx.x] = make_double2(v0, v1);
  }
  // slot j of the calling thread: block-local, no decoding of the flat index (most accesses are of this kind)
  __device__ __forceinline__ double ldo(const double* a, int j, int, int) const { return a[j * tcn() + (int)threadIdx.x]; }
  __device__ __forceinline__ double ldc(const double* a, int j, int, int t) const { return a[j * tcn() + (t & (tcn() - 1))]; }
  __device__ __forceinline__ void sto(double* a, int j, int, int, double v) const { a[j * tcn() + (int)threadIdx.x] = v; }
  __device__ __forceinline__ double dw(int k) const { return dW[k]; }
  __device__ __forceinline__ void set_dw(int k, double v) const { dW[k] = v; }
};
typedef SharedClusterT<0, false> SharedCluster;

// PP: second PCR row buffer in the exchange area (XB_SLOTS_PP slots; one barrier per PCR step) -- false in the
// two-CTAs-per-SM instantiations, which keep ONE row buffer (XB_SLOTS slots + a scratch area for the separator solve)
template <class ThT, int TCC = 0, bool ONEWARP = false, bool PP = true>
struct GpuExecClusterT {
  typedef ThT Th;
  static __device__ __forceinline__ int bdim() { return TCC > 0 ? TCC : (int)blockDim.x; }
  static __device__ __forceinline__ int nwarps() { return bdim() >> 5; }
  Th& th;
  double* red;          // [2][3*nw + 3]: warp partials + this CTA's aggregate, ping-pong
  int flip, g0, g1;
  unsigned rank, C;
  double rA, rc0, rc1;
  // push exchanges (single-buffer kernels, !PP): the communication area of this CTA, all-gathers done so far
  double* comm; int ngath;
  static constexpr bool PUSH = !PP;
  __device__ __forceinline__ GpuExecClusterT(Th& t, double* scratch) : th(t), red(scratch), flip(0), comm(nullptr), ngath(0) {}
  // once per launch, before the first cluster barrier of the kernel
  __device__ __forceinline__ void comm_init(double* area) {
    comm = area;
    if (PUSH && threadIdx.x == 0) {
      const unsigned mb = pkx::s32(comm + pkx::MB_OFF);
      pkx::mb_init(mb + 8 * (pkx::MB_GATH + 0), 1); pkx::mb_init(mb + 8 * (pkx::MB_GATH + 1), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      pkx::mb_arm(mb + 8 * (pkx::MB_GATH + 0), (int)C * 24); pkx::mb_arm(mb + 8 * (pkx::MB_GATH + 1), (int)C * 24);
    }
  }
  __device__ __forceinline__ bool barrier_free_gather() const { return PUSH; }
  // warp 0: every CTA's (v0, v1, v2) -- valid in lane 0 -- to every CTA; returns CTA `src`'s triple to lane `src` once
  // all C have arrived (src may be any permutation of the lanes below C)
  __device__ __forceinline__ void gather3(double& v0, double& v1, double& v2, unsigned src) {
    const unsigned lane = threadIdx.x & 31;
    const int ph = ngath & 1, par = (ngath >> 1) & 1;
    v0 = __shfl_sync(0xffffffffu, v0, 0); v1 = __shfl_sync(0xffffffffu, v1, 0); v2 = __shfl_sync(0xffffffffu, v2, 0);
    const unsigned base = pkx::s32(comm), mbl = base + 8 * (pkx::MB_OFF + pkx::MB_GATH + ph);
    if (lane < C) {
      const unsigned dst = pkx::mapa(base + 8 * (pkx::GATH_OFF + (ph * 16 + (int)rank) * 4), lane), mbr = pkx::mapa(mbl, lane);
      pkx::push2(dst, v0, v1, mbr); pkx::push1(dst + 16, v2, mbr);
    }
    pkx::mb_wait(mbl, par);
    v0 = 0.0; v1 = 0.0; v2 = 0.0;
    if (lane < C) {
      const double* q = comm + pkx::GATH_OFF + (ph * 16 + (int)src) * 4;
      v0 = q[0]; v1 = q[1]; v2 = q[2];
    }
    __syncwarp();
    if (lane == 0) pkx::mb_arm(mbl, (int)C * 24);       // phase ngath + 2 of this barrier
  }
  template <class F> __device__ __forceinline__ void each(F f) { f(th); }
  template <class F> __device__ __forceinline__ void each_tid(F f) { f(th, (int)threadIdx.x, bdim(), rank == 0, g0, g1); }
  // Cluster barrier.  cg's sync() lets EVERY thread arrive with release semantics, which ptxas turns into
  // MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in front of every UCGABAR_ARV: with 256-512 threads per SM doing that ~20 times
  // per timestep it was the largest single cost of the cluster kernels (a build whose threads arrive RELAXED behind a
  // CTA-scope fence runs 19 % faster on BASELINE config 4 -- profiles/r2z_cluster_ab.log -- but the PTX memory model
  // does not promise DSMEM readers anything for it).  The barrier below keeps the release, once per CTA: a block barrier
  // makes every thread's shared-memory accesses happen-before warp 0, warp 0 alone arrives with RELEASE semantics at
  // cluster scope (cumulative: it covers what it observed through the block barrier -- the pattern of a grid barrier,
  // bar.sync / one thread fences and signals / bar.sync), the other warps arrive relaxed, everybody waits with ACQUIRE.
  // Measured (profiles/r2za_cluster_ab.log): +4 % with two CTAs per SM (512 threads no longer issue the GPU-scope fence:
  // C4 5.42e5 -> 5.65e5), -2 % with one CTA per SM (the extra block barrier; C5 4.33e5 -> 4.23e5): ONEWARP is set by the
  // two-CTAs-per-SM instantiations only.  What remains is the latency of the one MEMBAR.ALL.GPU on the critical path.
  // -DPK_CG_CLUSTER_BARRIER restores cg::cluster_group::sync() everywhere.
  __device__ __forceinline__ void sync() {
#if defined(PK_RELAXED_CLUSTER_BARRIER)
    __threadfence_block();
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
#elif defined(PK_CG_CLUSTER_BARRIER)
    cg::this_cluster().sync();
#else
    if (ONEWARP) {
      __syncthreads();
      if (threadIdx.x < 32) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      else asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
      cg::this_cluster().sync();
    }
#endif
  }
#ifdef PK_PROFILE
  long long tlast; long long* tbuf;
  __device__ __forceinline__ void tick(int id) {
    if (threadIdx.x == 0) { const long long c = clock64(); tbuf[id] += c - tlast; tlast = c; }
  }
#else
  __device__ __forceinline__ void tick(int) {}
#endif
  template <class F> __device__ __forceinline__ void one(F f) { if (rank == 0 && threadIdx.x == 0) f(); }
  // a per-CTA flag in shared memory, raised by any thread: true if any CTA of the cluster has raised its own
  // (call after a cluster barrier that follows the writes)
  __device__ __forceinline__ bool any_flag(int* flag) const {
    cg::cluster_group cl = cg::this_cluster();
    int v = 0;
    for (unsigned b = 0; b < C; ++b) v |= *cl.map_shared_rank(flag, b);
    return v != 0;
  }
  __device__ __forceinline__ bool hybrid(int) const { return false; }     // rows are spread over the CTAs of the cluster
  __device__ __forceinline__ bool pingpong() const { return PP; }         // XB_SLOTS_PP exchange slots: one barrier per PCR step
  // Schur solve partitioned by CTA (pk_core.h, pcr_step_local): the PCR steps inside a CTA need block barriers and
  // block-local loads only; the C separator rows are solved by the leading threads.  part = 0 keeps the flat PCR
  // over all C*Tc rows (every step a cluster barrier and, from stride Tc/2 on, DSMEM-bandwidth-bound row loads).
  int part;
  __device__ __forceinline__ bool partitioned() const { return part != 0; }
  __device__ __forceinline__ bool redundant() const { return part == 1 || !PP; }      // part = 2: separators by the leading threads
  template <class F> __device__ __forceinline__ void each_cta_first(int n, F f) { if ((int)threadIdx.x < n) f(th); }
  __device__ __forceinline__ int tc() const { return bdim(); }
  __device__ __forceinline__ void sync_local() { __syncthreads(); }
  template <class F> __device__ __forceinline__ void each_first(int, F) {}
  template <class F> __device__ __forceinline__ void warp0_steps(int, F) {}
  __device__ __forceinline__ double* area() {
    double* a = red + flip * cluster_red_half(bdim());
    flip ^= 1;
    return a;
  }

  // collectives of the fused dump diagnostics: block-level result, then the C block aggregates through DSMEM
  // (scratch: doubles [0, 4 nw) values, [4 nw, 6 nw) indices, [6 nw, 9 nw) sums, this CTA's aggregate [9 nw, 9 nw + 9), the
  //  cluster's result [9 nw + 9, 9 nw + 18); minint2: ints from double offset 9 nw + 18)
  template <class F> __device__ __forceinline__ void argmax4sum3(F f, double* outv, int* outi, double* outs, double* scratch) {
    double v[4], s3[3] = {0.0, 0.0, 0.0}; int ix[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { v[c] = -INFINITY; ix[c] = 0x7fffffff; }
    f(th, v, ix, s3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double ov = __shfl_xor_sync(0xffffffffu, v[c], o);
        const int oi = __shfl_xor_sync(0xffffffffu, ix[c], o);
        if (ov > v[c] || (ov == v[c] && oi < ix[c])) { v[c] = ov; ix[c] = oi; }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) s3[c] += __shfl_xor_sync(0xffffffffu, s3[c], o);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = nwarps();
    int* si = reinterpret_cast<int*>(scratch + 4 * nw);
    double* ss = scratch + 6 * nw;
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) { scratch[c * nw + w] = v[c]; si[c * nw + w] = ix[c]; }
#pragma unroll
      for (int c = 0; c < 3; ++c) ss[c * nw + w] = s3[c];
    }
    __syncthreads();
    double* agg = scratch + 9 * nw;                       // 4 values | 4 indices (2 doubles) | 3 sums
    int* aggi = reinterpret_cast<int*>(agg + 4);
    if (threadIdx.x == 0) {
      for (int c = 0; c < 4; ++c) {
        double rv = scratch[c * nw]; int ri = si[c * nw];
        for (int k = 1; k < nw; ++k) {
          const double ov = scratch[c * nw + k]; const int oi = si[c * nw + k];
          if (ov > rv || (ov == rv && oi < ri)) { rv = ov; ri = oi; }
        }
        agg[c] = rv; aggi[c] = ri;
      }
      for (int c = 0; c < 3; ++c) {
        double t = 0.0;
        for (int k = 0; k < nw; ++k) t += ss[c * nw + k];
        agg[6 + c] = t;
      }
    }
    cg::cluster_group cl = cg::this_cluster();
    sync();
    // one warp fetches the C aggregates (lane b <- CTA b), combines them and broadcasts through local shared memory
    double* fin = agg + 9;                                // 4 values | 4 indices (2 doubles) | 3 sums
    int* fini = reinterpret_cast<int*>(fin + 4);
    if (w == 0) {
      double gv[4], gs[3] = {0.0, 0.0, 0.0}; int gi[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) { gv[c] = -INFINITY; gi[c] = 0x7fffffff; }
      if ((unsigned)lane < C) {
        const double* qa = cl.map_shared_rank(agg, lane);
        const int* qi = reinterpret_cast<const int*>(qa + 4);
#pragma unroll
        for (int c = 0; c < 4; ++c) { gv[c] = qa[c]; gi[c] = qi[c]; }
#pragma unroll
        for (int c = 0; c < 3; ++c) gs[c] = qa[6 + c];
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double ov = __shfl_xor_sync(0xffffffffu, gv[c], o);
          const int oi = __shfl_xor_sync(0xffffffffu, gi[c], o);
          if (ov > gv[c] || (ov == gv[c] && oi < gi[c])) { gv[c] = ov; gi[c] = oi; }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) gs[c] += __shfl_xor_sync(0xffffffffu, gs[c], o);
      }
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { fin[c] = gv[c]; fini[c] = gi[c]; }
#pragma unroll
        for (int c = 0; c < 3; ++c) fin[6 + c] = gs[c];
      }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 4; ++c) { outv[c] = fin[c]; outi[c] = fini[c]; }
#pragma unroll
    for (int c = 0; c < 3; ++c) outs[c] = fin[6 + c];
  }
  template <class F> __device__ __forceinline__ void minint2(F f, int* out, double* scratch) {
    int k[2] = {0x7fffffff, 0x7fffffff};
    f(th, k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      k[0] = min(k[0], __shfl_xor_sync(0xffffffffu, k[0], o));
      k[1] = min(k[1], __shfl_xor_sync(0xffffffffu, k[1], o));
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = nwarps();
    int* si = reinterpret_cast<int*>(scratch + 9 * nw + 18);
    if (lane == 0) { si[w] = k[0]; si[nw + w] = k[1]; }
    __syncthreads();
    int* agg = si + 2 * nw;                               // [0,1] this CTA's minima, [2,3] the cluster's
    if (threadIdx.x == 0) {
      int r0 = si[0], r1 = si[nw];
      for (int q = 1; q < nw; ++q) { r0 = min(r0, si[q]); r1 = min(r1, si[nw + q]); }
      agg[0] = r0; agg[1] = r1;
    }
    cg::cluster_group cl = cg::this_cluster();
    sync();
    if (w == 0) {
      int g0_ = 0x7fffffff, g1_ = 0x7fffffff;
      if ((unsigned)lane < C) { const int* q = cl.map_shared_rank(agg, lane); g0_ = q[0]; g1_ = q[1]; }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        g0_ = min(g0_, __shfl_xor_sync(0xffffffffu, g0_, o)); g1_ = min(g1_, __shfl_xor_sync(0xffffffffu, g1_, o));
      }
      if (lane == 0) { agg[2] = g0_; agg[3] = g1_; }
    }
    __syncthreads();
    out[0] = agg[2]; out[1] = agg[3];
  }

  // allreduce3: the sums AND a cluster-wide barrier (the callers order other exchanges with it); allreduce3_nb: the sums
  // only -- with push exchanges no barrier at all
  template <class F> __device__ __forceinline__ void allreduce3(F f, double* out) { allreduce3_t<true>(f, out); }
  template <class F> __device__ __forceinline__ void allreduce3_nb(F f, double* out) { allreduce3_t<false>(f, out); }
  template <bool BARRIER, class F> __device__ __forceinline__ void allreduce3_t(F f, double* out) {
    double v[3] = {0.0, 0.0, 0.0};
    f(th, v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
      v[1] += __shfl_xor_sync(0xffffffffu, v[1], o);
      v[2] += __shfl_xor_sync(0xffffffffu, v[2], o);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = nwarps();
    double* sc = area();
    if (lane == 0) { sc[w] = v[0]; sc[nw + w] = v[1]; sc[2 * nw + w] = v[2]; }
    __syncthreads();
    if (PUSH) {
      // warp 0: CTA aggregate (same summation order as below), pushed to every CTA; same combination tree
      if (w == 0) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        if (lane == 0) for (int i = 0; i < nw; ++i) { a0 += sc[i]; a1 += sc[nw + i]; a2 += sc[2 * nw + i]; }
        gather3(a0, a1, a2, (unsigned)lane);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        }
        if (lane == 0) { sc[3 * nw + 3] = a0; sc[3 * nw + 4] = a1; sc[3 * nw + 5] = a2; }
      }
      ++ngath;
      if (BARRIER) sync(); else __syncthreads();
      out[0] = sc[3 * nw + 3]; out[1] = sc[3 * nw + 4]; out[2] = sc[3 * nw + 5];
      return;
    }
    if (threadIdx.x == 0) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
      for (int i = 0; i < nw; ++i) { s0 += sc[i]; s1 += sc[nw + i]; s2 += sc[2 * nw + i]; }
      sc[3 * nw] = s0; sc[3 * nw + 1] = s1; sc[3 * nw + 2] = s2;
    }
    cg::cluster_group cl = cg::this_cluster();
    sync();
    // ONE warp fetches the C aggregates through DSMEM (lane b <- CTA b) and adds them up; the result reaches the other
    // threads through the CTA's own shared memory.  (Every thread reading all C aggregates remotely costs 3 C DSMEM
    // loads per thread -- 49 KB per CTA and collective at C = 8, ~2 400 cycles at the measured DSMEM bandwidth.)
    if (w == 0) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
      if ((unsigned)lane < C) { const double* q = cl.map_shared_rank(sc + 3 * nw, lane); a0 = q[0]; a1 = q[1]; a2 = q[2]; }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {      // C <= 16; the same tree in every CTA: identical bits everywhere
        a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o);
      }
      if (lane == 0) { sc[3 * nw + 3] = a0; sc[3 * nw + 4] = a1; sc[3 * nw + 5] = a2; }
    }
    __syncthreads();
    out[0] = sc[3 * nw + 3]; out[1] = sc[3 * nw + 4]; out[2] = sc[3 * nw + 5];
  }
  template <int DIR>
  __device__ __forceinline__ void cyc_scan(double A, double c0, double c1, double& in0, double& in1) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = nwarps();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double Ao, c0o, c1o;
      if (DIR > 0) {
        Ao = __shfl_up_sync(0xffffffffu, A, o); c0o = __shfl_up_sync(0xffffffffu, c0, o); c1o = __shfl_up_sync(0xffffffffu, c1, o);
      } else {
        Ao = __shfl_down_sync(0xffffffffu, A, o); c0o = __shfl_down_sync(0xffffffffu, c0, o); c1o = __shfl_down_sync(0xffffffffu, c1, o);
      }
      const bool ok = (DIR > 0) ? (lane >= o) : (lane + o < 32);
      if (ok) { c0 = A * c0o + c0; c1 = A * c1o + c1; A = A * Ao; }
    }
    const int last = (DIR > 0) ? 31 : 0;
    double* sc = area();
    if (lane == last) { sc[w] = A; sc[nw + w] = c0; sc[2 * nw + w] = c1; }
    __syncthreads();
    // prefix of the warps in front of mine inside this CTA, and the CTA aggregate
    double tA = 1.0, t0 = 0.0, t1 = 0.0, eA = 1.0, e0 = 0.0, e1 = 0.0;
    for (int k = 0; k < nw; ++k) {
      const int ww = (DIR > 0) ? k : nw - 1 - k;
      if (ww == w) { eA = tA; e0 = t0; e1 = t1; }
      const double a = sc[ww];
      t0 = a * t0 + sc[nw + ww]; t1 = a * t1 + sc[2 * nw + ww]; tA = a * tA;
    }
    cg::cluster_group cl = cg::this_cluster();
    if (!PUSH) {
      if (threadIdx.x == 0) { sc[3 * nw] = tA; sc[3 * nw + 1] = t0; sc[3 * nw + 2] = t1; }
      sync();
    }
    // ONE warp composes the C CTA aggregates (lane b <- the b-th CTA in scan order -- pushed into this CTA's shared
    // memory, or read through DSMEM behind the cluster barrier; Kogge-Stone over the lanes), closes the ring and leaves
    // the value entering THIS CTA in the CTA's own shared memory
    if (w == 0) {
      double mA = 1.0, m0 = 0.0, m1 = 0.0;
      const unsigned cc = (DIR > 0) ? (unsigned)lane : C - 1 - (unsigned)lane;
      if (PUSH) {
        mA = tA; m0 = t0; m1 = t1;                       // (every thread holds the CTA aggregate; lane 0's copy is sent)
        gather3(mA, m0, m1, cc);
        if ((unsigned)lane >= C) mA = 1.0;
      } else if ((unsigned)lane < C) {
        const double* q = cl.map_shared_rank(sc + 3 * nw, cc);
        mA = q[0]; m0 = q[1]; m1 = q[2];
      }
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        const double Ao = __shfl_up_sync(0xffffffffu, mA, o), c0o = __shfl_up_sync(0xffffffffu, m0, o), c1o = __shfl_up_sync(0xffffffffu, m1, o);
        if (lane >= o) { m0 = mA * c0o + m0; m1 = mA * c1o + m1; mA = mA * Ao; }
      }
      // inclusive composition up to lane b in (mA, m0, m1); the whole ring = lane C-1; exclusive prefix = lane b-1
      const double gA = __shfl_sync(0xffffffffu, mA, (int)C - 1), g0_ = __shfl_sync(0xffffffffu, m0, (int)C - 1), g1_ = __shfl_sync(0xffffffffu, m1, (int)C - 1);
      double pA = __shfl_up_sync(0xffffffffu, mA, 1), p0 = __shfl_up_sync(0xffffffffu, m0, 1), p1 = __shfl_up_sync(0xffffffffu, m1, 1);
      if (lane == 0) { pA = 1.0; p0 = 0.0; p1 = 0.0; }
      const double inv = PK_RCP(1.0 - gA);
      const double x0 = g0_ * inv, x1 = g1_ * inv;          // entering the first thread of the ring
      const int mine = (DIR > 0) ? (int)rank : (int)C - 1 - (int)rank;
      if (lane == mine) { sc[3 * nw + 3] = pA * x0 + p0; sc[3 * nw + 4] = pA * x1 + p1; }      // entering my CTA
    }
    if (PUSH) ++ngath;
    __syncthreads();
    const double y0 = sc[3 * nw + 3], y1 = sc[3 * nw + 4];
    const double w0 = eA * y0 + e0, w1 = eA * y1 + e1;    // entering my warp
    const double o0 = A * w0 + c0, o1 = A * w1 + c1;      // leaving my thread
    double q0, q1;
    if (DIR > 0) { q0 = __shfl_up_sync(0xffffffffu, o0, 1); q1 = __shfl_up_sync(0xffffffffu, o1, 1); }
    else { q0 = __shfl_down_sync(0xffffffffu, o0, 1); q1 = __shfl_down_sync(0xffffffffu, o1, 1); }
    const bool first = (DIR > 0) ? (lane == 0) : (lane == 31);
    in0 = first ? w0 : q0; in1 = first ? w1 : q1;
  }
  template <class G, class P> __device__ __forceinline__ void scan_fwd(G get, P put) {
    double A = 1.0, c0 = 0.0, c1 = 0.0, in0, in1;
    get(th, A, c0, c1);
    cyc_scan<+1>(A, c0, c1, in0, in1);
    rA = 1.0; rc0 = 0.0; rc1 = 0.0;
    put(th, in0, in1, rA, rc0, rc1);
  }
  template <class F> __device__ __forceinline__ void scan_rev(F fin) {
    double in0, in1;
    cyc_scan<-1>(rA, rc0, rc1, in0, in1);
    fin(th, in0, in1);
  }
};

template <int M, int MODE, int TCC = 0, bool ONEWARP = false, bool PP = true> using GpuExecCluster = GpuExecClusterT<Thread<M, MODE, 0>, TCC, ONEWARP, PP>;

// TCC = TB: launched with exactly TB threads per CTA (the production layouts); TCC = 0: any block size <= TB (forced
// cluster layouts of small meshes in the tests)
template <int M, int MODE, int TB, int MINB = 1, int TCC = 0>
__global__ void __launch_bounds__(TB, MINB) pk_step_cluster_kernel(const __grid_constant__ Params p, const __grid_constant__ StepArgs a, int lgTc) {
  extern __shared__ __align__(16) double smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank(), C = cluster.num_blocks();
  const int Tc = TCC > 0 ? TCC : (int)blockDim.x, n = Tc * M;
  SharedClusterT<TCC, (MINB > 1)> sh;
  sh.Tc_ = Tc; sh.lgTc_ = lgTc; sh.rank = rank;
  { int lg = 0; while ((1 << lg) < p.T) ++lg; sh.lgT = lg; }
  sh.U0 = smem; sh.U1 = smem + n; sh.F = smem + 2 * n;
  double* nxt = smem + 3 * n;
  if (MODE == NOISE_FULL) { sh.G = nxt; sh.X = nxt + n; sh.Xx = nxt + 2 * n; sh.Xxx = nxt + 3 * n; nxt += 4 * n; }
  else if (PK_NOISE_OF(MODE) == NOISE_X) { sh.G = nullptr; sh.X = nxt; sh.Xx = nullptr; sh.Xxx = nullptr; nxt += n; }
  else { sh.G = nullptr; sh.X = nullptr; sh.Xx = nullptr; sh.Xxx = nullptr; }
  sh.xb = nxt; nxt += cluster_xb_doubles(Tc, MINB > 1);      // two CTAs per SM: one row buffer + separator scratch
  sh.edge = nullptr; sh.edgeL = nullptr; sh.sepR = nullptr;
  if (MINB > 1) {
    sh.edge = sh.xb + XB_SLOTS * Tc + 30 * 16;
    sh.edgeL = cluster.map_shared_rank(sh.edge, (rank + C - 1) % C);
    sh.sepR = sh.xb + XB_SLOTS * Tc + 28 * (int)C + 2 * (int)((rank + 1) % C);      // sep_final's solutions (pk_core.h)
  }
  double* red = nxt; nxt += cluster_red_doubles(Tc);
  sh.dW = nxt; nxt += (p.nxi > 0 ? p.nxi : 1);
  double* comm = nxt + ((nxt - smem) & 1);                    // push exchanges (two CTAs per SM): 16-byte aligned

  const int r = blockIdx.x / C;
  const long long off = (long long)r * p.ne;
  const int t0 = (int)rank * Tc;
  const int g0 = chunk_start(p, t0);
  const int g1 = (rank == C - 1) ? p.ne : chunk_start(p, t0 + Tc);
  // (the rows [g0, g1) of this CTA: block-local slots, no decoding of the flat index)
  for (int i = g0 + threadIdx.x; i < g1; i += Tc) {
    int j, t; row_jt_c<M>(p, i, j, t);
    const int o = j * Tc + (t - t0);
    sh.U0[o] = a.u[off + i]; sh.U1[o] = a.u1[off + i]; sh.F[o] = a.dF[off + i];
    if (MODE == NOISE_FULL) sh.G[o] = a.dG[off + i];
  }

  Thread<M, MODE, 0> th;
  th.init(p, t0 + (int)threadIdx.x);
  sh.set_neighbours(smem, th.tl, th.tr);
  GpuExecCluster<M, MODE, TCC, (MINB > 1), (MINB <= 1)> ex(th, red);
  ex.rank = rank; ex.C = C; ex.g0 = g0; ex.g1 = g1; ex.part = a.part;
  ex.comm_init(comm);
  cluster.sync();       // the state is loaded and the mbarriers of the push exchanges are initialised everywhere
#ifdef PK_PROFILE
  __shared__ long long tb[16];
  if (threadIdx.x < 16) tb[threadIdx.x] = 0;
  __syncthreads();
  ex.tbuf = tb; ex.tlast = clock64();
#endif
  StepIO io;
  io.dW_in = a.dW_in; io.dW_out = a.dW_out; io.xi = a.xi; io.xi_x = a.xi_x; io.xi_xx = a.xi_xx;
  io.R = a.R; io.r = r; io.rid = a.rid0 + (unsigned)r; io.seed = a.seed; io.step0 = a.step0; io.nsteps = a.nsteps;
  io.n0 = 0; io.n1 = a.nsteps;
  io.fail_step = a.fail_step; io.its_last = a.its_last; io.its_total = a.its_total; io.fail_reason = nullptr;
  io.diag_every = a.diag_every; io.diag_out = a.diag_out; io.diag_idx = a.diag_idx; io.suite_vals = a.suite_vals; io.suite_ids = a.suite_ids;
  io.snap_every = a.snap_every; io.snap_u = a.snap_u; io.snap_flag = a.snap_flag;
  run_steps<M, MODE, 0, GpuExecCluster<M, MODE, TCC, (MINB > 1), (MINB <= 1)>, Thread<M, MODE, 0>, SharedClusterT<TCC, (MINB > 1)> >(ex, p, sh, io);
  cluster.sync();     // nobody leaves while a peer may still read its shared memory
#ifdef PK_PROFILE
  if (a.prof && rank == 0 && threadIdx.x < 16) a.prof[r * 16 + threadIdx.x] = tb[threadIdx.x];
#endif

  for (int i = g0 + threadIdx.x; i < g1; i += Tc) {
    int j, t; row_jt_c<M>(p, i, j, t);
    const int o = j * Tc + (t - t0);
    a.u[off + i] = sh.U0[o]; a.u1[off + i] = sh.U1[o]; a.dF[off + i] = sh.F[o];
    if (MODE == NOISE_FULL) a.dG[off + i] = sh.G[o];
  }
}

