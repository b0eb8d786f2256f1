// Micro-benchmark: all-gather of 3 doubles per CTA inside a thread-block cluster of 8 (the exchange step of the norm
// all-reduce and of the cyclic scans of the cluster kernels), two ways:
//   pull: every CTA writes its aggregate to its own shared memory, cluster barrier (release/acquire), one warp reads the
//         C aggregates through DSMEM (what pk_kernels.cuh did up to round 2);
//   push: one warp sends the aggregate to every CTA with st.async (completion counted by an mbarrier in the destination
//         CTA), the destination waits on its local mbarrier and reads its own shared memory -- no fence, no barrier.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/cluster_gather profiles/micro/cluster_gather.cu
#include <cooperative_groups.h>
#include <cstdint>
#include <cstdio>
namespace cg = cooperative_groups;
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r; }
__device__ __forceinline__ void mb_init(uint32_t mb, int cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mb), "r"(cnt) : "memory"); }
__device__ __forceinline__ void mb_arm(uint32_t mb, int tx) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"(tx) : "memory"); }
__device__ __forceinline__ void mb_wait(uint32_t mb, int ph) {
  uint32_t ok;
  do { asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(mb), "r"(ph) : "memory"); } while (!ok);
}
__device__ __forceinline__ void push2(uint32_t dst, double a, double b, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" :: "r"(dst), "d"(a), "d"(b), "r"(mbar) : "memory");
}
__device__ __forceinline__ void push1(uint32_t dst, double a, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" :: "r"(dst), "d"(a), "r"(mbar) : "memory");
}

__global__ void __cluster_dims__(8, 1, 1) k_push(double* out, long long* cyc, int iters) {
  __shared__ __align__(16) double inb[2][16][4];
  __shared__ __align__(8) unsigned long long mb[2];
  __shared__ double res;
  cg::cluster_group cl = cg::this_cluster();
  const unsigned rank = cl.block_rank(), C = cl.num_blocks();
  if (threadIdx.x == 0) {
    mb_init(s32(&mb[0]), 1); mb_init(s32(&mb[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mb_arm(s32(&mb[0]), C * 24); mb_arm(s32(&mb[1]), C * 24);
  }
  cl.sync();
  double acc = 0.0; int g = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int ph = g & 1, par = (g >> 1) & 1;
    if (threadIdx.x < 32) {
      const unsigned lane = threadIdx.x;
      if (lane < C) {
        const uint32_t d = mapa(s32(&inb[ph][rank][0]), lane), m = mapa(s32(&mb[ph]), lane);
        push2(d, (double)(rank + 1) + acc * 1e-3, 2.0 * rank + it, m); push1(d + 16, 3.0 * rank - it, m);
      }
      mb_wait(s32(&mb[ph]), par);
      double a = 0.0;
      if (lane < C) a = inb[ph][lane][0] + inb[ph][lane][1] + inb[ph][lane][2];
      for (int o = 8; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) { res = a; mb_arm(s32(&mb[ph]), C * 24); }
    }
    ++g;
    __syncthreads();
    acc = res;
    __syncthreads();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x] = acc; cyc[blockIdx.x] = t1 - t0; }
  cl.sync();
}

__global__ void __cluster_dims__(8, 1, 1) k_pull(double* out, long long* cyc, int iters) {
  __shared__ __align__(16) double agg[2][4];
  __shared__ double res;
  cg::cluster_group cl = cg::this_cluster();
  const unsigned rank = cl.block_rank(), C = cl.num_blocks();
  cl.sync();
  double acc = 0.0; int g = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int ph = g & 1;
    if (threadIdx.x == 0) { agg[ph][0] = (double)(rank + 1) + acc * 1e-3; agg[ph][1] = 2.0 * rank + it; agg[ph][2] = 3.0 * rank - it; }
    cl.sync();
    if (threadIdx.x < 32) {
      const unsigned lane = threadIdx.x;
      double a = 0.0;
      if (lane < C) { const double* q = cl.map_shared_rank(&agg[ph][0], lane); a = q[0] + q[1] + q[2]; }
      for (int o = 8; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) res = a;
    }
    ++g;
    __syncthreads();
    acc = res;
    __syncthreads();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x] = acc; cyc[blockIdx.x] = t1 - t0; }
  cl.sync();
}

int main() {
  const int iters = 2000, nclu = 18, nb = 8 * nclu;
  double *out, *h = new double[nb]; long long *cyc, *hc = new long long[nb];
  cudaMalloc(&out, nb * sizeof(double)); cudaMalloc(&cyc, nb * sizeof(long long));
  for (int variant = 0; variant < 2; ++variant) {
    for (int rep = 0; rep < 2; ++rep) {
      if (variant == 0) k_pull<<<nb, 256>>>(out, cyc, iters); else k_push<<<nb, 256>>>(out, cyc, iters);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, out, nb * sizeof(double), cudaMemcpyDeviceToHost); cudaMemcpy(hc, cyc, nb * sizeof(long long), cudaMemcpyDeviceToHost);
    bool same = true; for (int i = 1; i < nb; ++i) same = same && (h[i] == h[0]);
    printf("%s: result %.17g (identical in all %d CTAs: %s), %.0f cycles per gather + 2 block barriers\n", variant ? "push (st.async + mbarrier)" : "pull (cluster barrier + DSMEM loads)",
           h[0], nb, same ? "yes" : "NO", (double)hc[0] / iters);
  }
  return 0;
}
