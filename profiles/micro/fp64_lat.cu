// fp64 latency / throughput microbenchmark for B200 (sm_100a): informs the kernel design.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dep_chain(double* out, double a, double b, int n, long long* cyc) {
  double x = a + threadIdx.x * 1e-9;
  long long t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < n; ++i) x = fma(x, b, a);
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
template <int ILP>
__global__ void ilp_chain(double* out, double a, double b, int n, long long* cyc) {
  double x[ILP];
  #pragma unroll
  for (int k = 0; k < ILP; ++k) x[k] = a + k + threadIdx.x * 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    #pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], b, a);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  double s = 0; 
  #pragma unroll
  for (int k = 0; k < ILP; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void rcp_chain(double* out, double a, int n, long long* cyc) {
  double x = a + threadIdx.x * 1e-9;
  long long t0 = clock64();
  #pragma unroll 4
  for (int i = 0; i < n; ++i) x = __drcp_rn(x) + 0.5;
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void div_chain(double* out, double a, int n, long long* cyc) {
  double x = a + threadIdx.x * 1e-9;
  long long t0 = clock64();
  #pragma unroll 4
  for (int i = 0; i < n; ++i) x = 1.0 / x + 0.5;
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void lds_chain(double* out, int n, long long* cyc) {
  __shared__ double s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (double)((i * 7 + 1) & 1023);
  __syncthreads();
  double x = threadIdx.x;
  long long t0 = clock64();
  #pragma unroll 4
  for (int i = 0; i < n; ++i) x = s[((int)x) & 1023];
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void bar_chain(double* out, int n, long long* cyc) {
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = 1.0;
}
__global__ void shfl_chain(double* out, int n, long long* cyc) {
  double x = threadIdx.x;
  long long t0 = clock64();
  #pragma unroll 4
  for (int i = 0; i < n; ++i) x = __shfl_up_sync(0xffffffffu, x, 1) + 1.0;
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8 << 20); cudaMalloc(&cyc, 64);
  long long h; const int n = 4096;
  auto rd = [&]() { cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); return (double)h; };
  dep_chain<<<1, 32>>>(out, 1.0, 0.999, n, cyc); printf("DFMA dependent chain, 1 warp         : %.2f cyc/instr\n", rd() / n);
  for (int w : {1, 2, 4, 8, 16}) { dep_chain<<<1, 32 * 4 * w>>>(out, 1.0, 0.999, n, cyc); printf("DFMA dependent chain, %2d warps/SMSP   : %.2f cyc/instr/warp\n", w, rd() / n); }
  ilp_chain<2><<<1, 32>>>(out, 1.0, 0.999, n, cyc); printf("DFMA ILP2, 1 warp                    : %.2f cyc/instr\n", rd() / (2.0 * n));
  ilp_chain<4><<<1, 32>>>(out, 1.0, 0.999, n, cyc); printf("DFMA ILP4, 1 warp                    : %.2f cyc/instr\n", rd() / (4.0 * n));
  ilp_chain<8><<<1, 32>>>(out, 1.0, 0.999, n, cyc); printf("DFMA ILP8, 1 warp                    : %.2f cyc/instr\n", rd() / (8.0 * n));
  ilp_chain<8><<<1, 32 * 4 * 2>>>(out, 1.0, 0.999, n, cyc); printf("DFMA ILP8, 2 warps/SMSP              : %.2f cyc/instr/warp\n", rd() / (8.0 * n));
  ilp_chain<8><<<1, 32 * 4 * 4>>>(out, 1.0, 0.999, n, cyc); printf("DFMA ILP8, 4 warps/SMSP              : %.2f cyc/instr/warp\n", rd() / (8.0 * n));
  rcp_chain<<<1, 32>>>(out, 1.3, n, cyc); printf("__drcp_rn + DADD dependent           : %.1f cyc\n", rd() / n);
  div_chain<<<1, 32>>>(out, 1.3, n, cyc); printf("1.0/x + DADD dependent               : %.1f cyc\n", rd() / n);
  lds_chain<<<1, 32>>>(out, n, cyc); printf("LDS.64 dependent (+cvt)              : %.1f cyc\n", rd() / n);
  bar_chain<<<1, 128>>>(out, n, cyc); printf("__syncthreads, 128 threads           : %.1f cyc\n", rd() / n);
  bar_chain<<<1, 256>>>(out, n, cyc); printf("__syncthreads, 256 threads           : %.1f cyc\n", rd() / n);
  shfl_chain<<<1, 32>>>(out, n, cyc); printf("SHFL(64-bit)+DADD dependent          : %.1f cyc\n", rd() / n);
  return 0;
}
