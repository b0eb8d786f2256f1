// pk_kernels_m.cu -- instantiates the step kernels for ONE chunk size PK_M (5..8; 9 and 10: cluster kernels only).
#include "pk_kernels.cuh"
#include "pk_dispatch.h"

#ifndef PK_M
#error "compile with -DPK_M=5..10"
#endif


// exact = the launch uses exactly 256 / 512 threads per CTA: instantiation with compile-time shared-memory offsets
template <int M> static ckern_t pick_cmode(int mode, int Tc, int minb) {
  if (Tc <= 256) {
    // two clusters share the SMs when two CTAs fit one SM's shared memory (128 registers per thread)
    if (minb == 2 && mode == NOISE_SCALAR) return Tc == 256 ? pk_step_cluster_kernel<M, NOISE_SCALAR, 256, 2, 256> : pk_step_cluster_kernel<M, NOISE_SCALAR, 256, 2>;
    const char* gen = getenv("PK_CLUSTER_GENERIC");      // A/B: the runtime-offset instantiation for every block size
    if (Tc == 256 && !(gen && atoi(gen) == 1)) {
      switch (mode) {
        case NOISE_SCALAR: return pk_step_cluster_kernel<M, NOISE_SCALAR, 256, 1, 256>;
        case NOISE_X: return pk_step_cluster_kernel<M, NOISE_X, 256, 1, 256>;
        case TEST_SCALAR: return pk_step_cluster_kernel<M, TEST_SCALAR, 256, 1, 256>;
        case TEST_X: return pk_step_cluster_kernel<M, TEST_X, 256, 1, 256>;
        default: return pk_step_cluster_kernel<M, NOISE_FULL, 256, 1, 256>;
      }
    }
    switch (mode) {
      case NOISE_SCALAR: return pk_step_cluster_kernel<M, NOISE_SCALAR, 256>;
      case NOISE_X: return pk_step_cluster_kernel<M, NOISE_X, 256>;
      case TEST_SCALAR: return pk_step_cluster_kernel<M, TEST_SCALAR, 256>;
      case TEST_X: return pk_step_cluster_kernel<M, TEST_X, 256>;
      default: return pk_step_cluster_kernel<M, NOISE_FULL, 256>;
    }
  }
  switch (mode) {
    case NOISE_SCALAR: return pk_step_cluster_kernel<M, NOISE_SCALAR, 512>;
    case NOISE_X: return pk_step_cluster_kernel<M, NOISE_X, 512>;
    case TEST_SCALAR: return pk_step_cluster_kernel<M, TEST_SCALAR, 512>;
    case TEST_X: return pk_step_cluster_kernel<M, TEST_X, 512>;
    default: return pk_step_cluster_kernel<M, NOISE_FULL, 512>;
  }
}

#if PK_M <= 8
// feat = 0: lean time loop (no fused dumps, no warp-shared sub-updates): the headline launches; feat != 0: everything
template <int M, int MODE, int FEAT> static kern_t pick_tb(int block, int T, int minb) {
  if (T < 32) return pk_step_kernel<M, MODE, 32, 1, 0, 0, FEAT_ALL>;    // tiny meshes: some lanes of the single warp idle
  if (T == 128) {
#ifdef PK_EXPERIMENTAL_VARIANTS      // A/B builds only (profiles/exp_fs.py); see Thread<> in pk_core.h
    if (minb == 2) return pk_step_kernel<M, MODE, 128, 2, 128, 1>;    // FS variant (2 CTAs/SM)
    if (minb == 4) return pk_step_kernel<M, MODE, 128, 3, 128, 2>;    // rolled assembly, register factors
    if (minb == 5) return pk_step_kernel<M, MODE, 128, 3, 128, 4>;    // residual-only pass unrolled as well
    if (minb == 6) return pk_step_kernel<M, MODE, 128, 2, 128, 0>;    // default code, 2 CTAs/SM (up to 255 registers)
    if (minb == 7) return pk_step_kernel<M, MODE, 128, 4, 128, 0>;    // default code, 4 CTAs/SM (128 registers)
#else
    if (minb != 3) return nullptr;
#endif
    return pk_step_kernel<M, MODE, 128, 3, 128, 0, FEAT>;
  }
  if (block <= 128) return pk_step_kernel<M, MODE, 128, 2, 0, 0, FEAT_ALL>;
  if (block <= 256) return (M <= 5) ? pk_step_kernel<M, MODE, 256, 2, 256, 0, FEAT> : pk_step_kernel<M, MODE, 256, 1, 256, 0, FEAT>;
  return pk_step_kernel<M, MODE, 512, 1, 512, 0, FEAT_ALL>;
}
template <int M, int FEAT> static kern_t pick_mode(int mode, int block, int T, int minb) {
  switch (mode) {
    case NOISE_SCALAR: return pick_tb<M, NOISE_SCALAR, FEAT>(block, T, minb);
    case NOISE_X: return pick_tb<M, NOISE_X, FEAT>(block, T, minb);
    case TEST_SCALAR: return pick_tb<M, TEST_SCALAR, FEAT_ALL>(block, T, minb);
    case TEST_X: return pick_tb<M, TEST_X, FEAT_ALL>(block, T, minb);
    default: return pick_tb<M, NOISE_FULL, FEAT>(block, T, minb);
  }
}

#endif

#define PK_CAT2(a, b) a##b
#define PK_CAT(a, b) PK_CAT2(a, b)
#if PK_M <= 8
kern_t PK_CAT(pk_pick_kernel_m, PK_M)(int mode, int block, int T, int minb, int feat) {
  return feat ? pick_mode<PK_M, FEAT_ALL>(mode, block, T, minb) : pick_mode<PK_M, 0>(mode, block, T, minb);
}
#endif
ckern_t PK_CAT(pk_pick_cluster_kernel_m, PK_M)(int mode, int Tc, int minb) { return pick_cmode<PK_M>(mode, Tc, minb); }
