// pk_dispatch.h -- kernel pickers exported by the per-M translation units (pk_kernels_m.cu).
#pragma once
#include "pk_params.h"
struct StepArgs;
typedef void (*kern_t)(const pk::Params, const StepArgs);
typedef void (*ckern_t)(const pk::Params, const StepArgs, int);
kern_t pk_pick_kernel_m5(int mode, int block, int T, int minb, int feat);
kern_t pk_pick_kernel_m6(int mode, int block, int T, int minb, int feat);
kern_t pk_pick_kernel_m7(int mode, int block, int T, int minb, int feat);
kern_t pk_pick_kernel_m8(int mode, int block, int T, int minb, int feat);
ckern_t pk_pick_cluster_kernel_m5(int mode, int Tc, int minb);
ckern_t pk_pick_cluster_kernel_m6(int mode, int Tc, int minb);
ckern_t pk_pick_cluster_kernel_m7(int mode, int Tc, int minb);
ckern_t pk_pick_cluster_kernel_m8(int mode, int Tc, int minb);
ckern_t pk_pick_cluster_kernel_m9(int mode, int Tc, int minb);
ckern_t pk_pick_cluster_kernel_m10(int mode, int Tc, int minb);
