// pk_ctx.h -- the context object behind the C ABI (internal; shared by the translation units).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <cstdio>
#include <cstdarg>
#include "pk_params.h"
#include "../../include/peakon_b200.h"

struct pk_ctx {
  pk_config cfg;
  pk::Params p;
  const double *xi, *xi_x, *xi_xx;
  long long step;
  size_t smem;
  int block;
  int cluster;         // CTAs per realisation (1 = single-CTA kernel)
  int lgTc;
  int cminb;           // resident CTAs/SM the cluster kernel is compiled for (2: two clusters share their SMs)
  int fs;              // 1: FS variant of the single-CTA step kernel (factors in shared memory, rolled row loops)
  int slots;           // CTAs of the single-CTA step kernel resident at once (persistent grid), 0 = other kernel
  int sms;             // multiprocessors of the device
  int* d_queue;        // [PK_HOST_CHUNKS][1 + R] work counter + per-realisation finished-block counters (one set per concurrent launch)
  cudaStream_t hs[4];  // pk_step_host: streams of the chunk pipeline (H2D -> kernel -> D2H per chunk)
  cudaEvent_t hev;
  int last_nb, last_nblk, last_grid;   // work split of the last pk_step launch (pk_get_work_split)
  std::string err;
  // internal device buffers for the *_host entry points
  bool w1_resident;    // pk_step_host keeps (u1, dF, dG) in d_state between calls (u1 = dF = dG = NULL mode)
  double* d_state; double* d_dWin; double* d_dWout; int* d_cnt; size_t dW_cap;
  double* d_scratch;   // diagnostics scratch [rows,2,ne] (rows = R, or the rows of a batched dump)
  size_t scratch_cap;  // its size in doubles
  double* d_mom;       // moments partial sums [32, 2*ne]
  long long* d_prof;   // PK_PROFILE: [R,16] cycle buckets
  double* d_tab[3];    // chunk-interleaved copies of the noise tables [nxi, T*M]
  cudaEvent_t ev0, ev1;            // the event pair around the last step kernel (one of the ring below)
  cudaEvent_t evr[2][64];          // ring of event pairs: pk_step_kernel_ms_since() sums them without per-step syncs
  long long evn, evq;              // pairs recorded / pairs already summed
  bool timed;
  long long nlaunch;               // kernels of this library launched through the context (pk_launch_count)
  int* h_stage;                    // pinned [3 R] staging of the counters for pk_step_host
};

extern std::string g_create_err;

int pk_fail(pk_ctx* c, const char* fmt, ...);
#define fail pk_fail
int pk_ensure_scratch(pk_ctx* c, size_t doubles);     // (re)allocates d_scratch; 0 on success
#define CK(c, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(c, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

