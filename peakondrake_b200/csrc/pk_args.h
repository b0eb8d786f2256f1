// pk_args.h -- argument block of the step kernels.
#pragma once
struct StepArgs {
  double *u, *u1, *dF, *dG;
  const double *xi, *xi_x, *xi_xx;
  const double* dW_in;
  double* dW_out;
  int *fail_step, *its_last, *its_total;
  int R, step0, nsteps;   // R: realisations of the context (stride of dW_in / dW_out and of the state arrays)
  int r0, Rq;             // this launch advances realisations [r0, r0 + Rq)
  long long* prof;    // [R,16] cycle buckets, PK_PROFILE builds only
  unsigned rid0;
  unsigned long long seed;
  // single-CTA step kernel (persistent CTAs + work queue)
  int* queue;         // [1 + R]: work counter, then per realisation the number of finished step blocks
  int nblk;           // steps per work item
  int part;           // cluster kernel: 1 = Schur solve partitioned by CTA, separators solved redundantly in every CTA (default);
                      // 2 = same, separators solved by the leading threads (a cluster barrier per level); 0 = flat PCR
  int nb;             // step blocks per realisation: items = nb * R, item i = (block i / R, realisation i % R)
  // fused dumps (pk_step_dump): see StepIO in pk_core.h
  int diag_every;
  double* diag_out; int* diag_idx; double* suite_vals; int* suite_ids;
  // state snapshots (pk_step_snap)
  int snap_every; double* snap_u; int* snap_flag;
};
