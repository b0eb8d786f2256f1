/*
 * peakon_b200.h -- C ABI of the B200 ensemble integrator for peakondrake's stochastic
 * Camassa-Holm timestep.
 *
 * The reference (tommbendall/peakondrake) has no FFI layer: its hot path is reached through the
 * per-step operator protocol of its Python classes.  Each entry point
 * below names the reference interface it replaces (file:line into /root/reference/peakondrake/).
 * The Python binding a maintainer would add is shown in INTEGRATION.md and implemented in
 * peakondrake_b200/_lib.py.
 *
 * Conventions
 *   - plain C, no exceptions across the ABI: every function returns 0 on success, <0 on error;
 *     pk_last_error() gives the message.
 *   - all arrays are CALLER-OWNED.  Unless a function says "host", pointers are DEVICE pointers
 *     on the context's device; fp64, realisation-major: u[r*ne + i] is CG1 dof i of realisation r
 *     in the natural mesh order (node i at x = i*Ld/ne), i.e. u[r] equals the reference's
 *     `u.dat.data` for that seed.
 *   - work is enqueued on the caller's stream (a cudaStream_t passed as void*; NULL = default
 *     stream); no hidden synchronisation except in the *_host functions.
 *   - a context is not thread-safe; use one per host thread / device.
 */
#ifndef PEAKON_B200_H
#define PEAKON_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pk_ctx pk_ctx;

/* noise_mode */
#define PK_NOISE_SCALAR 0 /* num_Xis == 0, or Xi_family 'constant': dXi is a number, dXi_x = dXi_xx = 0 */
#define PK_NOISE_X 1      /* scheme 'no_gradient': dXi varies in x, dXi_x = dXi_xx = 0 (stochastic_functions.py:121,189,232) */
#define PK_NOISE_FULL 2   /* scheme 'hydrodynamic': dXi, dXi_x, dXi_xx from the tables */

/* diagnostics computed by pk_diagnostics(): column index in out[r*PK_NDIAG + k] */
enum {
  PK_D_ENERGY = 0,      /* outputting.py:110 */
  PK_D_L2U2,            /* ||u||_L2^2        (l2_u = sqrt, outputting.py:114) */
  PK_D_H1SEMI2,         /* |u|_H1^2 = u^T K u (h1_u = sqrt(L2^2 + this), :116) */
  PK_D_MASS_U,          /* :124 */
  PK_D_MAX_DU,          /* :140 */
  PK_D_MIN_DU,          /* :142 */
  PK_D_MAX_JUMP_LOCAL,  /* :128 */
  PK_D_MAX_U,           /* np.max(u), :158-160 (p_pde before the periodic factor) */
  PK_D_L2M2,            /* ||m||_L2^2, m from M m = H u (diagnostic_equations.py:236-243) */
  PK_D_H1SEMI_M2,       /* m^T K m */
  PK_D_MASS_M,          /* :120 */
  PK_D_M_MAX,           /* :164 */
  PK_D_UMU,             /* u^T M m   (E_2 = 1/2 of it, :174) */
  PK_D_ERR_SDE2,        /* ||u - u_sde||_L2^2 (:181), only if usde != NULL */
  PK_NDIAG
};
/* integer outputs of pk_diagnostics(): idx[r*PK_NIDX + k]; ties -> lowest index (NumPy argmax/argmin) */
enum {
  PK_I_ARGMAX_DU = 0,   /* find_max(du)  outputting.py:254-258 */
  PK_I_ARGMIN_DU,       /* find_min(du)  :248-252 */
  PK_I_ARGMAX_U,        /* np.argmax(u)  :162, :388 */
  PK_I_ARGMAX_JUMP,
  PK_NIDX
};

typedef struct pk_config {
  int32_t ne;            /* resolution: cells = CG1 dofs (experiment.py:171) */
  int32_t R;             /* realisations held by this context (this GPU's shard) */
  double Ld;             /* domain length */
  double dt;             /* time step */
  double alphasq;        /* alpha^2 > 0 (Camassa-Holm set-up, equations.py:26) */
  int32_t noise_mode;    /* PK_NOISE_* */
  int32_t num_xis;       /* number of noise modes (0 allowed) */
  int32_t nxi_updates;   /* nXi_updates, >= 1 (stochastic_functions.py:213-217); used by the Philox path */
  int32_t device;        /* CUDA device ordinal */
  uint64_t seed;         /* Philox key */
  uint32_t rid_offset;   /* global id of realisation 0 of this shard (Philox counter) */
  int32_t max_it;        /* SNES max_it   (0 -> PETSc default 50) */
  double rtol, atol, stol, divtol; /* SNES tolerances (0 -> PETSc defaults 1e-8, 1e-50, 1e-8, 1e4) */
  double xi_const;       /* PK_NOISE_SCALAR: the constant value of xi_0 (sigma for 'constant') */
  int32_t scheme;        /* 0: 'hydrodynamic' / 'no_gradient' with 'midpoint' (equations.py:63-130);
                            1: 'upwind' with 'ssprk3' (equations.py:28-61) -> pk_step_upwind;
                            2: 'test' with 'midpoint' (equations.py:132-145): one CG1 unknown, noise_mode SCALAR
                               or X, stepped by pk_step (dF/dG are ignored).
                            Every scheme runs one realisation per CTA while the mesh fits (ne <= 4096) and per thread-block
                            cluster beyond (ne <= 32768). */
} pk_config;

/* Replaces Equations.__init__ (equations.py:63-95) + PrognosticVariables (simulation.py:227-249)
 * for R realisations: validates the set-up, chooses the thread layout, allocates scratch. */
int pk_create(const pk_config* cfg, pk_ctx** out);
void pk_destroy(pk_ctx* ctx);
const char* pk_last_error(const pk_ctx* ctx);  /* ctx may be NULL: last create error */

/* Replaces StochasticFunctions.__init__'s tables (stochastic_functions.py:66-201):
 * xi, xi_x, xi_xx are [num_xis, ne] device arrays shared by all realisations (xi_x/xi_xx may be
 * NULL unless noise_mode == PK_NOISE_FULL).  The tables are copied (into the kernel's chunk-interleaved
 * order); the caller may free them afterwards.  Synchronises the device. */
int pk_set_tables(pk_ctx* ctx, const double* xi, const double* xi_x, const double* xi_xx);

/* Replaces `nsteps` passes of  Xis.update(t); equations.solve()  (simulation.py:96-108;
 * stochastic_functions.py:204-236; equations.py:193-199) for every realisation.
 *   u      [R, ne]  in/out  prognostic u
 *   u1,dF,dG [R, ne] in/out the Newton unknown w1 = (u1, dF, dG) the reference carries between
 *                   steps as warm start (equations.py:71,89); zero before the first step.
 *                   dG may be NULL unless noise_mode == PK_NOISE_FULL (it is identically zero).
 *   dW_in  [nsteps, R, num_xis] Brownian increments already scaled by sqrt(dt), or NULL to draw
 *                   them in-kernel (Philox4x32-10 + Box-Muller, counter = (global realisation id,
 *                   global step, mode, sub-update))
 *   dW_out [nsteps, R, num_xis] or NULL: the increments actually used (for path-wise comparison)
 *   fail_step [R] in/out  -1 = healthy; otherwise the global step index at which SNES diverged.
 *                   A failed realisation is frozen (simulation.py:99-106) but its increments are
 *                   still drawn (simulation.py:96).
 *   its_last [R] out, its_total [R] in/out  Newton iteration counts.
 * The context's global step counter advances by nsteps. */
int pk_step(pk_ctx* ctx, int32_t nsteps, double* u, double* u1, double* dF, double* dG,
            const double* dW_in, double* dW_out, int32_t* fail_step, int32_t* its_last,
            int32_t* its_total, void* stream);

/* pk_step with the dump diagnostics FUSED into the step kernel: after every dump_every-th step of the launch the
 * kernel evaluates, on the state it holds in shared memory, what DiagnosticEquations.solve() + dump_diagnostics
 * (simulation.py:118-125) compute on a dump step, and writes row d = (n+1)/dump_every - 1 (n = launch-local step):
 *   diag_out [nsteps/dump_every, R, PK_NDIAG], diag_idx [nsteps/dump_every, R, PK_NIDX]: the columns PK_D_ENERGY ..
 *            PK_D_MAX_U and the four indices of pk_diagnostics (bit-identical extrema / indices; sums differ in
 *            summation order only); the m- and SDE-based columns are NaN -- use pk_diagnostics for those;
 *   suite_vals [.., R, 2], suite_ids [.., R, 3] or both NULL: the output of pk_peakon_suite for the same state.
 * One launch serves all dumps of the chunk: a run that dumps every step (BASELINE config 4) needs no second
 * kernel and never re-reads u.  The step counter must be a multiple of dump_every on entry. */
int pk_step_dump(pk_ctx* ctx, int32_t nsteps, double* u, double* u1, double* dF, double* dG,
                 const double* dW_in, double* dW_out, int32_t* fail_step, int32_t* its_last,
                 int32_t* its_total, int32_t dump_every, double* diag_out, int32_t* diag_idx,
                 double* suite_vals, int32_t* suite_ids, void* stream);

/* pk_step with state SNAPSHOTS for the dumps: after every snap_every-th step of the launch the kernel stores u (natural
 * node order) into snap_u[d] and the realisation's failure state into snap_flag[d] (-1 healthy, else the failing step),
 * d = (n+1)/snap_every - 1.  The dump diagnostics of ALL dumps of the launch are then one batched, HBM-streaming call of
 * pk_diagnostics_rows / pk_peakon_suite_rows over the nsnap*R rows -- cheaper than evaluating reductions inside the
 * (latency-bound) step kernel: a dump costs the step kernel ~1 % instead of ~11 %, for
 * nsnap*R*ne*8 bytes of device memory.   snap_u [nsteps/snap_every, R, ne], snap_flag [nsteps/snap_every, R]. */
int pk_step_snap(pk_ctx* ctx, int32_t nsteps, double* u, double* u1, double* dF, double* dG,
                 const double* dW_in, double* dW_out, int32_t* fail_step, int32_t* its_last,
                 int32_t* its_total, int32_t snap_every, double* snap_u, int32_t* snap_flag, void* stream);

/* pk_diagnostics / pk_peakon_suite over an arbitrary number of rows (e.g. the snapshots of pk_step_snap): u [nrows, ne],
 * pq [nrows, 2] or NULL, out [nrows, PK_NDIAG], idx [nrows, PK_NIDX], row_flag [nrows] or NULL (>= 0: NaN row). */
int pk_diagnostics_rows(pk_ctx* ctx, const double* u, int32_t nrows, const double* pq, int32_t periodic, int32_t want_m,
                        double* out, int32_t* idx, double* du_out, const int32_t* row_flag, void* stream);
int pk_peakon_suite_rows(pk_ctx* ctx, const double* u, int32_t nrows, const int32_t* idx, const int32_t* row_flag,
                         double* vals, int32_t* ids, void* stream);

/* Same, with HOST buffers: copies state (and dW_in) to internal device buffers, steps, copies the
 * state, dW_out and the counters back, and synchronises (the realisations go through a 2-chunk
 * H2D / kernel / D2H pipeline on internal streams; pin the host buffers).  This is the call a host
 * application without device arrays makes; bench.py's "e2e" times it.
 * u1 = dF = dG = NULL selects the reference's own ownership: the Newton unknown w1 = (u1, dF, dG) is
 * internal state of the solver object there (equations.py:71,89: created as zero, never reset, never seen
 * by simulation()), so it stays resident in the context and only u and the counters cross the bus.
 * pk_reset_warm_start() zeroes it again (a new initial condition). */
int pk_step_host(pk_ctx* ctx, int32_t nsteps, double* u, double* u1, double* dF, double* dG,
                 const double* dW_in, double* dW_out, int32_t* fail_step, int32_t* its_last,
                 int32_t* its_total);

int pk_reset_warm_start(pk_ctx* ctx);

/* Replaces `nsteps` passes of  Xis.update(t); equations.solve()  for scheme 'upwind', timestepping
 * 'ssprk3' (equations.py:172-185): three SSPRK3 stages of the DG1 upwind transport of m with the
 * frozen velocity us = Dt*u + dXi, then the Helmholtz solve (M + alphasq K) u = P m.
 *   m [R, ne, 2] in/out: DG1 m, (left-node value, right-node value) of every cell
 *   u [R, ne]    in/out: CG1 u (the reference's vector-valued u, scalar on the interval)
 *   dW_in / dW_out / fail_step: as for pk_step.  The scheme is explicit: a realisation fails only on
 *   a non-finite increment (before the state is touched) or when its state has blown up. */
int pk_step_upwind(pk_ctx* ctx, int32_t nsteps, double* m, double* u, const double* dW_in,
                   double* dW_out, int32_t* fail_step, void* stream);

/* Diagnostics of the prognostic DG1 field m of the upwind scheme (outputting.py:112,118,120,164,174):
 * fills the columns PK_D_L2M2, PK_D_H1SEMI_M2, PK_D_MASS_M, PK_D_M_MAX, PK_D_UMU of `out`
 * ([R, PK_NDIAG], the table pk_diagnostics writes; call it afterwards on the same stream). */
int pk_diagnostics_dg1(pk_ctx* ctx, const double* m, const double* u, const int32_t* fail_step,
                       double* out, void* stream);

int pk_set_step_counter(pk_ctx* ctx, int64_t step);
int64_t pk_get_step_counter(const pk_ctx* ctx);

/* Replaces DiagnosticEquations.solve() for du, jump_du, m, u_sde (diagnostic_equations.py:94-113,
 * 236-243,254-279,373-378) and the reductions of Outputting.dump_diagnostics (outputting.py:106-196)
 * + find_min/find_max (:248-258), for every realisation.
 *   u [R, ne] in
 *   pq [R, 2] in or NULL: (p, q) of the "true" peakon for this dump (what update_constants reads from
 *       the true-peakon file, diagnostic_equations.py:377-378); enables PK_D_ERR_SDE2
 *   periodic: the `periodic` option (selects the u_sde formula)
 *   want_m: compute the m-based columns (costs two more passes); otherwise they are NaN
 *   out [R, PK_NDIAG] doubles, idx [R, PK_NIDX] int32
 *   du_out [R, ne] or NULL: the DG0 field du (cell i = [node i, node i+1]) for host-side find_mus
 *   fail_step [R] or NULL: rows of failed realisations are filled with NaN (outputting.py:107-108) */
int pk_diagnostics(pk_ctx* ctx, const double* u, const double* pq, int32_t periodic, int32_t want_m,
                   double* out, int32_t* idx, double* du_out, const int32_t* fail_step, void* stream);

/* Replaces the two sequential scans of peakon_diagnostics (outputting.py:401-435) for every
 * realisation: idx = output of pk_diagnostics for the same u.
 *   vals [R,2] = (peakon_min_du, peakon_max_du), NaN when the scan finds nothing within ne/2 cells
 *   ids  [R,3] = (argmax u, min_du_idx, max_du_idx): array positions as the reference's (possibly
 *               negative, Python-wrapped) indices resolve; -1 = not found.  The host turns them into
 *               peakon_loc / *_loc / mu / nu with its own coordinate array (bit-exact look-ups). */
int pk_peakon_suite(pk_ctx* ctx, const double* u, const int32_t* idx, const int32_t* fail_step,
                    double* vals, int32_t* ids, void* stream);

/* Replaces find_mus (outputting.py:260-370), the `mu` diagnostic, for every realisation: peaks = + -> - sign changes
 * of the DG0 slope whose CG1 value exceeds twice the mean of u; per peak zone the distance between the maximum and
 * the minimum slope (mu) and between the peak and the minimum slope (alt_mu), four highest peaks only; the quirks of
 * the reference are kept as written (bit-equal to the reference helper on its golden vectors).
 *   x_cells [ne]  DG0 coordinate array (cell centres), device memory; coordinates are looked up, never recomputed
 *   out [R,8]     mu_0..3, alt_mu_0..3 (NaN where the reference's lists are shorter or the realisation has failed)
 *   status [R]    number of peaks (>= 0); -1: more peaks than the device routine holds, -2: the reference itself
 *                 would raise for this field (run the host routine for that realisation), -3: failed realisation */
int pk_mu(pk_ctx* ctx, const double* u, const double* x_cells, const int32_t* fail_step, double* out,
          int32_t* status, void* stream);

/* Ensemble statistics for one dump (the quantities the reference accumulates across seeds off-line:
 * running mean of u, experiment.py:224-227; histograms plotted by plotting/plot_stochastic_histograms.py:61-67).
 *   diag/idx: outputs of pk_diagnostics for the same state
 *   packed [2*ne + 1 + 2*nbins] doubles out:
 *     sum_r u | sum_r u^2 | number of healthy realisations | histogram of q_pde on [q_lo,q_hi) |
 *     histogram of min_du on [d_lo,d_hi)      (integer counts stored as doubles; out-of-range -> end bins)
 * Sums are order-fixed (deterministic); a multi-GPU caller all-reduces `packed` once (ncclSum). */
int pk_ensemble_stats(pk_ctx* ctx, const double* u, const int32_t* fail_step, const double* diag,
                      const int32_t* idx, int32_t nbins, double q_lo, double q_hi, double d_lo,
                      double d_hi, double* packed, void* stream);

/* Per-diagnostic ensemble statistics of one dump (sum / sum of squares per scalar diagnostic, extrema,
 * fail count), over the healthy realisations of this shard, from the table pk_diagnostics wrote:
 *   sums [2*PK_NDIAG + 2] = sum_r v_k | sum_r v_k^2 | number of healthy realisations | number of failed ones
 *   ext  [2*PK_NDIAG]     = min_r v_k | -max_r v_k
 * A multi-GPU caller all-reduces `sums` with ncclSum and `ext` with ncclMin (the negated maxima make one
 * collective serve both extrema).  Fixed summation order; NaN columns (not requested) are skipped. */
int pk_ensemble_scalars(pk_ctx* ctx, const double* diag, const int32_t* fail_step, double* sums, double* ext,
                        void* stream);

/* Replaces `nsteps` calls of PeakonEquations.update() (peakon_equations.py:34-65) for every realisation: the
 * (p, q) SDE of a peakon driven by the SAME Brownian increments as the PDE.
 *   p, q [R]      in/out (peakon_equations.py:28-32 gives the initial values)
 *   dW_in [nsteps, R, num_xis] increments (e.g. the dW_out of pk_step), or NULL: drawn in-kernel with the Philox
 *                 counters pk_step uses for the global steps step0 .. step0+nsteps-1 (identical numbers)
 *   xi, xi_x, xi_xx [num_xis, ne]  the CG1 fields pure_xi, pure_xi_x, pure_xi_xx (stochastic_functions.py:175-182),
 *                 natural node order; evaluated at q by CG1 point evaluation (`.at`, peakon_equations.py:42-44)
 *   method        0 = 'ito_euler' (:39-47), 1 = 'milstein' (:49-57)
 *   E             exp(-Ld/alpha), computed by the caller (so that host and device agree on it to the bit)
 *   traj [nsteps/stride, R, 2] or NULL: (p, q) after every `stride` steps (the values dump_diagnostics would write,
 *                 outputting.py:216-218)
 * The arithmetic is the reference's, operation by operation, with individually rounded operations. */
int pk_peakon_sde_step(pk_ctx* ctx, int32_t nsteps, double* p, double* q, const double* dW_in, int64_t step0,
                       const double* xi, const double* xi_x, const double* xi_xx, int32_t method, double E,
                       double* traj, int32_t stride, void* stream);

/* Replaces Projector(du, du_smooth) (diagnostic_equations.py:115-119) + find_max / find_min on it
 * (outputting.py:135-138): L2 projection of the DG0 slope onto periodic CG5 by static condensation (constant
 * cyclic tridiagonal system for the vertex values), extrema with lowest-dof-index tie-break.
 *   consts [16] host: the condensed reference-element constants (peakondrake_b200/outputting.py: cg5_condensed)
 *   out [R,4] = (max, x of max, min, x of min), idx [R,2] = CG5 dof index (5*cell + local node) of max / min
 *   field [R, 5*ne] or NULL: du_smooth itself */
int pk_du_smooth_loc(pk_ctx* ctx, const double* u, const int32_t* fail_step, const double* consts, double* out,
                     int32_t* idx, double* field, void* stream);

/* number of kernels of this library launched through the context so far (bench.py's gpu_launches) */
int64_t pk_launch_count(const pk_ctx* ctx);

/* layout query (for tests / documentation): threads per realisation, rows per thread, remainder */
int pk_get_layout(const pk_ctx* ctx, int32_t* T, int32_t* M, int32_t* rem, int32_t* smem_bytes);

/* work split of the last pk_step launch of the single-CTA step kernel (persistent CTAs that pull items of
 * `nblk` consecutive steps of one realisation from a queue): resident CTA slots, step blocks per realisation,
 * steps per block, CTAs launched.  slots = 0 when the cluster / upwind kernels serve this context. */
int pk_get_work_split(const pk_ctx* ctx, int32_t* slots, int32_t* nb, int32_t* nblk, int32_t* grid);

/* measured duration (ms) of the kernels launched by the last pk_step(_host) call, from CUDA events
 * recorded on the launching stream; synchronises on the second event. */
double pk_last_step_kernel_ms(pk_ctx* ctx);
/* sum of the device durations (ms) of the step kernels launched by pk_step since the previous call of this function
 * (at most the last 64 launches; *count = how many), from CUDA event pairs recorded around each kernel on the
 * launching stream; synchronises once, on the last one -- lets a caller time a loop without a sync per step. */
double pk_step_kernel_ms_since(pk_ctx* ctx, int32_t* count);

#ifdef __cplusplus
}
#endif
#endif /* PEAKON_B200_H */
