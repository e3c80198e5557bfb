/* aligator_b200/gar.h -- C ABI of the B200-native batched Riccati sweep.
 *
 * Drop-in boundary for ONE path of Simple-Robotics/aligator: the linear-quadratic
 * subproblem solve behind gar::RiccatiSolverBase<double>
 * (include/aligator/gar/riccati-base.hpp:13-37), i.e. what
 * gar::ProximalRiccatiSolver (gar/proximal-riccati.hpp:12-47) does for
 * SolverProxDDPTpl::innerLoop (solvers/proxddp/solver-proxddp.hxx:605-632) and
 * for bench/gar-riccati.cpp:42-50 -- for a BATCH of independent problem
 * instances of identical dimensions, on one B200.
 *
 * Plain C: opaque handle, pointers and sizes, int status codes, no exceptions.
 * All matrices are fp64.  "column-major" / "row-major" are the reference's own
 * storage orders (Eigen default column-major; fb / Z are RowMatrixXs,
 * math.hpp:23-27, riccati-kernel.hpp:96-98).
 *
 * Data layout (identical on host and device; `batch` and knot indices lead):
 *
 *   stage knots  [batch][N][stage_record]   one record = the 11 buffers of
 *       LqrKnotTpl (gar/lqr-problem.hpp:60-65) concatenated, each in the
 *       reference's own column-major storage, nx2 = nx, nth = 0:
 *           [ A (nx*nx) | B (nx*nu) | f (nx) | Q (nx*nx) | S (nx*nu) | R (nu*nu)
 *             | q (nx) | r (nu) | C (nc*nx) | D (nc*nu) | d (nc) | pad to even ]
 *   terminal knot [batch][term_record] = [ Q | q | C (nct*nx) | d (nct) ]   (nu = 0)
 *   G0 [batch][nc0*nx] column-major, g0 [batch][nc0]   (lqr-problem.hpp:126-127)
 *
 *   outputs (StageFactor members, riccati-kernel.hpp:86-101):
 *   FF   [batch][N][nu+nc+nx]          ff  = [k; z; a]
 *   FB   [batch][N][(nu+nc+nx)*nx]     fb  = [K; Z; Ahat], ROW-major
 *   VXX  [batch][N+1][nx*nx]           vm.Vxx column-major; symmetric for t>=1,
 *                                      as computed for t=0
 *   VX   [batch][N+1][nx]              vm.vx
 *   FFT  [batch][nct], FBT [batch][nct*nx]   terminal knot's z, Z (row-major)
 *   KKT0 [batch][nx+nc0]               kkt0.ff = [x0; lbda0]
 *   XS [batch][N+1][nx]  US [batch][N][nu]  VS [batch][N][nc]  VST [batch][nct]
 *   LBD0 [batch][nc0]    LBDAS [batch][N][nx]  (lbdas[1..N])
 *   STATUS [batch] int   0 = ok, bit0 = a stage LDL^T failed (the reference throws
 *                        "Failed stage LDL factorization", riccati-kernel.hxx:239-241),
 *                        bit1 = the initial-stage factorisation failed,
 *                        bit2 = (parallel solver) a block of the condensed system failed to factor.
 */
#ifndef ALIGATOR_B200_GAR_H
#define ALIGATOR_B200_GAR_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ab2_gar_solver ab2_gar_solver;

enum {
  AB2_OK = 0,
  AB2_ERR_INVALID = 1,     /* bad argument */
  AB2_ERR_UNSUPPORTED = 2, /* dims not instantiated in this build */
  AB2_ERR_CUDA = 3,        /* CUDA runtime error; see ab2_gar_last_error() */
  AB2_ERR_STATE = 4        /* call order (e.g. forward before backward) */
};

enum { AB2_HOST = 0, AB2_DEVICE = 1 };

/* output selectors for ab2_gar_get / ab2_gar_output_doubles / ab2_gar_device_ptr */
enum {
  AB2_OUT_FF = 0,
  AB2_OUT_FB = 1,
  AB2_OUT_VXX = 2,
  AB2_OUT_VX = 3,
  AB2_OUT_FFT = 4,
  AB2_OUT_FBT = 5,
  AB2_OUT_KKT0 = 6,
  AB2_OUT_XS = 7,
  AB2_OUT_US = 8,
  AB2_OUT_VS = 9,
  AB2_OUT_VST = 10,
  AB2_OUT_LBD0 = 11,
  AB2_OUT_LBDAS = 12,
  /* parametric problems (nth > 0), riccati-kernel.hpp:86-101, proximal-riccati.hpp:40-43 */
  AB2_OUT_FTH = 13,     /* [batch][N][(nu+nc+nx)*nth]  row-major [Kth; Zth; Yth]      (StageFactor::fth) */
  AB2_OUT_VXT = 14,     /* [batch][N+1][nx*nth]        column-major                   (vm.Vxt) */
  AB2_OUT_VTT = 15,     /* [batch][N+1][nth*nth]                                      (vm.Vtt) */
  AB2_OUT_VT = 16,      /* [batch][N+1][nth]                                          (vm.vt)  */
  AB2_OUT_KKT0FTH = 17, /* [batch][(nx+nc0)*nth]       row-major                      (kkt0.fth) */
  AB2_OUT_THGRAD = 18,  /* [batch][nth]                                               (thGrad) */
  AB2_OUT_THHESS = 19,  /* [batch][nth*nth]            column-major                   (thHess) */
  AB2_OUT_COUNT = 20
};

typedef struct ab2_gar_dims {
  int nx;      /* state (tangent) dimension of every knot; nx2 = nx */
  int nu;      /* control dimension of the N stage knots (>= 1)     */
  int nc;      /* constraint rows of the stage knots                */
  int nct;     /* constraint rows of the terminal knot (nu = 0)     */
  int nc0;     /* rows of the initial condition G0 x0 + g0 = 0      */
  int horizon; /* N: the problem has N stage knots + 1 terminal     */
  int batch;   /* number of independent problem instances           */
  int device;  /* CUDA device ordinal                               */
} ab2_gar_dims;

/* launch tuning.  variant: -1 = automatic (the FP64 tensor-core formulation where the
 * shape allows it, else the lane-per-column one; the CTA-per-instance kernel for shapes
 * without a compile-time instantiation); 0..8 and 10 select a specific warp-per-instance build,
 * see csrc/riccati_launch.cuh; 9 forces the CTA-per-instance kernel (csrc/riccati_block.cuh). */
typedef struct ab2_gar_tuning {
  int variant;
  int stagger_ns;  /* > 0: start-up delay per resident warp slot (de-phases the warps of an SM) */
  int ctas_per_sm; /* > 0: cap on resident CTAs per SM (e.g. 7 -> 4096 instances = two full rounds) */
} ab2_gar_tuning;

/* Doubles in one stage / terminal record (stage includes the pad to even).
 * Replaces: the 11 ArenaMatrix members of LqrKnotTpl, gar/lqr-problem.hpp:60-65. */
size_t ab2_gar_stage_record_doubles(int nx, int nu, int nc);
size_t ab2_gar_term_record_doubles(int nx, int nct);
/* 1 if (nx,nu,nc,nc0) is served by a compile-time kernel instantiation of this build (one
 * warp or part of one per instance), 2 if by the run-time-dimension kernel (one CTA per
 * instance: any shape whose buffers fit 227 KB of shared memory and whose row counts
 * nx+1, nu+nc, nx+nc0, nu+nc+nx are <= 256), 0 if not served. */
int ab2_gar_supported(int nx, int nu, int nc, int nc0);

/* Replaces: ProximalRiccatiSolver(const LqrProblemTpl&), gar/proximal-riccati.hxx:13-31
 * (allocates all factor storage once; the hot calls below never allocate). */
int ab2_gar_create(const ab2_gar_dims *dims, ab2_gar_solver **out);
/* Parametric problems (LqrKnotTpl::Gth, Gx, Gu, Gv, gamma with nth > 0, gar/lqr-problem.hpp:66-71;
 * what ParallelRiccatiSolver's legs solve).  Same as ab2_gar_create with `nth` parameters: stage
 * records grow by [Gx nx*nth | Gu nu*nth | Gv nc*nth | Gth nth*nth | gamma nth], the terminal record
 * by [Gx | Gv nct*nth | Gth | gamma] (ab2_gar_*_record_doubles_th); backward also produces
 * AB2_OUT_FTH..THHESS; runs the CTA-per-instance kernel. */
int ab2_gar_create_parametric(const ab2_gar_dims *dims, int nth, ab2_gar_solver **out);
size_t ab2_gar_stage_record_doubles_th(int nx, int nu, int nc, int nth);
size_t ab2_gar_term_record_doubles_th(int nx, int nct, int nth);
/* forward(xs, us, vs, lbdas, theta) (riccati-base.hpp:21-24 with the optional theta): theta is
 * [batch][nth] in host or device memory, NULL = no parameter (like std::nullopt). */
int ab2_gar_forward_theta(ab2_gar_solver *s, const double *theta, int memspace, void *stream);
int ab2_gar_destroy(ab2_gar_solver *s);
int ab2_gar_set_tuning(ab2_gar_solver *s, const ab2_gar_tuning *t);

/* Replaces: RiccatiSolverDense(const LqrProblemTpl&), gar/dense-riccati.hxx:13-45 -- the reference's second solver
 * (LQSolverChoice::STAGEDENSE): per knot ONE Bunch-Kaufman factorisation of the (nu + nc + 2 nx)^2 matrix
 * [[R, D^T, B^T, 0],[D, -mu I, 0, 0],[B, 0, 0, -I],[0, 0, -I, P']] (gar/dense-kernel.hpp:98-113), one CTA per
 * instance.  Same problem layout and call sequence as ab2_gar_create; FF / FB have nu + nc + 2 nx rows
 * [k; z; l; y] / [K; Z; L; Y] (dense-kernel.hpp:28-31: u = k + K x, v = z + Z x, lbda' = l + L x, x' = y + Y x),
 * VXX / VX hold Pxx / px (not symmetrised).  Not the fast path: an independent algorithm on the device. */
int ab2_gar_create_dense(const ab2_gar_dims *dims, ab2_gar_solver **out);
/* Replaces: ParallelRiccatiSolver(LqrProblemTpl&, num_threads), gar/parallel-solver.hxx:32-82 -- the
 * parallel-in-time variant.  The horizon of EVERY instance is cut into `num_legs` legs
 * [i(N+1)/T, (i+1)(N+1)/T) (get_work, :23-28); backward() runs the legs of all instances as the work
 * items of one launch (each leg = the recursion of riccati-kernel.hxx:105-129 on its span, parametric
 * in the co-state at the next leg's head, nth = nx), then solves the condensed symmetric
 * block-tridiagonal system of every instance (:85-129, 166-203; block-tridiagonal.hpp:82-182) with at
 * most 5 refinement steps to 1e-10; forward() rolls the legs out in one launch (:209-243).
 * Unlike the reference this does NOT mutate the caller's problem (:52-60, :136-147): the records stay
 * the plain [A|B|f|Q|S|R|q|r|C|D|d] ones, the leg parameterisation (Gx = A^T, Gu = B^T, gamma = f on a
 * leg's last knot) is implicit.  Outputs: as ab2_gar_create plus FTH/VXT/VTT/VT with nth = nx (zero on
 * the last leg, which has no parameters).  Status bit2 = a block of the condensed system failed to
 * factor (the reference ignores that, :176-179).  num_legs < 2 is AB2_ERR_INVALID (the reference
 * throws, :42-46); horizon + 1 >= num_legs is required. */
int ab2_gar_create_parallel(const ab2_gar_dims *dims, int num_legs, ab2_gar_solver **out);
/* Replaces: RiccatiSolverBase::collapseFeedback(), riccati-base.hpp:32 (no-op for the serial solver)
 * / ParallelRiccatiSolver::collapseFeedback(), parallel-solver.hpp:41-51: K_0 -= Kth_0 * subdiagonal[1]
 * (restated as written: after the swap at parallel-solver.hxx:180-181 that block is Vxt_0^T). */
int ab2_gar_collapse_feedback(ab2_gar_solver *s, void *stream);

/* Give the solver the problem data.  Replaces the non-owning `problem_` pointer the
 * reference re-reads at every backward() (proximal-riccati.hpp:46; the knots are
 * rewritten in place by updateLQSubproblem, solver-proxddp.hxx:734-805).
 * memspace AB2_HOST: buffers are copied host->device on `stream` (pinned memory makes
 * the copy asynchronous).  AB2_DEVICE: the pointers are kept, non-owning, zero-copy.
 * Any of the four may be NULL to keep the previous one. */
int ab2_gar_set_problem(ab2_gar_solver *s, const double *stage, const double *term,
                        const double *G0, const double *g0, int memspace, void *stream);

/* Replaces: RiccatiSolverBase::backward(mueq), riccati-base.hpp:19
 * (terminal + stage recursion + initial saddle system, proximal-riccati.hxx:34-62). */
int ab2_gar_backward(ab2_gar_solver *s, double mueq, void *stream);
/* Replaces: RiccatiSolverBase::forward(xs,us,vs,lbdas), riccati-base.hpp:21-24
 * (riccati-kernel.hxx:196-207, 315-377; theta unsupported: nth = 0). */
int ab2_gar_forward(ab2_gar_solver *s, void *stream);
/* backward + forward in ONE persistent launch: the loop body of
 * bench/gar-riccati.cpp:46-49 and solver-proxddp.hxx:608-611. */
int ab2_gar_sweep(ab2_gar_solver *s, double mueq, void *stream);
/* Inputs of the batched LQ assembly: the derivative buffers SolverProxDDP::updateLQSubproblem
 * (solvers/proxddp/solver-proxddp.hxx:734-805) and computeProjectedJacobians (:25-69) read.
 * DEVICE pointers; stage arrays are [batch][N][block], terminal / initial arrays [batch][block],
 * blocks column-major like the reference's Eigen matrices.  Constraint sets are given per row
 * by bounds: a row is ACTIVE (kept by applyNormalConeProjectionJacobian, core/constraint-set.hxx:
 * 25-37) iff shifted > hi or shifted < lo -- equality rows: lo = +inf; negative orthant: lo = -inf,
 * hi = 0; box: its limits (computeActiveSet of equality-constraint.hpp:52, negative-orthant.hpp:30,
 * box-constraint.hpp:39). */
typedef struct ab2_lq_inputs {
  const double *Jx, *Ju, *slack;   /* dd.Jx() -> A, dd.Ju() -> B, dyn_slacks[t+1] -> f        (:755-757) */
  const double *Lxx, *Lxu, *Luu;   /* cd.Lxx_, Lxu_, Luu_ -> Q, S, R (+ preg on the diagonals) (:759-768) */
  const double *Lx, *Lu;           /* workspace Lxs[t], Lus[t] -> q, r                          (:764-765) */
  const double *Hxx, *Hxu, *Huu;   /* dd.Hxx_, Hxu_, Huu_ (HessianApprox::EXACT) or NULL        (:770-774) */
  const double *cJx, *cJu;         /* constraint Jacobians before projection [nc x nx], [nc x nu] (:40-41) */
  const double *Lv, *shifted;      /* workspace Lvs[t] -> d; shifted_constraints[t]              (:46,49,780) */
  const double *lo, *hi;           /* [nc] bounds of the stage constraint rows (shared by all knots) */
  const double *Lxx_N, *Lx_N;      /* terminal cost                                               (:787-790) */
  const double *cJx_N, *Lv_N, *shifted_N, *loN, *hiN; /* terminal constraints [nct ...]          (:55-68,791-794) */
  const double *G0, *g0;           /* init_data Jx(), value_                                      (:798-800) */
  const double *Hxx0;              /* init_data Hxx_ (added to stage 0's Q) or NULL               (:803-804) */
  double preg, mu_inv;
} ab2_lq_inputs;
/* updateLQSubproblem + computeProjectedJacobians for every instance and knot in one pass over
 * HBM: writes the solver-owned packed problem (the same bytes ab2_gar_set_problem uploads) and
 * makes it the current problem.  A device-resident caller never moves the knots over PCIe. */
int ab2_gar_assemble(ab2_gar_solver *s, const ab2_lq_inputs *in, void *stream);
/* Device address of the current packed problem: what = 0 stage, 1 term, 2 G0, 3 g0
 * (the bytes workspace_.lqr_problem holds after updateLQSubproblem). */
int ab2_gar_problem_ptr(ab2_gar_solver *s, int what, const double **out);
/* Copy of it (whole array) to `dst` in host or device memory. */
int ab2_gar_get_problem(ab2_gar_solver *s, int what, double *dst, int memspace, void *stream);

/* One whole iteration of the caller's loop with HOST buffers, pipelined over the batch:
 * upload the problem (what updateLQSubproblem rewrote, solver-proxddp.hxx:734-805), sweep,
 * and download `nwhat` result arrays (`whats[i]` -> `dsts[i]`, full-size host arrays laid out
 * like ab2_gar_get's; what solver-proxddp.hxx:610-632 reads back).  The batch is cut into
 * `nchunks` slices (0 = automatic) that travel on internal streams, so the upload of slice
 * i+1, the sweep of slice i and the download of slice i-1 overlap; PCIe is full duplex, so the
 * step costs max(upload, download) instead of their sum.  Host buffers should be pinned
 * (page-locked); pageable memory works but serialises.  Ordered after prior work on `stream`;
 * work enqueued on `stream` afterwards waits for it.  Results also stay on the device. */
int ab2_gar_sweep_host(ab2_gar_solver *s, const double *stage, const double *term, const double *G0,
                       const double *g0, double mueq, int nchunks, const int *whats,
                       double *const *dsts, int nwhat, void *stream);
/* The same with the symmetric blocks of every stage knot sent as LOWER TRIANGLES (what Eigen's
 * triangularView<Lower> of LqrKnotTpl::Q / ::R walks, lqr-problem.hpp:53-57; the Riccati recursion only ever
 * needs those): record [A | B | f | Qlow nx(nx+1)/2 | S | Rlow nu(nu+1)/2 | q | r | C | D | d], column j of a
 * triangle holding rows j..n-1, no padding.  The host path is PCIe-bound, so the 16 % fewer bytes at
 * config 2 are 16 % less time; the full records are rebuilt in HBM by one streaming kernel per slice.
 * Plain serial handles only (nth = 0).  pack_stage_sym is the host-side helper that derives the packed
 * records from full ones (an adapter packs its Eigen matrices straight into this layout instead). */
size_t ab2_gar_stage_record_doubles_sym(int nx, int nu, int nc);
int ab2_gar_pack_stage_sym(int nx, int nu, int nc, const double *stage, double *stage_sym, long nrec);
int ab2_gar_sweep_host_sym(ab2_gar_solver *s, const double *stage_sym, const double *term, const double *G0,
                           const double *g0, double mueq, int nchunks, const int *whats,
                           double *const *dsts, int nwhat, void *stream);

/* Replaces: getFeedforward(i)/getFeedback(i) (riccati-base.hpp:33-34), the public
 * `datas[i].vm` / `kkt0` members (proximal-riccati.hpp:40-43) and the caller-owned
 * xs/us/vs/lbdas vectors.  Copies the whole [batch][...] array `what` to dst. */
size_t ab2_gar_output_doubles(const ab2_gar_solver *s, int what);
int ab2_gar_get(ab2_gar_solver *s, int what, double *dst, int memspace, void *stream);
/* Sub-range copy: knots [t0, t0+nt) of instances [b0, b0+nb) (dense [nb][nt][...]). */
int ab2_gar_get_range(ab2_gar_solver *s, int what, int b0, int nb, int t0, int nt,
                      double *dst, int memspace, void *stream);
/* Device-resident consumers: raw device pointer of an output array. */
/* First-step policy of every instance, packed [batch][nu][nx+1] = [K_0 | k_0] (row-major) into
 * the DEVICE buffer `dst` by one small kernel: what a receding-horizon consumer applies
 * (results_.gains_[0], solver-proxddp.hxx:619-626) and the payload of the one all-gather when
 * the batch is sharded across GPUs. */
int ab2_gar_first_step_policy(ab2_gar_solver *s, double *dst, void *stream);
/* Gains in the layout of the caller's results: for every instance and stage knot a COLUMN-major
 * (nu+nc+nx) x (nx+1) block whose column 0 is the feedforward [k; z; a] and columns 1..nx the
 * feedback [K; Z; Ahat] -- what SolverProxDDP copies into results_.gains_[i] from
 * getFeedforward(i) / getFeedback(i) (solver-proxddp.hxx:619-626, results.hxx:23-38).
 * dst: [batch][N][(nu+nc+nx)*(nx+1)], host or device. */
int ab2_gar_get_gains(ab2_gar_solver *s, double *dst, int memspace, void *stream);
/* lqrComputeKktError (gar/utils.hxx:88-182) of the current problem and the solution of the last
 * forward pass, for every instance: dst[batch][3] = infinity norms of the dynamics (incl. the
 * initial condition), constraint (C x + D u + d - mu v) and stationarity residuals.  Computed on
 * the device (one warp per (instance, knot)); dst in host or device memory. */
int ab2_gar_kkt_error(ab2_gar_solver *s, double mueq, double *dst, int memspace, void *stream);
int ab2_gar_device_ptr(ab2_gar_solver *s, int what, double **out);

/* Multi-GPU (one process per GPU, the batch sharded by instance): the ONE exchange
 * of a sweep -- the all-gather of the first-step policy [K_0 | k_0] -- fused into the sweep over
 * NVLink peer memory instead of a separate NCCL collective.  Every rank owns a receive buffer
 * [world][batch][nu][nx+1] (three slots, alternating by step) that all peers map through CUDA IPC.
 * Once connected, every warp-per-instance backward / sweep launch stores each instance's block straight
 * into slot r of every rank's buffer as soon as that instance's backward pass is done (other kernels:
 * a pack kernel does the same stores); a step flag (release / acquire at system scope) publishes it.
 *   init:      allocate the local buffer; *ipc_handle_out = 64 bytes to hand to every peer
 *   connect:   all_handles = world x 64 bytes, rank order (exchange them with any host transport)
 *   allgather: after backward / sweep, on the same `stream` (all ranks, same batch): publish the step
 *              (pack + store first where the sweep has not done it); holds `stream` until every peer
 *              has consumed what the next step's slot held
 *   wait:      `stream` acknowledges the previous step as consumed, then waits until the blocks of every
 *              rank have arrived for the last allgather
 *   buffer:    device address of the slot holding the last allgather, [world][batch][nu][nx+1] */
int ab2_gar_peer_gather_init(ab2_gar_solver *s, int world, int rank, void *ipc_handle_out);
int ab2_gar_peer_gather_connect(ab2_gar_solver *s, const void *all_handles);
int ab2_gar_policy_allgather(ab2_gar_solver *s, void *stream);
int ab2_gar_policy_allgather_wait(ab2_gar_solver *s, void *stream);
int ab2_gar_peer_gather_buffer(ab2_gar_solver *s, double **out, long *step);
/* Per-instance status words (layout above). */
int ab2_gar_status(ab2_gar_solver *s, int *dst, int memspace, void *stream);
/* Pivot statistics of the last backward pass, one int per instance: bits 0-14 = number of
 * 2x2 pivots, bit 15 = the initial saddle system needed no interchange / 2x2 pivot and ran on the
 * register fast path, high 16 bits = number of symmetric interchanges the Bunch-Kaufman factorisations
 * of the stage KKT matrices and of the initial saddle system took (core/bunchkaufman.hpp:61-83,
 * what Eigen::BunchKaufman reports through pivots() / m_pivot_count, :158-163).  Diagnostics:
 * lets a caller (and the tests) see that the pivoted code paths actually ran. */
int ab2_gar_pivot_stats(ab2_gar_solver *s, int *dst, int memspace, void *stream);

/* The consumers of the step inside the caller's line search, batched over the
 * instances so that SolverProxDDP's inner loop (solver-proxddp.hxx:605-660) can stay on the device between
 * the sweep and the next model evaluation; the step (dxs, dus, dvs, dlams) is the solver's own last forward
 * pass.  All array arguments are DEVICE pointers laid out like the solver's outputs: xs [batch][N+1][nx],
 * us [batch][N][nu], vs [batch][N][nc], vsT [batch][nct], lam0 [batch][nc0], lams [batch][N][nx]. */
typedef struct ab2_ls_iterate {
  const double *xs, *us, *vs, *vsT, *lam0, *lams;
} ab2_ls_iterate;
typedef struct ab2_ls_trial {
  double *xs, *us, *vs, *vsT, *lam0, *lams;
} ab2_ls_trial;
/* Replaces: the vector part of SolverProxDDP::tryLinearStep, solver-proxddp.hxx:111-155: trial = results +
 * alpha * step for lams, vs (math::vectorMultiplyAdd, :121-124) and for xs, us with the vector-space
 * integrate (:139-150); a manifold's integrate and problem.evaluate() stay with the modelling library. */
int ab2_gar_linear_step(ab2_gar_solver *s, double alpha, const ab2_ls_iterate *current, const ab2_ls_trial *trial,
                        void *stream);
/* Replaces: ALFunction::directionalDerivative, merit-function.hxx:68-104 (Lxs [batch][N+1][nx], Lus [batch][N][nu]
 * = the Lagrangian gradients) and costDirectionalDerivative, :13-31 (pass the cost gradients): dst[batch] =
 * sum_t Lxs_t . dxs_t + sum_t Lus_t . dus_t.  dst in host or device memory. */
int ab2_gar_directional_derivative(ab2_gar_solver *s, const double *Lxs, const double *Lus, double *dst, int memspace,
                                   void *stream);
/* Replaces: ALFunction::evaluate, merit-function.hxx:33-66: dst[batch] = cost[batch] (NULL = 0) + 1/2 (mucstr
 * |lam0|^2 + mudyn sum |lams_t|^2 + mucstr sum |vs_t|^2 + mucstr |vsT|^2) of the multiplier estimates `plus`
 * (only lam0, lams, vs, vsT are read). */
int ab2_gar_al_value(ab2_gar_solver *s, const ab2_ls_iterate *plus, const double *cost, double mudyn, double mucstr,
                     double *dst, int memspace, void *stream);

/* SolverFDDPTpl::backwardPass, solvers/fddp/solver-fddp.hxx:204-277, for a batch: the
 * unconstrained recursion is the sweep's own stage step with A = Jx, B = Ju, f_i = fs[i+1], Q = Lxx + preg I,
 * S = Lxu, R = Luu + preg I, q = Lx, r = Lu (nc = 0; the LLT of Quu (:259-260) is the Bunch-Kaufman factorisation
 * on its all-1x1-pivots path), so this assembles the knots from FDDP's buffers on the device, runs backward()
 * and adds FDDP's own bookkeeping: Vx_i += Vxx_i fs[i] (:219-220, 274-276) and Quuks_i = Quu_i k_i (:264).
 * DEVICE pointers: Jx [batch][N][nx*nx], Ju [batch][N][nx*nu] (column-major), fs [batch][N+1][nx], Lxx, Lxu, Luu,
 * Lx, Lu per stage, Lxx_N [batch][nx*nx], Lx_N [batch][nx].  The solver must have nc = nct = 0, nc0 = nx.
 * After the call: K_i, k_i = the first nu rows of FB / FF (kkt_fb, kkt_ff), Vxx_i = VXX (symmetric from the lower
 * triangle, + preg on the diagonal carried by Q); Vx_out [batch][N+1][nx] and Quuks_out [batch][N][nu] (device,
 * may be NULL) receive FDDP's Vx_ (with the defect term) and Quuks_. */
typedef struct ab2_fddp_inputs {
  const double *Jx, *Ju, *fs, *Lxx, *Lxu, *Luu, *Lx, *Lu, *Lxx_N, *Lx_N;
  double preg;
} ab2_fddp_inputs;
int ab2_fddp_backward_pass(ab2_gar_solver *s, const ab2_fddp_inputs *in, double *Vx_out, double *Quuks_out, void *stream);

/* Replaces: cycleAppend(knot), proximal-riccati.hxx:79-86 + the problem rotation the
 * caller performs (solver-proxddp.hxx:202-209): factors and stage knots of every
 * instance shift one knot to the left; `new_last` ([batch][stage_record], memspace)
 * becomes stage knot N-1; its factor slot and kkt0 are zeroed.
 * O(1) in the horizon: the per-knot factor arrays and the solver-owned copy of the stage records are
 * rings -- the call advances a head index, zeroes ONE factor slot and writes ONE record per instance; the
 * getters (ab2_gar_get / get_range / get_gains / get_problem / first_step_policy) and the kernels apply the
 * head, and the next backward() rewrites the factors in plain knot order.  Raw device pointers
 * (ab2_gar_device_ptr, ab2_gar_problem_ptr) see the physical layout: stage knot t sits in slot
 * (t + head) mod N, heads from ab2_gar_ring_heads (0 except after a cycle).  The parallel solver drops
 * every factor instead (parallel-solver.hxx:246-258). */
int ab2_gar_cycle_append(ab2_gar_solver *s, const double *new_last, int memspace, void *stream);

int ab2_gar_ring_heads(const ab2_gar_solver *s, int *factor_head, int *stage_head);
/* Profiling aid (environment AB2_PHASE_CLOCKS=1 at create): clock64() cycles per phase of the CTA-per-instance
 * kernel, summed over the knots of instance 0 since the last call; 16 counters (csrc/riccati_block.cuh). */
int ab2_gar_phase_clocks(ab2_gar_solver *s, long long *dst16);

int ab2_gar_synchronize(ab2_gar_solver *s, void *stream);
/* Page-locked host memory for the buffers handed to set_problem / get / sweep_host: copies to and
 * from it are asynchronous (what the reference's mimalloc arena is to its hot loop,
 * solver-proxddp.hpp:177-178: allocate once, never in the loop).  free(NULL) is a no-op. */
int ab2_gar_pinned_alloc(size_t bytes, void **out);
void ab2_gar_pinned_free(void *p);
/* Kernels launched by this solver since creation (for bench accounting). */
long ab2_gar_launch_count(const ab2_gar_solver *s);
/* Shared memory per CTA / registers etc. of the kernel serving this solver
 * (*regs_per_thread: low 16 bits = registers, high 16 bits = resident CTAs per SM). */
int ab2_gar_kernel_info(const ab2_gar_solver *s, int *group_lanes, int *smem_bytes_per_cta,
                        int *threads_per_cta, int *grid, int *regs_per_thread);
const char *ab2_gar_last_error(void);
const char *ab2_gar_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ALIGATOR_B200_GAR_H */
