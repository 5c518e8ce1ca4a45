/*
 * srbd_b200.h — C-ABI of the B200-native SRBD-NMPC hot path.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, no torch /
 * Eigen / STL types.  Everything above it (the C++ facades in
 * srbd-nmpc-solver_b200/host/, the Python ctypes binding, bench.py) goes through these entry
 * points; everything below it is hand-written CUDA for sm_100a.  There is NO CPU fallback:
 * every compute entry point returns SRBD_ERR_CUDA when no device is usable.
 *
 * Reference interfaces replaced (paths relative to the reference checkout):
 *   - SRBDModel::GetShootingDynamic / GetContinuousDynamic   dynamics/SRBD_model.cpp:75-235
 *     + SO(3) helpers                                         dynamics/orientation_tool.h:55-227
 *         -> srbd_linearize()
 *   - SRBDModel::GetConstrain / Barrier                       dynamics/SRBD_model.cpp:237-295
 *     + NMPCSolver::prepareQpStructures                       NMPC_solver.cpp:276-314
 *         -> srbd_assemble()
 *   - hpipm::OcpQpIpmSolver::solve                            hpipm-cpp/src/ocp_qp_ipm_solver.cpp:181-414
 *     (d_ocp_qp_set_all + masks + d_ocp_qp_ipm_solve + getters, hpipm_d_ocp_qp_ipm.h:139-242)
 *         -> srbd_qp_upload() / srbd_qp_solve() / srbd_download_*()
 *   - NMPCSolver::linearSearch                                NMPC_solver.cpp:149-274
 *         -> srbd_line_search()
 *   - NMPCSolver::solveQpProblems + SQP loop                  NMPC_solver.cpp:316-330,367-375
 *         -> srbd_sqp_iterate()
 *
 * Conventions
 *   - all matrices are column-major doubles, exactly what Eigen's .data() hands to HPIPM
 *     (hpipm-cpp/src/ocp_qp_ipm_solver.cpp:227-281);
 *   - batches are contiguous: [B][stage][elements];
 *   - every function returns 0 on success, <0 on error (srbd_status_t); the last error string
 *     of a context is available through srbd_last_error();
 *   - per-QP solver outcome is a value of hpipm-cpp's HpipmStatus
 *     (hpipm-cpp/include/hpipm-cpp/ocp_qp_ipm_solver.hpp:24-30): 0 Success, 1 MaxIterReached,
 *     2 MinStepLengthReached, 3 NaNDetected, 4 UnknownFailure;
 *   - a context is NOT thread safe (the reference solver object is not either,
 *     ocp_qp_ipm_solver.hpp:157-158): one context per host thread / CUDA stream.
 */
#ifndef SRBD_B200_H_
#define SRBD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRBD_NX 12 /* SRBD_model.cpp:21 */
#define SRBD_NU 12 /* SRBD_model.cpp:22 */
#define SRBD_NG 24 /* SRBD_model.cpp:23 */
#define SRBD_NCONTACT 2 /* two 6-D wrench contacts, SRBD_model.h:55,67 */
#define SRBD_STAT_M 18 /* columns of the per-iteration statistics table, ocp_qp_ipm_solver.cpp:381 */

typedef enum srbd_status_t {
  SRBD_OK = 0,
  SRBD_ERR_ARG = -1,    /* bad argument / shape (the facades rethrow as std::runtime_error) */
  SRBD_ERR_CUDA = -2,   /* CUDA runtime error or no device */
  SRBD_ERR_STATE = -3,  /* call order (e.g. solve before upload) */
  SRBD_ERR_UNSUPPORTED = -4
} srbd_status_t;

/* Model + cost constants.  Defaults (srbd_model_params_default) are the reference's:
 * SRBD_model.cpp:12-23, NMPC_solver.cpp:56-58,334-338, config/mpc_option.yaml:2-18. */
typedef struct srbd_model_params {
  double mass;            /* SetMass(15.0) */
  double dt;              /* dt_MPC 0.015 */
  double inertia_inv[9];  /* Lbody_ = inertia^-1, column-major 3x3 (SRBD_model.cpp:46-49) */
  double foot_pos[6];     /* pf_: right foot (3), left foot (3) */
  double foot_rot[18];    /* Rf_: two column-major 3x3 foot rotations */
  double mu;              /* friction coefficient 0.5 */
  double Lfx, Lfz;        /* CoP / yaw-torque half sizes 0.05 */
  double fmax, fmin;      /* 1000, 0 */
  double gravity[3];      /* (0, 0, -9.8) */
  double Q[12];           /* diag(Q) */
  double Qf[12];          /* diag(Qf) ALREADY multiplied by the horizon (NMPC_solver.cpp:58) */
  double R;               /* R = R_read * I */
  double mu_b, theta_b;   /* relaxed log barrier (mpc_option.yaml:17-18) */
  double swing_fmax;      /* extension: fmax of a swing contact */
} srbd_model_params;

/* IPM arguments = hpipm-cpp's OcpQpIpmSolverSettings
 * (hpipm-cpp/include/hpipm-cpp/ocp_qp_ipm_solver_settings.hpp:21-92) plus the hidden HPIPM
 * SPEED-mode defaults that influence iterates. */
typedef struct srbd_ipm_args {
  int iter_max;
  double alpha_min;
  double mu0;
  double tol_stat, tol_eq, tol_ineq, tol_comp;
  double reg_prim;
  int warm_start;
  int pred_corr;
  int ric_alg;     /* 0 classical (what NMPC_solver.cpp:81 selects), 1 square-root */
  int split_step;
  /* hidden constants (named so they can be flipped once diffed against real HPIPM) */
  int cond_pred_corr;      /* 1 */
  double cond_factor;      /* 2.0: centering fallback fires when mu_corr > cond_factor * mu_aff */
  double thr0;             /* 0.1 initial slack threshold */
  double lam_min, t_min, tau_min; /* 1e-16 */
  int t_lam_min;           /* 2: clip lam,t after the update */
  int alpha_shorten;       /* 0: alpha*=0.995 ; 1: alpha*=((1-alpha)*0.99+alpha*0.9999999) */
  /* iterative refinement of the Newton steps (hpipm_d_ocp_qp_ipm.h:74-75; 0 / 0 in SPEED, 0 / 2 in BALANCE, 0 / 4 in
   * ROBUST mode): after the predictor (itref_pred_max) / corrector (itref_corr_max) solve, the residual of the LINEAR
   * KKT system is computed, and while it is neither below itref_abs x the exit tolerance nor below itref_rel x the
   * current nonlinear residual (component-wise: stat, eq, ineq, comp) the system is solved again for it with the same
   * factorization and the correction added to the step. */
  int itref_pred_max, itref_corr_max;
  double itref_abs, itref_rel; /* 1.0, 1e-3 */
} srbd_ipm_args;

/* Uniform OCP-QP dimensions (stage 0 has its state eliminated by the x0 embedding, stage N has
 * no input: hpipm-cpp/src/ocp_qp_ipm_solver.cpp:128-130, ocp_qp_dim.cpp:47-48). */
typedef struct srbd_qp_dims {
  int N;    /* horizon: stages 0..N */
  int nx;   /* states */
  int nu;   /* inputs (stages 0..N-1) */
  int nbx;  /* box-constrained states (stages 1..N) */
  int nbu;  /* box-constrained inputs (stages 0..N-1) */
  int ng;   /* general constraint rows (stages 0..N-1) */
  int ngN;  /* general constraint rows at stage N (C only) */
} srbd_qp_dims;

/* Host-side view of a batch of QPs, column-major blocks, [B][stage][...] contiguous.
 * NULL is allowed for: S (zeros), every *_mask (all ones), C (zeros), x_init/u_init. */
typedef struct srbd_qp_host {
  const double* A;   /* [B][N][nx*nx] */
  const double* Bm;  /* [B][N][nx*nu] */
  const double* b;   /* [B][N][nx] */
  const double* Q;   /* [B][N+1][nx*nx] */
  const double* S;   /* [B][N][nu*nx]  (nu x nx, ocp_qp.hpp) */
  const double* R;   /* [B][N][nu*nu] */
  const double* q;   /* [B][N+1][nx] */
  const double* r;   /* [B][N][nu] */
  const int* idxbx;  /* [nbx], shared by all stages and QPs */
  const double* lbx; /* [B][N+1][nbx] (stage 0 ignored: nbx[0]:=0, ocp_qp_ipm_solver.cpp:129) */
  const double* ubx;
  const double* lbx_mask;
  const double* ubx_mask;
  const int* idxbu;  /* [nbu] */
  const double* lbu; /* [B][N][nbu] */
  const double* ubu;
  const double* lbu_mask;
  const double* ubu_mask;
  const double* C;   /* [B][N][ng*nx]  (stage 0 ignored: nx[0]:=0) */
  const double* D;   /* [B][N][ng*nu] */
  const double* lg;  /* [B][N][ng] */
  const double* ug;
  const double* lg_mask;
  const double* ug_mask;
  const double* CN;  /* [B][ngN*nx] */
  const double* lgN; /* [B][ngN] */
  const double* ugN;
  const double* lgN_mask;
  const double* ugN_mask;
  const double* x0;     /* [B][nx] */
  const double* x_init; /* [B][N+1][nx] primal warm start (warm_start=1) */
  const double* u_init; /* [B][N][nu] */
} srbd_qp_host;

/* Host-side outputs; any pointer may be NULL (not downloaded). */
typedef struct srbd_sol_host {
  double* x;   /* [B][N+1][nx]; x[0] = x0 (ocp_qp_ipm_solver.cpp:337) */
  double* u;   /* [B][N][nu] */
  double* pi;  /* [B][N+1][nx] */
  double* lam; /* [B][nct]  per stage [lb lg ub ug] (hpipm_d_ocp_qp_sol.h:57-63 order) */
  double* t;   /* [B][nct] */
  double* P;   /* [B][N+1][nx*nx] */
  double* p;   /* [B][N+1][nx] */
  double* K;   /* [B][N][nu*nx] */
  double* k;   /* [B][N][nu] */
} srbd_sol_host;

typedef struct srbd_stats_host {
  int* iter;        /* [B] */
  int* status;      /* [B] HpipmStatus values */
  double* res_max;  /* [B][4] stat, eq, ineq, comp */
  double* stat;     /* [B][stat_rows][18] or NULL (stat_rows from srbd_ctx_stat_rows) */
} srbd_stats_host;

/* Batch-level statistics block, reduced on the device as the epilogue of the solve and gathered
 * across ranks by the host. */
#define SRBD_HIST_BINS 64
typedef struct srbd_batch_stats {
  long long solves;
  long long iter_sum;
  long long iter_hist[SRBD_HIST_BINS];
  long long status_count[5];
  double res_max[4];
} srbd_batch_stats;

typedef struct srbd_ctx srbd_ctx; /* opaque; owns device + pinned host buffers */

typedef enum srbd_assemble_mode {
  SRBD_BARRIER_SOFT = 0, /* reference: constraints folded into R,r (NMPC_solver.cpp:288-309) */
  SRBD_HARD_INEQ = 1     /* the variant commented out at NMPC_solver.cpp:300-304 */
} srbd_assemble_mode;

/* device buffer ids for srbd_ctx_device_ptr (zero-copy interop with torch) */
typedef enum srbd_buf {
  SRBD_BUF_TRAJ_X = 0,   /* [B][N+1][12] */
  SRBD_BUF_TRAJ_U = 1,   /* [B][N][12] */
  SRBD_BUF_TRAJ_XREF = 2,/* [B][N+1][12] */
  SRBD_BUF_X0 = 3,       /* [B][nx] */
  SRBD_BUF_CONTACT = 4,  /* uint8 [B][N][2] */
  SRBD_BUF_SOL_X = 5,    /* [B][N+1][nx] */
  SRBD_BUF_SOL_U = 6,    /* [B][N][nu] */
  SRBD_BUF_SOL_PI = 7,   /* [B][N+1][nx] */
  SRBD_BUF_SOL_LAM = 8,  /* [B][nct] */
  SRBD_BUF_SOL_T = 9,    /* [B][nct] */
  SRBD_BUF_ITER = 10,    /* int [B] */
  SRBD_BUF_STATUS = 11,  /* int [B] */
  SRBD_BUF_RESMAX = 12,  /* [B][4] */
  SRBD_BUF_BABT = 13,    /* packed panel-major stage records, see DESIGN.md */
  SRBD_BUF_RSQRQ = 14,
  SRBD_BUF_DCT = 15,
  SRBD_BUF_D = 16,
  SRBD_BUF_DMASK = 17,
  SRBD_BUF_DEFECT = 18,  /* [B][N][12] shooting defect f (SRBD_model.cpp:189-197) */
  SRBD_BUF_STAGE_REC = 19, /* [B][N+1][192] compact stage records written by srbd_assemble next to the dense ones:
                              [R tile 96 | pad 12 | gradient row 24 | pad 12 | lg 24 | lg mask 24], see DESIGN.md section 3 */
  SRBD_BUF_BABT_DYN = 20,  /* [B][N][72] the 36 16-byte chunks of each BAbt record that hold a stage-dependent element (written by
                              srbd_linearize next to the dense records; the SRBD variant of K3 streams these and keeps the model
                              constants of the record resident in shared memory), see DESIGN.md section 3 */
  SRBD_BUF_COUNT = 21
} srbd_buf;

/* ---- defaults ------------------------------------------------------------------------------ */
void srbd_model_params_default(srbd_model_params* p, int horizon);
/* The fields of OcpQpIpmSolverSettings get that struct's defaults (ocp_qp_ipm_solver_settings.hpp:26-86), the hidden
 * constants those of HPIPM's SPEED mode (d_ocp_qp_ipm_arg_set_default(SPEED)). */
void srbd_ipm_args_default(srbd_ipm_args* a);
/* The hidden constants of d_ocp_qp_ipm_arg_set_default(mode) (hpipm-cpp/src/ocp_qp_ipm_solver.cpp:103; mode = hpipm-cpp's
 * HpipmMode: 0 SpeedAbs, 1 Speed, 2 Balance, 3 Robust) that this implementation honours: cond_pred_corr (0 in SpeedAbs),
 * itref_pred_max / itref_corr_max (0 / 2 in Balance, 0 / 4 in Robust).  NOT implemented: the absolute-form iteration of
 * SpeedAbs (abs_form = 1: the delta form runs instead) and the LQ re-factorization of Balance / Robust (lq_fact = 1 / 2).
 * Returns SRBD_ERR_ARG for an unknown mode.  The public fields (iter_max, tolerances, ...) are left alone: hpipm-cpp
 * overrides them from OcpQpIpmSolverSettings right after (:104-116). */
int srbd_ipm_args_set_mode(srbd_ipm_args* a, int mode);
size_t srbd_qp_nct(const srbd_qp_dims* d); /* length of lam / t per QP */

/* ---- context ------------------------------------------------------------------------------- */
/* stream: a cudaStream_t (as void*) to run on, or NULL to let the context create its own. */
int srbd_ctx_create(int device, int batch, const srbd_qp_dims* dims, void* stream, srbd_ctx** out);
int srbd_ctx_destroy(srbd_ctx* ctx);
const char* srbd_last_error(const srbd_ctx* ctx);
int srbd_set_model(srbd_ctx* ctx, const srbd_model_params* p);
int srbd_set_ipm_args(srbd_ctx* ctx, const srbd_ipm_args* a);
/* optional outputs of the solve (off by default: they cost HBM): Riccati P,p,K,k (+ pi[0], which the
 * reference reconstructs from them, ocp_qp_ipm_solver.cpp:349-373) and the 18-column statistics table */
int srbd_set_outputs(srbd_ctx* ctx, int export_ric, int export_stat);
int srbd_ctx_stat_rows(const srbd_ctx* ctx);
void* srbd_ctx_stream(const srbd_ctx* ctx);
int srbd_ctx_device_ptr(srbd_ctx* ctx, int buf, void** ptr, size_t* bytes);
int srbd_ctx_sync(srbd_ctx* ctx);
/* number of kernels the context has launched so far (bench.py reports the delta) */
long long srbd_ctx_launch_count(const srbd_ctx* ctx);

/* ---- NMPC level (SRBD dims only: nx=nu=12, ng=24) -------------------------------------------- */
/* host -> device copy of the SQP iterate.  contact: uint8 [B][N][2], 1 = stance, NULL = all stance */
int srbd_upload_traj(srbd_ctx* ctx, const double* x, const double* u, const double* xref,
                     const double* x0, const uint8_t* contact);
int srbd_download_traj(srbd_ctx* ctx, double* x, double* u);
int srbd_linearize(srbd_ctx* ctx);           /* K1: A,B,b,defect for every (QP, stage) */
int srbd_assemble(srbd_ctx* ctx, int mode);  /* K2: packed RSQrq / DCt / d / masks + the compact stage records */
int srbd_download_linearization(srbd_ctx* ctx, double* A, double* Bm, double* b, double* defect);
/* unpack the packed stage records back to column-major hpipm-cpp fields (tests / facades) */
int srbd_download_qp(srbd_ctx* ctx, double* Q, double* S, double* R, double* q, double* r,
                     double* D, double* lg, double* lg_mask);

/* ---- QP level (any dims up to the compiled maxima) ------------------------------------------- */
/* H2D + pack (d_ocp_qp_set_all analog).  Host fields that lie in one contiguous range (a staging arena) travel in a single
 * copy.  With the SRBD dimensions (nx = nu = 12, ng = 24, no boxes) the batch is also checked ON THE DEVICE for the
 * structure srbd_assemble produces (S = 0, C = 0, one constant diagonal Q for stages 1..N-1, one constant D made of two
 * 12 x 6 blocks, ug masked); srbd_qp_solve then routes it to the tensor-core variant of K3 when the settings allow
 * (cold start, ric_alg = 0, no Riccati / statistics exports) -- the reference's own boundary reaches the fast kernel. */
int srbd_qp_upload(srbd_ctx* ctx, const srbd_qp_host* qp);
/* Layout of the host fields of the NEXT uploads of this context (sticky).  d_shared != 0: qp->D is ONE ng x nu matrix
 * shared by every stage of every QP -- what NMPCSolver::prepareQpStructures hands to hpipm-cpp (the friction-cone / force-box
 * matrix of SRBDModel::GetConstrain is a constant, NMPC_solver.cpp:290-300) -- instead of B x N copies of it.  Default 0. */
int srbd_qp_upload_layout(srbd_ctx* ctx, int d_shared);
/* K3: the whole IPM solve of every QP of the batch (d_ocp_qp_ipm_solve, hpipm_d_ocp_qp_ipm.h:238), asynchronous on
 * the context's stream.  QPs assembled by srbd_assemble() go through the tensor-core variant (one launch) followed by
 * the rescue launch of the generic kernel for the QPs that ran to iter_max (usually none);
 * everything else (srbd_qp_upload data, warm start, Riccati / statistics exports) through the generic kernel. */
int srbd_qp_solve(srbd_ctx* ctx);
/* P, p, K, k (and a meaningful pi[0]) exist only when the LAST solve ran with srbd_set_outputs(ctx, 1, ..): otherwise
 * asking for them returns SRBD_ERR_STATE (never the exports of an earlier solve). */
int srbd_download_solution(srbd_ctx* ctx, const srbd_sol_host* sol);
int srbd_download_stats(srbd_ctx* ctx, const srbd_stats_host* st);
/* Lr of stage 0, [B][nu*nu] column-major lower triangle (zeros above): the Cholesky factor of the last factorization's
 * stage-0 Hessian, i.e. d_ocp_qp_ipm_get_ric_Lr(qp, arg, ws, 0, Lr) -- what hpipm-cpp's own stage-0 reconstruction reads
 * (hpipm-cpp/src/ocp_qp_ipm_solver.cpp:352).  Exported together with P, p, K, k (srbd_set_outputs(ctx, 1, ..)). */
int srbd_download_ric_lr0(srbd_ctx* ctx, double* Lr0);
/* The same outputs in ONE device-to-host copy: x, u, pi, res_max, iter, status (and lam, t with with_duals = 1) live in one
 * device arena; srbd_out_layout returns their offsets IN DOUBLES inside it (order: x, u, pi, res_max [B][4], iter int32
 * [B], status int32 [B], lam, t) and its total length; srbd_download_packed copies the arena (up to `status` without
 * the duals) into a host buffer of that layout -- pinned (srbd_host_alloc) for an asynchronous copy -- and waits. */
int srbd_out_layout(const srbd_ctx* ctx, size_t offs[8], size_t* total_doubles);
int srbd_download_packed(srbd_ctx* ctx, double* arena, int with_duals);
/* pinned (page-locked) host memory for staging buffers that outlive a solve (the facades own one arena per solver:
 * the reference's wrappers own their workspaces the same way, detail/d_ocp_qp_ipm_ws_wrapper.cpp:141-155) */
int srbd_host_alloc(size_t bytes, void** ptr);
int srbd_host_free(void* ptr);
int srbd_batch_stats_get(srbd_ctx* ctx, srbd_batch_stats* out);

/* ---- closed-loop batched MPC --------------------------------------------------------------------*/
/* The MPC loop of the reference's example and test (hpipm-cpp/examples/example_mpc.cpp:99-119,
 * hpipm-cpp/test/ocp_qp_ipm_solver.cpp:298-314) for B robots at once, on the device: for t = 0 .. steps-1
 *     x0 := x(t);  solve the QP uploaded by srbd_qp_upload (only x0 changes: b0 <- A0 x0 + b0, r0 <- S0 x0 + r0 are
 *     re-embedded);  x(t+1) := A x(t) + B u0 + b.
 * With warm_start = 1 step 0 starts from the uploaded x_init / u_init and every later step from the previous solution,
 * exactly like passing `solution` back into OcpQpIpmSolver::solve.  No host round trip between the steps.
 * A, Bm, b: the plant, column-major [nx*nx], [nx*nu], [nx], one per robot ([B][...]) or shared (plant_shared = 1).
 * x_start [B][nx].  Outputs (any may be NULL): x_traj [steps+1][B][nx], u_traj [steps][B][nu], iter / status
 * [steps][B].  After the call srbd_download_solution returns the solution of the LAST step. */
int srbd_mpc_run(srbd_ctx* ctx, const double* A, const double* Bm, const double* b, int plant_shared,
                 const double* x_start, int steps, double* x_traj, double* u_traj, int* iter, int* status);

/* ---- SQP level --------------------------------------------------------------------------------*/
/* K4: filter line search on the device: updates the trajectory in place, per-QP alpha carried in
 * the context (NMPC_solver.h:104), writes converged[B] (NMPC_solver.cpp:267).  merit: [B][3] =
 * phi, dphi, theta or NULL. */
int srbd_line_search(srbd_ctx* ctx);
int srbd_download_sqp_state(srbd_ctx* ctx, double* alpha, int* converged, double* merit);
int srbd_reset_sqp_state(srbd_ctx* ctx);
/* linearize + assemble + solve (+ line search if do_line_search) with device-resident inputs */
int srbd_sqp_iterate(srbd_ctx* ctx, int mode, int do_line_search);
/* The outer SQP loop of NMPCSolver::controlLoop (NMPC_solver.cpp:367-375) ON THE DEVICE: up to max_iter iterations of
 * K1 + K2 + K3 + K4 enqueued without a host round trip.  Per problem the loop ends at its first "nmpc solve success"
 * (:372-374): a converged problem is frozen (K3 skips it, K4 leaves its trajectory alone), and once every problem of the
 * batch has converged the remaining launches return at once (a device-side counter gates them).  Resets the converged
 * flags, NOT the carried step length alpha (NMPC_solver.h:104; srbd_reset_sqp_state does).  Asynchronous; read the
 * outcome with srbd_download_sqp_state (converged flags, merit of the last iteration of each problem) and
 * srbd_download_sqp_iters (SQP iterations each problem took), the trajectories with srbd_download_traj. */
int srbd_sqp_solve(srbd_ctx* ctx, int mode, int max_iter);
int srbd_download_sqp_iters(srbd_ctx* ctx, int* iters);

/* ---- end to end (host buffers in, host buffers out; the `e2e` leg of bench.py) ----------------*/
int srbd_solve_host(srbd_ctx* ctx, int mode, const double* x, const double* u, const double* xref,
                    const double* x0, const uint8_t* contact, double* sol_x, double* sol_u,
                    int* iter, int* status);
/* The same, asynchronously: enqueues the H2D copies, the kernels and the D2H copies on the context's stream and
 * returns; srbd_wait() blocks until they are done.  Host buffers must be pinned and stay untouched until then.  Two
 * contexts driven alternately overlap the copies of one batch with the kernels of the other (streaming use: the
 * reference calls NMPCSolver::solveQpProblems once per control step, NMPC_solver.cpp:316-330). */
int srbd_solve_host_async(srbd_ctx* ctx, int mode, const double* x, const double* u, const double* xref,
                          const double* x0, const uint8_t* contact, double* sol_x, double* sol_u,
                          int* iter, int* status);
int srbd_wait(srbd_ctx* ctx);
/* The low-latency form of srbd_solve_host (BASELINE config 5: one problem per call, the reference's call site
 * NMPCSolver::solveQpProblems, NMPC_solver.cpp:316-330): the inputs are staged in pinned memory owned by the context and
 * the whole pipeline -- H2D copies, K1, K2, K3 (+ its rescue launches), D2H copies -- is replayed as ONE CUDA graph
 * launch (captured on the first call for a given mode).  Any host buffers; synchronous. */
int srbd_solve_host_graph(srbd_ctx* ctx, int mode, const double* x, const double* u, const double* xref,
                          const double* x0, const uint8_t* contact, double* sol_x, double* sol_u,
                          int* iter, int* status);

/* ---- measurement helpers ----------------------------------------------------------------------*/
/* FP64-pipe-saturating microbenchmark (DMMA m8n8k4 chains; DMMA and DFMA share one datapath on B200): returns the
 * achieved FP64 FLOP/s on the context's device.  A short burst: call it BEFORE a sustained run (power cap). */
int srbd_fp64_peak(srbd_ctx* ctx, double* flops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* SRBD_B200_H_ */
