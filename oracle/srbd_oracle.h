/*
 * srbd_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the reference's SRBD-NMPC hot path, used ONLY as the checker in
 * tests/, __graft_entry__.smoke() and as the `cpu_baseline` / `--impl reference` leg of bench.py.
 * The product path (srbd-nmpc-solver_b200/csrc) never links, imports or calls anything here.
 *
 * What is restated (paths relative to the reference checkout):
 *   dynamics/orientation_tool.h:55-227, dynamics/SRBD_model.cpp:75-295,
 *   NMPC_solver.cpp:149-314 (line search, QP assembly),
 *   hpipm-cpp/src/ocp_qp_ipm_solver.cpp:181-414 (facade semantics: x0 embedding, outputs,
 *   stage-0 Riccati reconstruction), and the HPIPM OCP-QP interior-point algorithm.
 *
 * PINNING STATUS
 *   - OCP-QP solve: pinned by the reference's own golden vectors hpipm-cpp/test/sol0..14.txt
 *     (committed as tests/golden/quadcopter_sol.npz) and by the analytic Riccati identities of
 *     hpipm-cpp/test/ocp_qp_ipm_solver.cpp:60-109 (tests/test_oracle_qp.py).
 *   - HPIPM itself (giaf/hpipm, giaf/blasfeo; unpinned HEAD in the reference's README.md:10-31)
 *     is NOT vendored in the reference (headers only) and cannot be built here, so its
 *     per-iteration path and iteration counts are "parity unpinned": the algorithm below is the
 *     published HPIPM algorithm restated from SURVEY.md Appendix C.
 *   - dynamics / Jacobians / constraint rows / barrier / assembly / line search: the reference has
 *     no tests or fixtures for them -> "parity unpinned"; checked here by finite differences and
 *     by an independent numpy restatement (tests/test_oracle_model.py).
 */
#ifndef SRBD_ORACLE_H_
#define SRBD_ORACLE_H_

#include "../include/srbd_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- SO(3) helpers, column-major 3x3 (orientation_tool.h) ------------------------------------ */
void orc_skew(const double v[3], double M[9]);
void orc_expm(const double r[3], double R[9]);
void orc_jl(const double r[3], double J[9]);
void orc_jlt(const double r[3], double J[9]);
void orc_djl(const double r[3], double dJ[27]);  /* three 3x3: d/dr_x, d/dr_y, d/dr_z */
void orc_djlt(const double r[3], double dJ[27]);

/* ---- SRBD model (SRBD_model.cpp) -------------------------------------------------------------- */
/* dx(12); jfx, jfu (12x12 col-major) may be NULL */
void orc_continuous(const srbd_model_params* m, const double* x, const double* u, double* dx,
                    double* jfx, double* jfu);
/* A,B (12x12), b(12), f(12); any output may be NULL (GetShootingDynamic, :143-235) */
void orc_shooting(const srbd_model_params* m, const double* x, const double* xn, const double* u,
                  double* A, double* B, double* b, double* f);
/* Ac (24x12 col-major), f(24) = Ac u + bc; stance[2]: 1 stance / 0 swing (fmax := swing_fmax) */
void orc_constraint(const srbd_model_params* m, const double* u, const uint8_t* stance, double* Ac,
                    double* f);
void orc_barrier(double v, double mu, double theta, double* b, double* db, double* ddb);

/* ---- SQP assembly (NMPC_solver.cpp:276-314) ---------------------------------------------------- */
/* One QP.  x[(N+1)*12], u[N*12], xref[(N+1)*12], contact[N*2] or NULL.
 * Outputs (column-major, hpipm-cpp OcpQp fields): A,B [N][144], b [N][12], Q [N+1][144],
 * S [N][144], R [N][144], q [N+1][12], r [N][12]; HARD_INEQ additionally D [N][24*12],
 * lg [N][24], lg_mask [N][24] (C = 0, ug masked); defect [N][12], fcon [N][24]. */
void orc_assemble(const srbd_model_params* m, int N, int mode, const double* x, const double* u,
                  const double* xref, const uint8_t* contact, double* A, double* B, double* b,
                  double* Q, double* S, double* R, double* q, double* r, double* D, double* lg,
                  double* lg_mask, double* defect, double* fcon);

/* ---- filter line search (NMPC_solver.cpp:149-274) ------------------------------------------------ */
/* Updates x,u in place with the accepted step, carries *alpha (NMPC_solver.h:104, never reset by
 * the reference).  merit[3] = phi, dphi, theta.  Returns 1 when converged (:267). */
int orc_line_search(const srbd_model_params* m, int N, double* x, double* u, const double* xref,
                    const uint8_t* contact, const double* dx, const double* du, double* alpha,
                    double* merit);
/* The same for a step that came from a QP assembled in `mode`: in SRBD_HARD_INEQ (the extension of
 * NMPC_solver.cpp:300-304) only the rows the assembly keeps as a relaxed barrier enter phi / dphi. */
int orc_line_search_mode(const srbd_model_params* m, int N, int mode, double* x, double* u,
                         const double* xref, const uint8_t* contact, const double* dx, const double* du,
                         double* alpha, double* merit);

/* ---- OCP-QP IPM (hpipm::OcpQpIpmSolver::solve semantics) --------------------------------------------- */
/* Solves ONE QP (index `which` of the batch views in qp / sol / st).  Returns the HpipmStatus. */
int orc_qp_solve_one(const srbd_qp_dims* d, const srbd_ipm_args* a, const srbd_qp_host* qp,
                     const srbd_sol_host* sol, const srbd_stats_host* st, int stat_rows, int which);
/* Whole batch, OpenMP over QPs with `threads` threads (<=0: all). Returns 0. */
int orc_qp_solve_batch(const srbd_qp_dims* d, const srbd_ipm_args* a, const srbd_qp_host* qp,
                       const srbd_sol_host* sol, const srbd_stats_host* st, int stat_rows,
                       int batch, int threads);

/* ---- full pipeline on trajectories (what bench.py's cpu_baseline times) ---------------------------- */
/* linearize + assemble + IPM solve for `batch` SRBD problems; sol_x [B][N+1][12], sol_u [B][N][12],
 * optional lam,t [B][N*48], pi [B][N+1][12]. */
int orc_pipeline_batch(const srbd_model_params* m, const srbd_ipm_args* a, int N, int mode, int batch,
                       const double* x, const double* u, const double* xref, const double* x0,
                       const uint8_t* contact, double* sol_x, double* sol_u, double* sol_pi,
                       double* sol_lam, double* sol_t, int* iter, int* status, double* res_max,
                       int threads);
/* batched assembly / line search helpers for the tests */
int orc_assemble_batch(const srbd_model_params* m, int N, int mode, int batch, const double* x,
                       const double* u, const double* xref, const uint8_t* contact, double* A,
                       double* B, double* b, double* Q, double* S, double* R, double* q, double* r,
                       double* D, double* lg, double* lg_mask, double* defect, double* fcon,
                       int threads);
int orc_num_threads(void);
size_t orc_qp_nct(const srbd_qp_dims* d);
void orc_model_params_default(srbd_model_params* p, int horizon);
void orc_ipm_args_default(srbd_ipm_args* a);

#ifdef __cplusplus
}
#endif
#endif
