/* hpipm_b200_compat.h -- SURVEY.md section 8(b), "Option A": the 54 HPIPM C symbols that the reference's hpipm-cpp wrapper
 * links against, backed by the B200 library.
 *
 * The reference's OCP-QP boundary has two levels: the C++ classes of hpipm-cpp (mirrored by host/hpipm-cpp/hpipm-cpp.hpp,
 * "Option B") and, underneath, HPIPM's C API, which hpipm-cpp calls from
 *     hpipm-cpp/src/ocp_qp_ipm_solver.cpp:103-116,123-130,283-289,295-319,330-345,350,376-407 and
 *     hpipm-cpp/src/detail/d_ocp_qp_{dim,,sol,ipm_arg,ipm_ws}_wrapper.cpp  (*_memsize / *_create / *_copy_all).
 * libsrbd_b200.so exports exactly those symbols with the signatures of the reference's vendored headers
 * (hpipm-cpp/include/include/hpipm_d_ocp_qp_dim.h:71-85, hpipm_d_ocp_qp.h:88-162, hpipm_d_ocp_qp_sol.h:68-104,
 * hpipm_d_ocp_qp_ipm.h:141-238), so an UNMODIFIED hpipm-cpp links against it in place of libhpipm + libblasfeo
 * (INTEGRATION.md shows the one-line CMake change).  d_ocp_qp_ipm_solve then runs K3 on the GPU: the single-QP,
 * host-memory-in / host-memory-out call of NMPCSolver::solveQpProblems (NMPC_solver.cpp:316-330).
 *
 * ABI.  HPIPM's memory model is kept: the caller asks *_memsize, allocates, and hands the block to *_create; nothing here
 * allocates host memory behind the caller's back, and there is no *_destroy.  The structs below have the member order and
 * types of the vendored headers (pointers to BLASFEO types are declared void*: same size and alignment), because
 * hpipm-cpp reads some members directly: d_ocp_qp_dim::N (d_ocp_qp_dim_wrapper.cpp:143), the scalar members of
 * d_ocp_qp_ipm_arg (d_ocp_qp_ipm_arg_wrapper.cpp:91-116) and d_ocp_qp_ipm_ws::stat (ocp_qp_ipm_solver.cpp:385-402).
 * What the members point to inside the caller's block is this library's business (plain column-major host arrays in
 * the layout of srbd_qp_host, not BLASFEO panels): do not mix these objects with a real libhpipm.
 * Device memory is owned by a small pool of contexts inside the library, keyed by the QP dimensions and released at
 * unload: constructing a new solver per SQP iteration (NMPC_solver.cpp:319) allocates nothing after the first.
 *
 * Supported shapes: what hpipm-cpp produces -- nx[0] = 0 after its x0 embedding, one nx for stages 1..N, one nu for stages
 * 0..N-1, one (nbx, idxbx) for stages 1..N, one (nbu, idxbu) and one ng for stages 0..N-1, ng[N] free, no soft
 * constraints (hpipm-cpp forces nsg = 0 and throws on soft data, ocp_qp_dim.cpp:43-45,217-245); nx, nu <= 12,
 * ng <= 24.  Anything else makes d_ocp_qp_ipm_solve report status 4 (hpipm-cpp maps it to HpipmStatus::UnknownFailure,
 * ocp_qp_ipm_solver.cpp:409-414) with the reason in hpipm_b200_last_error().  Not thread-safe, like the reference. */
#ifndef HPIPM_B200_COMPAT_H_
#define HPIPM_B200_COMPAT_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef size_t hpipm_size_t;                                  /* hpipm_common.h */
enum hpipm_mode { SPEED_ABS, SPEED, BALANCE, ROBUST };
enum hpipm_status { SUCCESS, MAX_ITER, MIN_STEP, NAN_SOL, INCONS_EQ };

struct d_ocp_qp_dim {                                         /* hpipm_d_ocp_qp_dim.h */
  int *nx, *nu, *nb, *nbx, *nbu, *ng, *ns, *nsbx, *nsbu, *nsg, *nbxe, *nbue, *nge;   /* [N+1] each, in the caller's block */
  int N;
  hpipm_size_t memsize;
};

struct d_ocp_qp {                                             /* hpipm_d_ocp_qp.h */
  struct d_ocp_qp_dim* dim;
  void *BAbt, *RSQrq, *DCt, *b, *rqz, *d, *d_mask, *m, *Z;    /* BAbt: this library's host store; the others stay NULL */
  int **idxb, **idxs_rev, **idxe;
  int* diag_H_flag;
  hpipm_size_t memsize;
};

struct d_ocp_qp_sol {                                         /* hpipm_d_ocp_qp_sol.h */
  struct d_ocp_qp_dim* dim;
  void *ux, *pi, *lam, *t;                                    /* ux: this library's host store */
  void* misc;
  hpipm_size_t memsize;
};

struct d_ocp_qp_ipm_arg {                                     /* hpipm_d_ocp_qp_ipm.h */
  double mu0, alpha_min, res_g_max, res_b_max, res_d_max, res_m_max, reg_prim, lam_min, t_min, tau_min;
  int iter_max, stat_max, pred_corr, cond_pred_corr, itref_pred_max, itref_corr_max, warm_start, square_root_alg, lq_fact,
      abs_form, comp_dual_sol_eq, comp_res_exit, comp_res_pred, split_step, var_init_scheme, t_lam_min, mode;
  hpipm_size_t memsize;
};

struct d_ocp_qp_ipm_ws {                                      /* hpipm_d_ocp_qp_ipm.h */
  double qp_res[4];                                           /* max residuals: stat, eq, ineq, comp */
  void* core_workspace;                                       /* this library's host store (Riccati exports of the last solve) */
  struct d_ocp_qp_dim* dim;
  void *res_workspace, *sol_step, *sol_itref, *qp_step, *qp_itref, *res_itref, *res;
  void *Gamma, *gamma, *tmp_nuxM, *tmp_nbgM, *tmp_nsM, *Pb, *Zs_inv, *tmp_m, *l;
  void *L, *Ls, *P, *Lh, *AL, *lq0, *tmp_nxM_nxM;
  double* stat;                                               /* [stat_max + 2][stat_m]: read directly by hpipm-cpp */
  int* use_hess_fact;
  void* lq_work0;
  int iter, stat_max, stat_m, use_Pb, status, square_root_alg, lq_fact, mask_constr, valid_ric_vec, valid_ric_p;
  hpipm_size_t memsize;
};

/* ---- dimensions (7) ---------------------------------------------------------------------------------------------------- */
hpipm_size_t d_ocp_qp_dim_memsize(int N);
void d_ocp_qp_dim_create(int N, struct d_ocp_qp_dim* dim, void* memory);
void d_ocp_qp_dim_copy_all(struct d_ocp_qp_dim* dim_orig, struct d_ocp_qp_dim* dim_dest);
void d_ocp_qp_dim_set_all(int* nx, int* nu, int* nbx, int* nbu, int* ng, int* nsbx, int* nsbu, int* nsg, struct d_ocp_qp_dim* dim);
void d_ocp_qp_dim_set_nx(int stage, int value, struct d_ocp_qp_dim* dim);
void d_ocp_qp_dim_set_nbx(int stage, int value, struct d_ocp_qp_dim* dim);
void d_ocp_qp_dim_set_nsbx(int stage, int value, struct d_ocp_qp_dim* dim);

/* ---- QP data (10) ------------------------------------------------------------------------------------------------------ */
hpipm_size_t d_ocp_qp_memsize(struct d_ocp_qp_dim* dim);
void d_ocp_qp_create(struct d_ocp_qp_dim* dim, struct d_ocp_qp* qp, void* memory);
void d_ocp_qp_copy_all(struct d_ocp_qp* qp_orig, struct d_ocp_qp* qp_dest);
/* per-stage pointers, column-major blocks as Eigen's .data() hands them over (ocp_qp_ipm_solver.cpp:227-289); the soft
 * constraint arguments are ignored (nsg = 0) */
void d_ocp_qp_set_all(double** A, double** B, double** b, double** Q, double** S, double** R, double** q, double** r,
                      int** idxbx, double** lbx, double** ubx, int** idxbu, double** lbu, double** ubu, double** C,
                      double** D, double** lg, double** ug, double** Zl, double** Zu, double** zl, double** zu, int** idxs,
                      double** ls, double** us, struct d_ocp_qp* qp);
void d_ocp_qp_set_lbx_mask(int stage, double* vec, struct d_ocp_qp* qp);
void d_ocp_qp_set_ubx_mask(int stage, double* vec, struct d_ocp_qp* qp);
void d_ocp_qp_set_lbu_mask(int stage, double* vec, struct d_ocp_qp* qp);
void d_ocp_qp_set_ubu_mask(int stage, double* vec, struct d_ocp_qp* qp);
void d_ocp_qp_set_lg_mask(int stage, double* vec, struct d_ocp_qp* qp);
void d_ocp_qp_set_ug_mask(int stage, double* vec, struct d_ocp_qp* qp);

/* ---- solution (8) ------------------------------------------------------------------------------------------------------ */
hpipm_size_t d_ocp_qp_sol_memsize(struct d_ocp_qp_dim* dim);
void d_ocp_qp_sol_create(struct d_ocp_qp_dim* dim, struct d_ocp_qp_sol* qp_sol, void* memory);
void d_ocp_qp_sol_copy_all(struct d_ocp_qp_sol* qp_sol_orig, struct d_ocp_qp_sol* qp_sol_dest);
void d_ocp_qp_sol_get_x(int stage, struct d_ocp_qp_sol* qp_sol, double* vec);
void d_ocp_qp_sol_get_u(int stage, struct d_ocp_qp_sol* qp_sol, double* vec);
void d_ocp_qp_sol_get_pi(int stage, struct d_ocp_qp_sol* qp_sol, double* vec);   /* multiplier of dynamics `stage`: nx[stage+1] */
void d_ocp_qp_sol_set_x(int stage, double* vec, struct d_ocp_qp_sol* qp_sol);    /* primal warm start */
void d_ocp_qp_sol_set_u(int stage, double* vec, struct d_ocp_qp_sol* qp_sol);

/* ---- solver arguments (15) --------------------------------------------------------------------------------------------- */
hpipm_size_t d_ocp_qp_ipm_arg_memsize(struct d_ocp_qp_dim* ocp_dim);
void d_ocp_qp_ipm_arg_create(struct d_ocp_qp_dim* ocp_dim, struct d_ocp_qp_ipm_arg* arg, void* mem);
void d_ocp_qp_ipm_arg_set_default(enum hpipm_mode mode, struct d_ocp_qp_ipm_arg* arg);
void d_ocp_qp_ipm_arg_set_mu0(double* mu0, struct d_ocp_qp_ipm_arg* arg);
void d_ocp_qp_ipm_arg_set_iter_max(int* iter_max, struct d_ocp_qp_ipm_arg* arg);
void d_ocp_qp_ipm_arg_set_alpha_min(double* alpha_min, struct d_ocp_qp_ipm_arg* arg);
void d_ocp_qp_ipm_arg_set_tol_stat(double* tol_stat, struct d_ocp_qp_ipm_arg* arg);
void d_ocp_qp_ipm_arg_set_tol_eq(double* tol_eq, struct d_ocp_qp_ipm_arg* arg);
void d_ocp_qp_ipm_arg_set_tol_ineq(double* tol_ineq, struct d_ocp_qp_ipm_arg* arg);
void d_ocp_qp_ipm_arg_set_tol_comp(double* tol_comp, struct d_ocp_qp_ipm_arg* arg);
void d_ocp_qp_ipm_arg_set_reg_prim(double* reg, struct d_ocp_qp_ipm_arg* arg);
void d_ocp_qp_ipm_arg_set_warm_start(int* warm_start, struct d_ocp_qp_ipm_arg* arg);
void d_ocp_qp_ipm_arg_set_pred_corr(int* pred_corr, struct d_ocp_qp_ipm_arg* arg);
void d_ocp_qp_ipm_arg_set_ric_alg(int* value, struct d_ocp_qp_ipm_arg* arg);
void d_ocp_qp_ipm_arg_set_split_step(int* value, struct d_ocp_qp_ipm_arg* arg);

/* ---- workspace, solve, getters (14) ------------------------------------------------------------------------------------ */
hpipm_size_t d_ocp_qp_ipm_ws_memsize(struct d_ocp_qp_dim* ocp_dim, struct d_ocp_qp_ipm_arg* arg);
void d_ocp_qp_ipm_ws_create(struct d_ocp_qp_dim* ocp_dim, struct d_ocp_qp_ipm_arg* arg, struct d_ocp_qp_ipm_ws* ws, void* mem);
/* H2D + pack, K3, D2H of x, u, pi, lam, t, the Riccati exports and the statistics table; synchronous */
void d_ocp_qp_ipm_solve(struct d_ocp_qp* qp, struct d_ocp_qp_sol* qp_sol, struct d_ocp_qp_ipm_arg* arg, struct d_ocp_qp_ipm_ws* ws);
void d_ocp_qp_ipm_get_iter(struct d_ocp_qp_ipm_ws* ws, int* iter);
void d_ocp_qp_ipm_get_status(struct d_ocp_qp_ipm_ws* ws, int* status);
void d_ocp_qp_ipm_get_max_res_stat(struct d_ocp_qp_ipm_ws* ws, double* res_stat);
void d_ocp_qp_ipm_get_max_res_eq(struct d_ocp_qp_ipm_ws* ws, double* res_eq);
void d_ocp_qp_ipm_get_max_res_ineq(struct d_ocp_qp_ipm_ws* ws, double* res_ineq);
void d_ocp_qp_ipm_get_max_res_comp(struct d_ocp_qp_ipm_ws* ws, double* res_comp);
/* column-major; Lr: lower Cholesky factor of the stage's barrier-augmented input Hessian (stage 0 only: the one hpipm-cpp
 * reads); P, p for stages 1..N; K (nu x nx) for stages 1..N-1 and k for stages 0..N-1 in the form u = K x + k */
void d_ocp_qp_ipm_get_ric_Lr(struct d_ocp_qp* qp, struct d_ocp_qp_ipm_arg* arg, struct d_ocp_qp_ipm_ws* ws, int stage, double* Lr);
void d_ocp_qp_ipm_get_ric_P(struct d_ocp_qp* qp, struct d_ocp_qp_ipm_arg* arg, struct d_ocp_qp_ipm_ws* ws, int stage, double* P);
void d_ocp_qp_ipm_get_ric_p(struct d_ocp_qp* qp, struct d_ocp_qp_ipm_arg* arg, struct d_ocp_qp_ipm_ws* ws, int stage, double* p);
void d_ocp_qp_ipm_get_ric_K(struct d_ocp_qp* qp, struct d_ocp_qp_ipm_arg* arg, struct d_ocp_qp_ipm_ws* ws, int stage, double* K);
void d_ocp_qp_ipm_get_ric_k(struct d_ocp_qp* qp, struct d_ocp_qp_ipm_arg* arg, struct d_ocp_qp_ipm_ws* ws, int stage, double* k);

/* ---- not part of HPIPM -------------------------------------------------------------------------------------------------- */
/* why the last d_ocp_qp_ipm_solve reported status 4 ("" if it did not); valid until the next call */
const char* hpipm_b200_last_error(void);
/* number of device contexts the library holds (tests: a second solver of the same shape must not add one) */
int hpipm_b200_pool_size(void);

#ifdef __cplusplus
}
#endif
#endif /* HPIPM_B200_COMPAT_H_ */
