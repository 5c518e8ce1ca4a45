/*
 * pipeline.c — CPU ORACLE (test infrastructure): batched drivers over the restated hot path
 * (linearize -> assemble -> OCP-QP IPM solve), OpenMP over independent QPs.  This is what bench.py's
 * `cpu_baseline` / `--impl reference` legs time; it stands in for "reference HPIPM/BLASFEO on CPU",
 * which cannot be built offline (see srbd_oracle.h).  Follows NMPC_solver.cpp:276-330.
 */
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "srbd_oracle.h"

int orc_assemble_batch(const srbd_model_params* m, int N, int mode, int batch, const double* x,
                       const double* u, const double* xref, const uint8_t* contact, double* A,
                       double* B, double* b, double* Q, double* S, double* R, double* q, double* r,
                       double* D, double* lg, double* lg_mask, double* defect, double* fcon,
                       int threads) {
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
  for (int i = 0; i < batch; ++i) {
    const size_t s = (size_t)i;
    orc_assemble(m, N, mode, x + s * (N + 1) * 12, u + s * N * 12, xref + s * (N + 1) * 12,
                 contact ? contact + s * N * 2 : NULL, A ? A + s * N * 144 : NULL,
                 B ? B + s * N * 144 : NULL, b ? b + s * N * 12 : NULL, Q ? Q + s * (N + 1) * 144 : NULL,
                 S ? S + s * N * 144 : NULL, R ? R + s * N * 144 : NULL, q ? q + s * (N + 1) * 12 : NULL,
                 r ? r + s * N * 12 : NULL, D ? D + s * N * 288 : NULL, lg ? lg + s * N * 24 : NULL,
                 lg_mask ? lg_mask + s * N * 24 : NULL, defect ? defect + s * N * 12 : NULL,
                 fcon ? fcon + s * N * 24 : NULL);
  }
  (void)threads;
  return 0;
}

/* One NMPC QP: assemble (prepareQpStructures) then solve (solveQpProblems) with the delta initial
 * state x0 - x_nmpc[:,0] (NMPC_solver.cpp:320). */
static void pipeline_one(const srbd_model_params* m, const srbd_ipm_args* a, int N, int mode,
                         const double* x, const double* u, const double* xref, const double* x0,
                         const uint8_t* contact, double* sol_x, double* sol_u, double* sol_pi,
                         double* sol_lam, double* sol_t, int* iter, int* status, double* res_max) {
  const size_t n1 = (size_t)(N + 1), n0 = (size_t)N;
  double* buf = (double*)calloc(n0 * 144 * 4 + n1 * 144 + n0 * 12 * 2 + n1 * 12 + n0 * 288 + n0 * 24 * 3 + 12, sizeof(double));
  double* A = buf;
  double* B = A + n0 * 144;
  double* S = B + n0 * 144;
  double* R = S + n0 * 144;
  double* Q = R + n0 * 144;
  double* b = Q + n1 * 144;
  double* r = b + n0 * 12;
  double* q = r + n0 * 12;
  double* D = q + n1 * 12;
  double* lg = D + n0 * 288;
  double* lgm = lg + n0 * 24;
  double* ug = lgm + n0 * 24;
  double* dx0 = ug + n0 * 24;
  orc_assemble(m, N, mode, x, u, xref, contact, A, B, b, Q, S, R, q, r, D, lg, lgm, NULL, NULL);
  for (int i = 0; i < 12; ++i) dx0[i] = x0[i] - x[i];
  srbd_qp_dims d;
  memset(&d, 0, sizeof(d));
  d.N = N; d.nx = 12; d.nu = 12;
  srbd_qp_host qp;
  memset(&qp, 0, sizeof(qp));
  qp.A = A; qp.Bm = B; qp.b = b; qp.Q = Q; qp.S = S; qp.R = R; qp.q = q; qp.r = r; qp.x0 = dx0;
  double* ugm = NULL;
  if (mode == SRBD_HARD_INEQ) {
    d.ng = 24;
    ugm = (double*)calloc(n0 * 24, sizeof(double)); /* upper side masked */
    qp.D = D; qp.lg = lg; qp.ug = ug; qp.lg_mask = lgm; qp.ug_mask = ugm;
  }
  srbd_sol_host sol;
  memset(&sol, 0, sizeof(sol));
  sol.x = sol_x; sol.u = sol_u; sol.pi = sol_pi; sol.lam = sol_lam; sol.t = sol_t;
  srbd_stats_host st;
  memset(&st, 0, sizeof(st));
  st.iter = iter; st.status = status; st.res_max = res_max;
  orc_qp_solve_one(&d, a, &qp, &sol, &st, 0, 0);
  free(ugm);
  free(buf);
}

int orc_pipeline_batch(const srbd_model_params* m, const srbd_ipm_args* a, int N, int mode, int batch,
                       const double* x, const double* u, const double* xref, const double* x0,
                       const uint8_t* contact, double* sol_x, double* sol_u, double* sol_pi,
                       double* sol_lam, double* sol_t, int* iter, int* status, double* res_max,
                       int threads) {
  const size_t nct = (size_t)N * 48;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads)
#endif
  for (int i = 0; i < batch; ++i) {
    const size_t s = (size_t)i;
    pipeline_one(m, a, N, mode, x + s * (N + 1) * 12, u + s * N * 12, xref + s * (N + 1) * 12, x0 + s * 12,
                 contact ? contact + s * N * 2 : NULL, sol_x ? sol_x + s * (N + 1) * 12 : NULL,
                 sol_u ? sol_u + s * N * 12 : NULL, sol_pi ? sol_pi + s * (N + 1) * 12 : NULL,
                 (sol_lam && mode == SRBD_HARD_INEQ) ? sol_lam + s * nct : NULL,
                 (sol_t && mode == SRBD_HARD_INEQ) ? sol_t + s * nct : NULL, iter ? iter + s : NULL,
                 status ? status + s : NULL, res_max ? res_max + 4 * s : NULL);
  }
  (void)threads;
  return 0;
}
