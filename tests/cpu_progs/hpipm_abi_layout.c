/* Prints sizeof / offsetof of the HPIPM structs as include/hpipm_b200_compat.h declares them.  The expected values in
 * tests/test_capi_cpu.py::test_hpipm_struct_layout_matches_the_vendored_headers were produced by compiling this same
 * file against the reference's vendored headers instead
 *   (gcc -I/root/reference/hpipm-cpp/include/include -DUSE_VENDORED hpipm_abi_layout.c)
 * in the build container: hpipm-cpp reads dim->N, the scalars of d_ocp_qp_ipm_arg and ws->stat directly, so an unmodified
 * hpipm-cpp compiled against ITS headers must find them where this library puts them. */
#include <stddef.h>
#include <stdio.h>
#ifdef USE_VENDORED
#include "hpipm_d_ocp_qp_dim.h"
#include "hpipm_d_ocp_qp.h"
#include "hpipm_d_ocp_qp_sol.h"
#include "hpipm_d_ocp_qp_ipm.h"
#else
#include "hpipm_b200_compat.h"
#endif
#define O(s, m) offsetof(struct s, m)
int main(void) {
  printf("dim %zu N %zu memsize %zu\n", sizeof(struct d_ocp_qp_dim), O(d_ocp_qp_dim, N), O(d_ocp_qp_dim, memsize));
  printf("qp %zu BAbt %zu idxb %zu diag_H_flag %zu memsize %zu\n", sizeof(struct d_ocp_qp), O(d_ocp_qp, BAbt), O(d_ocp_qp, idxb),
         O(d_ocp_qp, diag_H_flag), O(d_ocp_qp, memsize));
  printf("sol %zu ux %zu misc %zu memsize %zu\n", sizeof(struct d_ocp_qp_sol), O(d_ocp_qp_sol, ux), O(d_ocp_qp_sol, misc),
         O(d_ocp_qp_sol, memsize));
  printf("arg %zu mu0 %zu tau_min %zu iter_max %zu stat_max %zu pred_corr %zu warm_start %zu square_root_alg %zu lq_fact %zu "
         "split_step %zu t_lam_min %zu mode %zu memsize %zu\n",
         sizeof(struct d_ocp_qp_ipm_arg), O(d_ocp_qp_ipm_arg, mu0), O(d_ocp_qp_ipm_arg, tau_min), O(d_ocp_qp_ipm_arg, iter_max),
         O(d_ocp_qp_ipm_arg, stat_max), O(d_ocp_qp_ipm_arg, pred_corr), O(d_ocp_qp_ipm_arg, warm_start),
         O(d_ocp_qp_ipm_arg, square_root_alg), O(d_ocp_qp_ipm_arg, lq_fact), O(d_ocp_qp_ipm_arg, split_step),
         O(d_ocp_qp_ipm_arg, t_lam_min), O(d_ocp_qp_ipm_arg, mode), O(d_ocp_qp_ipm_arg, memsize));
  printf("ws %zu qp_res %zu core_workspace %zu dim %zu stat %zu iter %zu stat_max %zu stat_m %zu status %zu valid_ric_p %zu memsize %zu\n",
         sizeof(struct d_ocp_qp_ipm_ws), O(d_ocp_qp_ipm_ws, qp_res), O(d_ocp_qp_ipm_ws, core_workspace), O(d_ocp_qp_ipm_ws, dim),
         O(d_ocp_qp_ipm_ws, stat), O(d_ocp_qp_ipm_ws, iter), O(d_ocp_qp_ipm_ws, stat_max), O(d_ocp_qp_ipm_ws, stat_m),
         O(d_ocp_qp_ipm_ws, status), O(d_ocp_qp_ipm_ws, valid_ric_p), O(d_ocp_qp_ipm_ws, memsize));
  printf("enums %d %d %d %d | %d %d %d %d %d\n", SPEED_ABS, SPEED, BALANCE, ROBUST, SUCCESS, MAX_ITER, MIN_STEP, NAN_SOL, INCONS_EQ);
  return 0;
}
