// test_hpipm_compat.cpp -- include/hpipm_b200_compat.h ("Option A" of SURVEY.md 8(b)): the HPIPM C symbols hpipm-cpp links
// against, driven in exactly the order the reference's wrapper drives them:
//   hpipm-cpp/src/ocp_qp_ipm_solver.cpp:80-116  (settings: set_default(mode), then the public fields),
//   :120-146 (dimensions: set_all, then nx[0] = nbx[0] = nsbx[0] = 0; memsize / create of qp, sol, arg, ws),
//   :225-321 (b0 = A0 x0 + b0, r0 = S0 x0 + r0, set_all, the six mask setters),
//   :323-345 (warm start, solve, getters), :346-373 (stage 0 reconstructed from Lr0), :375-414 (statistics, status).
// ReplayedSolver below is that sequence without Eigen.  It is checked against the reference's golden vectors
// (test/ocp_qp_ipm_solver.cpp:298-314) and, field by field, against the C++ facade of this repository (Option B), which
// reaches the same kernels through the batch C-ABI.  Needs a CUDA device (run by pytest -m gpu).
// usage: test_hpipm_compat <tests/golden/quadcopter_sol.txt>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

// Two builds of this file:
//  * default: against include/hpipm_b200_compat.h, the five HPIPM objects in blocks this file allocates;
//  * -DSRBD_TEST_REFERENCE_WRAPPERS (the `ref` recipe of the test infrastructure; only where the reference tree exists): against the
//    reference's VENDORED HPIPM headers, with the five objects owned by the reference's OWN memory-management classes
//    hpipm::d_ocp_qp_{dim,,sol,ipm_arg,ipm_ws}_wrapper, whose unmodified sources (hpipm-cpp/src/detail/*.cpp) are
//    compiled from where they lie and linked against libsrbd_b200.so: reference object code running on this library.
#ifdef SRBD_TEST_REFERENCE_WRAPPERS
#include "hpipm-cpp/detail/d_ocp_qp_dim_wrapper.hpp"
#include "hpipm-cpp/detail/d_ocp_qp_ipm_arg_wrapper.hpp"
#include "hpipm-cpp/detail/d_ocp_qp_ipm_ws_wrapper.hpp"
#include "hpipm-cpp/detail/d_ocp_qp_sol_wrapper.hpp"
#include "hpipm-cpp/detail/d_ocp_qp_wrapper.hpp"
extern "C" {
const char* hpipm_b200_last_error(void);
int hpipm_b200_pool_size(void);
}
#else
#include "../../../include/hpipm_b200_compat.h"
#endif
#include "../hpipm-cpp/hpipm-cpp.hpp"

using Eigen::MatrixXd;
using Eigen::VectorXd;
static int g_fail = 0;
#define CHECK(c) do { if (!(c)) { std::printf("CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); ++g_fail; } } while (0)

static uint64_t g_seed = 0x9E3779B97F4A7C15ull;
static double rnd() { g_seed ^= g_seed << 13; g_seed ^= g_seed >> 7; g_seed ^= g_seed << 17; return (double)(g_seed >> 11) / 9007199254740992.0 * 2.0 - 1.0; }
static MatrixXd Rnd(int r, int c) { MatrixXd m(r, c); for (int i = 0; i < r * c; ++i) m.data()[i] = rnd(); return m; }
static VectorXd RndV(int n) { VectorXd v(n); for (int i = 0; i < n; ++i) v(i) = rnd(); return v; }
static MatrixXd mul(const MatrixXd& A, const MatrixXd& B) {
  MatrixXd C(A.rows(), B.cols());
  for (int i = 0; i < A.rows(); ++i) for (int j = 0; j < B.cols(); ++j) { double s = 0; for (int k = 0; k < A.cols(); ++k) s += A(i, k) * B(k, j); C(i, j) = s; }
  return C;
}
static MatrixXd tr(const MatrixXd& A) { MatrixXd T(A.cols(), A.rows()); for (int i = 0; i < A.rows(); ++i) for (int j = 0; j < A.cols(); ++j) T(j, i) = A(i, j); return T; }
static VectorXd mv(const MatrixXd& A, const VectorXd& x) { VectorXd y(A.rows()); for (int i = 0; i < A.rows(); ++i) { double s = 0; for (int j = 0; j < A.cols(); ++j) s += A(i, j) * x(j); y(i) = s; } return y; }
static double nrm(const double* a, int n) { double s = 0; for (int i = 0; i < n; ++i) s += a[i] * a[i]; return std::sqrt(s); }
static bool approx(const double* a, const double* b, int n, double prec) {  // Eigen isApprox
  double d = 0; for (int i = 0; i < n; ++i) d += (a[i] - b[i]) * (a[i] - b[i]);
  return std::sqrt(d) <= prec * std::min(nrm(a, n), nrm(b, n)) || std::sqrt(d) <= 1e-300;
}

// ---- the reference wrapper's call sequence ---------------------------------------------------------------------------------
struct ReplayedSolver {
  hpipm::OcpQpIpmSolverSettings settings;
  hpipm::OcpQpDim dim;
#ifdef SRBD_TEST_REFERENCE_WRAPPERS
  // the reference's wrapper_holder_ (ocp_qp_ipm_solver.cpp:15-37): dim and arg shared, qp / sol / ws by value
  std::shared_ptr<hpipm::d_ocp_qp_dim_wrapper> dim_w = std::make_shared<hpipm::d_ocp_qp_dim_wrapper>();
  std::shared_ptr<hpipm::d_ocp_qp_ipm_arg_wrapper> arg_w = std::make_shared<hpipm::d_ocp_qp_ipm_arg_wrapper>();
  hpipm::d_ocp_qp_wrapper qp_w;
  hpipm::d_ocp_qp_sol_wrapper sol_w;
  hpipm::d_ocp_qp_ipm_ws_wrapper ws_w;
#else
  d_ocp_qp_dim dim_h{};
  d_ocp_qp qp_h{};
  d_ocp_qp_sol sol_h{};
  d_ocp_qp_ipm_arg arg_h{};
  d_ocp_qp_ipm_ws ws_h{};
  std::vector<char> dim_mem, qp_mem, sol_mem, arg_mem, ws_mem;
#endif
  d_ocp_qp_dim* dim_p = nullptr;
  d_ocp_qp* qp_p = nullptr;
  d_ocp_qp_sol* sol_p = nullptr;
  d_ocp_qp_ipm_arg* arg_p = nullptr;
  d_ocp_qp_ipm_ws* ws_p = nullptr;
  hpipm::OcpQpIpmSolverStatistics stats;

  explicit ReplayedSolver(const hpipm::OcpQpIpmSolverSettings& s) : settings(s) {}

  void applySettings() {   // ocp_qp_ipm_solver.cpp:83-116
    const hpipm_mode m = settings.mode == hpipm::HpipmMode::SpeedAbs ? SPEED_ABS
                         : settings.mode == hpipm::HpipmMode::Balance ? BALANCE
                         : settings.mode == hpipm::HpipmMode::Robust ? ROBUST : SPEED;
    d_ocp_qp_ipm_arg_set_default(m, arg_p);
    d_ocp_qp_ipm_arg_set_mu0(&settings.mu0, arg_p);
    d_ocp_qp_ipm_arg_set_iter_max(&settings.iter_max, arg_p);
    d_ocp_qp_ipm_arg_set_alpha_min(&settings.alpha_min, arg_p);
    d_ocp_qp_ipm_arg_set_tol_stat(&settings.tol_stat, arg_p);
    d_ocp_qp_ipm_arg_set_tol_eq(&settings.tol_eq, arg_p);
    d_ocp_qp_ipm_arg_set_tol_ineq(&settings.tol_ineq, arg_p);
    d_ocp_qp_ipm_arg_set_tol_comp(&settings.tol_comp, arg_p);
    d_ocp_qp_ipm_arg_set_reg_prim(&settings.reg_prim, arg_p);
    d_ocp_qp_ipm_arg_set_warm_start(&settings.warm_start, arg_p);
    d_ocp_qp_ipm_arg_set_pred_corr(&settings.pred_corr, arg_p);
    d_ocp_qp_ipm_arg_set_ric_alg(&settings.ric_alg, arg_p);
    d_ocp_qp_ipm_arg_set_split_step(&settings.split_step, arg_p);
  }

  void resize(const std::vector<hpipm::OcpQp>& qp) {   // :120-146 + detail/*_wrapper.cpp (memsize -> malloc -> create)
    dim.resize(qp);
    const int N = (int)dim.N;
#ifdef SRBD_TEST_REFERENCE_WRAPPERS
    dim_w->resize(N);
    dim_p = dim_w->get();
    arg_p = arg_w->get();
#else
    dim_mem.assign(d_ocp_qp_dim_memsize(N), 0);
    d_ocp_qp_dim_create(N, &dim_h, dim_mem.data());
    dim_p = &dim_h;
    arg_mem.assign(d_ocp_qp_ipm_arg_memsize(dim_p), 0);
    arg_p = &arg_h;
    d_ocp_qp_ipm_arg_create(dim_p, arg_p, arg_mem.data());
#endif
    d_ocp_qp_dim_set_all(dim.nx.data(), dim.nu.data(), dim.nbx.data(), dim.nbu.data(), dim.ng.data(), dim.nsbx.data(),
                         dim.nsbu.data(), dim.nsg.data(), dim_p);
    d_ocp_qp_dim_set_nx(0, 0, dim_p);
    d_ocp_qp_dim_set_nbx(0, 0, dim_p);
    d_ocp_qp_dim_set_nsbx(0, 0, dim_p);
    applySettings();
#ifdef SRBD_TEST_REFERENCE_WRAPPERS
    qp_w.resize(dim_w);
    sol_w.resize(dim_w);
    ws_w.resize(dim_w, arg_w);
    qp_p = qp_w.get(); sol_p = sol_w.get(); ws_p = ws_w.get();
#else
    qp_p = &qp_h; sol_p = &sol_h; ws_p = &ws_h;
    qp_mem.assign(d_ocp_qp_memsize(dim_p), 0);
    d_ocp_qp_create(dim_p, qp_p, qp_mem.data());
    sol_mem.assign(d_ocp_qp_sol_memsize(dim_p), 0);
    d_ocp_qp_sol_create(dim_p, sol_p, sol_mem.data());
    ws_mem.assign(d_ocp_qp_ipm_ws_memsize(dim_p, arg_p), 0);
    d_ocp_qp_ipm_ws_create(dim_p, arg_p, ws_p, ws_mem.data());
#endif
  }

  hpipm::HpipmStatus solve(const VectorXd& x0, std::vector<hpipm::OcpQp>& qp, std::vector<hpipm::OcpQpSolution>& sol) {
    resize(qp);
    const int N = (int)dim.N;
    if ((int)sol.size() != N + 1) sol.resize(N + 1);
    for (int i = 0; i <= N; ++i) {
      if (!settings.warm_start) { sol[i].x.resize(dim.nx[i]); if (i < N) sol[i].u.resize(dim.nu[i]); }
      sol[i].pi.resize(dim.nx[i]); sol[i].P.resize(dim.nx[i], dim.nx[i]); sol[i].p.resize(dim.nx[i]);
      if (i < N) sol[i].K.resize(dim.nu[i], dim.nx[i]);
      sol[i].k.resize(dim.nu[i]);
    }
    // :225-281
    VectorXd b0 = mv(qp[0].A, x0), r0(dim.nu[0]);
    for (int i = 0; i < b0.size(); ++i) b0(i) += qp[0].b(i);
    if (qp[0].S.size()) r0 = mv(qp[0].S, x0);
    for (int i = 0; i < r0.size(); ++i) r0(i) += qp[0].r(i);
    std::vector<double*> A(N + 1), B(N + 1), b(N + 1), Q(N + 1), S(N + 1), R(N + 1), q(N + 1), r(N + 1), lbx(N + 1), ubx(N + 1),
        lbu(N + 1), ubu(N + 1), C(N + 1), D(N + 1), lg(N + 1), ug(N + 1), Zl(N + 1), Zu(N + 1), zl(N + 1), zu(N + 1), lls(N + 1), lus(N + 1);
    std::vector<int*> idxbx(N + 1), idxbu(N + 1), idxs(N + 1);
    for (int i = 0; i <= N; ++i) {
      if (i < N) {
        A[i] = qp[i].A.data(); B[i] = qp[i].B.data(); b[i] = i == 0 ? b0.data() : qp[i].b.data();
        S[i] = qp[i].S.data(); R[i] = qp[i].R.data(); r[i] = i == 0 ? r0.data() : qp[i].r.data();
        idxbu[i] = qp[i].idxbu.data(); lbu[i] = qp[i].lbu.data(); ubu[i] = qp[i].ubu.data();
        D[i] = qp[i].D.data();
      }
      Q[i] = qp[i].Q.data(); q[i] = qp[i].q.data();
      idxbx[i] = qp[i].idxbx.data(); lbx[i] = qp[i].lbx.data(); ubx[i] = qp[i].ubx.data();
      C[i] = qp[i].C.data(); lg[i] = qp[i].lg.data(); ug[i] = qp[i].ug.data();
      Zl[i] = qp[i].Zl.data(); Zu[i] = qp[i].Zu.data(); zl[i] = qp[i].zl.data(); zu[i] = qp[i].zu.data();
      idxs[i] = qp[i].idxs.data(); lls[i] = qp[i].lls.data(); lus[i] = qp[i].lus.data();
    }
    d_ocp_qp_set_all(A.data(), B.data(), b.data(), Q.data(), S.data(), R.data(), q.data(), r.data(), idxbx.data(), lbx.data(),
                     ubx.data(), idxbu.data(), lbu.data(), ubu.data(), C.data(), D.data(), lg.data(), ug.data(), Zl.data(),
                     Zu.data(), zl.data(), zu.data(), idxs.data(), lls.data(), lus.data(), qp_p);
    // :291-321
    for (int i = 1; i <= N; ++i) {
      if (qp[i].lbx_mask.size() == dim.nbx[i]) d_ocp_qp_set_lbx_mask(i, qp[i].lbx_mask.data(), qp_p);
      if (qp[i].ubx_mask.size() == dim.nbx[i]) d_ocp_qp_set_ubx_mask(i, qp[i].ubx_mask.data(), qp_p);
    }
    for (int i = 0; i < N; ++i) {
      if (qp[i].lbu_mask.size() == dim.nbu[i]) d_ocp_qp_set_lbu_mask(i, qp[i].lbu_mask.data(), qp_p);
      if (qp[i].ubu_mask.size() == dim.nbu[i]) d_ocp_qp_set_ubu_mask(i, qp[i].ubu_mask.data(), qp_p);
    }
    for (int i = 0; i <= N; ++i) {
      if (qp[i].lg_mask.size() == dim.ng[i]) d_ocp_qp_set_lg_mask(i, qp[i].lg_mask.data(), qp_p);
      if (qp[i].ug_mask.size() == dim.ng[i]) d_ocp_qp_set_ug_mask(i, qp[i].ug_mask.data(), qp_p);
    }
    // :323-345
    if (settings.warm_start)
      for (int i = 0; i < N; ++i) {
        d_ocp_qp_sol_set_x(i + 1, sol[i + 1].x.data(), sol_p);
        d_ocp_qp_sol_set_u(i, sol[i].u.data(), sol_p);
      }
    d_ocp_qp_ipm_solve(qp_p, sol_p, arg_p, ws_p);
    sol[0].x = x0;
    for (int i = 0; i < N; ++i) {
      d_ocp_qp_sol_get_x(i + 1, sol_p, sol[i + 1].x.data());
      d_ocp_qp_sol_get_u(i, sol_p, sol[i].u.data());
      d_ocp_qp_sol_get_pi(i, sol_p, sol[i + 1].pi.data());
      d_ocp_qp_ipm_get_ric_P(qp_p, arg_p, ws_p, i + 1, sol[i + 1].P.data());
      d_ocp_qp_ipm_get_ric_p(qp_p, arg_p, ws_p, i + 1, sol[i + 1].p.data());
      d_ocp_qp_ipm_get_ric_K(qp_p, arg_p, ws_p, i, sol[i].K.data());
      d_ocp_qp_ipm_get_ric_k(qp_p, arg_p, ws_p, i, sol[i].k.data());
    }
    // :346-373: stage 0 from Lr0 (Lr0^-1 by forward substitution, G0^-1 = Lr0^-T Lr0^-1)
    const int nu = dim.nu[0], nx = dim.nx[0];
    MatrixXd Lr0(nu, nu), Li(nu, nu);
    d_ocp_qp_ipm_get_ric_Lr(qp_p, arg_p, ws_p, 0, Lr0.data());
    for (int c = 0; c < nu; ++c)
      for (int i = c; i < nu; ++i) {
        double s = i == c ? 1.0 : 0.0;
        for (int m = c; m < i; ++m) s -= Lr0(i, m) * Li(m, c);
        Li(i, c) = s / Lr0(i, i);
      }
    const MatrixXd G0inv = mul(tr(Li), Li);
    MatrixXd H0 = qp[0].S.size() ? qp[0].S : MatrixXd(nu, nx);
    const MatrixXd BtP = mul(tr(qp[0].B), sol[1].P), H0b = mul(BtP, qp[0].A);
    for (int i = 0; i < H0.size(); ++i) H0.data()[i] += H0b.data()[i];
    const MatrixXd GH = mul(G0inv, H0);
    sol[0].K.resize(nu, nx);
    for (int i = 0; i < GH.size(); ++i) sol[0].K.data()[i] = -GH.data()[i];
    const VectorXd Kx = mv(sol[0].K, sol[0].x);
    sol[0].k = sol[0].u;
    for (int i = 0; i < nu; ++i) sol[0].k(i) -= Kx(i);
    const MatrixXd AtP = mul(tr(qp[0].A), sol[1].P), AtPA = mul(AtP, qp[0].A), HtGH = mul(tr(H0), GH);
    sol[0].P = qp[0].Q;
    for (int i = 0; i < nx * nx; ++i) sol[0].P.data()[i] += AtPA.data()[i] - HtGH.data()[i];
    const VectorXd t1 = mv(tr(qp[0].A), sol[1].p), t2 = mv(AtP, qp[0].b), t3 = mv(tr(H0), sol[0].k), Px = mv(sol[0].P, sol[0].x);
    sol[0].p = qp[0].q;
    for (int i = 0; i < nx; ++i) sol[0].p(i) += t1(i) + t2(i) + t3(i);
    sol[0].pi = sol[0].p;
    for (int i = 0; i < nx; ++i) sol[0].pi(i) += Px(i);
    // :375-414
    d_ocp_qp_ipm_get_iter(ws_p, &stats.iter);
    d_ocp_qp_ipm_get_max_res_stat(ws_p, &stats.max_res_stat);
    d_ocp_qp_ipm_get_max_res_eq(ws_p, &stats.max_res_eq);
    d_ocp_qp_ipm_get_max_res_ineq(ws_p, &stats.max_res_ineq);
    d_ocp_qp_ipm_get_max_res_comp(ws_p, &stats.max_res_comp);
    stats.clear();
    const int stat_m = 18;
    auto cols = stats.columns();
    for (int i = 0; i <= stats.iter + 1; ++i)
      for (int c = 0; c < stat_m; ++c) cols[c]->push_back(ws_p->stat[stat_m * i + c]);
    int st;
    d_ocp_qp_ipm_get_status(ws_p, &st);
    return 0 <= st && st <= 3 ? static_cast<hpipm::HpipmStatus>(st) : hpipm::HpipmStatus::UnknownFailure;
  }
};

// every field of two solutions of the same QP.  Stages >= 1 come from the same kernel outputs; stage 0 is reconstructed on
// the host here (from Lr0, like the reference) and on the device by the facade.  p_rest: Eigen-isApprox precision asked of
// the stages >= 1 (1e-12 when both paths see bit-identical data, i.e. x0 = 0: the facade embeds x0 on the device, the
// wrapper sequence on the host, and A0 x0 + b0 may round differently), stage 0: 1e-9 like the reference's own comparison.
static double relerr(const double* a, const double* b, int n) {
  double d = 0; for (int i = 0; i < n; ++i) d += (a[i] - b[i]) * (a[i] - b[i]);
  const double m = std::min(nrm(a, n), nrm(b, n));
  return d == 0.0 ? 0.0 : std::sqrt(d) / (m > 0 ? m : 1e-300);
}
static void compare(const std::vector<hpipm::OcpQpSolution>& a, const std::vector<hpipm::OcpQpSolution>& b, int N, const char* what,
                    double p_rest = 1e-12, double p0 = 1e-9) {
  int bad = 0;
  double worst0 = 0, worst = 0;
  auto one = [&](const double* x, const double* y, std::ptrdiff_t n, int stage) {
    const double e = relerr(x, y, (int)n);
    (stage == 0 ? worst0 : worst) = std::max(stage == 0 ? worst0 : worst, e);
    bad += e > (stage == 0 ? p0 : p_rest);
  };
  for (int i = 0; i <= N; ++i) {
    CHECK(a[i].x.size() == b[i].x.size() && a[i].P.size() == b[i].P.size() && a[i].K.size() == b[i].K.size());
    one(a[i].x.data(), b[i].x.data(), a[i].x.size(), i); one(a[i].pi.data(), b[i].pi.data(), a[i].pi.size(), i);
    one(a[i].P.data(), b[i].P.data(), a[i].P.size(), i); one(a[i].p.data(), b[i].p.data(), a[i].p.size(), i);
    if (i < N) {
      one(a[i].u.data(), b[i].u.data(), a[i].u.size(), i); one(a[i].K.data(), b[i].K.data(), a[i].K.size(), i);
      one(a[i].k.data(), b[i].k.data(), a[i].k.size(), i);
    }
  }
  std::printf("%s: HPIPM-symbol path vs facade, largest relative difference %.2e (stage 0: %.2e)%s\n", what, worst, worst0,
              bad ? "  <-- beyond the tolerance" : "");
  CHECK(bad == 0);
}

static std::vector<hpipm::OcpQp> quadcopterQp(int N);   // (the problem of hpipm-cpp/test/ocp_qp_ipm_solver.cpp:170-283)

static void test_golden(const std::string& golden) {
  const int N = 10;
  std::vector<hpipm::OcpQp> qp = quadcopterQp(N);
  const MatrixXd A = qp[0].A, B = qp[0].B;
  const double u0 = 10.5916;
  hpipm::OcpQpIpmSolverSettings s;
  s.mode = hpipm::HpipmMode::Balance; s.iter_max = 30; s.alpha_min = 1e-8; s.mu0 = 1e2;
  s.tol_stat = s.tol_eq = s.tol_ineq = s.tol_comp = 1e-10;
  s.reg_prim = 1e-12; s.warm_start = 1; s.pred_corr = 1; s.ric_alg = 0; s.split_step = 1;
  std::vector<hpipm::OcpQpSolution> sol(N + 1), solf(N + 1);
  VectorXd x(12);
  for (int i = 0; i <= N; ++i) { sol[i].x = x; if (i < N) { sol[i].u = VectorXd(4); sol[i].u.fill(u0); } }
  solf = sol;
  hpipm::OcpQpIpmSolver facade(qp, s);
  std::ifstream in(golden);
  CHECK(in.good());
  for (int t = 0; t < 15; ++t) {
    ReplayedSolver solver(s);   // a NEW solver object per solve, like NMPC_solver.cpp:319
    const VectorXd x0 = x;
    CHECK(solver.solve(x0, qp, sol) == hpipm::HpipmStatus::Success);
    CHECK(facade.solve(x0, qp, solf) == hpipm::HpipmStatus::Success);
    std::vector<double> cat, gold(172);
    for (int i = 0; i <= N; ++i) for (int k = 0; k < 12; ++k) cat.push_back(sol[i].x(k));
    for (int i = 0; i < N; ++i) for (int k = 0; k < 4; ++k) cat.push_back(sol[i].u(k));
    for (double& v : gold) in >> v;
    CHECK(approx(cat.data(), gold.data(), 172, 1.0e-09));
    compare(sol, solf, N, "golden MPC step", 1e-9);
    const auto& fs = facade.getSolverStatistics();
    CHECK(solver.stats.iter == fs.iter && solver.stats.mu.size() == fs.mu.size());
    // (the two paths embed x0 in different places -- host vs device --, so b0 may differ in the last bit: close, not equal)
    for (size_t i = 0; i < fs.mu.size() && i < solver.stats.mu.size(); ++i)
      CHECK(std::fabs(solver.stats.mu[i] - fs.mu[i]) <= 1e-6 * std::fabs(fs.mu[i]) + 1e-300 &&
            std::fabs(solver.stats.alpha_prim[i] - fs.alpha_prim[i]) <= 1e-6);
    x = mv(A, x);
    const VectorXd Bu = mv(B, sol[0].u);
    for (int i = 0; i < 12; ++i) x(i) += Bu(i);
  }
  CHECK(hpipm_b200_pool_size() == 1);   // fifteen solver objects of one shape: one device context
  std::printf("golden vectors through the HPIPM symbols: done\n");
}

static std::vector<hpipm::OcpQp> randomQp(int nx, int nu, int N, double a_scale) {
  std::vector<hpipm::OcpQp> qp(N + 1);
  for (int i = 0; i < N; ++i) {
    qp[i].A = Rnd(nx, nx); for (int k = 0; k < nx * nx; ++k) qp[i].A.data()[k] *= a_scale;
    qp[i].B = Rnd(nx, nu); qp[i].b = RndV(nx);
    const MatrixXd H = Rnd(nx + nu, nx + nu), HH = mul(H, tr(H));
    qp[i].Q = MatrixXd(nx, nx); qp[i].S = MatrixXd(nu, nx); qp[i].R = MatrixXd(nu, nu);
    for (int a = 0; a < nx; ++a) for (int c = 0; c < nx; ++c) qp[i].Q(a, c) = HH(nu + a, nu + c);
    for (int a = 0; a < nu; ++a) for (int c = 0; c < nx; ++c) qp[i].S(a, c) = HH(a, nu + c);
    for (int a = 0; a < nu; ++a) for (int c = 0; c < nu; ++c) qp[i].R(a, c) = HH(a, c) + (a == c ? 0.5 : 0.0);
    qp[i].q = RndV(nx); qp[i].r = RndV(nu);
  }
  const MatrixXd H = Rnd(nx, nx);
  qp[N].Q = mul(H, tr(H)); qp[N].q = RndV(nx);
  return qp;
}

static void test_random(bool constrained) {
  const int nx = 5, nu = 3, ng = 2, N = 20;
  auto qp = randomQp(nx, nu, N, constrained ? 0.4 : 1.0);
  const VectorXd x0 = RndV(nx);
  auto absv = [](VectorXd v, double s, double off) { for (int i = 0; i < v.size(); ++i) v(i) = s * (off + std::fabs(v(i))); return v; };
  if (constrained) {
    for (int i = 0; i < N; ++i) {
      qp[i].idxbu = {0, 1, 2};
      qp[i].lbu = absv(RndV(3), -1.0, 0.5); qp[i].ubu = absv(RndV(3), 1.0, 0.5);
      qp[i].C = Rnd(ng, nx); qp[i].D = Rnd(ng, nu);
      qp[i].lg = absv(RndV(ng), -10.0, 0.5); qp[i].ug = absv(RndV(ng), 10.0, 0.5);
      if (i % 3 == 0) { qp[i].ug_mask = VectorXd(ng); qp[i].ug_mask(0) = 1.0; }   // masks ug[1]
    }
    for (int i = 1; i <= N; ++i) {
      qp[i].idxbx = {1, 3};
      qp[i].lbx = absv(RndV(2), -10.0, 0.5); qp[i].ubx = absv(RndV(2), 10.0, 0.5);
      qp[i].lbx(0) += x0(1); qp[i].lbx(1) += x0(3); qp[i].ubx(0) += x0(1); qp[i].ubx(1) += x0(3);
    }
    qp[N].C = Rnd(ng, nx); qp[N].lg = absv(RndV(ng), -10.0, 0.5); qp[N].ug = absv(RndV(ng), 10.0, 0.5);
  }
  hpipm::OcpQpIpmSolverSettings s;
  if (constrained) { s.ric_alg = 0; s.iter_max = 40; s.tol_stat = 1e-6; }
  std::vector<hpipm::OcpQpSolution> sol(N + 1), solf(N + 1);
  ReplayedSolver solver(s);
  hpipm::OcpQpIpmSolver facade(qp, s);
  CHECK(solver.solve(x0, qp, sol) == hpipm::HpipmStatus::Success);
  CHECK(facade.solve(x0, qp, solf) == hpipm::HpipmStatus::Success);
  CHECK(constrained ? solver.stats.iter > 0 : solver.stats.iter == 0);   // test/ocp_qp_ipm_solver.cpp:56
  CHECK(solver.stats.iter == facade.getSolverStatistics().iter);
  CHECK((int)solver.stats.mu.size() == solver.stats.iter + 2);
  compare(sol, solf, N, constrained ? "random constrained QP" : "random unconstrained QP");
  std::printf("%s QP through the HPIPM symbols: done (%d iterations)\n", constrained ? "constrained" : "unconstrained", solver.stats.iter);
}

// an SRBD-shaped QP (nx = nu = 12, 24 general rows with the structure K2 produces): srbd_qp_upload detects the structure and
// the solve runs in the tensor-core kernel, whose exporting instantiation rebuilds Lr0 from its blocked factor panels
static std::vector<hpipm::OcpQp> srbdShapedQp(int N) {
  const int nx = 12, nu = 12, ng = 24;
  std::vector<hpipm::OcpQp> qp(N + 1);
  MatrixXd D(ng, nu);
  for (int leg = 0; leg < 2; ++leg)
    for (int g = 0; g < 12; ++g)
      for (int j = 0; j < 6; ++j) D(12 * leg + g, 6 * leg + j) = rnd();
  MatrixXd Q(nx, nx);
  for (int i = 0; i < nx; ++i) Q(i, i) = 0.5 + std::fabs(rnd());
  for (int i = 0; i <= N; ++i) {
    qp[i].Q = Q; qp[i].q = RndV(nx);
    if (i == N) { for (int k = 0; k < nx; ++k) qp[i].Q(k, k) = 3.0 + k; break; }
    qp[i].A = Rnd(nx, nx); for (int k = 0; k < nx * nx; ++k) qp[i].A.data()[k] *= 0.2;
    for (int k = 0; k < nx; ++k) qp[i].A(k, k) += 1.0;
    qp[i].B = Rnd(nx, nu); qp[i].b = RndV(nx); for (int k = 0; k < nx; ++k) qp[i].b(k) *= 0.1;
    qp[i].S = MatrixXd(nu, nx);
    qp[i].R = MatrixXd(nu, nu); for (int k = 0; k < nu; ++k) qp[i].R(k, k) = 0.3;
    qp[i].r = RndV(nu);
    qp[i].C = MatrixXd(ng, nx); qp[i].D = D;
    qp[i].lg = VectorXd(ng); qp[i].ug = VectorXd(ng); qp[i].ug_mask = VectorXd(ng);   // upper side masked
    for (int g = 0; g < ng; ++g) { qp[i].lg(g) = -1.0 - std::fabs(rnd()); qp[i].ug(g) = 1e10; }
  }
  return qp;
}

static void test_srbd_shape() {
  const int N = 8, nx = 12, nu = 12;
  std::vector<hpipm::OcpQp> qp = srbdShapedQp(N);
  hpipm::OcpQpIpmSolverSettings s;
  s.ric_alg = 0; s.iter_max = 40; s.split_step = 1;
  for (int pass = 0; pass < 2; ++pass) {
    const VectorXd x0 = pass == 0 ? VectorXd(nx) : RndV(nx);   // x0 = 0: both paths see bit-identical data
    std::vector<hpipm::OcpQpSolution> sol(N + 1), solf(N + 1);
    ReplayedSolver solver(s);
    hpipm::OcpQpIpmSolver facade(qp, s);
    CHECK(solver.solve(x0, qp, sol) == hpipm::HpipmStatus::Success);
    CHECK(facade.solve(x0, qp, solf) == hpipm::HpipmStatus::Success);
    CHECK(solver.stats.iter > 0 && solver.stats.iter == facade.getSolverStatistics().iter);
    if (pass == 0) compare(sol, solf, N, "SRBD-shaped QP, x0 = 0");
    else {
      // x0 != 0: b0 = A0 x0 + b0 is rounded on the host here and on the device in the facade; the primal solution moves by
      // ~1e-12, the Riccati matrices of the LAST barrier-augmented factorization (Gamma up to 1e10) by up to ~1e-5
      double ex = 0.0;
      for (int i = 0; i <= N; ++i) {
        ex = std::max(ex, relerr(sol[i].x.data(), solf[i].x.data(), nx));
        if (i < N) ex = std::max(ex, relerr(sol[i].u.data(), solf[i].u.data(), nu));
      }
      std::printf("SRBD-shaped QP, x0 != 0: x, u of the two paths within %.2e\n", ex);
      CHECK(ex <= 1e-8);
      compare(sol, solf, N, "SRBD-shaped QP, x0 != 0 (Riccati exports: sensitivity of the last factorization)", 1e-3, 1e-3);
    }
    std::printf("SRBD-shaped QP through the HPIPM symbols: done (%d iterations)\n", solver.stats.iter);
  }
}

// host -> host latency of one solve through the HPIPM symbols (a new solver object per call, like NMPC_solver.cpp:319)
static void report_latency() {
  for (int which = 0; which < 2; ++which) {
    const int N = which == 0 ? 10 : 20;
    std::vector<hpipm::OcpQp> qp = which == 0 ? quadcopterQp(N) : srbdShapedQp(N);
    hpipm::OcpQpIpmSolverSettings s;
    s.ric_alg = 0; s.iter_max = 40; s.split_step = 1;
    const VectorXd x0(12);
    std::vector<hpipm::OcpQpSolution> sol(N + 1);
    std::vector<double> ms;
    int iters = 0;
    for (int rep = 0; rep < 60; ++rep) {
      const auto t0 = std::chrono::steady_clock::now();
      ReplayedSolver solver(s);
      CHECK(solver.solve(x0, qp, sol) == hpipm::HpipmStatus::Success);
      const auto t1 = std::chrono::steady_clock::now();
      if (rep >= 10) ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
      iters = solver.stats.iter;
    }
    std::sort(ms.begin(), ms.end());
    std::printf("latency through the HPIPM symbols, %s N = %d (%d iterations, all exports): p50 %.3f ms, p90 %.3f ms\n",
                which == 0 ? "quadcopter QP" : "SRBD-shaped QP", N, iters, ms[ms.size() / 2], ms[ms.size() * 9 / 10]);
  }
}

static void test_unsupported_shape() {
  const int nx = 4, nu = 2, N = 5;
  auto qp = randomQp(nx, nu, N, 0.5);
  qp[3].idxbu = {0};   // nbu differs between the stages
  qp[3].lbu = VectorXd(1); qp[3].ubu = VectorXd(1); qp[3].lbu(0) = -1.0; qp[3].ubu(0) = 1.0;
  hpipm::OcpQpIpmSolverSettings s;
  std::vector<hpipm::OcpQpSolution> sol(N + 1);
  ReplayedSolver solver(s);
  const int pool0 = hpipm_b200_pool_size();
  CHECK(solver.solve(RndV(nx), qp, sol) == hpipm::HpipmStatus::UnknownFailure);
  CHECK(std::string(hpipm_b200_last_error()).find("unsupported QP shape") != std::string::npos);
  CHECK(hpipm_b200_pool_size() == pool0);
  std::printf("unsupported shape: status 4, \"%s\"\n", hpipm_b200_last_error());
}

int main(int argc, char** argv) {
  const std::string golden = argc > 1 ? argv[1] : "tests/golden/quadcopter_sol.txt";
  test_golden(golden);
  test_random(false);
  test_random(true);
  test_srbd_shape();
  test_unsupported_shape();
  report_latency();
  if (g_fail) { std::printf("FAILED: %d checks\n", g_fail); return 1; }
  std::printf("ALL OK\n");
  return 0;
}

static std::vector<hpipm::OcpQp> quadcopterQp(int N) {
  std::vector<hpipm::OcpQp> qp(N + 1);
  const double Ad[12][12] = {
      {1., 0., 0., 0., 0., 0., 0.1, 0., 0., 0., 0., 0.}, {0., 1., 0., 0., 0., 0., 0., 0.1, 0., 0., 0., 0.},
      {0., 0., 1., 0., 0., 0., 0., 0., 0.1, 0., 0., 0.}, {0.0488, 0., 0., 1., 0., 0., 0.0016, 0., 0., 0.0992, 0., 0.},
      {0., -0.0488, 0., 0., 1., 0., 0., -0.0016, 0., 0., 0.0992, 0.}, {0., 0., 0., 0., 0., 1., 0., 0., 0., 0., 0., 0.0992},
      {0., 0., 0., 0., 0., 0., 1., 0., 0., 0., 0., 0.}, {0., 0., 0., 0., 0., 0., 0., 1., 0., 0., 0., 0.},
      {0., 0., 0., 0., 0., 0., 0., 0., 1., 0., 0., 0.}, {0.9734, 0., 0., 0., 0., 0., 0.0488, 0., 0., 0.9846, 0., 0.},
      {0., -0.9734, 0., 0., 0., 0., 0., -0.0488, 0., 0., 0.9846, 0.}, {0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0.9846}};
  const double Bd[12][4] = {{0., -0.0726, 0., 0.0726}, {-0.0726, 0., 0.0726, 0.}, {-0.0152, 0.0152, -0.0152, 0.0152},
                            {-0., -0.0006, -0., 0.0006}, {0.0006, 0., -0.0006, 0.0000}, {0.0106, 0.0106, 0.0106, 0.0106},
                            {0, -1.4512, 0., 1.4512}, {-1.4512, 0., 1.4512, 0.}, {-0.3049, 0.3049, -0.3049, 0.3049},
                            {-0., -0.0236, 0., 0.0236}, {0.0236, 0., -0.0236, 0.}, {0.2107, 0.2107, 0.2107, 0.2107}};
  MatrixXd A(12, 12), B(12, 4), Q(12, 12), S(4, 12), R(4, 4);
  for (int i = 0; i < 12; ++i) { for (int j = 0; j < 12; ++j) A(i, j) = Ad[i][j]; for (int j = 0; j < 4; ++j) B(i, j) = Bd[i][j]; }
  const double qd[12] = {0, 0, 10., 10., 10., 10., 0, 0, 0, 5., 5., 5.};
  for (int i = 0; i < 12; ++i) Q(i, i) = qd[i];
  for (int i = 0; i < 4; ++i) R(i, i) = 0.1;
  VectorXd q(12); q(2) = -10.0;
  const double u0 = 10.5916, PI6 = M_PI / 6.0;
  for (int i = 0; i <= N; ++i) {
    qp[i].Q = Q; qp[i].q = q;
    if (i < N) { qp[i].A = A; qp[i].B = B; qp[i].b = VectorXd(12); qp[i].R = R; qp[i].S = S; qp[i].r = VectorXd(4); }
    if (i >= 1) {
      qp[i].idxbx = {0, 1, 5};
      qp[i].lbx = VectorXd(3); qp[i].ubx = VectorXd(3); qp[i].ubx_mask = VectorXd(3);
      qp[i].lbx(0) = -PI6; qp[i].lbx(1) = -PI6; qp[i].lbx(2) = -1.0;
      qp[i].ubx(0) = PI6; qp[i].ubx(1) = PI6; qp[i].ubx(2) = 1.0e10;
      qp[i].ubx_mask(0) = 1.0; qp[i].ubx_mask(1) = 1.0; qp[i].ubx_mask(2) = 0.0;
    }
    if (i < N) {
      qp[i].idxbu = {0, 1, 2, 3};
      qp[i].lbu = VectorXd(4); qp[i].ubu = VectorXd(4);
      qp[i].lbu.fill(9.6 - u0); qp[i].ubu.fill(13.0 - u0);
    }
  }
  return qp;
}
