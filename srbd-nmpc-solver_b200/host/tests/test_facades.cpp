// test_facades.cpp — the reference's own host-level tests, re-written against the B200-backed facades:
//   hpipm-cpp/test/ocp_qp_ipm_solver.cpp:22-110  (unconstrained: iter == 0, analytic Riccati, 1e-10)
//   hpipm-cpp/test/ocp_qp_ipm_solver.cpp:112-168 (constrained: Success, x[0] == x0)
//   hpipm-cpp/test/ocp_qp_ipm_solver.cpp:170-315 (compareResults: 15 golden vectors, isApprox 1e-9)
// plus SRBDModel / NMPCSolver drop-in checks (config 1: 11 SQP iterations) and the
// std::runtime_error behaviour of hpipm-cpp/src/ocp_qp_dim.cpp.  Needs a CUDA device (run by pytest -m gpu).
// usage: test_facades <tests/golden/quadcopter_sol.txt>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <sstream>

#include "../NMPC_solver.hpp"
#include "../SRBD_model.hpp"
#include "../hpipm-cpp/hpipm-cpp.hpp"

using Eigen::MatrixXd;
using Eigen::VectorXd;
static int g_fail = 0;
#define CHECK(c) do { if (!(c)) { std::printf("CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); ++g_fail; } } while (0)

// ---- tiny dense helpers (the shim has no arithmetic) ----------------------------------------------------
static uint64_t g_seed = 88172645463325252ull;
static double rnd() { g_seed ^= g_seed << 13; g_seed ^= g_seed >> 7; g_seed ^= g_seed << 17; return (double)(g_seed >> 11) / 9007199254740992.0 * 2.0 - 1.0; }
static MatrixXd Rnd(int r, int c) { MatrixXd m(r, c); for (int i = 0; i < r * c; ++i) m.data()[i] = rnd(); return m; }
static VectorXd RndV(int n) { VectorXd v(n); for (int i = 0; i < n; ++i) v(i) = rnd(); return v; }
static MatrixXd mul(const MatrixXd& A, const MatrixXd& B) {
  MatrixXd C(A.rows(), B.cols());
  for (int i = 0; i < A.rows(); ++i) for (int j = 0; j < B.cols(); ++j) { double s = 0; for (int k = 0; k < A.cols(); ++k) s += A(i, k) * B(k, j); C(i, j) = s; }
  return C;
}
static MatrixXd tr(const MatrixXd& A) { MatrixXd T(A.cols(), A.rows()); for (int i = 0; i < A.rows(); ++i) for (int j = 0; j < A.cols(); ++j) T(j, i) = A(i, j); return T; }
static MatrixXd add(const MatrixXd& A, const MatrixXd& B, double s = 1.0) { MatrixXd C(A.rows(), A.cols()); for (int i = 0; i < A.size(); ++i) C.data()[i] = A.data()[i] + s * B.data()[i]; return C; }
static MatrixXd col(const VectorXd& v) { MatrixXd m(v.size(), 1); for (int i = 0; i < v.size(); ++i) m(i, 0) = v(i); return m; }
static VectorXd vec(const MatrixXd& m) { VectorXd v(m.rows()); for (int i = 0; i < m.rows(); ++i) v(i) = m(i, 0); return v; }
static MatrixXd inv(MatrixXd A) {
  const int n = (int)A.rows();
  MatrixXd I = MatrixXd::Identity(n, n);
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r) if (std::fabs(A(r, c)) > std::fabs(A(piv, c))) piv = r;
    for (int j = 0; j < n; ++j) { std::swap(A(c, j), A(piv, j)); std::swap(I(c, j), I(piv, j)); }
    const double d = A(c, c);
    for (int j = 0; j < n; ++j) { A(c, j) /= d; I(c, j) /= d; }
    for (int r = 0; r < n; ++r) if (r != c) { const double f = A(r, c); for (int j = 0; j < n; ++j) { A(r, j) -= f * A(c, j); I(r, j) -= f * I(c, j); } }
  }
  return I;
}
static double nrm(const double* a, int n) { double s = 0; for (int i = 0; i < n; ++i) s += a[i] * a[i]; return std::sqrt(s); }
static bool approx(const double* a, const double* b, int n, double prec) {  // Eigen isApprox
  double d = 0; for (int i = 0; i < n; ++i) d += (a[i] - b[i]) * (a[i] - b[i]);
  return std::sqrt(d) <= prec * std::min(nrm(a, n), nrm(b, n));
}
static bool approxM(const MatrixXd& a, const MatrixXd& b, double p) { return a.size() == b.size() && approx(a.data(), b.data(), (int)a.size(), p); }
static bool approxV(const VectorXd& a, const VectorXd& b, double p) { return a.size() == b.size() && approx(a.data(), b.data(), (int)a.size(), p); }
static MatrixXd block(const MatrixXd& H, int r0, int c0, int r, int c) { MatrixXd m(r, c); for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) m(i, j) = H(r0 + i, c0 + j); return m; }

static std::vector<hpipm::OcpQp> randomQp(int nx, int nu, int N, double a_scale, bool boost_R) {
  std::vector<hpipm::OcpQp> qp(N + 1);
  for (int i = 0; i < N; ++i) {
    qp[i].A = Rnd(nx, nx); for (int k = 0; k < nx * nx; ++k) qp[i].A.data()[k] *= a_scale;
    qp[i].B = Rnd(nx, nu); qp[i].b = RndV(nx);
    const MatrixXd H = Rnd(nx + nu, nx + nu), HH = mul(H, tr(H));
    qp[i].Q = block(HH, nu, nu, nx, nx); qp[i].S = block(HH, 0, nu, nu, nx); qp[i].R = block(HH, 0, 0, nu, nu);
    if (boost_R) for (int k = 0; k < nu; ++k) qp[i].R(k, k) += std::fabs(rnd());
    qp[i].q = RndV(nx); qp[i].r = RndV(nu);
  }
  const MatrixXd H = Rnd(nx, nx);
  qp[N].Q = mul(H, tr(H)); qp[N].q = RndV(nx);
  return qp;
}

static void test_unconstrained() {
  const int nx = 5, nu = 3, N = 20;
  auto qp = randomQp(nx, nu, N, 1.0, true);
  const VectorXd x0 = RndV(nx);
  hpipm::OcpQpIpmSolverSettings s;  // defaults like the reference test: ric_alg = 1 (square-root Riccati)
  std::vector<hpipm::OcpQpSolution> sol(N + 1);
  hpipm::OcpQpIpmSolver solver(qp, s);
  const auto status = solver.solve(x0, qp, sol);
  CHECK(status == hpipm::HpipmStatus::Success);
  CHECK(solver.getSolverStatistics().iter == 0);
  CHECK(approxV(sol[0].x, x0, 1e-12));
  std::vector<MatrixXd> P(N + 1), K(N), sv(N + 1), kv(N);
  P[N] = qp[N].Q; sv[N] = col(qp[N].q); for (int i = 0; i < nx; ++i) sv[N](i, 0) = -sv[N](i, 0);
  for (int i = N - 1; i >= 0; --i) {
    const MatrixXd At = tr(qp[i].A), Bt = tr(qp[i].B);
    const MatrixXd F = add(qp[i].Q, mul(mul(At, P[i + 1]), qp[i].A));
    const MatrixXd H = add(qp[i].S, mul(mul(Bt, P[i + 1]), qp[i].A));
    const MatrixXd G = add(qp[i].R, mul(mul(Bt, P[i + 1]), qp[i].B));
    const MatrixXd Gi = inv(G);
    K[i] = mul(Gi, H); for (int k = 0; k < K[i].size(); ++k) K[i].data()[k] = -K[i].data()[k];
    MatrixXd t = add(add(mul(mul(Bt, P[i + 1]), col(qp[i].b)), mul(Bt, sv[i + 1]), -1.0), col(qp[i].r));
    kv[i] = mul(Gi, t); for (int k = 0; k < nu; ++k) kv[i](k, 0) = -kv[i](k, 0);
    P[i] = add(F, mul(mul(tr(K[i]), G), K[i]), -1.0);
    sv[i] = add(add(mul(At, add(sv[i + 1], mul(P[i + 1], col(qp[i].b)), -1.0)), col(qp[i].q), -1.0), mul(tr(H), kv[i]), -1.0);
  }
  std::vector<MatrixXd> x(N + 1), u(N);
  x[0] = col(x0);
  for (int i = 0; i < N; ++i) {
    u[i] = add(mul(K[i], x[i]), kv[i]);
    x[i + 1] = add(add(mul(qp[i].A, x[i]), mul(qp[i].B, u[i])), col(qp[i].b));
  }
  const double prec = 1e-10;
  for (int i = 0; i <= N; ++i) {
    CHECK(approxV(vec(x[i]), sol[i].x, prec));
    CHECK(approxV(vec(add(mul(P[i], x[i]), sv[i], -1.0)), sol[i].pi, prec));
    CHECK(approxM(P[i], sol[i].P, prec));
    VectorXd mp(nx); for (int k = 0; k < nx; ++k) mp(k) = -sol[i].p(k);
    CHECK(approxV(vec(sv[i]), mp, prec));
  }
  for (int i = 0; i < N; ++i) {
    CHECK(approxV(vec(u[i]), sol[i].u, prec));
    CHECK(approxM(K[i], sol[i].K, prec));
    CHECK(approxV(vec(kv[i]), sol[i].k, prec));
  }
  std::printf("unconstrained: done\n");
}

static void test_constrained() {
  const int nx = 5, nu = 3, ng = 2, N = 20;
  auto qp = randomQp(nx, nu, N, 0.4, true);
  const VectorXd x0 = RndV(nx);
  auto absv = [](VectorXd v, double s, double off) { for (int i = 0; i < v.size(); ++i) v(i) = s * (off + std::fabs(v(i))); return v; };
  for (int i = 0; i < N; ++i) {
    qp[i].idxbu = {0, 1, 2};
    qp[i].lbu = absv(RndV(3), -1.0, 0.5); qp[i].ubu = absv(RndV(3), 1.0, 0.5);
    qp[i].C = Rnd(ng, nx); qp[i].D = Rnd(ng, nu);
    qp[i].lg = absv(RndV(ng), -10.0, 0.5); qp[i].ug = absv(RndV(ng), 10.0, 0.5);
  }
  for (int i = 1; i <= N; ++i) {
    qp[i].idxbx = {1, 3};
    qp[i].lbx = absv(RndV(2), -10.0, 0.5); qp[i].ubx = absv(RndV(2), 10.0, 0.5);
    qp[i].lbx(0) += x0(1); qp[i].lbx(1) += x0(3); qp[i].ubx(0) += x0(1); qp[i].ubx(1) += x0(3);
  }
  qp[N].C = Rnd(ng, nx); qp[N].lg = absv(RndV(ng), -10.0, 0.5); qp[N].ug = absv(RndV(ng), 10.0, 0.5);
  hpipm::OcpQpIpmSolverSettings s;
  s.ric_alg = 0; s.iter_max = 40; s.tol_stat = 1e-6;
  std::vector<hpipm::OcpQpSolution> sol(N + 1);
  hpipm::OcpQpIpmSolver solver(qp, s);
  const auto status = solver.solve(x0, qp, sol);
  CHECK(status == hpipm::HpipmStatus::Success);
  CHECK(approxV(sol[0].x, x0, 1e-12));
  CHECK(solver.getSolverStatistics().iter > 0);
  CHECK((int)solver.getSolverStatistics().mu.size() == solver.getSolverStatistics().iter + 2);
  for (int i = 0; i < N; ++i) for (int k = 0; k < 3; ++k) CHECK(sol[i].u(k) >= qp[i].lbu(k) - 1e-7 && sol[i].u(k) <= qp[i].ubu(k) + 1e-7);
  std::printf("constrained: done (%d iterations)\n", solver.getSolverStatistics().iter);
}

// the quadcopter MPC problem of hpipm-cpp/test/ocp_qp_ipm_solver.cpp:170-283 (= examples/example_mpc.cpp:16-96)
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
  VectorXd q(12); q(2) = -10.0;  // -Q * x_ref, x_ref = e_2
  const double u0 = 10.5916, PI6 = M_PI / 6.0;
  for (int i = 0; i <= N; ++i) {
    qp[i].Q = Q; qp[i].q = q;
    if (i < N) { qp[i].A = A; qp[i].B = B; qp[i].b = VectorXd(12); qp[i].R = R; qp[i].S = S; qp[i].r = VectorXd(4); }
    if (i >= 1) {
      qp[i].idxbx = {0, 1, 5};
      qp[i].lbx = VectorXd(3); qp[i].ubx = VectorXd(3); qp[i].ubx_mask = VectorXd(3);
      qp[i].lbx(0) = -PI6; qp[i].lbx(1) = -PI6; qp[i].lbx(2) = -1.0;
      qp[i].ubx(0) = PI6; qp[i].ubx(1) = PI6; qp[i].ubx(2) = 1.0e10;
      qp[i].ubx_mask(0) = 1.0; qp[i].ubx_mask(1) = 1.0; qp[i].ubx_mask(2) = 0.0;  // disables ubx[2] (:232)
    }
    if (i < N) {
      qp[i].idxbu = {0, 1, 2, 3};
      qp[i].lbu = VectorXd(4); qp[i].ubu = VectorXd(4);
      qp[i].lbu.fill(9.6 - u0); qp[i].ubu.fill(13.0 - u0);
    }
  }
  return qp;
}

static void test_compare_results(const std::string& golden) {
  const int N = 10;
  std::vector<hpipm::OcpQp> qp = quadcopterQp(N);
  const MatrixXd A = qp[0].A, B = qp[0].B;
  const double u0 = 10.5916;
  hpipm::OcpQpIpmSolverSettings s;
  s.mode = hpipm::HpipmMode::Balance; s.iter_max = 30; s.alpha_min = 1e-8; s.mu0 = 1e2;
  s.tol_stat = s.tol_eq = s.tol_ineq = s.tol_comp = 1e-10;
  s.reg_prim = 1e-12; s.warm_start = 1; s.pred_corr = 1; s.ric_alg = 0; s.split_step = 1;
  std::vector<hpipm::OcpQpSolution> sol(N + 1);
  hpipm::OcpQpIpmSolver solver(qp, s);
  VectorXd x(12);
  for (int i = 0; i <= N; ++i) { sol[i].x = x; if (i < N) { sol[i].u = VectorXd(4); sol[i].u.fill(u0); } }
  std::ifstream in(golden);
  CHECK(in.good());
  for (int t = 0; t < 15; ++t) {
    const VectorXd x0 = x;
    CHECK(solver.solve(x0, qp, sol) == hpipm::HpipmStatus::Success);
    std::vector<double> cat, gold(172);
    for (int i = 0; i <= N; ++i) for (int k = 0; k < 12; ++k) cat.push_back(sol[i].x(k));
    for (int i = 0; i < N; ++i) for (int k = 0; k < 4; ++k) cat.push_back(sol[i].u(k));
    for (double& v : gold) in >> v;
    CHECK(approx(cat.data(), gold.data(), 172, 1.0e-09));
    VectorXd xn(12);
    for (int i = 0; i < 12; ++i) { double sacc = 0; for (int j = 0; j < 12; ++j) sacc += A(i, j) * x(j); for (int j = 0; j < 4; ++j) sacc += B(i, j) * sol[0].u(j); xn(i) = sacc; }
    x = xn;
  }
  std::printf("compareResults: done\n");
}

// NEW (extensions of the facade): the closed loop of compareResults on the device for several robots, and the pooled
// contexts behind "a new solver object per SQP iteration" (NMPC_solver.cpp:319)
static void test_closed_loop_and_pool(const std::string& golden) {
  const int N = 10, B = 3, steps = 15;
  std::vector<hpipm::OcpQp> qp1 = quadcopterQp(N);
  MatrixXd A = qp1[0].A, Bm = qp1[0].B;
  hpipm::OcpQpIpmSolverSettings s;
  s.mode = hpipm::HpipmMode::Balance; s.iter_max = 30; s.alpha_min = 1e-8; s.mu0 = 1e2;
  s.tol_stat = s.tol_eq = s.tol_ineq = s.tol_comp = 1e-10;
  s.reg_prim = 1e-12; s.warm_start = 1; s.pred_corr = 1; s.ric_alg = 0; s.split_step = 1;
  const double u0 = 10.5916;
  std::vector<std::vector<hpipm::OcpQp>> qps(B, qp1);
  std::vector<std::vector<hpipm::OcpQpSolution>> sols(B, std::vector<hpipm::OcpQpSolution>(N + 1));
  std::vector<VectorXd> x0(B, VectorXd(12));
  x0[1](0) = 0.2; x0[1](2) = -0.1; x0[2](1) = -0.3; x0[2](8) = 0.1;
  for (int b = 0; b < B; ++b)
    for (int i = 0; i <= N; ++i) { sols[b][i].x = x0[b]; if (i < N) { sols[b][i].u = VectorXd(4); sols[b][i].u.fill(u0); } }
  auto sols_loop = sols;
  hpipm::OcpQpIpmSolver solver(s);
  solver.setOutputs(false, false);
  const auto res = solver.solveClosedLoop(x0, qps, sols, A, Bm, VectorXd(12), steps);
  CHECK((int)res.x_traj.size() == steps + 1 && (int)res.u_traj.size() == steps);
  // robot 0 against the golden vectors: x(t) and u0(t) within 1e-9 |sol_t| (the criterion of :310 restricted to them)
  std::ifstream in(golden);
  CHECK(in.good());
  for (int t = 0; t < steps; ++t) {
    std::vector<double> gold(172);
    for (double& v : gold) in >> v;
    double nrm = 0, ex = 0, eu = 0;
    for (double v : gold) nrm += v * v;
    for (int k = 0; k < 12; ++k) ex += (res.x_traj[t][0](k) - gold[k]) * (res.x_traj[t][0](k) - gold[k]);
    for (int k = 0; k < 4; ++k) eu += (res.u_traj[t][0](k) - gold[132 + k]) * (res.u_traj[t][0](k) - gold[132 + k]);
    CHECK(std::sqrt(ex) <= 1e-9 * std::sqrt(nrm) && std::sqrt(eu) <= 1e-9 * std::sqrt(nrm));
    for (int b = 0; b < B; ++b) CHECK(res.status[t][b] == hpipm::HpipmStatus::Success);
  }
  // every robot against the host-driven loop (one solve() per step and robot, a NEW solver object each time)
  const long created0 = hpipm::contextsCreated();
  for (int b = 0; b < B; ++b) {
    VectorXd x = x0[b];
    for (int t = 0; t < steps; ++t) {
      hpipm::OcpQpIpmSolver per_step(qps[b], s);   // what NMPCSolver::solveQpProblems does (NMPC_solver.cpp:319)
      CHECK(per_step.solve(x, qps[b], sols_loop[b]) == hpipm::HpipmStatus::Success);
      CHECK(approxV(res.x_traj[t][b], x, 1e-8) || t == 0);
      CHECK(approxV(res.u_traj[t][b], sols_loop[b][0].u, 1e-7));
      VectorXd xn(12);
      for (int i = 0; i < 12; ++i) { double sacc = 0; for (int j = 0; j < 12; ++j) sacc += A(i, j) * x(j); for (int j = 0; j < 4; ++j) sacc += Bm(i, j) * sols_loop[b][0].u(j); xn(i) = sacc; }
      x = xn;
    }
  }
  // 45 solver objects of one shape: at most one new device context (the pool hands the same one out again)
  CHECK(hpipm::contextsCreated() - created0 <= 1);
  std::printf("closed loop + context pool: done (%ld contexts created for %d solver objects)\n",
              hpipm::contextsCreated() - created0, B * steps);
}

// NEW: QPs with the structure NMPCSolver::prepareQpStructures produces reach the tensor-core kernel through the
// reference's own boundary when the optional outputs are switched off; same iterates as the default path within 1e-9
static void test_srbd_structured_qp_fast_path() {
  const int N = 20, nx = 12, nu = 12, ng = 24;
  std::vector<hpipm::OcpQp> qp(N + 1);
  MatrixXd D(ng, nu);
  for (int leg = 0; leg < 2; ++leg) {
    const int r0 = 12 * leg, c0 = 6 * leg;
    D(r0 + 0, c0 + 0) = -1; D(r0 + 0, c0 + 2) = 0.5; D(r0 + 1, c0 + 1) = -1; D(r0 + 1, c0 + 2) = 0.5;
    D(r0 + 2, c0 + 0) = 1; D(r0 + 2, c0 + 2) = 0.5; D(r0 + 3, c0 + 1) = 1; D(r0 + 3, c0 + 2) = 0.5;
    D(r0 + 4, c0 + 2) = -1; D(r0 + 5, c0 + 2) = 1;
    for (int k = 6; k < 12; ++k) { D(r0 + k, c0 + 2) = 0.05; D(r0 + k, c0 + 3 + (k % 3)) = (k % 2) ? 1.0 : -1.0; }
  }
  for (int i = 0; i <= N; ++i) {
    qp[i].Q = MatrixXd(nx, nx); for (int k = 0; k < nx; ++k) qp[i].Q(k, k) = i < N ? (k == 11 ? 10.0 : 0.0) : 20.0 * (1 + k);
    qp[i].q = RndV(nx);
    if (i == N) break;
    qp[i].A = MatrixXd::Identity(nx, nx); for (int k = 0; k < 6; ++k) qp[i].A(k, 6 + k) = 0.015;
    qp[i].B = Rnd(nx, nu); for (int r = 0; r < nx; ++r) for (int c = 0; c < nu; ++c) qp[i].B(r, c) *= 0.02;
    qp[i].b = RndV(nx); for (int k = 0; k < nx; ++k) qp[i].b(k) *= 0.01;
    qp[i].R = MatrixXd(nu, nu); for (int k = 0; k < nu; ++k) qp[i].R(k, k) = 1e-2;
    qp[i].S = MatrixXd(nu, nx); qp[i].r = RndV(nu);
    qp[i].C = MatrixXd(ng, nx); qp[i].D = D;
    qp[i].lg = VectorXd(ng); for (int k = 0; k < ng; ++k) qp[i].lg(k) = -1.0 - 0.1 * k;
    qp[i].ug = VectorXd(ng); qp[i].ug_mask = VectorXd(ng);   // upper side masked (zeros)
    qp[i].lg_mask = VectorXd(ng); qp[i].lg_mask.fill(1.0); qp[i].lg_mask(10) = 0.0; qp[i].lg_mask(23) = 0.0;
  }
  const VectorXd x0 = RndV(nx);
  hpipm::OcpQpIpmSolverSettings s;
  s.ric_alg = 0; s.iter_max = 40; s.split_step = 1;
  std::vector<hpipm::OcpQpSolution> full(N + 1), fast(N + 1), gen(N + 1);
  hpipm::OcpQpIpmSolver a(qp, s), b(qp, s), c(qp, s);
  // default outputs (P, p, K, k, pi[0], statistics, like hpipm-cpp's solve()): the tensor-core kernel with its exports
  CHECK(a.solve(x0, qp, full) == hpipm::HpipmStatus::Success);
  b.setOutputs(false, false);
  CHECK(b.solve(x0, qp, fast) == hpipm::HpipmStatus::Success);   // tensor-core kernel, x / u / pi only
  setenv("SRBD_K3_GENERIC", "1", 1);
  CHECK(c.solve(x0, qp, gen) == hpipm::HpipmStatus::Success);    // the generic kernel on the same request as `a`
  unsetenv("SRBD_K3_GENERIC");
  CHECK(a.getSolverStatistics().iter == b.getSolverStatistics().iter && b.getSolverStatistics().iter > 3);
  CHECK(a.getSolverStatistics().iter == c.getSolverStatistics().iter);
  CHECK(full[1].P.rows() == nx && fast[1].P.rows() == 0 && gen[1].P.rows() == nx);
  bool differ = false, same = true;
  for (int i = 0; i <= N; ++i) {
    CHECK(approxV(gen[i].x, fast[i].x, 1e-9));
    CHECK(approxM(gen[i].P, full[i].P, 1e-6) && approxV(gen[i].p, full[i].p, 1e-6) && approxV(gen[i].pi, full[i].pi, 1e-6));
    for (int k = 0; k < nx; ++k) same = same && full[i].x(k) == fast[i].x(k);
    if (i < N) {
      CHECK(approxV(gen[i].u, fast[i].u, 1e-8));
      CHECK(approxM(gen[i].K, full[i].K, 1e-6) && approxV(gen[i].k, full[i].k, 1e-6));
      for (int k = 0; k < nu; ++k) { differ = differ || gen[i].u(k) != fast[i].u(k); same = same && full[i].u(k) == fast[i].u(k); }
    }
  }
  CHECK(differ);   // (two different kernels: equal to rounding, not bit for bit)
  CHECK(same);     // (one kernel, with and without the exports: bit for bit)
  {
    const auto& sa = a.getSolverStatistics(); const auto& sc = c.getSolverStatistics();
    CHECK(sa.alpha_prim.size() == sc.alpha_prim.size() && sa.obj.size() == sc.obj.size());
    for (size_t i = 0; i < sa.alpha_prim.size() && i < sc.alpha_prim.size(); ++i)
      CHECK(std::abs(sa.alpha_prim[i] - sc.alpha_prim[i]) <= 1e-6 && std::abs(sa.mu[i] - sc.mu[i]) <= 1e-6 * (1.0 + sc.mu[i]) &&
            std::abs(sa.obj[i] - sc.obj[i]) <= 1e-8 * (1.0 + std::abs(sc.obj[i])));
  }
  std::printf("SRBD-structured QP through the facade: done (%d iterations on both kernels)\n", b.getSolverStatistics().iter);
}

// NEW: solveBatch splits large batches into chunks that two pooled contexts take alternately (flatten / H2D / kernels of
// neighbouring chunks overlap).  Same results bit for bit as one launch: generic QPs (boxes, general rows, S != 0), chunks
// of 3 + 3 + 3 + 2, with and without the optional outputs; SRBD-structured QPs (S = 0, C = 0: the two fields do not travel)
// with one QP whose S is non-zero in the LAST chunk only (that chunk uploads S and takes the generic kernel).
static void test_pipelined_batch() {
  const int nx = 5, nu = 3, ng = 2, N = 8, B = 11;
  auto absv = [](VectorXd v, double s, double off) { for (int i = 0; i < v.size(); ++i) v(i) = s * (off + std::fabs(v(i))); return v; };
  std::vector<std::vector<hpipm::OcpQp>> qps;
  std::vector<VectorXd> x0;
  for (int b = 0; b < B; ++b) {
    auto qp = randomQp(nx, nu, N, 0.4, true);
    for (int i = 0; i < N; ++i) {
      qp[i].idxbu = {0, 2};
      qp[i].lbu = absv(RndV(2), -1.0, 0.5); qp[i].ubu = absv(RndV(2), 1.0, 0.5);
      qp[i].C = Rnd(ng, nx); qp[i].D = Rnd(ng, nu);
      qp[i].lg = absv(RndV(ng), -10.0, 0.5); qp[i].ug = absv(RndV(ng), 10.0, 0.5);
    }
    qps.push_back(qp);
    x0.push_back(RndV(nx));
  }
  hpipm::OcpQpIpmSolverSettings s;
  s.ric_alg = 0; s.iter_max = 40; s.tol_stat = 1e-6;
  auto same = [&](const std::vector<std::vector<hpipm::OcpQpSolution>>& a, const std::vector<std::vector<hpipm::OcpQpSolution>>& c) {
    bool ok = a.size() == c.size();
    for (size_t b = 0; ok && b < a.size(); ++b)
      for (size_t i = 0; i < a[b].size(); ++i) {
        const auto &p = a[b][i], &q = c[b][i];
        ok = ok && p.x.size() == q.x.size() && p.u.size() == q.u.size() && p.pi.size() == q.pi.size() && p.P.size() == q.P.size() && p.K.size() == q.K.size();
        if (!ok) break;
        ok = ok && std::memcmp(p.x.data(), q.x.data(), p.x.size() * 8) == 0 && std::memcmp(p.pi.data(), q.pi.data(), p.pi.size() * 8) == 0;
        if (p.u.size()) ok = ok && std::memcmp(p.u.data(), q.u.data(), p.u.size() * 8) == 0;
        if (p.P.size()) ok = ok && std::memcmp(p.P.data(), q.P.data(), p.P.size() * 8) == 0;
        if (p.K.size()) ok = ok && std::memcmp(p.K.data(), q.K.data(), p.K.size() * 8) == 0;
      }
    return ok;
  };
  for (int outputs = 0; outputs < 2; ++outputs) {
    std::vector<std::vector<hpipm::OcpQpSolution>> one, piped;
    hpipm::OcpQpIpmSolver a(s), b(s);
    a.setOutputs(outputs != 0, outputs != 0); b.setOutputs(outputs != 0, outputs != 0);
    hpipm::OcpQpIpmSolver::setBatchChunk(0);
    const auto st1 = a.solveBatch(x0, qps, one);
    hpipm::OcpQpIpmSolver::setBatchChunk(3);
    const auto st2 = b.solveBatch(x0, qps, piped);
    const auto st3 = b.solveBatch(x0, qps, piped);   // again: the workers' contexts are reused
    CHECK(st1 == st2 && st1 == st3 && (int)st1.size() == B);
    for (auto v : st1) CHECK(v == hpipm::HpipmStatus::Success);
    CHECK(same(one, piped));
    CHECK(a.getBatchIterations() == b.getBatchIterations() && a.getBatchMaxResiduals() == b.getBatchMaxResiduals());
    CHECK(a.getSolverStatistics().iter == b.getSolverStatistics().iter && a.getSolverStatistics().mu == b.getSolverStatistics().mu);
    if (outputs) {   // NEW: the statistics table of every QP of the batch (opt-in), one launch == pipelined == one QP per call
      a.setKeepBatchStatistics(true); b.setKeepBatchStatistics(true);
      hpipm::OcpQpIpmSolver::setBatchChunk(0);
      a.solveBatch(x0, qps, one);
      hpipm::OcpQpIpmSolver::setBatchChunk(3);
      b.solveBatch(x0, qps, piped);
      for (int q : {0, 4, B - 1}) {
        const auto sa = a.getBatchStatistics(q), sb = b.getBatchStatistics(q);
        std::vector<hpipm::OcpQpSolution> lone;
        hpipm::OcpQpIpmSolver c(s);
        c.solve(x0[q], qps[q], lone);
        const auto& sc = c.getSolverStatistics();
        CHECK(sa.iter == sb.iter && sa.iter == sc.iter && (int)sa.mu.size() == sa.iter + 2);
        CHECK(sa.mu == sb.mu && sa.mu == sc.mu && sa.res_stat == sc.res_stat && sa.alpha_prim == sb.alpha_prim && sa.obj == sc.obj);
        CHECK(sa.max_res_comp == sc.max_res_comp);
      }
      bool thrown = false;
      try { a.getBatchStatistics(B); } catch (const std::runtime_error&) { thrown = true; }
      CHECK(thrown);
    }
  }
  // ---- SRBD-structured QPs -------------------------------------------------------------------------------------------
  {
    const int Ns = 20, n = 12, g = 24, Bs = 7;
    MatrixXd D(g, n);
    for (int leg = 0; leg < 2; ++leg) {
      const int r0 = 12 * leg, c0 = 6 * leg;
      D(r0 + 0, c0 + 0) = -1; D(r0 + 0, c0 + 2) = 0.5; D(r0 + 1, c0 + 1) = -1; D(r0 + 1, c0 + 2) = 0.5;
      D(r0 + 2, c0 + 0) = 1; D(r0 + 2, c0 + 2) = 0.5; D(r0 + 3, c0 + 1) = 1; D(r0 + 3, c0 + 2) = 0.5;
      D(r0 + 4, c0 + 2) = -1; D(r0 + 5, c0 + 2) = 1;
      for (int k = 6; k < 12; ++k) { D(r0 + k, c0 + 2) = 0.05; D(r0 + k, c0 + 3 + (k % 3)) = (k % 2) ? 1.0 : -1.0; }
    }
    std::vector<std::vector<hpipm::OcpQp>> sq(Bs, std::vector<hpipm::OcpQp>(Ns + 1));
    std::vector<VectorXd> sx0;
    for (int b = 0; b < Bs; ++b) {
      auto& qp = sq[b];
      for (int i = 0; i <= Ns; ++i) {
        qp[i].Q = MatrixXd(n, n); for (int k = 0; k < n; ++k) qp[i].Q(k, k) = i < Ns ? (k == 11 ? 10.0 : 0.0) : 20.0 * (1 + k);
        qp[i].q = RndV(n);
        if (i == Ns) break;
        qp[i].A = MatrixXd::Identity(n, n); for (int k = 0; k < 6; ++k) qp[i].A(k, 6 + k) = 0.015;
        qp[i].B = Rnd(n, n); for (int e = 0; e < n * n; ++e) qp[i].B.data()[e] *= 0.02;
        qp[i].b = RndV(n); for (int k = 0; k < n; ++k) qp[i].b(k) *= 0.01;
        qp[i].R = MatrixXd(n, n); for (int k = 0; k < n; ++k) qp[i].R(k, k) = 1e-2;
        qp[i].S = MatrixXd(n, n); qp[i].r = RndV(n);
        qp[i].C = MatrixXd(g, n); qp[i].D = D;
        qp[i].lg = VectorXd(g); for (int k = 0; k < g; ++k) qp[i].lg(k) = -1.0 - 0.1 * k;
        qp[i].ug = VectorXd(g); qp[i].ug_mask = VectorXd(g);
        qp[i].lg_mask = VectorXd(g); qp[i].lg_mask.fill(1.0); qp[i].lg_mask(10) = 0.0;
      }
      sx0.push_back(RndV(n));
    }
    hpipm::OcpQpIpmSolverSettings ss;
    ss.ric_alg = 0; ss.iter_max = 40; ss.split_step = 1;
    for (int variant = 0; variant < 2; ++variant) {
      if (variant == 1) sq[Bs - 1][3].S(2, 5) = 1e-3;   // last chunk only: S travels for that chunk, generic kernel there
      std::vector<std::vector<hpipm::OcpQpSolution>> one, piped;
      hpipm::OcpQpIpmSolver a(ss), b(ss);
      a.setOutputs(false, false); b.setOutputs(false, false);
      hpipm::OcpQpIpmSolver::setBatchChunk(0);
      const auto st1 = a.solveBatch(sx0, sq, one);
      hpipm::OcpQpIpmSolver::setBatchChunk(2);
      const auto st2 = b.solveBatch(sx0, sq, piped);
      CHECK(st1 == st2);
      for (auto v : st1) CHECK(v == hpipm::HpipmStatus::Success);
      if (variant == 0) CHECK(same(one, piped));
      else {   // the one-launch batch takes the generic kernel as a whole, the pipelined one only in its last chunk
        CHECK(a.getBatchIterations() == b.getBatchIterations());
        for (int b2 = 0; b2 < Bs; ++b2) for (int i = 0; i <= Ns; ++i) CHECK(approxV(one[b2][i].x, piped[b2][i].x, 1e-9));
        std::vector<std::vector<hpipm::OcpQpSolution>> lone(1);
        std::vector<std::vector<hpipm::OcpQp>> lq{sq[Bs - 1]};
        std::vector<VectorXd> lx{sx0[Bs - 1]};
        hpipm::OcpQpIpmSolver c(ss);
        c.setOutputs(false, false);
        c.solveBatch(lx, lq, lone);
        for (int i = 0; i <= Ns; ++i) CHECK(std::memcmp(lone[0][i].x.data(), piped[Bs - 1][i].x.data(), n * 8) == 0);
      }
    }
  }
  hpipm::OcpQpIpmSolver::setBatchChunk(1024);
  std::printf("pipelined solveBatch: done (bit-identical to one launch)\n");
}

static void test_errors() {
  auto qp = randomQp(5, 3, 4, 1.0, true);
  bool thrown = false;
  try { qp[1].A = Rnd(4, 5); hpipm::OcpQpDim d(qp); } catch (const std::runtime_error& e) { thrown = std::string(e.what()).find("ocp_qp[1].A.rows() must be 5") != std::string::npos; }
  CHECK(thrown);
  thrown = false;
  try { std::vector<hpipm::OcpQp> e; hpipm::OcpQpDim d(e); } catch (const std::runtime_error&) { thrown = true; }
  CHECK(thrown);
  thrown = false;
  qp = randomQp(5, 3, 4, 1.0, true);
  hpipm::OcpQpIpmSolverSettings s; s.ric_alg = 0; s.warm_start = 1;
  try { hpipm::OcpQpIpmSolver solver(qp, s); std::vector<hpipm::OcpQpSolution> sol(5); solver.solve(RndV(5), qp, sol); }
  catch (const std::runtime_error& e) { thrown = std::string(e.what()).find("qp_sol[0].x.size() must be 5") != std::string::npos; }
  CHECK(thrown);
  thrown = false;
  try { hpipm::OcpQpIpmSolverSettings b; b.mu0 = -1; b.checkSettings(); } catch (const std::runtime_error&) { thrown = true; }
  CHECK(thrown);
  thrown = false;   // NEW: a batch entry with other dimensions than the first is refused (the flattened batch is uniform)
  const uint64_t seed_keep = g_seed;   // (the later tests keep the random problems they were tuned on)
  try {
    std::vector<std::vector<hpipm::OcpQp>> bq{randomQp(5, 3, 4, 1.0, true), randomQp(4, 3, 4, 1.0, true)};
    std::vector<VectorXd> bx{RndV(5), RndV(4)};
    std::vector<std::vector<hpipm::OcpQpSolution>> bs;
    hpipm::OcpQpIpmSolverSettings s2; s2.ric_alg = 0;
    hpipm::OcpQpIpmSolver solver(s2);
    solver.solveBatch(bx, bq, bs);
  } catch (const std::runtime_error& e) { thrown = std::string(e.what()).find("must have the dimensions of the first") != std::string::npos; }
  CHECK(thrown);
  {   // NEW: an empty batch is not an error
    std::vector<std::vector<hpipm::OcpQp>> bq; std::vector<VectorXd> bx; std::vector<std::vector<hpipm::OcpQpSolution>> bs;
    hpipm::OcpQpIpmSolver solver;
    CHECK(solver.solveBatch(bx, bq, bs).empty() && solver.getBatchIterations().empty());
  }
  g_seed = seed_keep;
  std::printf("errors: done\n");
}

static void test_srbd_model() {
  SRBDModel m;
  m.SetMass(15.0); m.SetMPCdt(0.015);
  MatrixXd L(3, 3); L(0, 0) = 0.541667; L(1, 1) = 0.516667; L(2, 2) = 1.0416667;
  m.SetInertia(L);
  VectorXd pr(3), pl(3); pr(1) = -0.1; pl(1) = 0.1;
  m.SetFoot(pr, pl, MatrixXd::Identity(3, 3), MatrixXd::Identity(3, 3));
  VectorXd x(12), xn(12), u(12);
  x(8) = 1.0; xn(8) = 1.0; u.fill(100.0);
  MatrixXd A, B, b, f;
  m.GetShootingDynamic(x, xn, u, &A, &B, &b, &f);
  CHECK(std::fabs(B(9, 0) - 0.015 / 15.0) < 1e-15 && std::fabs(A(6, 9) - 0.015) < 1e-15 && A(0, 0) > 0.99);
  // pdot = v: x_get(8) = 1 + dt*v integrated from vdot = 200/15 - 9.8 -> defect of p_z is -(dt^2/2)*vdot
  CHECK(std::fabs(f(8, 0) + 0.5 * 0.015 * 0.015 * (200.0 / 15.0 - 9.8)) < 1e-12);
  for (int i = 0; i < 12; ++i) CHECK(b(i, 0) == -f(i, 0));
  MatrixXd Ac, fc;
  m.GetConstrain(u, Ac, fc);
  CHECK(Ac.rows() == 24 && Ac.cols() == 12);
  CHECK(std::fabs(fc(0, 0) - (-100.0 + 0.5 * 100.0)) < 1e-12 && std::fabs(fc(4, 0) - 900.0) < 1e-12 && std::fabs(fc(5, 0) - 100.0) < 1e-12);
  std::printf("SRBDModel: done\n");
}

static void test_nmpc_solver() {
  NMPCConfig cfg;
  cfg.N_rep = 1;
  NMPCSolver nmpc("../mpc_option.yaml", cfg);
  nmpc.controlLoop();
  CHECK(nmpc.lastSqpIterations() == 11);
  const double u0[12] = {54.37, 48.28, 100.32, 4.46, 24.94, 5.55, 63.53, 59.05, 122.50, 4.46, 25.97, 6.28};
  for (int i = 0; i < 12; ++i) CHECK(std::fabs(nmpc.u()[i] - u0[i]) < 0.02);
  std::printf("NMPCSolver: done\n");
}

int main(int argc, char** argv) {
  const std::string golden = argc > 1 ? argv[1] : "tests/golden/quadcopter_sol.txt";
  try {
    test_errors();
    test_unconstrained();
    test_constrained();
    test_compare_results(golden);
    test_closed_loop_and_pool(golden);
    test_srbd_structured_qp_fast_path();
    test_pipelined_batch();
    test_srbd_model();
    test_nmpc_solver();
  } catch (const std::exception& e) {
    std::printf("An exception occurred: %s\n", e.what());
    return 2;
  }
  std::printf(g_fail ? "FAILED (%d checks)\n" : "ALL OK\n", g_fail);
  return g_fail ? 1 : 0;
}
