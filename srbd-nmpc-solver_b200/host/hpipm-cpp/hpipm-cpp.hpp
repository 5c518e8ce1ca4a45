// hpipm-cpp.hpp — B200-backed drop-in for the reference's vendored hpipm-cpp facade
// (hpipm-cpp/include/hpipm-cpp/{ocp_qp,ocp_qp_solution,ocp_qp_dim,ocp_qp_ipm_solver_settings,
//  ocp_qp_ipm_solver_statistics,ocp_qp_ipm_solver}.hpp; implementation hpipm-cpp/src/*.cpp).
//
// Same namespace, class names, member names, argument meaning and error behaviour (std::runtime_error for
// shape / usage errors, HpipmStatus as the solver outcome).  Where the reference calls into HPIPM's C API
// (d_ocp_qp_set_all ... d_ocp_qp_ipm_solve ... getters, hpipm-cpp/src/ocp_qp_ipm_solver.cpp:283-407) this
// facade calls the C-ABI of include/srbd_b200.h, i.e. the CUDA path.  There is no CPU solve in here.
// Additions (new, not in the reference): OcpQpIpmSolver::solveBatch for B independent QPs in one launch.
#pragma once
#include <cstring>
#include <iomanip>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/srbd_b200.h"
#include "../eigen_shim.hpp"

namespace hpipm {

// ---- ocp_qp.hpp:15-177 -----------------------------------------------------------------------------
struct OcpQp {
  Eigen::MatrixXd A, B;
  Eigen::VectorXd b;
  Eigen::MatrixXd Q, S, R;
  Eigen::VectorXd q, r;
  std::vector<int> idxbx;
  Eigen::VectorXd lbx, ubx, lbx_mask, ubx_mask;
  std::vector<int> idxbu;
  Eigen::VectorXd lbu, ubu, lbu_mask, ubu_mask;
  Eigen::MatrixXd C, D;
  Eigen::VectorXd lg, ug, lg_mask, ug_mask;
  Eigen::MatrixXd Zl, Zu;
  Eigen::VectorXd zl, zu;
  std::vector<int> idxs;
  Eigen::VectorXd lls, lus;
};

// ---- ocp_qp_solution.hpp:12-48 -----------------------------------------------------------------------
struct OcpQpSolution {
  Eigen::VectorXd x, u, pi;
  Eigen::MatrixXd P;
  Eigen::VectorXd p;
  Eigen::MatrixXd K;
  Eigen::VectorXd k;
};

// ---- ocp_qp_ipm_solver_settings.hpp:21-92, src/ocp_qp_ipm_solver_settings.cpp:7-38 ---------------------
enum class HpipmMode { SpeedAbs, Speed, Balance, Robust };

struct OcpQpIpmSolverSettings {
  HpipmMode mode = HpipmMode::Speed;
  int iter_max = 15;
  double alpha_min = 1.0e-08;
  double mu0 = 1.0e+02;
  double tol_stat = 1.0e-08;
  double tol_eq = 1.0e-08;
  double tol_ineq = 1.0e-08;
  double tol_comp = 1.0e-08;
  double reg_prim = 1.0e-12;
  int warm_start = 0;
  int pred_corr = 1;
  int ric_alg = 1;
  int split_step = 0;
  void checkSettings() const {
    auto bad = [](const char* m) { throw std::runtime_error(std::string("OcpQpIpmSolverSettings.") + m); };
    if (iter_max < 0) bad("iter_max must be non-negative");
    if (alpha_min <= 0) bad("alpha_min must be positive");
    if (alpha_min > 1.0) bad("alpha_min must be less than 1.0");
    if (mu0 <= 0.0) bad("mu0 must be positive");
    if (tol_stat <= 0.0) bad("tol_stat must be positive");
    if (tol_eq <= 0.0) bad("tol_eq must be positive");
    if (tol_ineq <= 0.0) bad("tol_ineq must be positive");
    if (tol_comp <= 0.0) bad("tol_comp must be positive");
    if (reg_prim < 0.0) bad("reg_prim must be non-negative");
  }
};

// ---- ocp_qp_ipm_solver_statistics.hpp:15-78 ------------------------------------------------------------
struct OcpQpIpmSolverStatistics {
  int iter = 0;
  double max_res_stat = 0.0, max_res_eq = 0.0, max_res_ineq = 0.0, max_res_comp = 0.0;
  std::vector<double> alpha_aff, mu_aff, sigma, alpha_prim, alpha_dual, mu, res_stat, res_eq, res_ineq, res_comp, obj,
      lq_fact, itref_pred, itref_corr, lin_res_stat, lin_res_eq, lin_res_ineq, lin_res_comp;
  std::vector<std::vector<double>*> columns() {
    return {&alpha_aff, &mu_aff, &sigma, &alpha_prim, &alpha_dual, &mu, &res_stat, &res_eq, &res_ineq, &res_comp, &obj,
            &lq_fact, &itref_pred, &itref_corr, &lin_res_stat, &lin_res_eq, &lin_res_ineq, &lin_res_comp};
  }
  void resize(const size_t size) { for (auto* c : columns()) c->resize(size); }
  void reserve(const size_t size) { for (auto* c : columns()) c->reserve(size); }
  void clear() { for (auto* c : columns()) c->clear(); }
  void disp(std::ostream& os) const {
    os << "iterations: " << iter << "\nmax residuals (stat, eq, ineq, comp): " << max_res_stat << ", " << max_res_eq
       << ", " << max_res_ineq << ", " << max_res_comp << "\n";
    os << " it  alpha_aff     mu_aff      sigma alpha_prim alpha_dual         mu   res_stat     res_eq   res_ineq   res_comp\n";
    for (size_t i = 0; i < mu.size(); ++i) {
      os << std::setw(3) << i;
      for (double v : {alpha_aff[i], mu_aff[i], sigma[i], alpha_prim[i], alpha_dual[i], mu[i], res_stat[i], res_eq[i],
                       res_ineq[i], res_comp[i]})
        os << " " << std::setw(10) << std::scientific << std::setprecision(3) << v;
      os << "\n";
    }
  }
};
inline std::ostream& operator<<(std::ostream& os, const OcpQpIpmSolverStatistics& s) { s.disp(os); return os; }

// ---- ocp_qp_dim.hpp:15-121, src/ocp_qp_dim.cpp:32-246 ---------------------------------------------------
struct OcpQpDim {
  unsigned int N = 0;
  std::vector<int> nx, nu, nbx, nbu, ng, nsbx, nsbu, nsg;
  OcpQpDim() = default;
  explicit OcpQpDim(const unsigned int N_) { resize(N_); }
  explicit OcpQpDim(const std::vector<OcpQp>& ocp_qp) { resize(ocp_qp); }
  void resize(const unsigned int N_) {
    N = N_;
    for (auto* v : {&nx, &nu, &nbx, &nbu, &ng, &nsbx, &nsbu, &nsg}) v->assign(N + 1, 0);
  }
  // dimensions are INFERRED from the data (src/ocp_qp_dim.cpp:37-54)
  void resize(const std::vector<OcpQp>& ocp_qp) {
    if (ocp_qp.empty()) throw std::runtime_error("ocp_qp.size() must not be empty");
    resize(static_cast<unsigned int>(ocp_qp.size() - 1));
    for (unsigned int i = 0; i <= N; ++i) {
      nx[i] = static_cast<int>(ocp_qp[i].q.size());
      nu[i] = i < N ? static_cast<int>(ocp_qp[i].r.size()) : 0;
      nbx[i] = static_cast<int>(ocp_qp[i].idxbx.size());
      nbu[i] = i < N ? static_cast<int>(ocp_qp[i].idxbu.size()) : 0;
      ng[i] = static_cast<int>(ocp_qp[i].lg.size());
      nsbx[i] = static_cast<int>(ocp_qp[i].idxs.size());
    }
    checkSize(ocp_qp);
  }
  void checkSize(const std::vector<OcpQp>& ocp_qp) const {
    auto need = [](bool ok, unsigned int i, const char* what, int v) {
      if (!ok) throw std::runtime_error("ocp_qp[" + std::to_string(i) + "]." + what + " must be " + std::to_string(v));
    };
    if (ocp_qp.size() != N + 1) throw std::runtime_error("ocp_qp.size() must be " + std::to_string(N + 1));
    for (unsigned int i = 0; i < N; ++i) {
      need(ocp_qp[i].A.rows() == nx[i + 1], i, "A.rows()", nx[i + 1]);
      need(ocp_qp[i].A.cols() == nx[i], i, "A.cols()", nx[i]);
      need(ocp_qp[i].B.rows() == nx[i + 1], i, "B.rows()", nx[i + 1]);
      need(ocp_qp[i].B.cols() == nu[i], i, "B.cols()", nu[i]);
      need(ocp_qp[i].b.size() == nx[i], i, "b.size()", nx[i]);  // quirk kept: nx[i], not nx[i+1] (ocp_qp_dim.cpp:77)
      need(ocp_qp[i].S.rows() == nu[i] && ocp_qp[i].S.cols() == nx[i], i, "S (nu x nx) rows()", nu[i]);
      need(ocp_qp[i].R.rows() == nu[i] && ocp_qp[i].R.cols() == nu[i], i, "R.rows()", nu[i]);
      need(ocp_qp[i].r.size() == nu[i], i, "r.size()", nu[i]);
    }
    for (unsigned int i = 0; i <= N; ++i) {
      need(ocp_qp[i].Q.rows() == nx[i] && ocp_qp[i].Q.cols() == nx[i], i, "Q.rows()", nx[i]);
      need(ocp_qp[i].q.size() == nx[i], i, "q.size()", nx[i]);
      need(ocp_qp[i].lbx.size() == nbx[i], i, "lbx.size()", nbx[i]);
      need(ocp_qp[i].ubx.size() == nbx[i], i, "ubx.size()", nbx[i]);
      need(ocp_qp[i].ug.size() == ng[i], i, "ug.size()", ng[i]);
      if (ng[i] > 0) need(ocp_qp[i].C.rows() == ng[i] && ocp_qp[i].C.cols() == nx[i], i, "C.rows()", ng[i]);
      // soft constraints: nsg is forced to 0 by the reference (ocp_qp_dim.cpp:43-45,217-245): any slack data throws
      need(ocp_qp[i].Zl.size() == 0 && ocp_qp[i].Zu.size() == 0 && ocp_qp[i].zl.size() == 0 && ocp_qp[i].zu.size() == 0 &&
               ocp_qp[i].idxs.empty(), i, "Zl/Zu/zl/zu/idxs size (soft constraints unsupported)", 0);
    }
    for (unsigned int i = 0; i < N; ++i) {
      need(ocp_qp[i].lbu.size() == nbu[i], i, "lbu.size()", nbu[i]);
      need(ocp_qp[i].ubu.size() == nbu[i], i, "ubu.size()", nbu[i]);
      if (ng[i] > 0) need(ocp_qp[i].D.rows() == ng[i] && ocp_qp[i].D.cols() == nu[i], i, "D.rows()", ng[i]);
    }
  }
};

// ---- ocp_qp_ipm_solver.hpp:24-159, src/ocp_qp_ipm_solver.cpp ----------------------------------------------
enum class HpipmStatus { Success = 0, MaxIterReached = 1, MinStepLengthReached = 2, NaNDetected = 3, UnknownFailure = 4 };

inline std::string to_string(const HpipmStatus& s) {
  switch (s) {
    case HpipmStatus::Success: return "HpipmStatus::Success";
    case HpipmStatus::MaxIterReached: return "HpipmStatus::MaxIterReached";
    case HpipmStatus::MinStepLengthReached: return "HpipmStatus::MinStepLengthReached";
    case HpipmStatus::NaNDetected: return "HpipmStatus::NaNDetected";
    default: return "HpipmStatus::UnknownFailure";
  }
}
inline std::ostream& operator<<(std::ostream& os, const HpipmStatus& s) { return os << to_string(s); }

class OcpQpIpmSolver {
 public:
  OcpQpIpmSolver(const std::vector<OcpQp>& ocp_qp, const OcpQpIpmSolverSettings& s = OcpQpIpmSolverSettings()) {
    setSolverSettings(s);
    resize(ocp_qp);
  }
  explicit OcpQpIpmSolver(const OcpQpIpmSolverSettings& s = OcpQpIpmSolverSettings()) { setSolverSettings(s); }
  ~OcpQpIpmSolver() { release(); }
  OcpQpIpmSolver(const OcpQpIpmSolver&) = delete;
  OcpQpIpmSolver& operator=(const OcpQpIpmSolver&) = delete;
  OcpQpIpmSolver(OcpQpIpmSolver&& o) noexcept { *this = std::move(o); }
  OcpQpIpmSolver& operator=(OcpQpIpmSolver&& o) noexcept {
    if (this != &o) {
      release();
      solver_settings_ = o.solver_settings_; solver_statistics_ = o.solver_statistics_; dim_ = o.dim_;
      ctx_ = o.ctx_; ctx_dims_ = o.ctx_dims_; ctx_batch_ = o.ctx_batch_; device_ = o.device_;
      o.ctx_ = nullptr;
    }
    return *this;
  }

  void setSolverSettings(const OcpQpIpmSolverSettings& s) { solver_settings_ = s; }
  void setDevice(int device) { device_ = device; }

  // sizes the (grow-only) device workspace like the reference's wrappers do (detail/d_ocp_qp_ipm_ws_wrapper.cpp:141-155)
  void resize(const std::vector<OcpQp>& ocp_qp) {
    dim_.resize(ocp_qp);
    ensureContext(1);
  }

  HpipmStatus solve(const Eigen::VectorXd& x0, std::vector<OcpQp>& ocp_qp, std::vector<OcpQpSolution>& qp_sol) {
    std::vector<const std::vector<OcpQp>*> qps{&ocp_qp};
    std::vector<std::vector<OcpQpSolution>*> sols{&qp_sol};
    std::vector<const Eigen::VectorXd*> x0s{&x0};
    std::vector<HpipmStatus> st;
    solveImpl(x0s, qps, sols, st, /*collect_stats=*/true);
    return st[0];
  }

  // NEW: B independent QPs of identical dimensions in one launch; statistics are those of the last QP
  std::vector<HpipmStatus> solveBatch(const std::vector<Eigen::VectorXd>& x0, std::vector<std::vector<OcpQp>>& ocp_qp,
                                      std::vector<std::vector<OcpQpSolution>>& qp_sol) {
    if (x0.size() != ocp_qp.size()) throw std::runtime_error("x0.size() must be " + std::to_string(ocp_qp.size()));
    if (qp_sol.size() != ocp_qp.size()) qp_sol.resize(ocp_qp.size());
    std::vector<const std::vector<OcpQp>*> qps;
    std::vector<std::vector<OcpQpSolution>*> sols;
    std::vector<const Eigen::VectorXd*> x0s;
    for (size_t i = 0; i < ocp_qp.size(); ++i) { qps.push_back(&ocp_qp[i]); sols.push_back(&qp_sol[i]); x0s.push_back(&x0[i]); }
    std::vector<HpipmStatus> st;
    solveImpl(x0s, qps, sols, st, true);
    return st;
  }

  const OcpQpIpmSolverSettings& getIpmSolverSettings() const { return solver_settings_; }
  const OcpQpIpmSolverStatistics& getSolverStatistics() const { return solver_statistics_; }

 private:
  OcpQpIpmSolverSettings solver_settings_;
  OcpQpIpmSolverStatistics solver_statistics_;
  OcpQpDim dim_;
  srbd_ctx* ctx_ = nullptr;
  srbd_qp_dims ctx_dims_{};
  int ctx_batch_ = 0;
  int device_ = 0;

  void release() {
    if (ctx_) srbd_ctx_destroy(ctx_);
    ctx_ = nullptr;
  }
  void check(int rc, const char* what) const {
    if (rc != 0) throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + srbd_last_error(ctx_));
  }
  srbd_qp_dims uniformDims() const {
    // the GPU path takes uniform stage dimensions (every reference workload and test has them)
    const unsigned int N = dim_.N;
    if (N < 1) throw std::runtime_error("ocp_qp.size() must be at least 2");
    srbd_qp_dims d{};
    d.N = static_cast<int>(N); d.nx = dim_.nx[N]; d.nu = dim_.nu[0]; d.nbx = dim_.nbx[N]; d.nbu = dim_.nbu[0];
    d.ng = dim_.ng[0]; d.ngN = dim_.ng[N];
    for (unsigned int i = 0; i <= N; ++i) {
      bool ok = dim_.nx[i] == d.nx && (i == 0 || dim_.nbx[i] == d.nbx);
      if (i < N) ok = ok && dim_.nu[i] == d.nu && dim_.nbu[i] == d.nbu && dim_.ng[i] == d.ng;
      if (!ok) throw std::runtime_error("ocp_qp[" + std::to_string(i) + "]: the B200 path needs uniform stage dimensions");
    }
    return d;
  }
  void ensureContext(int batch) {
    const srbd_qp_dims d = uniformDims();
    if (ctx_ && std::memcmp(&d, &ctx_dims_, sizeof(d)) == 0 && batch == ctx_batch_) return;
    release();
    const int rc = srbd_ctx_create(device_, batch, &d, nullptr, &ctx_);
    if (rc != 0) throw std::runtime_error("srbd_ctx_create failed (" + std::to_string(rc) + "): no usable CUDA device or "
                                          "dimensions beyond the compiled maxima");
    ctx_dims_ = d;
    ctx_batch_ = batch;
  }

  void solveImpl(const std::vector<const Eigen::VectorXd*>& x0s, const std::vector<const std::vector<OcpQp>*>& qps,
                 const std::vector<std::vector<OcpQpSolution>*>& sols, std::vector<HpipmStatus>& status, bool collect_stats) {
    solver_settings_.checkSettings();
    const int B = static_cast<int>(qps.size());
    dim_.resize(*qps[0]);  // resize(ocp_qp) on every call like the reference (ocp_qp_ipm_solver.cpp:185)
    for (int b = 1; b < B; ++b) { OcpQpDim chk(*qps[b]); (void)chk; }
    ensureContext(B);
    const srbd_qp_dims d = ctx_dims_;
    const size_t N = d.N, nx = d.nx, nu = d.nu, nbx = d.nbx, nbu = d.nbu, ng = d.ng, ngN = d.ngN;
    // warm start needs pre-sized solutions (ocp_qp_ipm_solver.cpp:189-208)
    for (int b = 0; b < B; ++b) {
      auto& s = *sols[b];
      if (s.size() != N + 1) s.resize(N + 1);
      for (size_t i = 0; i <= N; ++i) {
        if (solver_settings_.warm_start) {
          if (static_cast<size_t>(s[i].x.size()) != nx) throw std::runtime_error("qp_sol[" + std::to_string(i) + "].x.size() must be " + std::to_string(nx));
          if (i < N && static_cast<size_t>(s[i].u.size()) != nu) throw std::runtime_error("qp_sol[" + std::to_string(i) + "].u.size() must be " + std::to_string(nu));
        }
      }
    }
    // flatten into the batch-contiguous column-major layout of srbd_qp_host
    auto cat = [&](size_t per, auto getter, size_t stages, size_t first = 0) {
      std::vector<double> v(static_cast<size_t>(B) * stages * per, 0.0);
      for (int b = 0; b < B; ++b)
        for (size_t i = first; i < stages; ++i) {
          const double* src = getter((*qps[b])[i]);
          if (src && per) std::memcpy(v.data() + (static_cast<size_t>(b) * stages + i) * per, src, per * sizeof(double));
        }
      return v;
    };
    auto A = cat(nx * nx, [](const OcpQp& q) { return q.A.data(); }, N);
    auto Bm = cat(nx * nu, [](const OcpQp& q) { return q.B.data(); }, N);
    auto bv = cat(nx, [](const OcpQp& q) { return q.b.data(); }, N);
    auto Q = cat(nx * nx, [](const OcpQp& q) { return q.Q.data(); }, N + 1);
    auto S = cat(nu * nx, [](const OcpQp& q) { return q.S.data(); }, N);
    auto R = cat(nu * nu, [](const OcpQp& q) { return q.R.data(); }, N);
    auto qv = cat(nx, [](const OcpQp& q) { return q.q.data(); }, N + 1);
    auto rv = cat(nu, [](const OcpQp& q) { return q.r.data(); }, N);
    // (stage 0 is skipped: nbx[0] := 0 in the solver, and uniformDims() lets nbx[0] differ from nbx)
    auto lbx = cat(nbx, [&](const OcpQp& q) { return q.lbx.size() ? q.lbx.data() : nullptr; }, N + 1, 1);
    auto ubx = cat(nbx, [&](const OcpQp& q) { return q.ubx.size() ? q.ubx.data() : nullptr; }, N + 1, 1);
    auto lbu = cat(nbu, [](const OcpQp& q) { return q.lbu.data(); }, N);
    auto ubu = cat(nbu, [](const OcpQp& q) { return q.ubu.data(); }, N);
    auto C = cat(ng * nx, [&](const OcpQp& q) { return q.C.size() ? q.C.data() : nullptr; }, N);
    auto D = cat(ng * nu, [](const OcpQp& q) { return q.D.data(); }, N);
    auto lg = cat(ng, [](const OcpQp& q) { return q.lg.data(); }, N);
    auto ug = cat(ng, [](const OcpQp& q) { return q.ug.data(); }, N);
    // masks apply only when their size matches (ocp_qp_ipm_solver.cpp:292-321); default is "all ones"
    auto mask = [&](size_t per, auto vec, size_t stages, size_t first) {
      std::vector<double> v(static_cast<size_t>(B) * stages * per, 1.0);
      for (int b = 0; b < B; ++b)
        for (size_t i = first; i < stages; ++i) {
          const Eigen::VectorXd& m = vec((*qps[b])[i]);
          if (per && static_cast<size_t>(m.size()) == per)
            std::memcpy(v.data() + (static_cast<size_t>(b) * stages + i) * per, m.data(), per * sizeof(double));
        }
      return v;
    };
    auto lbxm = mask(nbx, [](const OcpQp& q) -> const Eigen::VectorXd& { return q.lbx_mask; }, N + 1, 1);
    auto ubxm = mask(nbx, [](const OcpQp& q) -> const Eigen::VectorXd& { return q.ubx_mask; }, N + 1, 1);
    auto lbum = mask(nbu, [](const OcpQp& q) -> const Eigen::VectorXd& { return q.lbu_mask; }, N, 0);
    auto ubum = mask(nbu, [](const OcpQp& q) -> const Eigen::VectorXd& { return q.ubu_mask; }, N, 0);
    auto lgm = mask(ng, [](const OcpQp& q) -> const Eigen::VectorXd& { return q.lg_mask; }, N, 0);
    auto ugm = mask(ng, [](const OcpQp& q) -> const Eigen::VectorXd& { return q.ug_mask; }, N, 0);
    std::vector<double> CN(B * ngN * nx), lgN(B * ngN), ugN(B * ngN), lgNm(B * ngN, 1.0), ugNm(B * ngN, 1.0), x0v(B * nx);
    std::vector<double> xin(B * (N + 1) * nx, 0.0), uin(B * N * nu, 0.0);
    for (int b = 0; b < B; ++b) {
      const OcpQp& qN = (*qps[b])[N];
      if (ngN) {
        std::memcpy(CN.data() + b * ngN * nx, qN.C.data(), ngN * nx * sizeof(double));
        std::memcpy(lgN.data() + b * ngN, qN.lg.data(), ngN * sizeof(double));
        std::memcpy(ugN.data() + b * ngN, qN.ug.data(), ngN * sizeof(double));
        if (static_cast<size_t>(qN.lg_mask.size()) == ngN) std::memcpy(lgNm.data() + b * ngN, qN.lg_mask.data(), ngN * sizeof(double));
        if (static_cast<size_t>(qN.ug_mask.size()) == ngN) std::memcpy(ugNm.data() + b * ngN, qN.ug_mask.data(), ngN * sizeof(double));
      }
      if (static_cast<size_t>(x0s[b]->size()) != nx) throw std::runtime_error("x0.size() must be " + std::to_string(nx));
      std::memcpy(x0v.data() + b * nx, x0s[b]->data(), nx * sizeof(double));
      if (solver_settings_.warm_start)
        for (size_t i = 0; i <= N; ++i) {
          std::memcpy(xin.data() + (b * (N + 1) + i) * nx, (*sols[b])[i].x.data(), nx * sizeof(double));
          if (i < N) std::memcpy(uin.data() + (b * N + i) * nu, (*sols[b])[i].u.data(), nu * sizeof(double));
        }
    }
    // The C-ABI takes ONE index set for idxbx (stages 1..N) and one for idxbu (stages 0..N-1), shared by the batch;
    // the reference passes them per stage (ocp_qp_ipm_solver.cpp:263-272).  Differing sets would silently put the
    // bounds on the wrong variables: refuse them.
    for (int b = 0; b < B; ++b)
      for (size_t i = 0; i <= N; ++i) {
        const OcpQp& s = (*qps[b])[i];
        if (i >= 1 && nbx && s.idxbx != (*qps[0])[N].idxbx)
          throw std::runtime_error("ocp_qp[" + std::to_string(i) + "].idxbx differs between stages / batch entries: the "
                                   "B200 path needs one idxbx for stages 1..N");
        if (i < N && nbu && s.idxbu != (*qps[0])[0].idxbu)
          throw std::runtime_error("ocp_qp[" + std::to_string(i) + "].idxbu differs between stages / batch entries: the "
                                   "B200 path needs one idxbu for stages 0..N-1");
      }
    srbd_qp_host h{};
    h.A = A.data(); h.Bm = Bm.data(); h.b = bv.data(); h.Q = Q.data(); h.S = S.data(); h.R = R.data(); h.q = qv.data(); h.r = rv.data();
    h.idxbx = nbx ? (*qps[0])[N].idxbx.data() : nullptr; h.lbx = lbx.data(); h.ubx = ubx.data(); h.lbx_mask = lbxm.data(); h.ubx_mask = ubxm.data();
    h.idxbu = nbu ? (*qps[0])[0].idxbu.data() : nullptr; h.lbu = lbu.data(); h.ubu = ubu.data(); h.lbu_mask = lbum.data(); h.ubu_mask = ubum.data();
    h.C = C.data(); h.D = D.data(); h.lg = lg.data(); h.ug = ug.data(); h.lg_mask = lgm.data(); h.ug_mask = ugm.data();
    h.CN = CN.data(); h.lgN = lgN.data(); h.ugN = ugN.data(); h.lgN_mask = lgNm.data(); h.ugN_mask = ugNm.data();
    h.x0 = x0v.data();
    if (solver_settings_.warm_start) { h.x_init = xin.data(); h.u_init = uin.data(); }

    srbd_ipm_args a;
    srbd_ipm_args_default(&a);
    a.iter_max = solver_settings_.iter_max; a.alpha_min = solver_settings_.alpha_min; a.mu0 = solver_settings_.mu0;
    a.tol_stat = solver_settings_.tol_stat; a.tol_eq = solver_settings_.tol_eq; a.tol_ineq = solver_settings_.tol_ineq;
    a.tol_comp = solver_settings_.tol_comp; a.reg_prim = solver_settings_.reg_prim; a.warm_start = solver_settings_.warm_start;
    a.pred_corr = solver_settings_.pred_corr; a.ric_alg = solver_settings_.ric_alg; a.split_step = solver_settings_.split_step;
    check(srbd_set_ipm_args(ctx_, &a), "srbd_set_ipm_args");
    check(srbd_set_outputs(ctx_, 1, collect_stats ? 1 : 0), "srbd_set_outputs");
    check(srbd_qp_upload(ctx_, &h), "srbd_qp_upload");
    check(srbd_qp_solve(ctx_), "srbd_qp_solve");

    std::vector<double> x(B * (N + 1) * nx), u(B * N * nu), pi(B * (N + 1) * nx), P(B * (N + 1) * nx * nx), p(B * (N + 1) * nx),
        K(B * N * nu * nx), k(B * N * nu);
    srbd_sol_host so{};
    so.x = x.data(); so.u = u.data(); so.pi = pi.data(); so.P = P.data(); so.p = p.data(); so.K = K.data(); so.k = k.data();
    check(srbd_download_solution(ctx_, &so), "srbd_download_solution");
    const int rows = srbd_ctx_stat_rows(ctx_);
    std::vector<int> it(B), st(B);
    std::vector<double> rm(B * 4), tab(static_cast<size_t>(B) * rows * SRBD_STAT_M);
    srbd_stats_host sh{};
    sh.iter = it.data(); sh.status = st.data(); sh.res_max = rm.data(); sh.stat = collect_stats ? tab.data() : nullptr;
    check(srbd_download_stats(ctx_, &sh), "srbd_download_stats");

    status.resize(B);
    for (int b = 0; b < B; ++b) {
      auto& s = *sols[b];
      for (size_t i = 0; i <= N; ++i) {
        s[i].x.resize(nx); s[i].pi.resize(nx); s[i].P.resize(nx, nx); s[i].p.resize(nx);
        std::memcpy(s[i].x.data(), x.data() + (b * (N + 1) + i) * nx, nx * sizeof(double));
        std::memcpy(s[i].pi.data(), pi.data() + (b * (N + 1) + i) * nx, nx * sizeof(double));
        std::memcpy(s[i].P.data(), P.data() + (b * (N + 1) + i) * nx * nx, nx * nx * sizeof(double));
        std::memcpy(s[i].p.data(), p.data() + (b * (N + 1) + i) * nx, nx * sizeof(double));
        if (i < N) {
          s[i].u.resize(nu); s[i].K.resize(nu, nx); s[i].k.resize(nu);
          std::memcpy(s[i].u.data(), u.data() + (b * N + i) * nu, nu * sizeof(double));
          std::memcpy(s[i].K.data(), K.data() + (b * N + i) * nu * nx, nu * nx * sizeof(double));
          std::memcpy(s[i].k.data(), k.data() + (b * N + i) * nu, nu * sizeof(double));
        }
      }
      status[b] = (st[b] >= 0 && st[b] <= 3) ? static_cast<HpipmStatus>(st[b]) : HpipmStatus::UnknownFailure;
    }
    // statistics of the (last) QP: iter, 4 max residuals, rows 0..iter+1 of the 18-column table (:376-403)
    const int b = B - 1;
    solver_statistics_.iter = it[b];
    solver_statistics_.max_res_stat = rm[4 * b + 0]; solver_statistics_.max_res_eq = rm[4 * b + 1];
    solver_statistics_.max_res_ineq = rm[4 * b + 2]; solver_statistics_.max_res_comp = rm[4 * b + 3];
    solver_statistics_.clear();
    if (collect_stats) {
      auto cols = solver_statistics_.columns();
      for (int i = 0; i <= it[b] + 1 && i < rows; ++i)
        for (int c = 0; c < SRBD_STAT_M; ++c) cols[c]->push_back(tab[(static_cast<size_t>(b) * rows + i) * SRBD_STAT_M + c]);
    }
  }
};

}  // namespace hpipm
