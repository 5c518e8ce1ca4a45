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
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <iomanip>
#include <iostream>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

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
  // NEW (solveBatch): another QP of the batch must have exactly these dimensions (the flattened batch is uniform);
  // same messages as checkSize, no allocation
  void checkBatchEntry(const std::vector<OcpQp>& ocp_qp) const {
    if (ocp_qp.size() != N + 1) throw std::runtime_error("ocp_qp.size() must be " + std::to_string(N + 1));
    for (unsigned int i = 0; i <= N; ++i) {
      const OcpQp& s = ocp_qp[i];
      const bool ok = static_cast<int>(s.q.size()) == nx[i] && (i == N || static_cast<int>(s.r.size()) == nu[i]) &&
                      static_cast<int>(s.idxbx.size()) == nbx[i] && (i == N || static_cast<int>(s.idxbu.size()) == nbu[i]) &&
                      static_cast<int>(s.lg.size()) == ng[i];
      if (!ok) throw std::runtime_error("ocp_qp[" + std::to_string(i) + "]: every QP of a batch must have the dimensions of the first");
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

namespace detail {
// Process-wide pool of B200 contexts (device buffers, stream, pinned staging arena), keyed by (device, batch, dims).
// The reference constructs a NEW OcpQpIpmSolver in every SQP iteration (NMPC_solver.cpp:319) and its wrappers malloc
// their workspaces each time (detail/d_ocp_qp_ipm_ws_wrapper.cpp:141-155); here a solver object borrows a context from
// the pool and gives it back in its destructor, so nothing is allocated after the first solver of a shape.
struct PooledContext {
  srbd_ctx* ctx = nullptr;
  int device = 0, batch = 0;
  srbd_qp_dims dims{};
  double* arena = nullptr;      // pinned host staging: [QP fields ... | outputs]
  size_t arena_doubles = 0;
};
class ContextPool {
 public:
  static ContextPool& instance() { static ContextPool* p = new ContextPool(); return *p; }  // never destroyed: CUDA may be gone at exit
  PooledContext acquire(int device, int batch, const srbd_qp_dims& d) {
    {
      std::lock_guard<std::mutex> lock(m_);
      for (size_t i = 0; i < free_.size(); ++i)
        if (free_[i].device == device && free_[i].batch == batch && std::memcmp(&free_[i].dims, &d, sizeof(d)) == 0) {
          PooledContext e = free_[i];
          free_.erase(free_.begin() + static_cast<long>(i));
          return e;
        }
    }
    PooledContext e;
    e.device = device; e.batch = batch; e.dims = d;
    const int rc = srbd_ctx_create(device, batch, &d, nullptr, &e.ctx);
    if (rc != 0) throw std::runtime_error("srbd_ctx_create failed (" + std::to_string(rc) + "): no usable CUDA device or "
                                          "dimensions beyond the compiled maxima");
    ++created_;
    return e;
  }
  void release(PooledContext e) {
    if (!e.ctx) return;
    std::lock_guard<std::mutex> lock(m_);
    free_.push_back(e);
  }
  void clear() {  // explicit teardown (tests, long-running hosts that change shapes)
    std::lock_guard<std::mutex> lock(m_);
    for (auto& e : free_) { if (e.arena) srbd_host_free(e.arena); srbd_ctx_destroy(e.ctx); }
    free_.clear();
  }
  long created() const { return created_; }
 private:
  std::mutex m_;
  std::vector<PooledContext> free_;
  long created_ = 0;
};
// Host-side phase times of solve / solveBatch, accumulated when SRBD_FACADE_PROFILE=1 (ms; bench_facade prints them)
struct FacadeProfile {
  double validate = 0, flatten = 0, enqueue = 0, wait = 0, scatter = 0;
  static FacadeProfile& get() { static FacadeProfile p; return p; }
  static bool on() { static const bool v = [] { const char* e = std::getenv("SRBD_FACADE_PROFILE"); return e && std::atoi(e) != 0; }(); return v; }
  static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
};
// Copy of one matrix block into the pinned staging arena.  The arena is written once and read by the DMA engine only:
// blocks of whole cache lines go out with non-temporal stores (no read-for-ownership of the destination lines, a third
// less memory traffic per copied byte); SRBD_FACADE_NT=0 falls back to memcpy.
inline bool streamStores() {
  static const bool on = [] { const char* e = std::getenv("SRBD_FACADE_NT"); return !e || std::atoi(e) != 0; }();
  return on;
}
inline void copyOut(double* dst, const double* src, size_t n) {
#if defined(__SSE2__)
  if (n >= 64 && n % 8 == 0 && (reinterpret_cast<uintptr_t>(dst) & 63) == 0 && streamStores()) {
    for (size_t e = 0; e < n; e += 2) _mm_stream_pd(dst + e, _mm_loadu_pd(src + e));
    return;
  }
#endif
  std::memcpy(dst, src, n * sizeof(double));
}
inline void storeFence() {
#if defined(__SSE2__)
  _mm_sfence();
#endif
}
// Flattening B x (N+1) Eigen objects into the staging arena (and scattering the solutions back) is plain memory
// traffic: 228 KB per N = 20 SRBD QP.  Large batches split it over a few host threads.
template <class F>
inline void parallelFor(size_t n, size_t grain, F&& fn) {
  size_t nt = std::thread::hardware_concurrency();
  if (nt > 16) nt = 16;
  if (nt * grain > n) nt = n / grain;
  if (nt < 2) { fn(size_t(0), n); return; }
  std::vector<std::thread> th;
  std::exception_ptr err;   // the first exception of a worker is rethrown on the calling thread
  std::mutex err_m;
  const size_t chunk = (n + nt - 1) / nt;
  for (size_t t = 0; t < nt; ++t) {
    const size_t lo = t * chunk, hi = lo + chunk < n ? lo + chunk : n;
    if (lo < hi) th.emplace_back([&fn, &err, &err_m, lo, hi] {
      try { fn(lo, hi); } catch (...) { std::lock_guard<std::mutex> lock(err_m); if (!err) err = std::current_exception(); }
    });
  }
  for (auto& t : th) t.join();
  if (err) std::rethrow_exception(err);
}
}  // namespace detail

// number of device contexts created so far (tests: a solver per SQP iteration must not create more than one)
inline long contextsCreated() { return detail::ContextPool::instance().created(); }
inline void releaseCachedContexts() { detail::ContextPool::instance().clear(); }

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
      pc_ = o.pc_; device_ = o.device_; want_ric_ = o.want_ric_; want_stat_ = o.want_stat_;
      o.pc_ = detail::PooledContext();
      worker_[0] = std::move(o.worker_[0]); worker_[1] = std::move(o.worker_[1]);
      batch_iter_ = std::move(o.batch_iter_); batch_res_ = std::move(o.batch_res_);
      keep_batch_stats_ = o.keep_batch_stats_; batch_tab_ = std::move(o.batch_tab_); batch_tab_rows_ = o.batch_tab_rows_;
    }
    return *this;
  }

  void setSolverSettings(const OcpQpIpmSolverSettings& s) { solver_settings_ = s; }
  void setDevice(int device) { device_ = device; }
  // NEW.  Which optional outputs solve() produces.  Default (true, true) = the reference: P, p, K, k of every stage and
  // the per-iteration statistics table.  (false, false) fills x, u, pi[1..N] only (pi[0] is reconstructed from the Riccati
  // data, ocp_qp_ipm_solver.cpp:349-373) and lets QPs with the structure NMPCSolver::prepareQpStructures produces take the
  // tensor-core kernel (include/srbd_b200.h: srbd_qp_upload) -- what NMPCSolver::solveQpProblems needs (:322-329).
  void setOutputs(bool riccati, bool statistics) { want_ric_ = riccati; want_stat_ = statistics; }
  // NEW.  solveBatch pipelines batches of at least two chunks (host flattening / H2D copy / kernels of neighbouring chunks
  // overlap on two pooled contexts); chunk size in QPs, 0 = one launch for the whole batch.  Default 1024 (SRBD_FACADE_CHUNK).
  static void setBatchChunk(size_t qps) { pipelineChunk() = qps; }

  // sizes the (grow-only, pooled) device workspace like the reference's wrappers do (detail/d_ocp_qp_ipm_ws_wrapper.cpp:141-155)
  void resize(const std::vector<OcpQp>& ocp_qp) {
    dim_.resize(ocp_qp);
    ensureContext(1);
  }

  HpipmStatus solve(const Eigen::VectorXd& x0, std::vector<OcpQp>& ocp_qp, std::vector<OcpQpSolution>& qp_sol) {
    std::vector<const std::vector<OcpQp>*> qps{&ocp_qp};
    std::vector<std::vector<OcpQpSolution>*> sols{&qp_sol};
    std::vector<const Eigen::VectorXd*> x0s{&x0};
    std::vector<HpipmStatus> st;
    solveImpl(x0s, qps, sols, st, nullptr);
    return st[0];
  }

  // NEW: B independent QPs of identical dimensions in one launch.  getSolverStatistics() describes the LAST QP of the
  // batch; getBatchIterations() / getBatchMaxResiduals() hold iter and the four max residuals of every QP.
  std::vector<HpipmStatus> solveBatch(const std::vector<Eigen::VectorXd>& x0, std::vector<std::vector<OcpQp>>& ocp_qp,
                                      std::vector<std::vector<OcpQpSolution>>& qp_sol) {
    if (x0.size() != ocp_qp.size()) throw std::runtime_error("x0.size() must be " + std::to_string(ocp_qp.size()));
    if (qp_sol.size() != ocp_qp.size()) qp_sol.resize(ocp_qp.size());
    std::vector<HpipmStatus> st;
    if (ocp_qp.empty()) { batch_iter_.clear(); batch_res_.clear(); return st; }   // an empty batch is solved: nothing to do
    std::vector<const std::vector<OcpQp>*> qps;
    std::vector<std::vector<OcpQpSolution>*> sols;
    std::vector<const Eigen::VectorXd*> x0s;
    for (size_t i = 0; i < ocp_qp.size(); ++i) { qps.push_back(&ocp_qp[i]); sols.push_back(&qp_sol[i]); x0s.push_back(&x0[i]); }
    solveImpl(x0s, qps, sols, st, nullptr);
    return st;
  }

  // NEW: the closed MPC loop of examples/example_mpc.cpp:99-119 / test/ocp_qp_ipm_solver.cpp:298-314 for B robots, on the
  // device (srbd_mpc_run): for t < sim_steps: x0 := x(t); solve(x0, qp, solution) with `solution` passed back in as the
  // warm start; x(t+1) := A x(t) + B u0 + b.  x0[i] is robot i's initial state, qp_sol[i] its initial guess (when
  // warm_start = 1) and, on return, the solution of its LAST step.  x_traj[t][i] (t <= sim_steps) / u_traj[t][i] are the
  // closed-loop states and applied inputs; the returned status matrix is [sim_steps][B].
  struct ClosedLoopResult {
    std::vector<std::vector<Eigen::VectorXd>> x_traj, u_traj;
    std::vector<std::vector<HpipmStatus>> status;
    std::vector<std::vector<int>> iter;
  };
  ClosedLoopResult solveClosedLoop(const std::vector<Eigen::VectorXd>& x0, std::vector<std::vector<OcpQp>>& ocp_qp,
                                   std::vector<std::vector<OcpQpSolution>>& qp_sol, const Eigen::MatrixXd& A,
                                   const Eigen::MatrixXd& B, const Eigen::VectorXd& b, int sim_steps) {
    if (x0.size() != ocp_qp.size()) throw std::runtime_error("x0.size() must be " + std::to_string(ocp_qp.size()));
    if (qp_sol.size() != ocp_qp.size()) qp_sol.resize(ocp_qp.size());
    if (sim_steps < 1) throw std::runtime_error("sim_steps must be positive");
    if (ocp_qp.empty()) throw std::runtime_error("solveClosedLoop needs at least one robot");
    std::vector<const std::vector<OcpQp>*> qps;
    std::vector<std::vector<OcpQpSolution>*> sols;
    std::vector<const Eigen::VectorXd*> x0s;
    for (size_t i = 0; i < ocp_qp.size(); ++i) { qps.push_back(&ocp_qp[i]); sols.push_back(&qp_sol[i]); x0s.push_back(&x0[i]); }
    ClosedLoop cl{&A, &B, &b, sim_steps, ClosedLoopResult()};
    std::vector<HpipmStatus> st;
    solveImpl(x0s, qps, sols, st, &cl);
    return cl.out;
  }

  const OcpQpIpmSolverSettings& getIpmSolverSettings() const { return solver_settings_; }
  const OcpQpIpmSolverStatistics& getSolverStatistics() const { return solver_statistics_; }
  const std::vector<int>& getBatchIterations() const { return batch_iter_; }
  const std::vector<double>& getBatchMaxResiduals() const { return batch_res_; }  // [B][4]: stat, eq, ineq, comp
  // NEW.  The per-iteration statistics table of EVERY QP of a batch (getSolverStatistics() describes the last one only).
  // Opt-in because the tables are copied out of the staging arena (rows x 18 doubles per QP); needs setOutputs(.., true).
  void setKeepBatchStatistics(bool keep) { keep_batch_stats_ = keep; }
  OcpQpIpmSolverStatistics getBatchStatistics(size_t b) const {
    if (b >= batch_iter_.size()) throw std::runtime_error("getBatchStatistics: the last batch had " + std::to_string(batch_iter_.size()) + " QPs");
    OcpQpIpmSolverStatistics st;
    st.iter = batch_iter_[b];
    st.max_res_stat = batch_res_[4 * b + 0]; st.max_res_eq = batch_res_[4 * b + 1];
    st.max_res_ineq = batch_res_[4 * b + 2]; st.max_res_comp = batch_res_[4 * b + 3];
    if (batch_tab_.empty()) throw std::runtime_error("getBatchStatistics: call setKeepBatchStatistics(true) and setOutputs(.., true) before solveBatch");
    auto cols = st.columns();
    const size_t rows = static_cast<size_t>(batch_tab_rows_);
    for (size_t i = 0; i <= static_cast<size_t>(st.iter) + 1 && i < rows; ++i)
      for (int c = 0; c < SRBD_STAT_M; ++c) cols[c]->push_back(batch_tab_[(b * rows + i) * SRBD_STAT_M + c]);
    return st;
  }

 private:
  struct ClosedLoop {
    const Eigen::MatrixXd *A, *B;
    const Eigen::VectorXd* b;
    int steps;
    ClosedLoopResult out;
  };
  OcpQpIpmSolverSettings solver_settings_;
  OcpQpIpmSolverStatistics solver_statistics_;
  OcpQpDim dim_;
  detail::PooledContext pc_;
  int device_ = 0;
  bool want_ric_ = true, want_stat_ = true;
  std::vector<int> batch_iter_;
  std::vector<double> batch_res_;
  bool keep_batch_stats_ = false;
  std::vector<double> batch_tab_;   // [B][rows][SRBD_STAT_M] when kept
  int batch_tab_rows_ = 0;
  // solveBatch on large batches: two worker solvers (each with its own pooled context, stream and pinned arena) take
  // chunks of the batch alternately, so that flattening / scattering on the host, the H2D copy and the kernels of
  // neighbouring chunks overlap (solvePipelined)
  std::unique_ptr<OcpQpIpmSolver> worker_[2];
  struct Pending {   // what submit() leaves for collect()
    std::vector<std::vector<OcpQpSolution>*> sols;
    size_t in_total = 0;
    bool ric = false, stat = false;
  } pend_;

  void release() {
    detail::ContextPool::instance().release(pc_);
    pc_ = detail::PooledContext();
  }
  void check(int rc, const char* what) const {
    if (rc != 0) throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + srbd_last_error(pc_.ctx));
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
    if (pc_.ctx && std::memcmp(&d, &pc_.dims, sizeof(d)) == 0 && batch == pc_.batch) return;
    release();
    pc_ = detail::ContextPool::instance().acquire(device_, batch, d);
  }
  double* arena(size_t doubles) {  // pinned, grow-only, owned by the pooled context
    if (doubles > pc_.arena_doubles) {
      if (pc_.arena) srbd_host_free(pc_.arena);
      pc_.arena = nullptr; pc_.arena_doubles = 0;
      void* p = nullptr;
      if (srbd_host_alloc(doubles * sizeof(double), &p) != 0) throw std::runtime_error("srbd_host_alloc failed");
      pc_.arena = static_cast<double*>(p);
      pc_.arena_doubles = doubles;
    }
    return pc_.arena;
  }

  // QPs per chunk of the pipelined solveBatch (SRBD_FACADE_CHUNK; 0 = never pipeline).  A chunk of 1024 N = 20 SRBD QPs is
  // 235 MB of QP fields: a few ms each of host flattening, H2D copy and kernels.
  static size_t& pipelineChunk() {
    static size_t c = [] { const char* e = std::getenv("SRBD_FACADE_CHUNK"); const long v = e ? std::atol(e) : 1024; return static_cast<size_t>(v > 0 ? v : 0); }();
    return c;
  }
  static bool anyNonZero(const double* v, size_t n) {
    bool nz = false;
    for (size_t e = 0; e < n; ++e) nz |= v[e] != 0.0;
    return nz;
  }

  void solveImpl(const std::vector<const Eigen::VectorXd*>& x0s, const std::vector<const std::vector<OcpQp>*>& qps,
                 const std::vector<std::vector<OcpQpSolution>*>& sols, std::vector<HpipmStatus>& status, ClosedLoop* loop) {
    const size_t chunk = pipelineChunk();
    if (!loop && chunk && qps.size() / 2 >= chunk) { solvePipelined(x0s, qps, sols, status, chunk); return; }
    submit(x0s, qps, sols, loop);
    collect(status, loop);
  }

  // Chunks of the batch go alternately to two worker solvers: while the host scatters chunk c - 2 and flattens chunk c,
  // the H2D copy and the kernels of chunk c - 1 run on the other worker's stream.  Same results as one big launch (QPs
  // are independent; every chunk is validated, uploaded and solved exactly like a batch of its own).
  void solvePipelined(const std::vector<const Eigen::VectorXd*>& x0s, const std::vector<const std::vector<OcpQp>*>& qps,
                      const std::vector<std::vector<OcpQpSolution>*>& sols, std::vector<HpipmStatus>& status, size_t chunk) {
    const size_t B = qps.size(), nchunks = (B + chunk - 1) / chunk;
    for (auto& w : worker_) {
      if (!w) w.reset(new OcpQpIpmSolver(solver_settings_));
      w->setSolverSettings(solver_settings_);
      w->setDevice(device_);
      w->setOutputs(want_ric_, want_stat_);
      w->setKeepBatchStatistics(keep_batch_stats_);
    }
    batch_tab_.clear();
    batch_tab_rows_ = 0;
    status.assign(B, HpipmStatus::UnknownFailure);
    batch_iter_.assign(B, 0);
    batch_res_.assign(4 * B, 0.0);
    std::vector<HpipmStatus> st;
    auto finish = [&](size_t c) {
      OcpQpIpmSolver& w = *worker_[c & 1];
      const size_t lo = c * chunk;
      w.collect(st, nullptr);
      std::copy(st.begin(), st.end(), status.begin() + static_cast<long>(lo));
      std::copy(w.batch_iter_.begin(), w.batch_iter_.end(), batch_iter_.begin() + static_cast<long>(lo));
      std::copy(w.batch_res_.begin(), w.batch_res_.end(), batch_res_.begin() + static_cast<long>(4 * lo));
      if (!w.batch_tab_.empty()) {   // (the number of rows depends on iter_max only: the same in every chunk)
        batch_tab_rows_ = w.batch_tab_rows_;
        const size_t per = static_cast<size_t>(batch_tab_rows_) * SRBD_STAT_M;
        if (batch_tab_.size() != B * per) batch_tab_.assign(B * per, 0.0);
        std::copy(w.batch_tab_.begin(), w.batch_tab_.end(), batch_tab_.begin() + static_cast<long>(lo * per));
      }
    };
    try {
      for (size_t c = 0; c < nchunks; ++c) {
        if (c >= 2) finish(c - 2);
        const size_t lo = c * chunk, hi = lo + chunk < B ? lo + chunk : B;
        const std::vector<const Eigen::VectorXd*> cx(x0s.begin() + static_cast<long>(lo), x0s.begin() + static_cast<long>(hi));
        const std::vector<const std::vector<OcpQp>*> cq(qps.begin() + static_cast<long>(lo), qps.begin() + static_cast<long>(hi));
        const std::vector<std::vector<OcpQpSolution>*> cs(sols.begin() + static_cast<long>(lo), sols.begin() + static_cast<long>(hi));
        worker_[c & 1]->submit(cx, cq, cs, nullptr);
      }
      for (size_t c = nchunks >= 2 ? nchunks - 2 : 0; c < nchunks; ++c) finish(c);
    } catch (...) {   // leave no work in flight on a context that goes back to the pool
      for (auto& w : worker_) if (w && w->pc_.ctx) srbd_ctx_sync(w->pc_.ctx);
      throw;
    }
    const OcpQpIpmSolver& last = *worker_[(nchunks - 1) & 1];
    dim_ = last.dim_;
    solver_statistics_ = last.solver_statistics_;
  }

  void submit(const std::vector<const Eigen::VectorXd*>& x0s, const std::vector<const std::vector<OcpQp>*>& qps,
              const std::vector<std::vector<OcpQpSolution>*>& sols, ClosedLoop* loop) {
    const bool prof = detail::FacadeProfile::on();
    double t_ = prof ? detail::FacadeProfile::now() : 0.0;
    auto lap = [&](double detail::FacadeProfile::*m) {
      if (!prof) return;
      const double t1 = detail::FacadeProfile::now();
      detail::FacadeProfile::get().*m += t1 - t_;
      t_ = t1;
    };
    solver_settings_.checkSettings();
    const int B = static_cast<int>(qps.size());
    dim_.resize(*qps[0]);  // resize(ocp_qp) on every call like the reference (ocp_qp_ipm_solver.cpp:185)
    ensureContext(B);
    srbd_ctx* ctx = pc_.ctx;
    const srbd_qp_dims d = pc_.dims;
    const size_t N = d.N, nx = d.nx, nu = d.nu, nbx = d.nbx, nbu = d.nbu, ng = d.ng, ngN = d.ngN, Bz = static_cast<size_t>(B);
    const bool warm = solver_settings_.warm_start != 0;
    // Per QP, spread over host threads (an exception of a worker is rethrown here): the dimensions of every batch entry
    // equal the first's; warm start needs pre-sized solutions (ocp_qp_ipm_solver.cpp:189-208); and the C-ABI takes ONE
    // index set for idxbx (stages 1..N) and one for idxbu (stages 0..N-1), shared by the batch, where the reference passes
    // them per stage (ocp_qp_ipm_solver.cpp:263-272) -- differing sets would silently put the bounds on the wrong
    // variables: refuse them.
    detail::parallelFor(Bz, 64, [&](size_t lo, size_t hi) {
      for (size_t b = lo; b < hi; ++b) {
        if (b > 0) dim_.checkBatchEntry(*qps[b]);
        auto& s = *sols[b];
        if (s.size() != N + 1) s.resize(N + 1);
        if (warm)
          for (size_t i = 0; i <= N; ++i) {
            if (static_cast<size_t>(s[i].x.size()) != nx) throw std::runtime_error("qp_sol[" + std::to_string(i) + "].x.size() must be " + std::to_string(nx));
            if (i < N && static_cast<size_t>(s[i].u.size()) != nu) throw std::runtime_error("qp_sol[" + std::to_string(i) + "].u.size() must be " + std::to_string(nu));
          }
        if (static_cast<size_t>(x0s[b]->size()) != nx) throw std::runtime_error("x0.size() must be " + std::to_string(nx));
        if (nbx || nbu)
          for (size_t i = 0; i <= N; ++i) {
            const OcpQp& q = (*qps[b])[i];
            if (i >= 1 && nbx && q.idxbx != (*qps[0])[N].idxbx)
              throw std::runtime_error("ocp_qp[" + std::to_string(i) + "].idxbx differs between stages / batch entries: the "
                                       "B200 path needs one idxbx for stages 1..N");
            if (i < N && nbu && q.idxbu != (*qps[0])[0].idxbu)
              throw std::runtime_error("ocp_qp[" + std::to_string(i) + "].idxbu differs between stages / batch entries: the "
                                       "B200 path needs one idxbu for stages 0..N-1");
          }
      }
    });
    lap(&detail::FacadeProfile::validate);
    // ---- staging arena: every QP field batch-contiguous and column-major (srbd_qp_host), one after the other, then the
    // outputs in the device's own layout: ONE H2D copy up, ONE D2H copy down -------------------------------------------
    size_t out_off[8], out_total = 0;
    check(srbd_out_layout(ctx, out_off, &out_total), "srbd_out_layout");
    const size_t in_sizes[31] = {Bz * N * nx * nx, Bz * N * nx * nu, Bz * N * nx, Bz * (N + 1) * nx * nx, Bz * N * nu * nx,
                                 Bz * N * nu * nu, Bz * (N + 1) * nx, Bz * N * nu,
                                 Bz * (N + 1) * nbx, Bz * (N + 1) * nbx, Bz * (N + 1) * nbx, Bz * (N + 1) * nbx,
                                 Bz * N * nbu, Bz * N * nbu, Bz * N * nbu, Bz * N * nbu,
                                 Bz * N * ng * nx, Bz * N * ng * nu, Bz * N * ng, Bz * N * ng, Bz * N * ng, Bz * N * ng,
                                 Bz * ngN * nx, Bz * ngN, Bz * ngN, Bz * ngN, Bz * ngN, Bz * nx,
                                 warm ? Bz * (N + 1) * nx : 0, warm ? Bz * N * nu : 0, ng * nu};
    enum { fA, fB, fb, fQ, fS, fR, fq, fr, flbx, fubx, flbxm, fubxm, flbu, fubu, flbum, fubum, fC, fD, flg, fug, flgm, fugm,
           fCN, flgN, fugN, flgNm, fugNm, fx0, fxin, fuin, fDs };
    // S and C lie at the END of the input area: when every S_k and C_k of the batch is zero (what
    // NMPCSolver::prepareQpStructures hands over, NMPC_solver.cpp:277-314) they are passed as "absent" and the single
    // H2D copy of srbd_qp_upload (the span of the fields that are present) ends before them: 30 % fewer bytes over PCIe
    // The same for D: NMPCSolver::prepareQpStructures hands over ONE constraint matrix in every stage of every QP
    // (NMPC_solver.cpp:290-300).  While every D_k equals the first bit for bit, only that one is staged (slot fDs) and the
    // upload is told so (srbd_qp_upload_layout); the per-stage copies are written by the second pass otherwise.
    const int order[31] = {fA, fB, fb, fQ, fR, fq, fr, flbx, fubx, flbxm, fubxm, flbu, fubu, flbum, fubum, fDs, flg, fug, flgm, fugm,
                           fCN, flgN, fugN, flgNm, fugNm, fx0, fxin, fuin, fD, fS, fC};
    size_t in_off[31], in_total = 0;
    for (int i = 0; i < 31; ++i) { in_off[order[i]] = in_total; in_total += in_sizes[order[i]]; }
    srbd_ipm_args a;
    srbd_ipm_args_default(&a);
    srbd_ipm_args_set_mode(&a, static_cast<int>(solver_settings_.mode));  // d_ocp_qp_ipm_arg_set_default(mode), :103
    a.iter_max = solver_settings_.iter_max; a.alpha_min = solver_settings_.alpha_min; a.mu0 = solver_settings_.mu0;
    a.tol_stat = solver_settings_.tol_stat; a.tol_eq = solver_settings_.tol_eq; a.tol_ineq = solver_settings_.tol_ineq;
    a.tol_comp = solver_settings_.tol_comp; a.reg_prim = solver_settings_.reg_prim; a.warm_start = solver_settings_.warm_start;
    a.pred_corr = solver_settings_.pred_corr; a.ric_alg = solver_settings_.ric_alg; a.split_step = solver_settings_.split_step;
    check(srbd_set_ipm_args(ctx, &a), "srbd_set_ipm_args");
    const bool ric = want_ric_, stat = want_stat_ && !loop;
    check(srbd_set_outputs(ctx, ric ? 1 : 0, stat ? 1 : 0), "srbd_set_outputs");
    // the optional outputs come down into the pinned arena as well (behind the packed outputs): no pageable staging
    const size_t ric_total = ric ? Bz * ((N + 1) * (nx * nx + nx) + N * (nu * nx + nu)) : 0;
    const size_t tab_total = stat ? Bz * static_cast<size_t>(srbd_ctx_stat_rows(ctx)) * SRBD_STAT_M : 0;
    double* ar = arena(in_total + out_total + ric_total + tab_total);
    std::atomic<int> nz_S{0}, nz_C{0}, var_D{0};
    const double* D0 = ng ? (*qps[0])[0].D.data() : nullptr;
    if (ng) std::memcpy(ar + in_off[fDs], D0, ng * nu * sizeof(double));
    auto put = [&](int f, size_t b, size_t stages, size_t i, size_t per, const double* src) {
      if (per) detail::copyOut(ar + in_off[f] + (b * stages + i) * per, src, per);
    };
    auto fill = [&](int f, size_t b, size_t stages, size_t i, size_t per, double v) {
      double* dst = ar + in_off[f] + (b * stages + i) * per;
      for (size_t e = 0; e < per; ++e) dst[e] = v;
    };
    // masks apply only when their size matches (ocp_qp_ipm_solver.cpp:292-321); default is "all ones"
    auto mask = [&](int f, size_t b, size_t stages, size_t i, size_t per, const Eigen::VectorXd& m) {
      if (static_cast<size_t>(m.size()) == per) put(f, b, stages, i, per, m.data());
      else fill(f, b, stages, i, per, 1.0);
    };
    detail::parallelFor(Bz, 64, [&](size_t b_lo, size_t b_hi) {
    for (size_t b = b_lo; b < b_hi; ++b) {
      const std::vector<OcpQp>& qp = *qps[b];
      for (size_t i = 0; i <= N; ++i) {
        const OcpQp& s = qp[i];
        put(fQ, b, N + 1, i, nx * nx, s.Q.data());
        put(fq, b, N + 1, i, nx, s.q.data());
        // (stage 0 is skipped: nbx[0] := 0 in the solver, and uniformDims() lets nbx[0] differ from nbx)
        if (i >= 1) {
          put(flbx, b, N + 1, i, nbx, s.lbx.data()); put(fubx, b, N + 1, i, nbx, s.ubx.data());
          mask(flbxm, b, N + 1, i, nbx, s.lbx_mask); mask(fubxm, b, N + 1, i, nbx, s.ubx_mask);
        } else {
          fill(flbx, b, N + 1, i, nbx, 0.0); fill(fubx, b, N + 1, i, nbx, 0.0);
          fill(flbxm, b, N + 1, i, nbx, 1.0); fill(fubxm, b, N + 1, i, nbx, 1.0);
        }
        if (i == N) break;
        put(fA, b, N, i, nx * nx, s.A.data()); put(fB, b, N, i, nx * nu, s.B.data()); put(fb, b, N, i, nx, s.b.data());
        put(fR, b, N, i, nu * nu, s.R.data()); put(fr, b, N, i, nu, s.r.data());
        // S and C are only LOOKED at here; they are copied by a second pass below if any of them is non-zero
        if (!nz_S.load(std::memory_order_relaxed) && anyNonZero(s.S.data(), nu * nx)) nz_S.store(1, std::memory_order_relaxed);
        put(flbu, b, N, i, nbu, s.lbu.data()); put(fubu, b, N, i, nbu, s.ubu.data());
        mask(flbum, b, N, i, nbu, s.lbu_mask); mask(fubum, b, N, i, nbu, s.ubu_mask);
        if (ng) {
          // (C_0 is dropped by the x0 embedding, nx[0] := 0, so it does not count)
          if (i >= 1 && s.C.size() && !nz_C.load(std::memory_order_relaxed) && anyNonZero(s.C.data(), ng * nx)) nz_C.store(1, std::memory_order_relaxed);
          if (!var_D.load(std::memory_order_relaxed) && std::memcmp(s.D.data(), D0, ng * nu * sizeof(double)) != 0)
            var_D.store(1, std::memory_order_relaxed);
          put(flg, b, N, i, ng, s.lg.data()); put(fug, b, N, i, ng, s.ug.data());
          mask(flgm, b, N, i, ng, s.lg_mask); mask(fugm, b, N, i, ng, s.ug_mask);
        }
      }
      const OcpQp& qN = qp[N];
      if (ngN) {
        put(fCN, b, 1, 0, ngN * nx, qN.C.data()); put(flgN, b, 1, 0, ngN, qN.lg.data()); put(fugN, b, 1, 0, ngN, qN.ug.data());
        mask(flgNm, b, 1, 0, ngN, qN.lg_mask); mask(fugNm, b, 1, 0, ngN, qN.ug_mask);
      }
      put(fx0, b, 1, 0, nx, x0s[b]->data());
      if (warm)
        for (size_t i = 0; i <= N; ++i) {
          put(fxin, b, N + 1, i, nx, (*sols[b])[i].x.data());
          if (i < N) put(fuin, b, N, i, nu, (*sols[b])[i].u.data());
        }
    }
    detail::storeFence();
    });
    if (nz_S.load() || nz_C.load() || var_D.load())
      detail::parallelFor(Bz, 64, [&](size_t b_lo, size_t b_hi) {
        for (size_t b = b_lo; b < b_hi; ++b)
          for (size_t i = 0; i < N; ++i) {
            const OcpQp& s = (*qps[b])[i];
            if (var_D.load() && ng) put(fD, b, N, i, ng * nu, s.D.data());
            if (nz_S.load()) put(fS, b, N, i, nu * nx, s.S.data());
            if (nz_C.load() && ng) { if (s.C.size()) put(fC, b, N, i, ng * nx, s.C.data()); else fill(fC, b, N, i, ng * nx, 0.0); }
          }
        detail::storeFence();
      });
    auto at = [&](int f) -> const double* { return in_sizes[f] ? ar + in_off[f] : nullptr; };
    srbd_qp_host h{};
    h.A = at(fA); h.Bm = at(fB); h.b = at(fb); h.Q = at(fQ); h.S = nz_S.load() ? at(fS) : nullptr; h.R = at(fR); h.q = at(fq); h.r = at(fr);
    h.idxbx = nbx ? (*qps[0])[N].idxbx.data() : nullptr; h.lbx = at(flbx); h.ubx = at(fubx); h.lbx_mask = at(flbxm); h.ubx_mask = at(fubxm);
    h.idxbu = nbu ? (*qps[0])[0].idxbu.data() : nullptr; h.lbu = at(flbu); h.ubu = at(fubu); h.lbu_mask = at(flbum); h.ubu_mask = at(fubum);
    h.C = nz_C.load() ? at(fC) : nullptr; h.D = var_D.load() ? at(fD) : at(fDs); h.lg = at(flg); h.ug = at(fug); h.lg_mask = at(flgm); h.ug_mask = at(fugm);
    h.CN = at(fCN); h.lgN = at(flgN); h.ugN = at(fugN); h.lgN_mask = at(flgNm); h.ugN_mask = at(fugNm);
    h.x0 = at(fx0);
    if (warm) { h.x_init = at(fxin); h.u_init = at(fuin); }
    lap(&detail::FacadeProfile::flatten);

    check(srbd_qp_upload_layout(ctx, var_D.load() ? 0 : 1), "srbd_qp_upload_layout");
    check(srbd_qp_upload(ctx, &h), "srbd_qp_upload");
    if (loop) {
      const int steps = loop->steps;
      if (static_cast<size_t>(loop->A->rows()) != nx || static_cast<size_t>(loop->A->cols()) != nx ||
          static_cast<size_t>(loop->B->rows()) != nx || static_cast<size_t>(loop->B->cols()) != nu ||
          static_cast<size_t>(loop->b->size()) != nx)
        throw std::runtime_error("solveClosedLoop: the plant must be A (nx x nx), B (nx x nu), b (nx)");
      std::vector<double> xs(Bz * nx), xt((steps + 1) * Bz * nx), ut(steps * Bz * nu);
      std::vector<int> it(steps * Bz), st(steps * Bz);
      for (size_t b = 0; b < Bz; ++b) std::memcpy(xs.data() + b * nx, x0s[b]->data(), nx * sizeof(double));
      check(srbd_mpc_run(ctx, loop->A->data(), loop->B->data(), loop->b->data(), 1, xs.data(), steps, xt.data(), ut.data(),
                         it.data(), st.data()), "srbd_mpc_run");
      auto& o = loop->out;
      o.x_traj.assign(steps + 1, std::vector<Eigen::VectorXd>(Bz, Eigen::VectorXd(static_cast<int>(nx))));
      o.u_traj.assign(steps, std::vector<Eigen::VectorXd>(Bz, Eigen::VectorXd(static_cast<int>(nu))));
      o.status.assign(steps, std::vector<HpipmStatus>(Bz));
      o.iter.assign(steps, std::vector<int>(Bz));
      for (int t = 0; t <= steps; ++t)
        for (size_t b = 0; b < Bz; ++b) {
          std::memcpy(o.x_traj[t][b].data(), xt.data() + (t * Bz + b) * nx, nx * sizeof(double));
          if (t == steps) continue;
          std::memcpy(o.u_traj[t][b].data(), ut.data() + (t * Bz + b) * nu, nu * sizeof(double));
          const int sv = st[t * Bz + b];
          o.status[t][b] = (sv >= 0 && sv <= 3) ? static_cast<HpipmStatus>(sv) : HpipmStatus::UnknownFailure;
          o.iter[t][b] = it[t * Bz + b];
        }
    } else {
      check(srbd_qp_solve(ctx), "srbd_qp_solve");   // (enqueued on the context's stream: collect() waits for it)
    }
    lap(&detail::FacadeProfile::enqueue);
    pend_.sols = sols;
    pend_.in_total = in_total;
    pend_.ric = ric;
    pend_.stat = stat;
  }

  // second half of a solve: the outputs of the submitted batch come down in one D2H copy and are scattered into qp_sol
  void collect(std::vector<HpipmStatus>& status, ClosedLoop* loop) {
    (void)loop;
    srbd_ctx* ctx = pc_.ctx;
    const srbd_qp_dims d = pc_.dims;
    const std::vector<std::vector<OcpQpSolution>*>& sols = pend_.sols;
    const int B = static_cast<int>(sols.size());
    const size_t N = d.N, nx = d.nx, nu = d.nu, Bz = static_cast<size_t>(B);
    const bool ric = pend_.ric, stat = pend_.stat;
    size_t out_off[8], out_total = 0;
    check(srbd_out_layout(ctx, out_off, &out_total), "srbd_out_layout");
    double* out = pc_.arena + pend_.in_total;
    const bool prof = detail::FacadeProfile::on();
    const double t0_ = prof ? detail::FacadeProfile::now() : 0.0;
    check(srbd_download_packed(ctx, out, 0), "srbd_download_packed");
    const double t1_ = prof ? detail::FacadeProfile::now() : 0.0;
    const double *x = out + out_off[0], *u = out + out_off[1], *pi = out + out_off[2], *rm = out + out_off[3];
    const int* it = reinterpret_cast<const int*>(out + out_off[4]);
    const int* st = reinterpret_cast<const int*>(out + out_off[5]);
    double *P = out + out_total, *p = P + Bz * (N + 1) * nx * nx, *K = p + Bz * (N + 1) * nx, *k = K + Bz * N * nu * nx;
    double* tab = out + out_total + (ric ? Bz * ((N + 1) * (nx * nx + nx) + N * (nu * nx + nu)) : 0);
    if (ric) {
      srbd_sol_host so{};
      so.P = P; so.p = p; so.K = K; so.k = k;
      check(srbd_download_solution(ctx, &so), "srbd_download_solution");
    }
    const int rows = srbd_ctx_stat_rows(ctx);
    if (stat) {
      srbd_stats_host sh{};
      sh.stat = tab;
      check(srbd_download_stats(ctx, &sh), "srbd_download_stats");
    }
    status.resize(B);
    batch_iter_.assign(it, it + B);
    batch_res_.assign(rm, rm + 4 * Bz);
    detail::parallelFor(Bz, 64, [&](size_t b_lo, size_t b_hi) {
    for (size_t b = b_lo; b < b_hi; ++b) {
      auto& s = *sols[b];
      for (size_t i = 0; i <= N; ++i) {
        s[i].x.resize(nx); s[i].pi.resize(nx);
        std::memcpy(s[i].x.data(), x + (b * (N + 1) + i) * nx, nx * sizeof(double));
        std::memcpy(s[i].pi.data(), pi + (b * (N + 1) + i) * nx, nx * sizeof(double));
        if (i == 0 && !ric) s[0].pi.setZero();   // pi[0] exists only together with the Riccati outputs (srbd_b200.h): not stale data
        if (ric) {
          s[i].P.resize(nx, nx); s[i].p.resize(nx);
          std::memcpy(s[i].P.data(), P + (b * (N + 1) + i) * nx * nx, nx * nx * sizeof(double));
          std::memcpy(s[i].p.data(), p + (b * (N + 1) + i) * nx, nx * sizeof(double));
        }
        if (i < N) {
          s[i].u.resize(nu);
          std::memcpy(s[i].u.data(), u + (b * N + i) * nu, nu * sizeof(double));
          if (ric) {
            s[i].K.resize(nu, nx); s[i].k.resize(nu);
            std::memcpy(s[i].K.data(), K + (b * N + i) * nu * nx, nu * nx * sizeof(double));
            std::memcpy(s[i].k.data(), k + (b * N + i) * nu, nu * sizeof(double));
          }
        }
      }
      status[b] = (st[b] >= 0 && st[b] <= 3) ? static_cast<HpipmStatus>(st[b]) : HpipmStatus::UnknownFailure;
    }
    });
    if (prof) {
      detail::FacadeProfile::get().wait += t1_ - t0_;
      detail::FacadeProfile::get().scatter += detail::FacadeProfile::now() - t1_;
    }
    // statistics of the (last) QP: iter, 4 max residuals, rows 0..iter+1 of the 18-column table (:376-403)
    const size_t bl = Bz - 1;
    solver_statistics_.iter = it[bl];
    solver_statistics_.max_res_stat = rm[4 * bl + 0]; solver_statistics_.max_res_eq = rm[4 * bl + 1];
    solver_statistics_.max_res_ineq = rm[4 * bl + 2]; solver_statistics_.max_res_comp = rm[4 * bl + 3];
    solver_statistics_.clear();
    if (stat && keep_batch_stats_) { batch_tab_.assign(tab, tab + Bz * static_cast<size_t>(rows) * SRBD_STAT_M); batch_tab_rows_ = rows; }
    else { batch_tab_.clear(); batch_tab_rows_ = 0; }
    if (stat) {
      auto cols = solver_statistics_.columns();
      for (int i = 0; i <= it[bl] + 1 && i < rows; ++i)
        for (int c = 0; c < SRBD_STAT_M; ++c) cols[c]->push_back(tab[(bl * rows + i) * SRBD_STAT_M + c]);
    }
  }
};

}  // namespace hpipm
