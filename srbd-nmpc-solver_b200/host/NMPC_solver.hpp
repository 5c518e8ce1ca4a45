// NMPC_solver.hpp — B200-backed drop-in for the reference's NMPCSolver (NMPC_solver.h:20-108, NMPC_solver.cpp).
// Same public surface: NMPCSolver(config), initialize(), controlLoop(); the private steps keep their names
// (setupDynamics, setupReference, prepareQpStructures, solveQpProblems, checkConvergence / linearSearch) and
// are each one call into the C-ABI: K1+K2 (prepareQpStructures, NMPC_solver.cpp:276-314), K3
// (solveQpProblems, :316-330) and K4 (linearSearch, :149-274).  New: `batch` independent problems per object.
// yaml-cpp is not available here, so the values of config/mpc_option.yaml are fields of NMPCConfig (same names,
// same defaults); the constructor's path argument is accepted and ignored like in the reference (:23).
#pragma once
#include <chrono>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/srbd_b200.h"

struct NMPCConfig {  // config/mpc_option.yaml:1-18
  double Q[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 10};
  double Qf[12] = {0.5, 0.5, 0.5, 0.01, 0.01, 0.01, 100, 100, 100, 0.0, 0.0, 100.0};
  double R = 0.0001;
  double dt_MPC = 0.015;
  int horizon_MPC = 20;
  int sqp_max_loop = 15;
  double Lbody[3] = {0.541667, 0.516667, 1.0416667};
  int N_rep = 100;
  double mu_b = 0.1;
  double theta_b = 5.0;
  int batch = 1;          // new: independent problems solved together
  int assemble_mode = SRBD_BARRIER_SOFT;  // the reference's formulation (constraints as relaxed barriers)
  int device = 0;
  bool device_sqp_loop = true;  // new: the outer SQP loop runs on the device (false: host-driven, one read-back per iteration)
};

class NMPCSolver {
 public:
  explicit NMPCSolver(const std::string& /*config_file*/, const NMPCConfig& cfg = NMPCConfig()) : cfg_(cfg) { initialize(); }
  ~NMPCSolver() { if (ctx_) srbd_ctx_destroy(ctx_); }
  NMPCSolver(const NMPCSolver&) = delete;
  NMPCSolver& operator=(const NMPCSolver&) = delete;

  void initialize() {  // NMPC_solver.cpp:53-111
    N_ = cfg_.horizon_MPC;
    B_ = cfg_.batch;
    x_nmpc_.assign(static_cast<size_t>(B_) * (N_ + 1) * 12, 0.0);
    u_nmpc_.assign(static_cast<size_t>(B_) * N_ * 12, 100.0);
    x_ref_.assign(static_cast<size_t>(B_) * (N_ + 1) * 12, 0.0);
    x0_.assign(static_cast<size_t>(B_) * 12, 0.0);
    srbd_ipm_args_default(&args_);  // hpipm params, NMPC_solver.cpp:70-82
    args_.iter_max = 30; args_.alpha_min = 1e-8; args_.mu0 = 1e2;
    args_.tol_stat = args_.tol_eq = args_.tol_ineq = args_.tol_comp = 1e-4;
    args_.reg_prim = 1e-12; args_.warm_start = 0; args_.pred_corr = 1; args_.ric_alg = 0; args_.split_step = 1;
    if (ctx_) srbd_ctx_destroy(ctx_);
    ctx_ = nullptr;
    srbd_qp_dims d{N_, 12, 12, 0, 0, 24, 0};
    if (srbd_ctx_create(cfg_.device, B_, &d, nullptr, &ctx_) != 0)
      throw std::runtime_error("Failed to create the B200 context (no usable CUDA device)");
    check(srbd_set_ipm_args(ctx_, &args_), "srbd_set_ipm_args");
  }

  void controlLoop() {  // NMPC_solver.cpp:353-380 (state is NOT reset between repetitions, like the reference)
    const auto t0 = std::chrono::steady_clock::now();
    for (int nrep = 0; nrep < cfg_.N_rep; ++nrep) {
      setupDynamics();
      setupReference();
      last_sqp_iters_ = 0;
      check(srbd_upload_traj(ctx_, x_nmpc_.data(), u_nmpc_.data(), x_ref_.data(), x0_.data(), nullptr), "srbd_upload_traj");
      if (cfg_.device_sqp_loop) {
        // the whole SQP loop (:367-375) on the device: no host round trip between the iterations, every problem leaves
        // the loop at its own first "nmpc solve success" (srbd_sqp_solve)
        check(srbd_sqp_solve(ctx_, cfg_.assemble_mode, cfg_.sqp_max_loop), "srbd_sqp_solve");
        std::vector<int> it(B_), conv(B_);
        check(srbd_download_sqp_iters(ctx_, it.data()), "srbd_download_sqp_iters");
        check(srbd_download_sqp_state(ctx_, nullptr, conv.data(), nullptr), "srbd_download_sqp_state");
        bool all = true;
        for (int b = 0; b < B_; ++b) { if (it[b] > last_sqp_iters_) last_sqp_iters_ = it[b]; all = all && conv[b]; }
        if (all) std::cout << "nmpc solve success!" << std::endl;
      } else {
        for (int i = 0; i < cfg_.sqp_max_loop; ++i) {
          prepareQpStructures();
          solveQpProblems();
          if (checkConvergence()) break;
        }
      }
      check(srbd_download_traj(ctx_, x_nmpc_.data(), u_nmpc_.data()), "srbd_download_traj");
    }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "-----------------------\nTesting repetitions: " << cfg_.N_rep << "\nNMPC horizon: " << N_
              << "\nNMPC dt: " << cfg_.dt_MPC << "\nbatch: " << B_ << std::endl;
    std::cout << "Average NMPC solution time = " << ms / double(cfg_.N_rep) << "ms" << std::endl;
  }

  int lastSqpIterations() const { return last_sqp_iters_; }
  const std::vector<double>& x() const { return x_nmpc_; }
  const std::vector<double>& u() const { return u_nmpc_; }
  std::vector<double>& x0() { return x0_; }
  std::vector<double>& xref() { return x_ref_; }

 private:
  NMPCConfig cfg_;
  int N_ = 0, B_ = 0, last_sqp_iters_ = 0;
  srbd_ctx* ctx_ = nullptr;
  srbd_ipm_args args_{};
  std::vector<double> x_nmpc_, u_nmpc_, x_ref_, x0_;
  bool user_reference_ = false;

  void check(int rc, const char* what) { if (rc != 0) throw std::runtime_error(std::string(what) + " failed: " + srbd_last_error(ctx_)); }

  void setupDynamics() {  // NMPC_solver.cpp:332-339
    srbd_model_params p;
    srbd_model_params_default(&p, N_);
    p.mass = 15.0; p.dt = cfg_.dt_MPC;
    for (int i = 0; i < 9; ++i) p.inertia_inv[i] = 0.0;
    for (int i = 0; i < 3; ++i) p.inertia_inv[i + 3 * i] = 1.0 / cfg_.Lbody[i];
    for (int i = 0; i < 12; ++i) { p.Q[i] = cfg_.Q[i]; p.Qf[i] = double(N_) * cfg_.Qf[i]; }
    p.R = cfg_.R; p.mu_b = cfg_.mu_b; p.theta_b = cfg_.theta_b;
    check(srbd_set_model(ctx_, &p), "srbd_set_model");
  }
  void setupReference() {  // NMPC_solver.cpp:341-351
    const double x0[12] = {0, 0, 0, 0, 0, 0, 0, 0, 1.0, 0, 0, 0};
    const double xr[12] = {0, 0, 0.2, 0, 0, 0, 0.5, 0, 1.0, 0, 0, 0};
    for (int b = 0; b < B_; ++b) {
      for (int i = 0; i < 12; ++i) x0_[b * 12 + i] = x0[i];
      for (int k = 0; k <= N_; ++k)
        for (int i = 0; i < 12; ++i) x_ref_[(static_cast<size_t>(b) * (N_ + 1) + k) * 12 + i] = xr[i];
    }
  }
  void prepareQpStructures() {  // K1 + K2
    check(srbd_linearize(ctx_), "srbd_linearize");
    check(srbd_assemble(ctx_, cfg_.assemble_mode), "srbd_assemble");
  }
  void solveQpProblems() { check(srbd_qp_solve(ctx_), "srbd_qp_solve"); }  // K3 (status ignored like :320)
  bool checkConvergence() { return linearSearch(); }
  bool linearSearch() {  // K4; "nmpc solve success" when every problem of the batch converged
    check(srbd_line_search(ctx_), "srbd_line_search");
    std::vector<int> conv(B_);
    check(srbd_download_sqp_state(ctx_, nullptr, conv.data(), nullptr), "srbd_download_sqp_state");
    ++last_sqp_iters_;
    for (int c : conv) if (!c) return false;
    std::cout << "nmpc solve success!" << std::endl;
    return true;
  }
};
