// SRBD_model.hpp — B200-backed drop-in for the reference's SRBDModel (dynamics/SRBD_model.h:13-106,
// dynamics/SRBD_model.cpp).  Same class, method names and argument meaning; the arithmetic of
// GetShootingDynamic / GetConstrain runs in the CUDA kernels K1 / K2 behind the C-ABI (include/srbd_b200.h).
// The single-call methods below exist for drop-in compatibility; throughput comes from the batched calls
// (srbd_upload_traj / srbd_linearize / srbd_assemble on B x N stages at once, see NMPC_solver.hpp).
// Not mirrored: GetContinuousDynamic — the reference only calls it from inside GetShootingDynamic
// (SRBD_model.cpp:174-177); it is internal to K1 and not part of the boundary.
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/srbd_b200.h"
#include "eigen_shim.hpp"

class SRBDModel {
 public:
  using scale_t = double;
  using Mat = Eigen::MatrixXd;
  using Vec = Eigen::VectorXd;

  SRBDModel() { srbd_model_params_default(&p_, 1); p_.mass = 0.0; p_.dt = 0.0; for (int i = 0; i < 9; ++i) p_.inertia_inv[i] = (i % 4 == 0); }
  ~SRBDModel() { if (ctx_) srbd_ctx_destroy(ctx_); }
  SRBDModel(const SRBDModel&) = delete;
  SRBDModel& operator=(const SRBDModel&) = delete;

  void SetFoot(const Vec& pr, const Vec& pl, const Mat& R0, const Mat& R1) {  // SRBD_model.cpp:28-34
    for (int i = 0; i < 3; ++i) { p_.foot_pos[i] = pr(i); p_.foot_pos[3 + i] = pl(i); }
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 3; ++i) { p_.foot_rot[i + 3 * j] = R0(i, j); p_.foot_rot[9 + i + 3 * j] = R1(i, j); }
    dirty_ = true;
  }
  void SetMass(scale_t m) { p_.mass = m; dirty_ = true; }
  void SetMPCdt(scale_t dt) { p_.dt = dt; dirty_ = true; }
  // takes the inertia and stores its inverse (SRBD_model.cpp:46-49); diagonal or general 3x3
  void SetInertia(const Mat& L) {
    const double a = L(0, 0), b = L(0, 1), c = L(0, 2), d = L(1, 0), e = L(1, 1), f = L(1, 2), g = L(2, 0), h = L(2, 1), i = L(2, 2);
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    const double inv[9] = {(e * i - f * h) / det, (f * g - d * i) / det, (d * h - e * g) / det,   // column 0
                           (c * h - b * i) / det, (a * i - c * g) / det, (b * g - a * h) / det,   // column 1
                           (b * f - c * e) / det, (c * d - a * f) / det, (a * e - b * d) / det};  // column 2
    for (int k = 0; k < 9; ++k) p_.inertia_inv[k] = inv[k];
    dirty_ = true;
  }
  Vec GetFoot(int N) { Vec v(3); for (int i = 0; i < 3; ++i) v(i) = p_.foot_pos[(N == 0 ? 0 : 3) + i]; return v; }
  Mat GetFootR(int N) { Mat R(3, 3); for (int k = 0; k < 9; ++k) R.data()[k] = p_.foot_rot[(N == 0 ? 0 : 9) + k]; return R; }

  // SRBD_model.cpp:143-235 -> K1 (one stage: B = 1, N = 1)
  void GetShootingDynamic(const Vec& x, const Vec& x_next, const Vec& u, Mat* pA, Mat* pB, Mat* pb, Mat* pf) {
    stage(x, x_next, u);
    check(srbd_linearize(ctx_), "srbd_linearize");
    double A[144], B[144], b[12], f[12];
    check(srbd_download_linearization(ctx_, A, B, b, f), "srbd_download_linearization");
    auto put = [](Mat* m, const double* src, int r, int c) { if (m) { m->resize(r, c); for (int k = 0; k < r * c; ++k) m->data()[k] = src[k]; } };
    put(pA, A, 12, 12); put(pB, B, 12, 12); put(pb, b, 12, 1); put(pf, f, 12, 1);
  }
  // SRBD_model.cpp:237-260 -> K2 (rows of the hard-inequality assembly: D = Ac, lg = -(Ac u + b))
  void GetConstrain(const Vec& u, Mat& Ac, Mat& f) {
    Vec x = Vec::Zero(12);   // (the constraint rows do not depend on the state; zero-initialised for real Eigen too)
    stage(x, x, u);
    check(srbd_linearize(ctx_), "srbd_linearize");
    check(srbd_assemble(ctx_, SRBD_HARD_INEQ), "srbd_assemble");
    double D[288], lg[24];
    check(srbd_download_qp(ctx_, nullptr, nullptr, nullptr, nullptr, nullptr, D, lg, nullptr), "srbd_download_qp");
    Ac.resize(24, 12); f.resize(24, 1);
    for (int k = 0; k < 288; ++k) Ac.data()[k] = D[k];
    for (int g = 0; g < 24; ++g) f.data()[g] = -lg[g];
  }
  const srbd_model_params& params() const { return p_; }
  srbd_model_params& params() { dirty_ = true; return p_; }

 private:
  srbd_model_params p_{};
  srbd_ctx* ctx_ = nullptr;
  bool dirty_ = true;
  void check(int rc, const char* what) { if (rc != 0) throw std::runtime_error(std::string(what) + " failed: " + srbd_last_error(ctx_)); }
  void stage(const Vec& x, const Vec& xn, const Vec& u) {
    if (x.size() != 12 || xn.size() != 12 || u.size() != 12) throw std::runtime_error("SRBDModel: x, x_next and u must have 12 entries");
    if (!ctx_) {
      srbd_qp_dims d{1, 12, 12, 0, 0, 24, 0};
      if (srbd_ctx_create(0, 1, &d, nullptr, &ctx_) != 0) throw std::runtime_error("srbd_ctx_create failed: no usable CUDA device");
      dirty_ = true;
    }
    if (dirty_) { check(srbd_set_model(ctx_, &p_), "srbd_set_model"); dirty_ = false; }
    double xs[24], zero[24] = {0};
    for (int i = 0; i < 12; ++i) { xs[i] = x(i); xs[12 + i] = xn(i); }
    check(srbd_upload_traj(ctx_, xs, u.data(), zero, xs, nullptr), "srbd_upload_traj");
  }
};
