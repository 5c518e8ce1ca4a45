// bench_facade.cpp — the reference's OWN boundary as the measured path: B SRBD QPs (N = 20, all stance, HARD_INEQ, the QP
// fields that NMPCSolver::prepareQpStructures hands to hpipm-cpp) through hpipm::OcpQpIpmSolver::solveBatch, host
// std::vector<OcpQp> in, std::vector<OcpQpSolution> out.  Prints one JSON object (bench.py embeds it as `facade_e2e`).
//   bench_facade [B = 4096] [reps = 3]
// The QP-level interface moves 228 KB per N = 20 QP across PCIe (A, B, Q, S, R, C, D of every stage) against
// 6 KB for the NMPC-level calls (trajectories in, K1 / K2 on the device): this figure is bounded by the interface, not by K3
// (host phases with SRBD_FACADE_PROFILE=1: flattening 230 MB of scattered Eigen blocks per 1024 QPs is 60 % of the time).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../NMPC_solver.hpp"
#include "../hpipm-cpp/hpipm-cpp.hpp"

using Eigen::MatrixXd;
using Eigen::VectorXd;
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
  const int B = argc > 1 ? std::atoi(argv[1]) : 4096, reps = argc > 2 ? std::atoi(argv[2]) : 3, N = 20;
  // ---- the QPs: K1 + K2 on the device (NMPC-level C-ABI), downloaded as hpipm-cpp fields -------------------------------
  srbd_qp_dims d{N, 12, 12, 0, 0, 24, 0};
  srbd_ctx* g = nullptr;
  if (srbd_ctx_create(0, B, &d, nullptr, &g) != 0) { std::printf("{\"error\": \"no CUDA device\"}\n"); return 2; }
  std::vector<double> x((size_t)B * (N + 1) * 12), u((size_t)B * N * 12), xr((size_t)B * (N + 1) * 12), x0((size_t)B * 12);
  unsigned long long s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) * (1.0 / 9007199254740992.0); };
  for (int b = 0; b < B; ++b) {
    double st[12], rf[12] = {0};
    const double lo[12] = {-.3, -.3, -.3, -.5, -.5, -.5, -.2, -.2, .8, -.5, -.5, -.5}, hi[12] = {.3, .3, .3, .5, .5, .5, .2, .2, 1.2, .5, .5, .5};
    for (int i = 0; i < 12; ++i) st[i] = lo[i] + (hi[i] - lo[i]) * rnd();
    rf[2] = -.3 + .6 * rnd(); rf[6] = -.5 + rnd(); rf[7] = -.5 + rnd(); rf[8] = 1.0;
    for (int i = 0; i < 12; ++i) x0[(size_t)b * 12 + i] = st[i];
    for (int k = 0; k <= N; ++k)
      for (int i = 0; i < 12; ++i) { x[((size_t)b * (N + 1) + k) * 12 + i] = st[i]; xr[((size_t)b * (N + 1) + k) * 12 + i] = rf[i]; }
    for (int k = 0; k < N; ++k) { u[((size_t)b * N + k) * 12 + 2] = 15.0 * 9.8 / 2; u[((size_t)b * N + k) * 12 + 8] = 15.0 * 9.8 / 2; }
  }
  srbd_upload_traj(g, x.data(), u.data(), xr.data(), x0.data(), nullptr);
  srbd_linearize(g); srbd_assemble(g, SRBD_HARD_INEQ);
  const size_t BN = (size_t)B * N, BS = (size_t)B * (N + 1);
  std::vector<double> A(BN * 144), Bm(BN * 144), bb(BN * 12), Q(BS * 144), S(BN * 144), R(BN * 144), q(BS * 12), r(BN * 12),
      D(BN * 288), lg(BN * 24), lgm(BN * 24);
  if (srbd_download_linearization(g, A.data(), Bm.data(), bb.data(), nullptr) != 0 ||
      srbd_download_qp(g, Q.data(), S.data(), R.data(), q.data(), r.data(), D.data(), lg.data(), lgm.data()) != 0) {
    std::printf("{\"error\": \"%s\"}\n", srbd_last_error(g)); return 2;
  }
  srbd_ctx_destroy(g);
  std::vector<std::vector<hpipm::OcpQp>> qps(B, std::vector<hpipm::OcpQp>(N + 1));
  std::vector<VectorXd> x0s(B, VectorXd(12));   // delta form: x0 - x_nmpc[0] = 0
  auto mat = [](const double* src, int rws, int cls) { MatrixXd m(rws, cls); for (int e = 0; e < rws * cls; ++e) m.data()[e] = src[e]; return m; };
  auto vec = [](const double* src, int n) { VectorXd v(n); for (int e = 0; e < n; ++e) v(e) = src[e]; return v; };
  for (int b = 0; b < B; ++b)
    for (int k = 0; k <= N; ++k) {
      hpipm::OcpQp& o = qps[b][k];
      o.Q = mat(Q.data() + ((size_t)b * (N + 1) + k) * 144, 12, 12); o.q = vec(q.data() + ((size_t)b * (N + 1) + k) * 12, 12);
      if (k == N) break;
      const size_t i = (size_t)b * N + k;
      o.A = mat(A.data() + i * 144, 12, 12); o.B = mat(Bm.data() + i * 144, 12, 12); o.b = vec(bb.data() + i * 12, 12);
      o.S = mat(S.data() + i * 144, 12, 12); o.R = mat(R.data() + i * 144, 12, 12); o.r = vec(r.data() + i * 12, 12);
      o.C = MatrixXd(24, 12); o.D = mat(D.data() + i * 288, 24, 12);
      o.lg = vec(lg.data() + i * 24, 24); o.ug = VectorXd(24); o.lg_mask = vec(lgm.data() + i * 24, 24); o.ug_mask = VectorXd(24);
    }
  hpipm::OcpQpIpmSolverSettings set;
  set.iter_max = 30; set.alpha_min = 1e-8; set.mu0 = 1e2; set.tol_stat = set.tol_eq = set.tol_ineq = set.tol_comp = 1e-8;
  set.reg_prim = 1e-12; set.warm_start = 0; set.pred_corr = 1; set.ric_alg = 0; set.split_step = 1;
  std::vector<std::vector<hpipm::OcpQpSolution>> sols(B);
  double best_fast = 1e30, best_full = 1e30;
  long it_sum = 0; int conv = 0;
  for (int full = 0; full < 2; ++full) {
    hpipm::OcpQpIpmSolver solver(set);
    solver.setOutputs(full != 0, full != 0);
    for (int rep = 0; rep < reps + 1; ++rep) {
      const double t0 = now_ms();
      const auto st = solver.solveBatch(x0s, qps, sols);
      const double dt = now_ms() - t0;
      if (rep == 0) { hpipm::detail::FacadeProfile::get() = hpipm::detail::FacadeProfile(); continue; }   // first call: context, pinned arena
      (full ? best_full : best_fast) = dt < (full ? best_full : best_fast) ? dt : (full ? best_full : best_fast);
      if (!full && rep == reps) { for (int v : solver.getBatchIterations()) it_sum += v; for (auto v : st) conv += v == hpipm::HpipmStatus::Success; }
    }
    if (hpipm::detail::FacadeProfile::on()) {
      const auto& pf = hpipm::detail::FacadeProfile::get();
      std::fprintf(stderr, "facade host phases, %s outputs, per call (ms): validate %.1f flatten %.1f enqueue %.1f wait %.1f scatter %.1f\n",
                   full ? "reference" : "x/u/pi", pf.validate / reps, pf.flatten / reps, pf.enqueue / reps, pf.wait / reps, pf.scatter / reps);
    }
  }
  // S and C of these QPs are all zero and D is one constant matrix: the facade sends neither S nor C and D once per chunk
  // (hpipm-cpp.hpp, submit)
  const double bytes_up = ((double)B * (N * (4 * 144 + 36 + 4 * 24) + 144 + 12 + 12) + 288.0 * ((B + 1023) / 1024)) * 8, bytes_down = (double)B * ((N + 1) * 24 + N * 12) * 8;
  std::printf("{\"qps\": %d, \"horizon\": %d, \"ms_per_batch\": %.3f, \"value\": %.1f, \"unit\": \"solves/s\", "
              "\"ms_per_batch_reference_outputs\": %.3f, \"value_reference_outputs\": %.1f, \"converged\": %d, \"iter_mean\": %.3f, "
              "\"h2d_bytes\": %.0f, \"d2h_bytes\": %.0f, "
              "\"how\": \"hpipm::OcpQpIpmSolver::solveBatch on host std::vector<OcpQp> (setOutputs(false,false): x, u, pi only; "
              "_reference_outputs: + P,p,K,k,pi[0] + statistics table like the reference's solve(); both run the tensor-core kernel); best of %d after one warm-up; includes "
              "flattening the Eigen fields into the pinned arena (all-zero S / C stay behind, the constant D travels once), H2D copy, pack + structure detection + K3, D2H copy, scattering "
              "into OcpQpSolution; batches of >= 2048 QPs are pipelined in chunks of 1024 over two pooled contexts\"}\n",
              B, N, best_fast, B / (best_fast * 1e-3), best_full, B / (best_full * 1e-3), conv, (double)it_sum / B, bytes_up, bytes_down, reps);
  return 0;
}
