// aux_kernels.cuh — pack (d_ocp_qp_set_all analog), K4 (filter line search) and the FP64 peak probe.
#pragma once
#include <cuda_runtime.h>

#include "ipm_solve.cuh"
#include "srbd_model.cuh"

namespace srbd {

// ---------------------------------------------------------------------------------------------------
// pack: column-major hpipm-cpp OcpQp fields -> packed panel-major stage records, with the x0 embedding
// b0 <- A0 x0 + b0, r0 <- S0 x0 + r0, nx[0] := 0, nbx[0] := 0
// (hpipm-cpp/src/ocp_qp_ipm_solver.cpp:128-130,225,236; d_ocp_qp_set_all hpipm_d_ocp_qp.h:90 and the
// six mask setters :124-162; upper bounds are stored negated like HPIPM's d vector).
// One warp per (QP, stage); every store is coalesced over the dense record.
// ---------------------------------------------------------------------------------------------------
struct PackParams {
  QpLayout L;
  int B;
  srbd_qp_host qp;  // DEVICE pointers (same shapes as the host view)
  double *babt, *rsq, *dct, *d, *dmask, *raw0;
  int raw0_stride;
  double* r0raw;  // [B][nu] the un-embedded r0 (the closed-loop driver re-embeds x0 every step) or null
  int d_stride;   // doubles between the D of consecutive (QP, stage) items: ng * nu, or 0 for ONE shared D (srbd_qp_upload_layout)
};

__global__ void __launch_bounds__(128) pack_kernel(const PackParams p) {
  const QpLayout& L = p.L;
  const int lane = threadIdx.x & 31;
  const long long it = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int S = L.N + 1;
  if (it >= (long long)p.B * S) return;
  const int q = (int)(it / S), k = (int)(it % S);
  const int N = L.N, nx = L.nx, nu = L.nu;
  const int nuk = stage_nu(L, k), nxk = stage_nx(L, k), n = nuk + nxk;
  const size_t qN = (size_t)q * N + k, qS = (size_t)q * S + k;
  const double* x0 = p.qp.x0 + (size_t)q * nx;
  // ---- BAbt ----
  if (k < N) {
    const double* A = p.qp.A + qN * nx * nx;
    const double* Bm = p.qp.Bm + qN * nx * nu;
    const double* b = p.qp.b + qN * nx;
    double* dst = p.babt + qN * L.babt_stride;
    const int cn = L.babt_cn;
    for (int e = lane; e < L.babt_stride; e += 32) {
      const int pnl = e / (4 * cn), rem = e - pnl * 4 * cn;
      const int j = rem >> 2, i = 4 * pnl + (rem & 3);
      double v = 0.0;
      if (j < nx) {
        if (i < nuk) v = Bm[j + nx * i];
        else if (i < n) v = A[j + nx * (i - nuk)];
        else if (i == n) {
          v = b[j];
          if (k == 0) {
            double s = 0.0;
            for (int l = 0; l < nx; ++l) s += A[j + nx * l] * x0[l];
            v = s + b[j];
          }
        }
      }
      dst[e] = v;
    }
  }
  // ---- RSQrq ----
  {
    const double* Q = p.qp.Q + qS * nx * nx;
    const double* qq = p.qp.q + qS * nx;
    const double* R = k < N ? p.qp.R + qN * nu * nu : nullptr;
    const double* r = k < N ? p.qp.r + qN * nu : nullptr;
    const double* Sm = (k < N && p.qp.S) ? p.qp.S + qN * nu * nx : nullptr;
    double* dst = p.rsq + qS * L.rsq_stride;
    const int cn = L.rsq_cn;
    for (int e = lane; e < L.rsq_stride; e += 32) {
      const int pnl = e / (4 * cn), rem = e - pnl * 4 * cn;
      const int j = rem >> 2, i = 4 * pnl + (rem & 3);
      double v = 0.0;
      if (j < n) {
        if (i < n) {
          if (i < nuk && j < nuk) v = R[i + nu * j];
          else if (i >= nuk && j >= nuk) v = Q[(i - nuk) + nx * (j - nuk)];
          else if (Sm) v = (i >= nuk) ? Sm[j + nu * (i - nuk)] : Sm[i + nu * (j - nuk)];
        } else if (i == n) {
          if (j < nuk) {
            v = r[j];
            if (k == 0 && Sm) {
              double s = 0.0;
              for (int l = 0; l < nx; ++l) s += Sm[j + nu * l] * x0[l];
              v = s + r[j];
            }
          } else {
            v = qq[j - nuk];
          }
        }
      }
      dst[e] = v;
    }
  }
  // ---- DCt, d, masks ----
  {
    const int nbu = k < N ? L.nbu : 0, nbx = k > 0 ? L.nbx : 0, nb = nbu + nbx;
    const int ng = stage_ng(L, k);
    double* dst = p.dct + qS * L.dct_stride;
    const int cn = L.dct_cn;
    const double* Dm = (k < N && ng > 0) ? p.qp.D + qN * p.d_stride : nullptr;
    const double* Cm = nullptr;
    int ldc = L.ng;
    if (ng > 0) {
      if (k < N) Cm = (k > 0 && p.qp.C) ? p.qp.C + qN * L.ng * nx : nullptr;  // C0 is dropped (nx[0] := 0)
      else { Cm = p.qp.CN + (size_t)q * L.ngN * nx; ldc = L.ngN; }
    }
    for (int e = lane; e < L.dct_stride; e += 32) {
      const int pnl = e / (4 * cn), rem = e - pnl * 4 * cn;
      const int g = rem >> 2, i = 4 * pnl + (rem & 3);
      double v = 0.0;
      if (g < ng && i < n) {
        if (i < nuk) v = Dm ? Dm[g + L.ng * i] : 0.0;
        else v = Cm ? Cm[g + ldc * (i - nuk)] : 0.0;
      }
      dst[e] = v;
    }
    double* dv = p.d + qS * L.d_stride;
    double* dk = p.dmask + qS * L.d_stride;
    for (int e = lane; e < L.d_stride; e += 32) {
      const bool lower = e < L.ncm;
      const int j = lower ? e : e - L.ncm;
      double v = 0.0, m = 0.0;
      if (j < nb + ng) {
        const double *lo, *up, *ml, *mu;
        size_t o;
        if (j < nbu) { o = qN * L.nbu + j; lo = p.qp.lbu; up = p.qp.ubu; ml = p.qp.lbu_mask; mu = p.qp.ubu_mask; }
        else if (j < nb) { o = qS * L.nbx + (j - nbu); lo = p.qp.lbx; up = p.qp.ubx; ml = p.qp.lbx_mask; mu = p.qp.ubx_mask; }
        else if (k < N) { o = qN * L.ng + (j - nb); lo = p.qp.lg; up = p.qp.ug; ml = p.qp.lg_mask; mu = p.qp.ug_mask; }
        else { o = (size_t)q * L.ngN + (j - nb); lo = p.qp.lgN; up = p.qp.ugN; ml = p.qp.lgN_mask; mu = p.qp.ugN_mask; }
        v = lower ? lo[o] : -up[o];
        const double* mk = lower ? ml : mu;
        m = mk ? (mk[o] != 0.0 ? 1.0 : 0.0) : 1.0;
      }
      dv[e] = v;
      dk[e] = m;
    }
  }
  // ---- raw stage-0 blocks for the facade's stage-0 reconstruction ----
  if (k == 0) {
    double* raw = p.raw0 + (size_t)q * p.raw0_stride;
    const int oA = 0, oB = nx * nx, ob = oB + nx * nu, oS = ob + nx, oQ = oS + nu * nx, oq = oQ + nx * nx;
    const size_t q0N = (size_t)q * N, q0S = (size_t)q * S;
    for (int e = lane; e < nx * nx; e += 32) { raw[oA + e] = p.qp.A[q0N * nx * nx + e]; raw[oQ + e] = p.qp.Q[q0S * nx * nx + e]; }
    for (int e = lane; e < nx * nu; e += 32) {
      raw[oB + e] = p.qp.Bm[q0N * nx * nu + e];
      raw[oS + e] = p.qp.S ? p.qp.S[q0N * nu * nx + e] : 0.0;
    }
    for (int e = lane; e < nx; e += 32) { raw[ob + e] = p.qp.b[q0N * nx + e]; raw[oq + e] = p.qp.q[q0S * nx + e]; }
    if (p.r0raw)
      for (int e = lane; e < nu; e += 32) p.r0raw[(size_t)q * nu + e] = p.qp.r[q0N * nu + e];
  }
}

// ---------------------------------------------------------------------------------------------------
// Structure detection for QP-level uploads with the SRBD dimensions (nx = nu = 12, 24 general rows, no boxes): does the
// batch have the structure K2 guarantees and the tensor-core variant of K3 relies on (ipm_srbd.cuh)?
//   S = 0, C = 0 (stages >= 1; C0 is dropped anyway), Q_k = one constant diagonal matrix for 1 <= k < N, D_k = one
//   constant matrix made of two 12 x 6 blocks, the upper side of every row masked (ug_mask = 0).
// R_k, Q_N (symmetric, lower triangle used), r, q, lg, lg_mask, A, B, b are free.  One warp per (QP, stage): any violation
// raises *bad; the same warp writes the compact stage record (srbd_model.cuh: R tile, gradient row, lg, lg mask) the
// variant reads, and item (0, 0) fills the model block (Ac = D, diag(Q)).  No host round trip: both K3 kernels are
// launched, each gated on *bad (capi.cu).
// ---------------------------------------------------------------------------------------------------
struct DetectParams {
  int B, N;
  srbd_qp_host qp;   // DEVICE pointers
  double* srec;      // [B][N+1][kSrec]
  ModelDev* model;
  int* bad;
  int d_stride;      // 288, or 0 for one shared D
};

__global__ void __launch_bounds__(128) detect_srbd_kernel(const DetectParams p) {
  const int lane = threadIdx.x & 31;
  const long long it = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int S = p.N + 1, N = p.N;
  if (it >= (long long)p.B * S) return;
  const int q = (int)(it / S), k = (int)(it % S);
  const size_t qN = (size_t)q * N + k, qS = (size_t)q * S + k;
  bool bad = false;
  if (k < N) {
    if (p.qp.S)
      for (int e = lane; e < 144; e += 32) bad |= p.qp.S[qN * 144 + e] != 0.0;
    if (p.qp.C && k > 0)
      for (int e = lane; e < 288; e += 32) bad |= p.qp.C[qN * 288 + e] != 0.0;
    const double* D = p.qp.D + qN * p.d_stride;   // column-major 24 x 12: D(g, j) at g + 24 j
    for (int e = lane; e < 288; e += 32) {
      const int g = e % 24, j = e / 24;
      const double v = D[e];
      bad |= v != p.qp.D[e];                               // constant over stages and QPs
      bad |= v != 0.0 && (j / 6) != (g / 12);              // two 12 x 6 blocks
    }
    if (!p.qp.ug_mask) bad = true;                         // no mask array = every row active on both sides
    else if (lane < 24) bad |= p.qp.ug_mask[qN * 24 + lane] != 0.0;
  }
  if (k > 0 && k < N) {
    const double* Q = p.qp.Q + qS * 144;
    const double* Qr = p.qp.Q + 144;                       // QP 0, stage 1
    for (int e = lane; e < 144; e += 32) {
      const int i = e % 12, j = e / 12;
      bad |= (i == j) ? (Q[e] != Qr[e]) : (Q[e] != 0.0);
    }
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicExch(p.bad, 1);
  // ---- compact stage record ------------------------------------------------------------------------------------------
  double* sr = p.srec + qS * kSrec;
  const double* H = k < N ? p.qp.R + qN * 144 : p.qp.Q + qS * 144;   // lower 12 x 12 block of rows 0..11
  for (int e = lane; e < 96; e += 32) {
    const int pnl = e < 16 ? 0 : (e < 48 ? 1 : 2), off = e - (pnl == 0 ? 0 : (pnl == 1 ? 16 : 48));
    const int j = off >> 2, i = 4 * pnl + (off & 3);
    sr[e] = H[i + 12 * j];
  }
  if (lane < 12) { sr[96 + lane] = 0.0; sr[132 + lane] = 0.0; }
  if (lane < 24) {
    double g = 0.0;
    if (k < N) g = lane < 12 ? p.qp.r[qN * 12 + lane] : (k > 0 ? p.qp.q[qS * 12 + lane - 12] : 0.0);
    else g = lane < 12 ? p.qp.q[qS * 12 + lane] : 0.0;
    sr[108 + lane] = g;
    sr[144 + lane] = k < N ? p.qp.lg[qN * 24 + lane] : 0.0;
    sr[168 + lane] = k < N ? (p.qp.lg_mask ? (p.qp.lg_mask[qN * 24 + lane] != 0.0 ? 1.0 : 0.0) : 1.0) : 0.0;
  }
  if (it == 0) {
    for (int e = lane; e < 288; e += 32) p.model->Ac[(e % 24) * 12 + e / 24] = p.qp.D[e];
    if (lane < 12) p.model->m.Q[lane] = N > 1 ? p.qp.Q[144 + 13 * lane] : 0.0;
    if (lane == 0) p.model->m.R = 0.0;
  }
}

// ---------------------------------------------------------------------------------------------------
// Closed-loop batched MPC (hpipm-cpp/examples/example_mpc.cpp:99-119, test/ocp_qp_ipm_solver.cpp:298-314): the QP data
// stay on the device, every step only the initial state changes.  mpc_embed_kernel redoes the x0 embedding of stage 0
// for the current plant state (same arithmetic as pack_kernel: b0 <- A0 x0 + b0, r0 <- S0 x0 + r0), mpc_plant_kernel
// applies u0 to the plant, x <- A x + B u0 + b, and records the closed-loop trajectory.  One thread per robot.
// ---------------------------------------------------------------------------------------------------
struct MpcParams {
  QpLayout L;
  int B, t;
  const double* raw0; int raw0_stride;
  const double* r0raw;
  double* xcur;            // [B][nx] plant state
  double* x0;              // [B][nx] initial state of the QP
  double *babt, *rsq;
  const double *A, *Bm, *b;  // plant: [B or 1][nx*nx], [nx*nu], [nx] column-major
  int plant_shared;
  const double* sol_u;     // [B][N][nu]
  const int *iter, *status;
  double *x_traj, *u_traj; // [steps+1][B][nx], [steps][B][nu]
  int *iter_traj, *status_traj;
};

__global__ void __launch_bounds__(128) mpc_embed_kernel(const MpcParams p) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= p.B) return;
  const QpLayout& L = p.L;
  const int nx = L.nx, nu = L.nu;
  const double* raw = p.raw0 + (size_t)q * p.raw0_stride;
  const double *A0 = raw, *b0 = raw + nx * nx + nx * nu, *S0 = b0 + nx;
  const double* x0 = p.xcur + (size_t)q * nx;
  double* ba = p.babt + (size_t)q * L.N * L.babt_stride;
  double* rs = p.rsq + (size_t)q * (L.N + 1) * L.rsq_stride;
  for (int j = 0; j < nx; ++j) {
    double s = 0.0;
    for (int l = 0; l < nx; ++l) s += A0[j + nx * l] * x0[l];
    ba[pm_index(nu, j, L.babt_cn)] = s + b0[j];
    p.x0[(size_t)q * nx + j] = x0[j];
    if (p.t == 0) p.x_traj[(size_t)q * nx + j] = x0[j];
  }
  for (int j = 0; j < nu; ++j) {
    double s = 0.0;
    for (int l = 0; l < nx; ++l) s += S0[j + nu * l] * x0[l];
    rs[pm_index(nu, j, L.rsq_cn)] = s + p.r0raw[(size_t)q * nu + j];
  }
}

__global__ void __launch_bounds__(128) mpc_plant_kernel(const MpcParams p) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= p.B) return;
  const QpLayout& L = p.L;
  const int nx = L.nx, nu = L.nu;
  const size_t po = p.plant_shared ? 0 : (size_t)q;
  const double *A = p.A + po * nx * nx, *Bm = p.Bm + po * nx * nu, *b = p.b + po * nx;
  const double* u0 = p.sol_u + (size_t)q * L.N * nu;
  double* x = p.xcur + (size_t)q * nx;
  double xn[kMaxNX];
  for (int i = 0; i < nx; ++i) {
    double s1 = 0.0, s2 = 0.0;
    for (int l = 0; l < nx; ++l) s1 += A[i + nx * l] * x[l];
    for (int l = 0; l < nu; ++l) s2 += Bm[i + nx * l] * u0[l];
    xn[i] = (s1 + s2) + b[i];
  }
  const size_t B = p.B;
  for (int i = 0; i < nx; ++i) {
    x[i] = xn[i];
    p.x_traj[((size_t)(p.t + 1) * B + q) * nx + i] = xn[i];
  }
  for (int i = 0; i < nu; ++i) p.u_traj[((size_t)p.t * B + q) * nu + i] = u0[i];
  p.iter_traj[(size_t)p.t * B + q] = p.iter[q];
  p.status_traj[(size_t)p.t * B + q] = p.status[q];
}

// ---------------------------------------------------------------------------------------------------
// K4: filter line search (NMPCSolver::linearSearch, NMPC_solver.cpp:149-274).  One warp per QP, one
// lane per stage (strided for N+1 > 32); merit / violation / directional derivative are warp sums.
// alpha is carried per QP across calls exactly like the member alpha_ (NMPC_solver.h:104).
// ---------------------------------------------------------------------------------------------------
struct LsParams {
  int B, N;
  int mode;              // assemble mode of the QP that produced the step: in SRBD_HARD_INEQ only the rows K2 kept as a
                         // relaxed barrier (row_soft_in_hard_mode) enter the merit function and its gradient
  double* x;             // [B][N+1][12] in/out
  double* u;             // [B][N][12]   in/out
  const double* xref;
  const uint8_t* contact;
  const double* dx;      // QP solution x [B][N+1][12]
  const double* du;      // [B][N][12]
  double* alpha;         // [B]
  int* converged;        // [B]
  double* merit;         // [B][3] phi, dphi, theta
  // device-side SQP loop (srbd_sqp_solve): a QP whose converged flag is already set is left alone (the reference leaves its
  // loop at the first "nmpc solve success", NMPC_solver.cpp:372-374); the others bump their iteration count and, if still
  // not converged, the counter that gates the next iteration's kernels
  int freeze;
  const int* run_gate;
  int* active_next;
  int* sqp_iter;         // [B]
};

// phi / theta (and optionally the cost gradient dotted with the step) of one stage at (x + a dx, u + a du)
__device__ __forceinline__ void stage_merit(const srbd_model_params& m, const double* Ac, const LsParams& p, int q,
                                            int k, double a, double& phi, double& theta, double* dphi) {
  const int N = p.N;
  double x[12], e[12];
  const size_t ox = ((size_t)q * (N + 1) + k) * 12;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    x[i] = p.x[ox + i] + a * p.dx[ox + i];
    e[i] = x[i] - p.xref[ox + i];
  }
  if (k == N) {
    double s = 0.0, dd = 0.0;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      s += e[i] * (m.Qf[i] * e[i]);
      dd += p.dx[ox + i] * (m.Qf[i] * e[i]);
    }
    phi += 0.5 * s;
    if (dphi) *dphi += dd;
    return;
  }
  double xn[12], u[12];
  const size_t ou = ((size_t)q * N + k) * 12;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    xn[i] = p.x[ox + 12 + i] + a * p.dx[ox + 12 + i];
    u[i] = p.u[ou + i] + a * p.du[ou + i];
  }
  double k1[12], k2[12], k3[12], k4[12], xt[12];
  const double dt = m.dt;
  srbd_f(m, x, u, k1, nullptr, nullptr, nullptr, nullptr);
#pragma unroll
  for (int i = 0; i < 12; ++i) xt[i] = x[i] + (0.5 * dt) * k1[i];
  srbd_f(m, xt, u, k2, nullptr, nullptr, nullptr, nullptr);
#pragma unroll
  for (int i = 0; i < 12; ++i) xt[i] = x[i] + (0.5 * dt) * k2[i];
  srbd_f(m, xt, u, k3, nullptr, nullptr, nullptr, nullptr);
#pragma unroll
  for (int i = 0; i < 12; ++i) xt[i] = x[i] + dt * k3[i];
  srbd_f(m, xt, u, k4, nullptr, nullptr, nullptr, nullptr);
  double ss = 0.0, s = 0.0, dd = 0.0;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const double xg = x[i] + (dt / 6.0) * (((k1[i] + 2.0 * k2[i]) + 2.0 * k3[i]) + k4[i]);
    const double f = xn[i] - xg;
    ss += f * f;
    s += e[i] * (m.Q[i] * e[i]);
    dd += p.dx[ox + i] * (m.Q[i] * e[i]);
  }
  theta += 0.5 * ss;
  phi += 0.5 * s;
  double bsum = 0.0, Ju[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) Ju[i] = 0.0;
  for (int g = 0; g < 24; ++g) {
    // hard rows are constraints of the QP, not cost terms: phi / dphi must describe the cost the QP step minimized
    if (p.mode == SRBD_HARD_INEQ && !row_soft_in_hard_mode(g)) continue;
    double v = 0.0;
#pragma unroll
    for (int j = 0; j < 12; ++j) v += Ac[g * 12 + j] * u[j];
    const int leg = g / 12, rr = g % 12;
    if (rr == 4) v += (p.contact ? (int)p.contact[((size_t)q * N + k) * 2 + leg] : 1) ? m.fmax : m.swing_fmax;
    else if (rr == 5) v += -m.fmin;
    double b, db;
    if (v > m.theta_b) {
      b = -m.mu_b * log(v);
      db = -m.mu_b / v;
    } else {
      const double t = (v - 2.0 * m.theta_b) / m.theta_b;
      b = 0.5 * m.mu_b * (t * t - 1.0) - m.mu_b * log(m.theta_b);
      db = m.mu_b * (v - 2.0 * m.theta_b) / (m.theta_b * m.theta_b);
    }
    bsum += b;
    if (dphi) {
#pragma unroll
      for (int j = 0; j < 12; ++j) Ju[j] += Ac[g * 12 + j] * db;
    }
  }
  double uu = 0.0;
#pragma unroll
  for (int i = 0; i < 12; ++i) uu += u[i] * (m.R * u[i]);
  phi += bsum + 0.5 * uu;
  if (dphi) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 12; ++i) t += p.du[ou + i] * (Ju[i] + m.R * u[i]);
    *dphi += dd + t;
  }
}

__global__ void __launch_bounds__(128) line_search_kernel(const LsParams p, const ModelDev* __restrict__ md) {
  __shared__ srbd_model_params sm;
  __shared__ double sAc[288];
  {
    const int nw = sizeof(srbd_model_params) / sizeof(double);
    const double* src = reinterpret_cast<const double*>(&md->m);
    double* dst = reinterpret_cast<double*>(&sm);
    for (int i = threadIdx.x; i < nw; i += blockDim.x) dst[i] = src[i];
    for (int i = threadIdx.x; i < 288; i += blockDim.x) sAc[i] = md->Ac[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (q >= p.B) return;
  if (p.run_gate && *p.run_gate == 0) return;
  if (p.freeze && p.converged[q]) return;
  const int N = p.N;
  const double theta_max = 1e-6, theta_min = 5e-10, eta = 1e-4, byta_phi = 1e-6, byta_theta = 1e-6,
               byta_alpha = 0.5, alpha_min = 1e-4;  // NMPC_solver.h:97-103
  double phi = 0.0, theta = 0.0, dphi = 0.0;
  for (int k = lane; k <= N; k += 32) stage_merit(sm, sAc, p, q, k, 0.0, phi, theta, &dphi);
  phi = warp_sum(phi); theta = warp_sum(theta); dphi = warp_sum(dphi);
  double alpha = p.alpha[q];
  while (alpha > alpha_min) {
    double phi_a = 0.0, theta_a = 0.0;
    for (int k = lane; k <= N; k += 32) stage_merit(sm, sAc, p, q, k, alpha, phi_a, theta_a, nullptr);
    phi_a = warp_sum(phi_a); theta_a = warp_sum(theta_a);
    bool accept = false;
    if (theta_a > theta_max) {
      accept = theta_a < (1.0 - byta_theta) * theta;
    } else if ((fmax(theta_a, theta) < theta_min) && (dphi < 0.0)) {
      accept = phi_a < phi + eta * alpha * dphi;
    } else {
      accept = (phi_a < phi - byta_phi * theta) || (theta_a < (1.0 - byta_theta) * theta);
    }
    if (accept) {
      for (int e = lane; e < (N + 1) * 12; e += 32) {
        const size_t o = (size_t)q * (N + 1) * 12 + e;
        p.x[o] = p.x[o] + alpha * p.dx[o];
      }
      for (int e = lane; e < N * 12; e += 32) {
        const size_t o = (size_t)q * N * 12 + e;
        p.u[o] = p.u[o] + alpha * p.du[o];
      }
      break;
    }
    alpha = byta_alpha * alpha;
  }
  if (lane == 0) {
    p.alpha[q] = alpha;
    const int conv = (dphi > -1e-3 && theta < 1e-6) ? 1 : 0;  // NMPC_solver.cpp:267
    p.converged[q] = conv;
    if (p.sqp_iter) p.sqp_iter[q] += 1;
    if (p.active_next && !conv) atomicAdd(p.active_next, 1);
    p.merit[3 * (size_t)q + 0] = phi;
    p.merit[3 * (size_t)q + 1] = dphi;
    p.merit[3 * (size_t)q + 2] = theta;
  }
}

// ---------------------------------------------------------------------------------------------------
// FP64 peak probe (the roofline denominator MEASURED_PEAKS.json does not carry): 8 independent
// DMMA m8n8k4 chains per warp (256 FMAs per instruction), 32 resident warps per SM.  DMMA and DFMA share ONE FP64
// datapath on B200 (scripts/microbench/fp64_pipes.cu: 36.9 vs 35.8 TFLOP/s); the DMMA form saturates it with fewer
// issue slots, so it is the stabler figure.  Run as a short burst BEFORE the timed loop (bench.py): after seconds of
// sustained load the board sits on its 1000 W cap and the same probe reads 30-35.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double seed) {
  double c0[8], c1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c0[i] = seed + i; c1[i] = i; }
  const double a = 1.0000001, b = 0.5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace srbd
