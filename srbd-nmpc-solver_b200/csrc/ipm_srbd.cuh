// ipm_srbd.cuh — K3, throughput variant for QPs assembled by K2 (nx = nu = 12, 24 general rows on u only).
//
// Same algorithm, same constants and the same per-iteration sequence as ipm_solve.cuh (the generic kernel,
// which stays the path for arbitrary hpipm::OcpQp data); this variant exploits what K2 guarantees
// (NMPC_solver.cpp:300-309, SRBD_model.cpp:244-255):
//   * C = 0, S = 0, Q diagonal, and D = Ac is ONE constant 24x12 matrix made of two 12x6 blocks (one per
//     contact)                       -> Ac lives in shared memory; D^T Gamma D = sum_g Gamma_g W_g with the 21
//                                       lower-triangle products W_g of each row's 6-vector precomputed per CTA;
//   * only the lower side of the rows is active (ug masked), no box constraints, cold start.
// Mapping: 4 independent warps per CTA, one QP per warp, persistent grid + atomic work counter.  Row i of the
// stage matrix [H~; g~^T] lives in the REGISTERS of lane i (m[24]), G rows too; broadcast operands come from
// shared memory four at a time (BLASFEO-style 4-wide panels, LDS.128): AL = G P, M += AL G^T, the Schur
// complement and the left-looking row-parallel Cholesky.
// Memory: the solve is a sequence of sweeps over the stages (S1 backward factorization, S2 forward rollout
// fused with dlam/dt and the step length, S3 mu_aff, S4 vector-only backward with the centering correction
// applied on the fly, S5 = S2, S6 variable update fused with the residuals).  Every sweep is SOFTWARE
// PIPELINED: while stage k is computed, the tiles of the next stage (BAbt record, P / L^-1 / Ls^T factors,
// the R block of RSQrq) stream into the other half of a shared-memory double buffer with cp.async, and the
// next stage's per-row vectors are prefetched into registers, so HBM/L2 latency overlaps the FP64 work.
#pragma once
#include <cuda_runtime.h>

#include "ipm_solve.cuh"
#include "srbd_model.cuh"

namespace srbd {

struct SrbdIpmParams {
  int B, N;
  srbd_ipm_args a;
  const double* babt;   // [B][N][336]
  const double* rsq;    // [B][N+1][672]
  const double* d;      // [B][N+1][48]   (lower part used)
  const double* dmask;  // [B][N+1][48]
  const double* x0;     // [B][12]
  const ModelDev* model;
  double* ws;           // [gridDim.x*4][ws_size]
  int ws_size;
  int* counter;
  double *sol_x, *sol_u, *sol_pi, *sol_lam, *sol_t;
  int* iter;
  int* status;
  double* res_max;
  srbd_batch_stats* bstats;
};

namespace v2 {
constexpr int kWarps = 4;
// per-stage workspace block (doubles)
constexpr int oZ = 0, oDZ = 24, oRG = 48, oLAM = 72, oT = 96, oDLAM = 120, oDT = 144, oRD = 168, oRM = 192, oRMB = 216,
              oPI = 240, oDPI = 252, oRB = 264, oPV = 276, oLV = 288, oP = 300, oLI = 444, oLST = 588, kStage = 732;
// shared memory (doubles)
constexpr int kGP = 52;            // padded panel stride of the BAbt tile (4 rows x 12 cols + 4)
constexpr int kLT = 26;            // row stride of L^T
constexpr int kW2 = 22;            // row stride of W (21 lower-triangle products of a constraint row's 6-vector)
constexpr int sAC = 0;             // [24][12] constraint Jacobian, row-major
constexpr int sW = sAC + 288;      // [24][22]
constexpr int kCtaShared = sW + 24 * kW2;
constexpr int kGT = 7 * kGP;       // 364: one BAbt tile
constexpr int kFT = 432;           // [P 144 | Linv 144 | Ls^T 144]
constexpr int kRT = 132;           // [R lower-panel prefixes 96 | Q diag 12 | rq row 24]
constexpr int wG0 = 0, wG1 = kGT;
constexpr int wF0 = 2 * kGT, wF1 = wF0 + kFT;
constexpr int wR0 = wF1 + kFT, wR1 = wR0 + kRT;
constexpr int wS = wR1 + kRT;      // 44
constexpr int wQX = wS + 44;       // 24 Gamma
constexpr int wqx = wQX + 24;      // 24 gamma
constexpr int wSG = wqx + 24;      // 24 gradient
constexpr int wSX = wSG + 24;      // 24 z of the current stage
constexpr int wXN = wSX + 24;      // 12 x_{k+1} / pi_k
constexpr int wT = wXN + 12;       // 12 t / lv
constexpr int wPV = wT + 12;       // 12 p_{k+1}
constexpr int wDI = wPV + 12;      // 12 1/L_jj
constexpr int wLAM = wDI + 12;     // 24 lam (residual sweep)
constexpr int kWarpShared = wLAM + 24;
constexpr int kSmemBytes = (kCtaShared + kWarps * kWarpShared) * 8;
// in the factorization sweep the factor buffers are free: L^T and the running P_{k+1} live there
constexpr int wLT = wF0;           // [12][26] = 312 <= 432
constexpr int wP = wF1;            // [13][12] = 156 <= 432
}  // namespace v2

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

struct SrbdSolver {
  const SrbdIpmParams& p;
  int lane, q, N;
  double* W;          // workspace of this warp
  const double* cAc;  // shared: Ac [24][12]
  const double* cW;   // shared: W [24][22]
  double* sm;         // this warp's shared block
  const double* sG;   // current BAbt tile
  const double* sF;   // current factor tile
  const double* sR;   // current R tile

  __device__ SrbdSolver(const SrbdIpmParams& p_, double* cta, double* warp_sm, int warp_global)
      : p(p_), lane(threadIdx.x & 31), q(0), N(p_.N) {
    W = p.ws + (size_t)warp_global * p.ws_size;
    cAc = cta + v2::sAC;
    cW = cta + v2::sW;
    sm = warp_sm;
    sG = sm; sF = sm + v2::wF0; sR = sm + v2::wR0;
  }
  __device__ __forceinline__ double* ws(int k, int off) const { return W + (size_t)k * v2::kStage + off; }
  __device__ __forceinline__ const double* gBAbt(int k) const { return p.babt + ((size_t)q * N + k) * 336; }
  __device__ __forceinline__ const double* gRSQ(int k) const { return p.rsq + ((size_t)q * (N + 1) + k) * 672; }
  __device__ __forceinline__ const double* gD(int k) const { return p.d + ((size_t)q * (N + 1) + k) * 48; }
  __device__ __forceinline__ const double* gMask(int k) const { return p.dmask + ((size_t)q * (N + 1) + k) * 48; }

  // ---- asynchronous tile prefetch (cp.async, no registers) -------------------------------------------
  // BAbt record (7 panels of 48 doubles) -> padded panels (bank-conflict-free row access)
  __device__ __forceinline__ void prefetch_G(int k, int b) {
    const double* src = gBAbt(k);
    double* dst = sm + (b ? v2::wG1 : v2::wG0);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int c = lane + 32 * i;
      if (c < 168) {
        const int pnl = c / 24, o = c - pnl * 24;
        cp_async16(dst + pnl * v2::kGP + 2 * o, src + 2 * c);
      }
    }
  }
  // P_{kP} and Linv | Ls^T of stage kL (contiguous 288 doubles in the workspace)
  __device__ __forceinline__ void prefetch_F(int kP, int kL, int b) {
    double* dst = sm + (b ? v2::wF1 : v2::wF0);
    const double* Ps = ws(kP, v2::oP);
    const double* Ls = ws(kL, v2::oLI);
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const int c = lane + 32 * i;
      if (c < 72) cp_async16(dst + 2 * c, Ps + 2 * c);
      else if (c < 216) cp_async16(dst + 2 * c, Ls + 2 * (c - 72));
    }
  }
  // R block of RSQrq (rows 0..11, lower: prefixes of panels 0..2), the diagonal of Q, the gradient row n
  __device__ __forceinline__ void prefetch_R(int k, int b) {
    const double* src = gRSQ(k);
    double* dst = sm + (b ? v2::wR1 : v2::wR0);
    const int n = (k < N ? 12 : 0) + (k > 0 ? 12 : 0);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = lane + 32 * i;
      if (c < 48) {
        const int pnl = c < 8 ? 0 : (c < 24 ? 1 : 2);
        const int o = c - (pnl == 0 ? 0 : (pnl == 1 ? 8 : 24));
        cp_async16(dst + (pnl == 0 ? 0 : (pnl == 1 ? 16 : 48)) + 2 * o, src + pnl * 96 + 2 * o);
      } else if (c < 60) {
        const int i2 = 12 + (c - 48);
        if (n == 24) cp_async8(dst + 96 + (c - 48), src + pm_index(i2, i2, 24));
      } else if (c < 84) {
        const int cc = c - 60;
        if (cc < n) cp_async8(dst + 108 + cc, src + pm_index(n, cc, 24));
      }
    }
  }
  __device__ __forceinline__ double Gel(int i, int l) const { return sG[(i >> 2) * v2::kGP + 4 * l + (i & 3)]; }
  // R(i,c), i >= c, i < 12
  __device__ __forceinline__ double Rel(int i, int c) const {
    const int pnl = i >> 2;
    return sR[(pnl == 0 ? 0 : (pnl == 1 ? 16 : 48)) + 4 * c + (i & 3)];
  }
  __device__ __forceinline__ void set_bufs(int b) {
    sG = sm + (b ? v2::wG1 : v2::wG0);
    sF = sm + (b ? v2::wF1 : v2::wF0);
    sR = sm + (b ? v2::wR1 : v2::wR0);
  }

  // ------------------------------------------------------------------------------------------------
  // S1: backward Riccati factorization sweep
  // ------------------------------------------------------------------------------------------------
  struct S1v { double mk, lam, t, rm, rd, rg, rb; };
  __device__ __forceinline__ S1v load_s1(int k) const {
    const int lc = lane < 24 ? lane : 0, l12 = lane < 12 ? lane : 0;
    S1v v;
    v.mk = __ldg(gMask(k) + lc); v.lam = ws(k, v2::oLAM)[lc]; v.t = ws(k, v2::oT)[lc];
    v.rm = ws(k, v2::oRM)[lc]; v.rd = ws(k, v2::oRD)[lc]; v.rg = ws(k, v2::oRG)[lc]; v.rb = ws(k, v2::oRB)[l12];
    return v;
  }
  __device__ void sweep_factor() {
    const double reg = p.a.reg_prim;
    double* sP = sm + v2::wP;
    double* sLt = sm + v2::wLT;
    // start streaming stage N-1 while stage N is handled
    prefetch_G(N - 1, 0);
    prefetch_R(N - 1, 0);
    S1v cur = load_s1(N - 1);
    // ---- stage N: P_N = Q_N + reg I, p_N = rg_N ------------------------------------------------------
    {
      const double* rs = gRSQ(N);
      if (lane < 12) {
#pragma unroll
        for (int c = 0; c < 12; ++c) {
          if (c <= lane) {
            double v = __ldg(rs + pm_index(lane, c, 24));
            if (c == lane) v += reg;
            sP[lane * 12 + c] = v;
            sP[c * 12 + lane] = v;
          }
        }
        sP[144 + lane] = ws(N, v2::oRG)[lane];
      }
      __syncwarp();
      for (int e = lane; e < 144; e += 32) ws(N, v2::oP)[e] = sP[e];
      if (lane < 12) ws(N, v2::oPV)[lane] = sP[144 + lane];
    }
    for (int k = N - 1; k >= 0; --k) {
      const int b = (N - 1 - k) & 1;
      const int nx = k > 0 ? 12 : 0, n = 12 + nx;  // nu = 12; gradient row lives in lane n
      cp_async_wait_all();
      __syncwarp();  // stage k's tiles have landed; every lane is done with the other buffers
      set_bufs(b);
      S1v nxt = cur;
      if (k > 0) {
        prefetch_G(k - 1, b ^ 1);
        prefetch_R(k - 1, b ^ 1);
        nxt = load_s1(k - 1);
      }
      // ---- own row of H (lower): R block rows from the tile, Q diagonal on the x rows ---------------------
      double m[24];
#pragma unroll
      for (int c = 0; c < 24; ++c) m[c] = 0.0;
      if (lane < 12) {
#pragma unroll
        for (int c = 0; c < 12; ++c)
          if (c <= lane) m[c] = Rel(lane, c);
      } else if (lane < n) {
#pragma unroll
        for (int c = 12; c < 24; ++c)
          if (c == lane) m[c] = sR[96 + (c - 12)];
      }
      if (lane < 24) {
        const double ti = 1.0 / cur.t;
        sm[v2::wQX + lane] = (ti * cur.lam) * cur.mk;
        sm[v2::wqx + lane] = (ti * (cur.rm - cur.lam * cur.rd)) * cur.mk;
      }
      double grow = (lane < n) ? cur.rg : 0.0;  // gradient entry c = lane
      const double rbv = (lane < 12) ? cur.rb : 0.0;
      __syncwarp();
      // ---- gradient: rg + D^T gamma (u part; a contact's rows only touch its own 6 inputs) ------------------
      if (lane < 12) {
        const int g0 = lane < 6 ? 0 : 12;
        double acc = 0.0;
#pragma unroll
        for (int g = 0; g < 12; ++g) acc += cAc[(g0 + g) * 12 + lane] * sm[v2::wqx + g0 + g];
        grow += acc;
      }
      if (lane < n) sm[v2::wSG + lane] = grow;
      // ---- D^T Gamma D: two 6x6 blocks, 21 lower-triangle entries each, spread over the warp -----------------
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int e2 = lane + 32 * r;
        if (e2 < 42) {
          const int leg = e2 >= 21 ? 1 : 0, e = e2 - 21 * leg;
          double acc = 0.0;
#pragma unroll
          for (int g = 0; g < 12; ++g) acc += sm[v2::wQX + 12 * leg + g] * cW[(12 * leg + g) * v2::kW2 + e];
          sm[v2::wS + e2] = acc;
        }
      }
      __syncwarp();
      if (lane == n) {
#pragma unroll
        for (int c = 0; c < 24; ++c) m[c] = (c < n) ? sm[v2::wSG + c] : 0.0;
      }
      if (lane < 12) {
        const int leg = lane < 6 ? 0 : 1, ii = lane - 6 * leg;
        const int base = 21 * leg + ii * (ii + 1) / 2;
#pragma unroll
        for (int c = 0; c < 12; ++c) {
          const int cc = c - 6 * leg;
          if (cc >= 0 && cc <= ii) m[c] += sm[v2::wS + base + cc];
        }
      }
      // ---- own G row (lane n: the rb row), AL = G P_{k+1} (+ p_{k+1} on the gradient row) ------------------------
      double g[12], al[12];
#pragma unroll
      for (int l = 0; l < 12; ++l) g[l] = (lane < n) ? Gel(lane, l) : 0.0;
#pragma unroll
      for (int l = 0; l < 12; ++l) {
        const double rbl = __shfl_sync(kFull, rbv, l);
        if (lane == n) g[l] = rbl;
      }
#pragma unroll
      for (int j = 0; j < 12; ++j) al[j] = (lane == n) ? sP[144 + j] : 0.0;
#pragma unroll
      for (int l = 0; l < 12; ++l) {
#pragma unroll
        for (int j4 = 0; j4 < 3; ++j4) {
          const double2 p01 = *reinterpret_cast<const double2*>(sP + l * 12 + 4 * j4);
          const double2 p23 = *reinterpret_cast<const double2*>(sP + l * 12 + 4 * j4 + 2);
          al[4 * j4 + 0] += g[l] * p01.x;
          al[4 * j4 + 1] += g[l] * p01.y;
          al[4 * j4 + 2] += g[l] * p23.x;
          al[4 * j4 + 3] += g[l] * p23.y;
        }
      }
      // ---- M += AL G^T, four columns (= four G rows of one panel) per pair of LDS.128 ------------------------
#pragma unroll
      for (int pc = 0; pc < 6; ++pc) {
        if (4 * pc < n) {
          double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
          for (int l = 0; l < 12; ++l) {
            const double2 g01 = *reinterpret_cast<const double2*>(sG + pc * v2::kGP + 4 * l);
            const double2 g23 = *reinterpret_cast<const double2*>(sG + pc * v2::kGP + 4 * l + 2);
            a0 += al[l] * g01.x; a1 += al[l] * g01.y; a2 += al[l] * g23.x; a3 += al[l] * g23.y;
          }
          m[4 * pc + 0] += a0; m[4 * pc + 1] += a1; m[4 * pc + 2] += a2; m[4 * pc + 3] += a3;
        }
      }
#pragma unroll
      for (int c = 0; c < 24; ++c)
        if (c == lane && lane < n) m[c] += reg;
      // ---- partial Cholesky of the 12 u-columns over rows 0..n (left-looking, row-parallel) -------------------
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        double s = m[j];
#pragma unroll
        for (int l = 0; l < j; ++l) s -= m[l] * sLt[l * v2::kLT + j];
        const double dj = __shfl_sync(kFull, s, j);
        const double inv = dj > 0.0 ? rsqrt(dj) : 0.0;
        const double val = (lane == j) ? dj * inv : s * inv;
        m[j] = val;
        if (lane >= j && lane <= n) sLt[j * v2::kLT + lane] = val;
        if (lane == j) sm[v2::wDI + j] = inv;
        __syncwarp();
      }
      // ---- Linv (column-parallel substitution on the identity), straight to the workspace -----------------------
      if (lane < 12) {
        double xc[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) xc[i] = (i == lane) ? sm[v2::wDI + i] : 0.0;
#pragma unroll
        for (int i = 1; i < 12; ++i) {
          double acc = 0.0;
#pragma unroll
          for (int l = 0; l < i; ++l) acc += sLt[l * v2::kLT + i] * xc[l];
          if (i > lane) xc[i] = -acc * sm[v2::wDI + i];
        }
        double* Li = ws(k, v2::oLI);
#pragma unroll
        for (int i = 0; i < 12; ++i) Li[i * 12 + lane] = xc[i];
        ws(k, v2::oLV)[lane] = sLt[lane * v2::kLT + n];  // lv = row n of L
      }
      // ---- Schur complement on the x rows (lanes 12..23) and the gradient row (lane n) --------------------------
      if (nx > 0) {
        double pr[12];
#pragma unroll
        for (int c = 0; c < 12; ++c) pr[c] = m[12 + c];
        if (lane >= 12 && lane <= n) {
#pragma unroll
          for (int c4 = 0; c4 < 3; ++c4) {
#pragma unroll
            for (int l = 0; l < 12; ++l) {
              const double2 l01 = *reinterpret_cast<const double2*>(sLt + l * v2::kLT + 12 + 4 * c4);
              const double2 l23 = *reinterpret_cast<const double2*>(sLt + l * v2::kLT + 12 + 4 * c4 + 2);
              pr[4 * c4 + 0] -= m[l] * l01.x; pr[4 * c4 + 1] -= m[l] * l01.y;
              pr[4 * c4 + 2] -= m[l] * l23.x; pr[4 * c4 + 3] -= m[l] * l23.y;
            }
          }
        }
        __syncwarp();  // everyone is done reading P_{k+1}
        if (lane >= 12 && lane < 24) {
          const int i = lane - 12;
#pragma unroll
          for (int c = 0; c < 12; ++c)
            if (c <= i) { sP[i * 12 + c] = pr[c]; sP[c * 12 + i] = pr[c]; }
        } else if (lane == 24) {
#pragma unroll
          for (int c = 0; c < 12; ++c) sP[144 + c] = pr[c];
        }
        __syncwarp();
        for (int e = lane; e < 144; e += 32) {
          ws(k, v2::oP)[e] = sP[e];
          ws(k, v2::oLST)[e] = sLt[(e / 12) * v2::kLT + 12 + (e % 12)];  // Ls^T[l][i]
        }
        if (lane < 12) ws(k, v2::oPV)[lane] = sP[144 + lane];
      }
      cur = nxt;
    }
    __syncwarp();
  }

  // ------------------------------------------------------------------------------------------------
  // S4: vector-only backward sweep (gradient recursion with the stored factors).  mode 1: centering
  // correction, mode 2: centering only; sm_ = sigma*mu (clamped by the caller)
  // ------------------------------------------------------------------------------------------------
  struct S4v { double mk, rmb, dt, dlam, lam, t, rd, rg, rb; };
  __device__ __forceinline__ S4v load_s4(int k) const {
    const int lc = lane < 24 ? lane : 0, l12 = lane < 12 ? lane : 0;
    S4v v;
    v.mk = __ldg(gMask(k) + lc); v.rmb = ws(k, v2::oRMB)[lc]; v.dt = ws(k, v2::oDT)[lc]; v.dlam = ws(k, v2::oDLAM)[lc];
    v.lam = ws(k, v2::oLAM)[lc]; v.t = ws(k, v2::oT)[lc]; v.rd = ws(k, v2::oRD)[lc]; v.rg = ws(k, v2::oRG)[lc];
    v.rb = ws(k, v2::oRB)[l12];
    return v;
  }
  __device__ void sweep_backvec(int mode, double sm_) {
    double* sPV = sm + v2::wPV;
    prefetch_G(N - 1, 0);
    prefetch_F(N, N - 1, 0);
    S4v cur = load_s4(N - 1);
    if (lane < 12) {
      const double v = ws(N, v2::oRG)[lane];
      ws(N, v2::oPV)[lane] = v;
      sPV[lane] = v;
    }
    for (int k = N - 1; k >= 0; --k) {
      const int b = (N - 1 - k) & 1;
      const int nx = k > 0 ? 12 : 0, n = 12 + nx;
      cp_async_wait_all();
      __syncwarp();
      set_bufs(b);
      S4v nxt = cur;
      if (k > 0) {
        prefetch_G(k - 1, b ^ 1);
        prefetch_F(k, k - 1, b ^ 1);
        nxt = load_s4(k - 1);
      }
      if (lane < 24) {
        double rm = cur.rmb;
        if (mode == 1) rm += cur.dt * cur.dlam;
        rm = (rm - sm_) * cur.mk;
        ws(k, v2::oRM)[lane] = rm;
        const double ti = 1.0 / cur.t;
        sm[v2::wqx + lane] = (ti * (rm - cur.lam * cur.rd)) * cur.mk;
      }
      if (lane < 12) sm[v2::wXN + lane] = cur.rb;
      double grow = (lane < n) ? cur.rg : 0.0;
      __syncwarp();
      if (lane < 12) {  // t = P_{k+1} rb + p_{k+1}
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 12; ++j) acc += sF[lane * 12 + j] * sm[v2::wXN + j];
        sm[v2::wT + lane] = acc + sPV[lane];
      }
      if (lane < 12) {
        const int g0 = lane < 6 ? 0 : 12;
        double acc = 0.0;
#pragma unroll
        for (int g = 0; g < 12; ++g) acc += cAc[(g0 + g) * 12 + lane] * sm[v2::wqx + g0 + g];
        grow += acc;
      }
      __syncwarp();
      if (lane < n) {
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < 12; ++l) acc += Gel(lane, l) * sm[v2::wT + l];
        sm[v2::wSG + lane] = grow + acc;
      }
      __syncwarp();
      if (lane < 12) {  // lv = Linv g_u
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 12; ++j)
          if (j <= lane) acc += sF[144 + lane * 12 + j] * sm[v2::wSG + j];
        sm[v2::wT + lane] = acc;
        ws(k, v2::oLV)[lane] = acc;
      }
      __syncwarp();
      if (lane < 12 && nx > 0) {  // p = g_x - Ls lv   (Ls[i][l] = LsT[l][i])
        double acc = sm[v2::wSG + 12 + lane];
#pragma unroll
        for (int l = 0; l < 12; ++l) acc -= sF[288 + l * 12 + lane] * sm[v2::wT + l];
        sPV[lane] = acc;
        ws(k, v2::oPV)[lane] = acc;
      }
      cur = nxt;
    }
    __syncwarp();
  }

  // ------------------------------------------------------------------------------------------------
  // S2/S5: forward rollout fused with dt / dlam and the step length
  // ------------------------------------------------------------------------------------------------
  struct S2v { double lv, pv, rb, mk, t, lam, rd, rm; };
  __device__ __forceinline__ S2v load_s2(int k) const {
    const int lc = lane < 24 ? lane : 0, l12 = lane < 12 ? lane : 0;
    S2v v;
    v.lv = ws(k, v2::oLV)[l12]; v.pv = ws(k + 1, v2::oPV)[l12]; v.rb = ws(k, v2::oRB)[l12];
    v.mk = __ldg(gMask(k) + lc); v.t = ws(k, v2::oT)[lc]; v.lam = ws(k, v2::oLAM)[lc];
    v.rd = ws(k, v2::oRD)[lc]; v.rm = ws(k, v2::oRM)[lc];
    return v;
  }
  __device__ void sweep_forward(double& ap, double& ad) {
    double a_p = 1.0, a_d = 1.0;
    prefetch_G(0, 0);
    prefetch_F(1, 0, 0);
    S2v cur = load_s2(0);
    for (int k = 0; k < N; ++k) {
      const int b = k & 1;
      const int nx = k > 0 ? 12 : 0, n = 12 + nx;
      cp_async_wait_all();
      __syncwarp();
      set_bufs(b);
      S2v nxt = cur;
      if (k + 1 < N) {
        prefetch_G(k + 1, b ^ 1);
        prefetch_F(k + 2, k + 1, b ^ 1);
        nxt = load_s2(k + 1);
      }
      // x part of this stage was left in wXN by the previous stage
      if (lane < 12 && nx > 0) {
        const double xv = sm[v2::wXN + lane];
        sm[v2::wSX + 12 + lane] = xv;
        ws(k, v2::oDZ)[12 + lane] = xv;
      }
      __syncwarp();
      if (lane < 12) {  // t = Ls^T x + lv
        double acc = 0.0;
        if (nx > 0) {
#pragma unroll
          for (int i = 0; i < 12; ++i) acc += sF[288 + lane * 12 + i] * sm[v2::wSX + 12 + i];
        }
        sm[v2::wT + lane] = acc + cur.lv;
      }
      __syncwarp();
      if (lane < 12) {  // u = -Linv^T t
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 12; ++j)
          if (j >= lane) acc += sF[144 + j * 12 + lane] * sm[v2::wT + j];
        sm[v2::wSX + lane] = -acc;
        ws(k, v2::oDZ)[lane] = -acc;
      }
      __syncwarp();
      if (lane < 12) {  // x+ = G^T z + rb
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < 24; ++i)
          if (i < n) acc += Gel(i, lane) * sm[v2::wSX + i];
        sm[v2::wXN + lane] = acc + cur.rb;
      }
      if (lane < 24) {  // v = D du ; dt, dlam, step lengths
        const int j0 = lane < 12 ? 0 : 6;
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) v += cAc[lane * 12 + j0 + j] * sm[v2::wSX + j0 + j];
        const double dt = (v - cur.rd) * cur.mk;
        const double dlam = (-(cur.lam * dt + cur.rm) / cur.t) * cur.mk;
        ws(k, v2::oDT)[lane] = dt;
        ws(k, v2::oDLAM)[lane] = dlam;
        if (dt < 0.0) a_p = fmin(a_p, -cur.t / dt);
        if (dlam < 0.0) a_d = fmin(a_d, -cur.lam / dlam);
      }
      __syncwarp();
      if (lane < 12) {  // dpi = P_{k+1} x+ + p_{k+1}
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 12; ++j) acc += sF[lane * 12 + j] * sm[v2::wXN + j];
        ws(k, v2::oDPI)[lane] = acc + cur.pv;
      }
      cur = nxt;
    }
    __syncwarp();
    if (lane < 12) ws(N, v2::oDZ)[lane] = sm[v2::wXN + lane];  // x_N
    ap = warp_min(a_p);
    ad = warp_min(a_d);
  }

  __device__ double mu_aff(double alpha, int nc_mask) {
    double acc = 0.0;
    if (lane < 24) {
#pragma unroll 5
      for (int k = 0; k < N; ++k)
        acc += (ws(k, v2::oLAM)[lane] + alpha * ws(k, v2::oDLAM)[lane]) * (ws(k, v2::oT)[lane] + alpha * ws(k, v2::oDT)[lane]);
    }
    return warp_sum(acc) / (double)nc_mask;
  }

  // ------------------------------------------------------------------------------------------------
  // S6: d_update_var_qp fused with d_ocp_qp_res_compute + inf norms.  With do_update the new iterate
  // (z,t) += sp*(dz,dt), (pi,lam) += sd*(dpi,dlam) is formed on the fly from the prefetched vectors (and
  // stored), so the update costs no extra pass; res_m is backed up (BACKUP_RES_M) in the same pass.
  // ------------------------------------------------------------------------------------------------
  struct S6v { double z, pi, lam, t, xn, lo, mk; };
  __device__ __forceinline__ S6v load_s6(int k, bool do_update, double sp, double sd) const {
    const int lc = lane < 24 ? lane : 0, l12 = lane < 12 ? lane : 0;
    S6v v;
    v.z = ws(k, v2::oZ)[lc];
    v.pi = 0.0; v.lam = 0.0; v.t = 1.0; v.xn = 0.0; v.lo = 0.0; v.mk = 0.0;
    const int xo = (k + 1 < N ? 12 : 0);
    if (k < N) {
      v.pi = ws(k, v2::oPI)[l12]; v.lam = ws(k, v2::oLAM)[lc]; v.t = ws(k, v2::oT)[lc];
      v.xn = ws(k + 1, v2::oZ)[xo + l12]; v.lo = __ldg(gD(k) + lc); v.mk = __ldg(gMask(k) + lc);
    }
    if (do_update) {
      v.z += sp * ws(k, v2::oDZ)[lc];
      if (k < N) {
        v.pi += sd * ws(k, v2::oDPI)[l12];
        v.xn += sp * ws(k + 1, v2::oDZ)[xo + l12];
        v.t += sp * ws(k, v2::oDT)[lc];
        v.lam += sd * ws(k, v2::oDLAM)[lc];
        if (p.a.t_lam_min == 2 && v.mk != 0.0) {
          v.t = v.t < p.a.t_min ? p.a.t_min : v.t;
          v.lam = v.lam < p.a.lam_min ? p.a.lam_min : v.lam;
        }
      }
    }
    return v;
  }
  __device__ void residuals(double res[4], double& mu, int nc_mask, bool do_update, double sp, double sd) {
    double ng_ = 0.0, nb_ = 0.0, nd_ = 0.0, nm_ = 0.0, smu = 0.0;
    prefetch_G(0, 0);
    prefetch_R(0, 0);
    S6v cur = load_s6(0, do_update, sp, sd);
    double pi_prev = 0.0;  // updated pi_{k-1}, held by lanes < 12
    for (int k = 0; k <= N; ++k) {
      const int b = k & 1;
      const int nu = k < N ? 12 : 0, nx = k > 0 ? 12 : 0, n = nu + nx;
      cp_async_wait_all();
      __syncwarp();
      set_bufs(b);
      S6v nxt = cur;
      if (k < N) {
        if (k + 1 < N) prefetch_G(k + 1, b ^ 1);
        prefetch_R(k + 1, b ^ 1);
        nxt = load_s6(k + 1, do_update, sp, sd);
      }
      // pi_{k-1} for the x rows (lane nu + l takes it from lane l)
      const double pim = __shfl_sync(kFull, pi_prev, (lane - nu) & 31);
      if (lane < n) sm[v2::wSX + lane] = cur.z;
      if (k < N) {
        if (lane < 12) sm[v2::wXN + lane] = cur.pi;
        if (lane < 24) sm[v2::wLAM + lane] = cur.lam;
      }
      if (do_update) {
        if (lane < n) ws(k, v2::oZ)[lane] = cur.z;
        if (k < N) {
          if (lane < 12) ws(k, v2::oPI)[lane] = cur.pi;
          if (lane < 24) { ws(k, v2::oT)[lane] = cur.t; ws(k, v2::oLAM)[lane] = cur.lam; }
        }
      }
      __syncwarp();
      if (lane < n) {
        // H z: the R block (k < N: rows 0..11) or Q_N's block (k == N) is dense lower; Q is diagonal, S = 0
        double acc = 0.0;
        const bool dense_rows = (k == N) || (lane < 12);
        if (dense_rows) {
#pragma unroll
          for (int j = 0; j < 12; ++j) {
            const int a = lane >= j ? lane : j, c = lane >= j ? j : lane;
            acc += Rel(a, c) * sm[v2::wSX + j];
          }
        } else {
          acc = sR[96 + (lane - 12)] * sm[v2::wSX + lane];
        }
        double r = acc + sR[108 + lane];
        if (k < N) {
          double a2 = 0.0;
#pragma unroll
          for (int l = 0; l < 12; ++l) a2 += Gel(lane, l) * sm[v2::wXN + l];
          r += a2;
        }
        if (k > 0 && lane >= nu) r -= pim;
        if (k < N && lane < 12) {  // J^T (lam_u - lam_l) = -D^T lam
          const int g0 = lane < 6 ? 0 : 12;
          double a3 = 0.0;
#pragma unroll
          for (int g = 0; g < 12; ++g) a3 += cAc[(g0 + g) * 12 + lane] * (0.0 - sm[v2::wLAM + g0 + g]);
          r += a3;
        }
        ws(k, v2::oRG)[lane] = r;
        ng_ = amax_nan(ng_, r);
      }
      if (k < N) {
        if (lane < 12) {
          double acc = 0.0;
#pragma unroll
          for (int i = 0; i < 24; ++i)
            if (i < n) acc += Gel(i, lane) * sm[v2::wSX + i];
          const double r = (acc + Gel(n, lane)) - cur.xn;
          ws(k, v2::oRB)[lane] = r;
          nb_ = amax_nan(nb_, r);
        }
        if (lane < 24) {
          const int j0 = lane < 12 ? 0 : 6;
          double v = 0.0;
#pragma unroll
          for (int j = 0; j < 6; ++j) v += cAc[lane * 12 + j0 + j] * sm[v2::wSX + j0 + j];
          const double rd = ((cur.lo - v) + cur.t) * cur.mk;
          const double rm = (cur.lam * cur.t) * cur.mk;
          ws(k, v2::oRD)[lane] = rd;
          ws(k, v2::oRM)[lane] = rm;
          ws(k, v2::oRMB)[lane] = rm;
          smu += rm;
          nd_ = amax_nan(nd_, rd);
          nm_ = amax_nan(nm_, rm);
        }
      }
      pi_prev = cur.pi;
      cur = nxt;
    }
    const double flag = warp_sum((ng_ != ng_ || nb_ != nb_ || nd_ != nd_ || nm_ != nm_) ? 1.0 : 0.0);
    res[0] = warp_max(ng_ == ng_ ? ng_ : 0.0);
    res[1] = warp_max(nb_ == nb_ ? nb_ : 0.0);
    res[2] = warp_max(nd_ == nd_ ? nd_ : 0.0);
    res[3] = warp_max(nm_ == nm_ ? nm_ : 0.0);
    if (flag > 0.0) res[0] = res[0] + __longlong_as_double(0x7ff8000000000000LL);
    mu = warp_sum(smu) / (double)nc_mask;
    __syncwarp();
  }

  __device__ __forceinline__ double shorten(double alpha) const {
    if (alpha < 1.0) {
      if (p.a.alpha_shorten == 0) return alpha * 0.995;
      return alpha * ((1.0 - alpha) * 0.99 + alpha * 0.9999999);
    }
    return alpha;
  }

  __device__ void solve_one(int qp) {
    q = qp;
    const srbd_ipm_args& a = p.a;
    // ---- d_ocp_qp_init_var (cold start): z = 0, pi = 0, t = max(thr0, -lo), lam = mu0/t (masked rows 0) ------
    int nmask = 0;
    for (int k = 0; k <= N; ++k) {
      if (lane < 24) ws(k, v2::oZ)[lane] = 0.0;
      if (k < N) {
        if (lane < 12) ws(k, v2::oPI)[lane] = 0.0;
        if (lane < 24) {
          const double lo = __ldg(gD(k) + lane), mk = __ldg(gMask(k) + lane);
          double tl = 0.0 - lo;
          tl = a.thr0 > tl ? a.thr0 : tl;
          ws(k, v2::oT)[lane] = tl;
          ws(k, v2::oLAM)[lane] = (a.mu0 / tl) * mk;
          nmask += (mk != 0.0);
          // the masked upper side never moves: t_u = max(thr0, up - v) with up = 0, v = 0; lam_u = 0
          const double tu = a.thr0 > 0.0 ? a.thr0 : 0.0;
          p.sol_t[(size_t)q * N * 48 + k * 48 + 24 + lane] = tu;
          p.sol_lam[(size_t)q * N * 48 + k * 48 + 24 + lane] = 0.0;
        }
      }
    }
    const int nc_mask = warp_sum_i(nmask);
    __syncwarp();
    double res[4], mu;
    residuals(res, mu, nc_mask, false, 0.0, 0.0);
    double alpha = 1.0;
    int kk = 0;
    for (; kk < a.iter_max && alpha > a.alpha_min &&
           (res[0] > a.tol_stat || res[1] > a.tol_eq || res[2] > a.tol_ineq || res[3] > a.tol_comp);
         ++kk) {
      sweep_factor();
      double ap, ad;
      sweep_forward(ap, ad);
      if (a.pred_corr == 1) {
        const double alpha_aff = fmin(ap, ad);
        const double mua = mu_aff(alpha_aff, nc_mask);
        const double tmp = mua / mu;
        const double sigma = tmp * tmp * tmp;
        double smv = sigma * mu;
        smv = smv > a.tau_min ? smv : a.tau_min;
        sweep_backvec(1, smv);
        sweep_forward(ap, ad);
        if (a.cond_pred_corr == 1) {
          const double muc = mu_aff(fmin(ap, ad), nc_mask);
          if (muc > a.cond_factor * mua) {
            sweep_backvec(2, sigma * mu);
            sweep_forward(ap, ad);
          }
        }
      }
      if (!a.split_step) {
        const double al = fmin(ap, ad);
        ap = al; ad = al;
      }
      alpha = fmin(ap, ad);
      residuals(res, mu, nc_mask, true, shorten(ap), shorten(ad));
    }
    int status;
    const bool nan = (res[0] != res[0]) || (mu != mu);
    if (kk == a.iter_max) status = 1;
    else if (alpha <= a.alpha_min) status = 2;
    else if (nan) status = 3;
    else status = 0;
    // ---- outputs ---------------------------------------------------------------------------------------------
    for (int k = 0; k <= N; ++k) {
      const int nu = k < N ? 12 : 0;
      if (lane < 12) {
        if (k == 0) p.sol_x[((size_t)q * (N + 1)) * 12 + lane] = p.x0[(size_t)q * 12 + lane];
        else {
          p.sol_x[((size_t)q * (N + 1) + k) * 12 + lane] = ws(k, v2::oZ)[nu + lane];
          p.sol_pi[((size_t)q * (N + 1) + k) * 12 + lane] = ws(k - 1, v2::oPI)[lane];
        }
        if (k < N) p.sol_u[((size_t)q * N + k) * 12 + lane] = ws(k, v2::oZ)[lane];
      }
      if (k < N && lane < 24) {
        p.sol_lam[(size_t)q * N * 48 + k * 48 + lane] = ws(k, v2::oLAM)[lane];
        p.sol_t[(size_t)q * N * 48 + k * 48 + lane] = ws(k, v2::oT)[lane];
      }
    }
    if (lane == 0) {
      p.iter[q] = kk;
      p.status[q] = status;
      for (int i = 0; i < 4; ++i) p.res_max[4 * (size_t)q + i] = res[i];
    }
    __syncwarp();
  }
};

__global__ void __launch_bounds__(128, 3) ipm_srbd_kernel(const SrbdIpmParams p) {
  extern __shared__ double2 smem2[];
  double* smem = reinterpret_cast<double*>(smem2);
  __shared__ int s_next[v2::kWarps];
  // CTA-shared constants: Ac and, per constraint row g, the 21 lower-triangle products of its 6-vector
  for (int i = threadIdx.x; i < 288; i += blockDim.x) smem[v2::sAC + i] = p.model->Ac[i];
  __syncthreads();
  for (int i = threadIdx.x; i < 24 * v2::kW2; i += blockDim.x) {
    const int g = i / v2::kW2, e = i - g * v2::kW2;
    double v = 0.0;
    if (e < 21) {
      int r = 0;
      while ((r + 1) * (r + 2) / 2 <= e) ++r;
      const int c = e - r * (r + 1) / 2, j0 = g < 12 ? 0 : 6;
      v = smem[v2::sAC + g * 12 + j0 + r] * smem[v2::sAC + g * 12 + j0 + c];
    }
    smem[v2::sW + i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  SrbdSolver S(p, smem, smem + v2::kCtaShared + warp * v2::kWarpShared, blockIdx.x * v2::kWarps + warp);
  long long it_sum = 0, solves = 0;
  int st_cnt[5] = {0, 0, 0, 0, 0};
  double rmax[4] = {0.0, 0.0, 0.0, 0.0};
  for (;;) {
    if (lane == 0) s_next[warp] = atomicAdd(p.counter, 1);
    __syncwarp();
    const int qp = s_next[warp];
    __syncwarp();
    if (qp >= p.B) break;
    S.solve_one(qp);
    if (lane == 0) {
      const int it = p.iter[qp], st = p.status[qp];
      it_sum += it;
      solves += 1;
      st_cnt[st < 0 || st > 4 ? 4 : st] += 1;
      atomicAdd((unsigned long long*)&p.bstats->iter_hist[it < SRBD_HIST_BINS ? it : SRBD_HIST_BINS - 1], 1ull);
      for (int i = 0; i < 4; ++i) rmax[i] = fmax(rmax[i], p.res_max[4 * (size_t)qp + i]);
    }
  }
  if (lane == 0 && solves > 0) {
    atomicAdd((unsigned long long*)&p.bstats->solves, (unsigned long long)solves);
    atomicAdd((unsigned long long*)&p.bstats->iter_sum, (unsigned long long)it_sum);
    for (int i = 0; i < 5; ++i)
      if (st_cnt[i]) atomicAdd((unsigned long long*)&p.bstats->status_count[i], (unsigned long long)st_cnt[i]);
    for (int i = 0; i < 4; ++i)
      atomicMax((unsigned long long*)&p.bstats->res_max[i], (unsigned long long)__double_as_longlong(rmax[i]));
  }
}

}  // namespace srbd
