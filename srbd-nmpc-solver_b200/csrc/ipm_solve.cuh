// ipm_solve.cuh — K3: the whole OCP-QP interior-point solve as ONE persistent sm_100a kernel.
//
// Replaces d_ocp_qp_ipm_solve and everything it calls (hpipm_d_ocp_qp_ipm.h:232-238,
// hpipm_d_ocp_qp_kkt.h:54-60, hpipm_d_ocp_qp_res.h:90-94, hpipm_d_core_qp_ipm_aux.h:44-62 of the
// reference's vendored headers; algorithm = SURVEY.md Appendix C, classical Riccati ric_alg = 0,
// which is what NMPC_solver.cpp:81 selects).
//
// Mapping: one warp owns one QP for its whole life (data-dependent trip count), warps pull QP
// indices from a global atomic counter (persistent grid, divergent iteration counts balance
// themselves).  Inside a stage the rows of the (nu+nx+1) x (nu+nx) Riccati matrix
// [H~ ; g~^T] live one-per-lane: the partial Cholesky / trsm panel (BLASFEO potrf_l_mn with the
// "+1 row" gradient trick) is row-parallel, the syrk/gemm pieces are row-parallel FP64 FMA loops
// with broadcast operands from shared memory, and every reduction (step length, duality gap,
// residual inf-norms) is a warp shuffle.  The stage matrix, the [G | AL] / DCt tile and P_{k+1}
// stay in shared memory; packed QP data is streamed from HBM/L2 with coalesced 16-byte loads.
#pragma once
#include <cuda_runtime.h>

#include "layout.cuh"

namespace srbd {

struct IpmParams {
  QpLayout L;
  srbd_ipm_args a;
  int B;
  const double* babt;
  const double* rsq;
  const double* dct;
  const double* d;
  const double* dmask;
  const double* x_init;  // [B][N+1][nx] or null
  const double* u_init;  // [B][N][nu] or null
  const double* x0;      // [B][nx]
  const double* raw0;    // [B][raw0_stride]: A0,B0,b0,S0,Q0,q0 (column-major) for the stage-0 reconstruction
  double* ws;            // [gridDim.x][ws_size]
  int* counter;
  double *sol_x, *sol_u, *sol_pi, *sol_lam, *sol_t;
  double *ric_P, *ric_p, *ric_K, *ric_k;  // optional
  double* ric_Lr0;  // optional (with ric_P): [B][nu*nu] column-major Cholesky factor Lr of stage 0 (d_ocp_qp_ipm_get_ric_Lr(.., 0, ..))
  int* iter;
  int* status;
  double* res_max;
  double* stat;  // [B][stat_rows][18] or null
  int stat_rows;
  srbd_batch_stats* bstats;
  // optional work list: solve only the QPs qlist[0 .. *qcount) (the rescue pass behind the SRBD variant, capi.cu)
  const int* qlist;
  const int* qcount;
  // device-side dispatch (QP-level uploads, capi.cu): run only if *gate == gate_value
  const int* gate;
  int gate_value;
  // Small launches (fewer CTAs than SMs: single QPs through the facade, the rescue list): the per-QP workspace lives in
  // the CTA's DYNAMIC SHARED MEMORY instead of global memory.  One warp alone pays a full L2 round trip at the top of
  // every phase of every stage (about sixty per stage and iteration: 14.1 ms for one N = 20 SRBD QP against 1.4 ms in
  // the tensor-core variant); with the iterate, the residuals and the factors on chip those are shared-memory loads.
  int ws_in_smem;
  // device-side SQP loop (srbd_sqp_solve): return at once if *run_gate == 0; skip QPs whose frozen[] flag is set
  const int* run_gate;
  const int* frozen;
};

// compile-time dimension policy (loops unroll, index math folds) ...
template <int NX, int NU, int NBX, int NBU, int NG, int NGN>
struct SDims {
  static constexpr int kNX = NX, kNU = NU, kNM = NX + NU;
  static constexpr int kNGM = NG > NGN ? NG : NGN;
  static constexpr int kNCM = (NBU + NBX + kNGM) > 0 ? (NBU + NBX + kNGM) : 1;
  __device__ explicit SDims(const QpLayout&) {}
  __device__ constexpr int nx() const { return NX; }
  __device__ constexpr int nu() const { return NU; }
  __device__ constexpr int nbx() const { return NBX; }
  __device__ constexpr int nbu() const { return NBU; }
  __device__ constexpr int ng() const { return NG; }
  __device__ constexpr int ngN() const { return NGN; }
  static bool matches(const QpLayout& L) {
    return L.nx == NX && L.nu == NU && L.nbx == NBX && L.nbu == NBU && L.ng == NG && L.ngN == NGN;
  }
};
// ... and the run-time fallback (any dims up to the compiled maxima)
struct DDims {
  static constexpr int kNX = kMaxNX, kNU = kMaxNU, kNM = kMaxN, kNGM = kMaxNG, kNCM = kMaxNC;
  int nx_, nu_, nbx_, nbu_, ng_, ngN_;
  __device__ explicit DDims(const QpLayout& L)
      : nx_(L.nx), nu_(L.nu), nbx_(L.nbx), nbu_(L.nbu), ng_(L.ng), ngN_(L.ngN) {}
  __device__ int nx() const { return nx_; }
  __device__ int nu() const { return nu_; }
  __device__ int nbx() const { return nbx_; }
  __device__ int nbu() const { return nbu_; }
  __device__ int ng() const { return ng_; }
  __device__ int ngN() const { return ngN_; }
  static bool matches(const QpLayout&) { return true; }
};

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
// Division by a Cholesky pivot in the triangular solves: BLASFEO's trsv multiplies by the inverse diagonal that potrf
// stored, which is 0 for a non-positive pivot (blasfeo_common.h: dA / use_dA) -- the component becomes 0 instead of
// inf / nan.  For a positive pivot: the plain division.
__device__ __forceinline__ double pivot_div(double s, double d) { return d > 0.0 ? s / d : 0.0; }
// NaN-propagating inf-norm accumulate (fmax would swallow NaNs)
__device__ __forceinline__ double amax_nan(double acc, double v) {
  const double a = fabs(v);
  return (a > acc || a != a) ? a : acc;
}

template <class D>
struct Solver {
  // row stride of the stage tiles: holds n+1 columns of M, [G | AL] (2 nx) and the DCt tile (ng); odd so
  // that 64-bit row-strided accesses are bank-conflict free
  static constexpr int kW0 = (D::kNM + 1) > (2 * D::kNX) ? (D::kNM + 1) : (2 * D::kNX);
  static constexpr int kW1 = kW0 > D::kNGM ? kW0 : D::kNGM;
  static constexpr int LD = (kW1 & 1) ? kW1 : (kW1 + 1);
  static constexpr int LDP = ((D::kNX + 1) & 1) ? (D::kNX + 1) : (D::kNX + 2);
  static constexpr int LDI = (D::kNU & 1) ? D::kNU : (D::kNU + 1);
  static constexpr int kSmemDoubles = 2 * (D::kNM + 1) * LD + (D::kNX + 1) * LDP + D::kNU * LDI +
                                      2 * D::kNCM + 4 * (D::kNM + 1) + 2 * (D::kNX + 1) + D::kNX * LDP;
  struct St {
    int k, nu, nx, n, nb, ng, nc, nxn, nun;
    const int* idxb;
  };

  const IpmParams& p;
  const QpLayout& L;
  D dm;
  int lane, q, N;
  double* W;
  double *sM, *sB, *sP, *sLi, *sQx, *sqx, *sg, *st, *sx, *sxn, *sdinv, *spn, *sLx;
  int* sIdx;
  // which workspace vectors the KKT solve reads as its right-hand sides (the step solves: the residuals; the
  // refinement solves: the residual of the linear system)
  struct Rhs { int rg, rb, rdl, rdu, rml, rmu; } rhs;

  __device__ Solver(const IpmParams& p_, double* smem, int* sidx, double* ws_smem)
      : p(p_), L(p_.L), dm(p_.L), lane(threadIdx.x & 31), q(0), N(p_.L.N) {
    rhs = Rhs{L.ws_rg, L.ws_rb, L.ws_rdl, L.ws_rdu, L.ws_rml, L.ws_rmu};
    W = p.ws_in_smem ? ws_smem : p.ws + (size_t)blockIdx.x * L.ws_size;
    sM = smem;
    sB = sM + (D::kNM + 1) * LD;
    sP = sB + (D::kNM + 1) * LD;
    sLi = sP + (D::kNX + 1) * LDP;
    sQx = sLi + D::kNU * LDI;
    sqx = sQx + D::kNCM;
    sg = sqx + D::kNCM;            // kNM+1
    st = sg + (D::kNM + 1);        // kNM+1
    sx = st + (D::kNM + 1);        // kNM+1  (dz of the current stage)
    sdinv = sx + (D::kNM + 1);     // kNM+1
    sxn = sdinv + (D::kNM + 1);    // kNX+1
    spn = sxn + (D::kNX + 1);      // kNX+1
    sLx = spn + (D::kNX + 1);      // kNX * LDP: Lxx_{k+1} (square-root Riccati, ric_alg = 1)
    sIdx = sidx;
    for (int j = lane; j < kMaxNB; j += 32) {
      sIdx[j] = L.idxb0[j];
      sIdx[kMaxNB + j] = L.idxb1[j];
      sIdx[2 * kMaxNB + j] = L.idxbN[j];
    }
    __syncwarp();
  }

  __device__ __forceinline__ St stage(int k) const {
    St s;
    s.k = k;
    s.nu = k < N ? dm.nu() : 0;
    s.nx = k > 0 ? dm.nx() : 0;
    s.n = s.nu + s.nx;
    s.nb = (k < N ? dm.nbu() : 0) + (k > 0 ? dm.nbx() : 0);
    s.ng = k < N ? dm.ng() : dm.ngN();
    s.nc = s.nb + s.ng;
    s.nxn = k < N ? dm.nx() : 0;
    s.nun = (k + 1 < N) ? dm.nu() : 0;
    s.idxb = sIdx + (k == 0 ? 0 : (k < N ? kMaxNB : 2 * kMaxNB));
    return s;
  }
  // strides (compile-time when D is static)
  __device__ __forceinline__ int nm() const { return dm.nu() + dm.nx(); }
  __device__ __forceinline__ int ngm() const { return dm.ng() > dm.ngN() ? dm.ng() : dm.ngN(); }
  __device__ __forceinline__ int ncm() const {
    const int v = dm.nbu() + dm.nbx() + ngm();
    return v > 0 ? v : 1;
  }
  __device__ __forceinline__ int babt_cn() const { return round4(dm.nx()); }
  __device__ __forceinline__ int babt_stride() const { return round4(nm() + 1) * babt_cn(); }
  __device__ __forceinline__ int rsq_cn() const { return round4(nm()); }
  __device__ __forceinline__ int rsq_stride() const { return round4(nm() + 1) * rsq_cn(); }
  __device__ __forceinline__ int dct_cn() const { return round4(ngm() > 0 ? ngm() : 1); }
  __device__ __forceinline__ int dct_stride() const { return round4(nm()) * dct_cn(); }

  __device__ __forceinline__ const double* gBAbt(int k) const { return p.babt + ((size_t)q * N + k) * babt_stride(); }
  __device__ __forceinline__ const double* gRSQ(int k) const { return p.rsq + ((size_t)q * (N + 1) + k) * rsq_stride(); }
  __device__ __forceinline__ const double* gDCt(int k) const { return p.dct + ((size_t)q * (N + 1) + k) * dct_stride(); }
  __device__ __forceinline__ const double* gD(int k) const { return p.d + ((size_t)q * (N + 1) + k) * 2 * ncm(); }
  __device__ __forceinline__ const double* gMask(int k) const { return p.dmask + ((size_t)q * (N + 1) + k) * 2 * ncm(); }

  // workspace accessors
  __device__ __forceinline__ double* wN(int off, int k) const { return W + off + k * nm(); }
  __device__ __forceinline__ double* wX(int off, int k) const { return W + off + k * dm.nx(); }
  __device__ __forceinline__ double* wC(int off, int k) const { return W + off + k * ncm(); }
  __device__ __forceinline__ double* wLi(int k) const { return W + L.ws_Li + k * dm.nu() * dm.nu(); }
  __device__ __forceinline__ double* wLr(int k) const { return W + L.ws_Lr + k * dm.nu() * dm.nu(); }
  __device__ __forceinline__ double* wLs(int k) const { return W + L.ws_Ls + k * dm.nx() * dm.nu(); }
  __device__ __forceinline__ double* wlv(int k) const { return W + L.ws_lv + k * dm.nu(); }
  __device__ __forceinline__ double* wP(int k) const { return W + L.ws_P + k * dm.nx() * dm.nx(); }
  __device__ __forceinline__ double* wp(int k) const { return W + L.ws_p + k * dm.nx(); }

  // panel-major global tile -> shared row-major (ld), coalesced 16-byte loads (two rows of one column)
  __device__ __forceinline__ void load_pm(double* dst, int ld, const double* __restrict__ src, int rows,
                                          int cols, int cn) const {
    const int half = round4(rows) * cn / 2;
    const double2* s2 = reinterpret_cast<const double2*>(src);
    for (int h = lane; h < half; h += 32) {
      const int idx = 2 * h;
      const int pnl = idx / (4 * cn), rem = idx - pnl * 4 * cn;
      const int j = rem >> 2, i = 4 * pnl + (rem & 3);
      if (j < cols && i < rows) {
        const double2 v = __ldg(s2 + h);
        dst[i * ld + j] = v.x;
        if (i + 1 < rows) dst[(i + 1) * ld + j] = v.y;
      }
    }
  }

  // v[j] = (J z)_j for j < nc: box rows pick, general rows dot with DCt columns (global, L1-cached)
  __device__ __forceinline__ double Jz_row(const St& s, const double* __restrict__ dct, const double* z, int j) const {
    if (j < s.nb) return z[s.idxb[j]];
    const int c = j - s.nb, cn = dct_cn();
    double acc = 0.0;
    for (int i = 0; i < s.n; ++i) acc += __ldg(dct + pm_index(i, c, cn)) * z[i];
    return acc;
  }
  // (J^T v)_i for i < n
  __device__ __forceinline__ double Jtv_row(const St& s, const double* __restrict__ dct, const double* v, int i) const {
    const int cn = dct_cn();
    double acc = 0.0;
    for (int j = 0; j < s.nb; ++j)
      if (s.idxb[j] == i) acc += v[j];
    for (int j = 0; j < s.ng; ++j) acc += __ldg(dct + pm_index(i, j, cn)) * v[s.nb + j];
    return acc;
  }

  // ------------------------------------------------------------------------------------------------
  // Gamma / gamma of one stage into sQx = Gamma_l + Gamma_u, sqx = gamma_l - gamma_u
  // (d_compute_Gamma_gamma_qp / d_compute_gamma_qp)
  // ------------------------------------------------------------------------------------------------
  __device__ __forceinline__ void gamma_stage(const St& s) {
    const double* mk = gMask(s.k);
    const int nc_m = ncm();
    for (int j = lane; j < s.nc; j += 32) {
      const double ml = __ldg(mk + j), mu = __ldg(mk + nc_m + j);
      const double ll = wC(L.ws_ll, s.k)[j], lu = wC(L.ws_lu, s.k)[j];
      const double til = 1.0 / wC(L.ws_tl, s.k)[j], tiu = 1.0 / wC(L.ws_tu, s.k)[j];
      const double Gl = (til * ll) * ml, Gu = (tiu * lu) * mu;
      const double gl = (til * (wC(rhs.rml, s.k)[j] - ll * wC(rhs.rdl, s.k)[j])) * ml;
      const double gu = (tiu * (wC(rhs.rmu, s.k)[j] - lu * wC(rhs.rdu, s.k)[j])) * mu;
      sQx[j] = Gl + Gu;
      sqx[j] = gl - gu;
    }
    __syncwarp();
  }

  // ------------------------------------------------------------------------------------------------
  // Backward Riccati sweep.  fact: factorize (stores Linv, Ls, P) else vector part only.
  // constr: add the barrier terms.  absolute: solve the QP itself (rhs = [r;q] rows and b rows of the
  // packed data) instead of the step QP (rhs = res_g, res_b from the workspace).
  // ------------------------------------------------------------------------------------------------
  __device__ void backward(bool fact, bool constr, bool absolute) {
    const bool sq = p.a.ric_alg == 1;
    for (int k = N; k >= 0; --k) {
      const St s = stage(k);
      const int n = s.n, nu = s.nu, nx = s.nx, nxn = s.nxn;
      if (constr) gamma_stage(s);
      const double* rsq = gRSQ(k);
      const double* dct = gDCt(k);
      // ---- rhs gradient row --------------------------------------------------------------------
      if (lane < n) {
        double g = absolute ? __ldg(rsq + pm_index(n, lane, rsq_cn())) : wN(rhs.rg, k)[lane];
        if (constr) g += Jtv_row(s, dct, sqx, lane);
        sg[lane] = g;
      }
      // rb / b row of this stage into sxn, t = P_{k+1} rb + p_{k+1} into st[0..nxn)
      if (k < N) {
        if (lane < nxn)
          sxn[lane] = absolute ? __ldg(gBAbt(k) + pm_index(n, lane, babt_cn())) : wX(rhs.rb, k)[lane];
        if (!fact) {  // P_{k+1}, p_{k+1} from the workspace (fact keeps them in sP from the previous stage)
          const double* Pn = wP(k + 1);
          for (int e = lane; e < nxn * nxn; e += 32) sP[(e / nxn) * LDP + (e % nxn)] = Pn[e];
          if (lane < nxn) sP[nxn * LDP + lane] = wp(k + 1)[lane];
        }
        __syncwarp();
        if (lane < nxn) {
          double acc = 0.0;
          for (int j = 0; j < nxn; ++j) acc += sP[lane * LDP + j] * sxn[j];
          st[lane] = acc + sP[nxn * LDP + lane];
        }
      }
      __syncwarp();
      if (fact) {
        // ---- M = H (lower) ; gradient row n -------------------------------------------------------
        load_pm(sM, LD, rsq, n, n, rsq_cn());
        __syncwarp();
        if (constr) {
          // box rows on the diagonal, then DCt diag(Qx) DCt^T (row-parallel, lower triangle)
          for (int j = lane; j < s.nb; j += 32) sM[s.idxb[j] * LD + s.idxb[j]] += sQx[j];
          if (s.ng > 0) {
            load_pm(sB, LD, dct, n, s.ng, dct_cn());
            __syncwarp();
            if (lane < n) {
              for (int c = 0; c <= lane; ++c) {
                double acc = 0.0;
                for (int j = 0; j < s.ng; ++j) acc += (sB[lane * LD + j] * sQx[s.nb + j]) * sB[c * LD + j];
                sM[lane * LD + c] += acc;
              }
            }
          }
          __syncwarp();
        }
        if (k < N) {
          // [G | AL]: G (n x nxn) from BAbt.  classical (ric_alg 0): AL = G P_{k+1}, M += AL G^T (gemm_nt +
          // syrk_ln_mn); square root (ric_alg 1): AL = G Lxx_{k+1} (trmm_rlnn), M += AL AL^T (syrk_dpotrf_ln_mn)
          load_pm(sB, LD, gBAbt(k), n, nxn, babt_cn());
          __syncwarp();
          if (lane < n) {
            for (int j = 0; j < nxn; ++j) {
              double acc = 0.0;
              if (sq) { for (int l = j; l < nxn; ++l) acc += sB[lane * LD + l] * sLx[l * LDP + j]; }
              else { for (int l = 0; l < nxn; ++l) acc += sB[lane * LD + l] * sP[l * LDP + j]; }
              sB[lane * LD + nxn + j] = acc;
            }
          }
          __syncwarp();
          if (lane < n) {
            const int off = sq ? nxn : 0;
            for (int c = 0; c <= lane; ++c) {
              double acc = 0.0;
              for (int l = 0; l < nxn; ++l) acc += sB[lane * LD + nxn + l] * sB[c * LD + off + l];
              sM[lane * LD + c] += acc;
            }
          }
          __syncwarp();
        }
        if (lane < n) sM[lane * LD + lane] += p.a.reg_prim;
      }
      // ---- gradient += G t ------------------------------------------------------------------------
      if (k < N) {
        if (!fact) {  // G straight from global (vector-only pass)
          const double* ba = gBAbt(k);
          if (lane < n) {
            double acc = 0.0;
            for (int l = 0; l < nxn; ++l) acc += __ldg(ba + pm_index(lane, l, babt_cn())) * st[l];
            sg[lane] += acc;
          }
        } else if (lane < n) {
          double acc = 0.0;
          for (int l = 0; l < nxn; ++l) acc += sB[lane * LD + l] * st[l];
          sg[lane] += acc;
        }
      }
      __syncwarp();
      if (fact) {
        // gradient row n of M (the "+1 row" of potrf_l_mn)
        if (lane < n) sM[n * LD + lane] = sg[lane];
        __syncwarp();
        // ---- Cholesky over rows 0..n (left-looking, row-parallel): the leading nu columns (classical:
        //      potrf_l_mn(n+1, nu)) or all n columns (square root)
        const int ncol = sq ? n : nu;
        for (int j = 0; j < ncol; ++j) {
          double sacc = 0.0;
          if (lane >= j && lane <= n) {
            sacc = sM[lane * LD + j];
            for (int l = 0; l < j; ++l) sacc -= sM[lane * LD + l] * sM[j * LD + l];
          }
          const double dj = __shfl_sync(kFull, sacc, j);
          const double inv = dj > 0.0 ? rsqrt(dj) : 0.0;
          if (lane == j) {
            sM[j * LD + j] = dj * inv;
            if (j < nu) sdinv[j] = inv;
          } else if (lane > j && lane <= n) {
            sM[lane * LD + j] = sacc * inv;
          }
          __syncwarp();
        }
        // ---- Linv = Lr^-1 (column-parallel forward substitution on the identity) -------------------
        if (lane < nu) {
          for (int i = 0; i < nu; ++i) sLi[i * LDI + lane] = (i == lane) ? sdinv[lane] : 0.0;
          for (int i = 1; i < nu; ++i) {
            if (i > lane) {
              double acc = 0.0;
              for (int l = 0; l < i; ++l) acc += sM[i * LD + l] * sLi[l * LDI + lane];
              sLi[i * LDI + lane] = -acc * sdinv[i];
            }
          }
        }
        __syncwarp();
        // ---- Schur complement: P_k = M_xx - Ls Ls^T, p_k = g_x - Ls lv (rows nu..n of M) -------------
        {
          const int tot = (nx + 1) * nx;
          for (int e = lane; e < tot; e += 32) {
            const int i = e / nx, c = e - i * nx;  // i == nx: gradient row
            // lower triangle of M holds the data: element (a,b) with a >= b
            const int a = nu + i, b = nu + c;
            double acc;
            if (sq) {  // P = Lxx Lxx^T, p = Lxx l_x  (M now holds the full factor)
              acc = 0.0;
              const int lmax = (i == nx) ? c : (i < c ? i : c);
              for (int l = 0; l <= lmax; ++l) acc += sM[a * LD + nu + l] * sM[b * LD + nu + l];
            } else {
              acc = (i == nx || i >= c) ? sM[a * LD + b] : sM[b * LD + a];
              for (int l = 0; l < nu; ++l) acc -= sM[a * LD + l] * sM[b * LD + l];
            }
            sP[i * LDP + c] = acc;
          }
          if (sq)
            for (int e = lane; e < nx * nx; e += 32) {
              const int i = e / nx, c = e - i * nx;
              sLx[i * LDP + c] = (c <= i) ? sM[(nu + i) * LD + nu + c] : 0.0;
            }
        }
        __syncwarp();
        // ---- store factors ------------------------------------------------------------------------------
        {
          double* Li = wLi(k);
          double* Ls = wLs(k);
          double* Pk = wP(k);
          for (int e = lane; e < nu * nu; e += 32) Li[e] = sLi[(e / nu) * LDI + (e % nu)];   // row-major
          double* Lrk = wLr(k);  // the triangular factor itself: the vector solves substitute with it (trsv)
          for (int e = lane; e < nu * nu; e += 32) Lrk[e] = sM[(e / nu) * LD + (e % nu)];
          for (int e = lane; e < nx * nu; e += 32) Ls[e] = sM[(nu + e / nu) * LD + (e % nu)];  // row-major nx x nu
          for (int e = lane; e < nx * nx; e += 32) Pk[e] = sP[(e / nx) * LDP + (e % nx)];
          if (lane < nu) wlv(k)[lane] = sM[n * LD + lane];
          if (lane < nx) wp(k)[lane] = sP[nx * LDP + lane];
        }
        __syncwarp();
      } else {
        // ---- vector part only: lv = Linv g_u ; p = g_x - Ls lv -------------------------------------------
        // lv = Lr^-1 g_u by forward substitution with Lr itself, row by row in BLASFEO's trsv
        // operation order: an explicit Lr^-1 leaves a residual ~ eps cond(H~uu) |g| in the Newton system, which at
        // tol 1e-8 decides between convergence and stagnation on the rounding floor
        const double* Lrk = wLr(k);
        const double* Ls = wLs(k);
        for (int e = lane; e < nu * nu; e += 32) sLi[(e / nu) * LDI + (e % nu)] = Lrk[e];
        __syncwarp();
        for (int i = 0; i < nu; ++i) {
          if (lane == i) {
            double acc = sg[i];
            for (int j = 0; j < i; ++j) acc -= sLi[i * LDI + j] * st[j];
            acc = pivot_div(acc, sLi[i * LDI + i]);
            st[i] = acc;
            wlv(k)[i] = acc;
          }
          __syncwarp();
        }
        if (lane < nx) {
          double acc = sg[nu + lane];
          for (int l = 0; l < nu; ++l) acc -= Ls[lane * nu + l] * st[l];
          wp(k)[lane] = acc;
        }
        __syncwarp();
      }
    }
  }

  // ------------------------------------------------------------------------------------------------
  // Forward rollout: oz (ws offset) <- minimizer, opi <- multipliers.  rhs b: step (res_b) or absolute.
  // ------------------------------------------------------------------------------------------------
  __device__ void forward(int off_z, int off_pi, bool absolute) {
    for (int k = 0; k <= N; ++k) {
      const St s = stage(k);
      const int n = s.n, nu = s.nu, nx = s.nx, nxn = s.nxn;
      double* z = wN(off_z, k);
      const double* Lrk = wLr(k);
      const double* Ls = wLs(k);
      // sx[nu..n) = x part (written by the previous stage into z)
      if (lane < nx) sx[nu + lane] = z[nu + lane];
      for (int e = lane; e < nu * nu; e += 32) sLi[(e / nu) * LDI + (e % nu)] = Lrk[e];
      __syncwarp();
      if (lane < nu) {  // t = Ls^T x + lv
        double acc = 0.0;
        for (int j = 0; j < nx; ++j) acc += Ls[j * nu + lane] * sx[nu + j];
        st[lane] = acc + wlv(k)[lane];
      }
      __syncwarp();
      // u = -Lr^-T t by back substitution (row by row, ascending inner index like BLASFEO's trsv)
      for (int i = nu - 1; i >= 0; --i) {
        if (lane == i) {
          double acc = st[i];
          for (int j = i + 1; j < nu; ++j) acc -= sLi[j * LDI + i] * st[j];
          st[i] = pivot_div(acc, sLi[i * LDI + i]);
        }
        __syncwarp();
      }
      if (lane < nu) {
        const double acc = st[lane];
        sx[lane] = -acc;
        z[lane] = -acc;
      }
      __syncwarp();
      if (k < N) {
        const double* ba = gBAbt(k);
        double* zn = wN(off_z, k + 1);
        if (lane < nxn) {
          double acc = 0.0;
          for (int i = 0; i < n; ++i) acc += __ldg(ba + pm_index(i, lane, babt_cn())) * sx[i];
          const double rb = absolute ? __ldg(ba + pm_index(n, lane, babt_cn())) : wX(rhs.rb, k)[lane];
          acc += rb;
          sxn[lane] = acc;
          zn[s.nun + lane] = acc;
        }
        __syncwarp();
        if (lane < nxn) {
          const double* Pn = wP(k + 1);
          double acc = 0.0;
          for (int j = 0; j < nxn; ++j) acc += Pn[lane * nxn + j] * sxn[j];
          wX(off_pi, k)[lane] = acc + wp(k + 1)[lane];
        }
      }
      __syncwarp();
    }
  }

  // ------------------------------------------------------------------------------------------------
  // d_ocp_qp_init_var
  // ------------------------------------------------------------------------------------------------
  __device__ void init_var() {
    const double thr0 = p.a.thr0, mu0 = p.a.mu0;
    const int nc_m = ncm();
    for (int k = 0; k <= N; ++k) {
      const St s = stage(k);
      double* z = wN(L.ws_z, k);
      // primal start: warm (x[k], u[k] given) or cold (zeros)
      if (lane < s.n) {
        double v = 0.0;
        if (p.a.warm_start && p.x_init && p.u_init) {
          v = lane < s.nu ? p.u_init[((size_t)q * N + k) * dm.nu() + lane]
                          : p.x_init[((size_t)q * (N + 1) + k) * dm.nx() + (lane - s.nu)];
        }
        z[lane] = v;
      }
      if (k < N && lane < dm.nx()) wX(L.ws_pi, k)[lane] = 0.0;
      __syncwarp();
      const double* dd = gD(k);
      const double* mk = gMask(k);
      double tl = 0.0, tu = 0.0, lo = 0.0, up = 0.0;
      if (lane < s.nb) {  // box rows (distinct idxb -> no write conflicts)
        const int i = s.idxb[lane];
        lo = __ldg(dd + lane);
        up = -__ldg(dd + nc_m + lane);
        const double zi = z[i];
        tl = -lo + zi;
        tu = up - zi;
        if (tl < thr0) {
          if (tu < thr0) {
            z[i] = 0.5 * (lo + up);
            tl = thr0; tu = thr0;
          } else {
            tl = thr0;
            z[i] = lo + thr0;
          }
        } else if (tu < thr0) {
          tu = thr0;
          z[i] = up - thr0;
        }
      }
      __syncwarp();
      for (int j = lane; j < s.nc; j += 32) {
        if (j >= s.nb) {
          lo = __ldg(dd + j);
          up = -__ldg(dd + nc_m + j);
          const double v = Jz_row(s, gDCt(k), z, j);
          tl = v - lo;
          tu = up - v;
          tl = thr0 > tl ? thr0 : tl;
          tu = thr0 > tu ? thr0 : tu;
        }
        wC(L.ws_tl, k)[j] = tl;
        wC(L.ws_tu, k)[j] = tu;
        wC(L.ws_ll, k)[j] = (mu0 / tl) * __ldg(mk + j);
        wC(L.ws_lu, k)[j] = (mu0 / tu) * __ldg(mk + nc_m + j);
      }
      __syncwarp();
    }
  }

  // ------------------------------------------------------------------------------------------------
  // d_ocp_qp_res_compute + inf norms.  out: res[4], mu, obj (same value in every lane)
  // ------------------------------------------------------------------------------------------------
  __device__ void residuals(double res[4], double& mu, double& obj, int nc_mask) {
    double ng_ = 0.0, nb_ = 0.0, nd_ = 0.0, nm_ = 0.0, sm = 0.0, ob = 0.0;
    const int nc_m = ncm();
    for (int k = 0; k <= N; ++k) {
      const St s = stage(k);
      const int n = s.n, nu = s.nu, nx = s.nx;
      const double* z = wN(L.ws_z, k);
      const double* rsq = gRSQ(k);
      const double* dct = gDCt(k);
      const double* dd = gD(k);
      const double* mk = gMask(k);
      // stage vectors into shared memory: z -> sx, lam_u - lam_l -> sqx
      if (lane < n) sx[lane] = z[lane];
      for (int j = lane; j < s.nc; j += 32) sqx[j] = wC(L.ws_lu, k)[j] - wC(L.ws_ll, k)[j];
      if (k < N && lane < dm.nx()) sxn[lane] = wX(L.ws_pi, k)[lane];
      __syncwarp();
      if (lane < n) {
        const int cn = rsq_cn();
        double acc = 0.0;
        for (int j = 0; j < n; ++j) {
          const int a = lane >= j ? lane : j, b = lane >= j ? j : lane;
          acc += __ldg(rsq + pm_index(a, b, cn)) * sx[j];
        }
        const double g = __ldg(rsq + pm_index(n, lane, cn));
        ob += 0.5 * sx[lane] * acc + g * sx[lane];
        double r = acc + g;
        if (k < N) {
          const double* ba = gBAbt(k);
          double a2 = 0.0;
          for (int j = 0; j < s.nxn; ++j) a2 += __ldg(ba + pm_index(lane, j, babt_cn())) * sxn[j];
          r += a2;
        }
        if (k > 0 && lane >= nu) r -= wX(L.ws_pi, k - 1)[lane - nu];
        r += Jtv_row(s, dct, sqx, lane);
        wN(L.ws_rg, k)[lane] = r;
        ng_ = amax_nan(ng_, r);
      }
      if (k < N && lane < s.nxn) {
        const double* ba = gBAbt(k);
        double acc = 0.0;
        for (int i = 0; i < n; ++i) acc += __ldg(ba + pm_index(i, lane, babt_cn())) * sx[i];
        const double r = (acc + __ldg(ba + pm_index(n, lane, babt_cn()))) - wN(L.ws_z, k + 1)[s.nun + lane];
        wX(L.ws_rb, k)[lane] = r;
        nb_ = amax_nan(nb_, r);
      }
      for (int j = lane; j < s.nc; j += 32) {
        const double v = Jz_row(s, dct, sx, j);
        const double ml = __ldg(mk + j), mu_ = __ldg(mk + nc_m + j);
        const double lo = __ldg(dd + j), nup = __ldg(dd + nc_m + j);  // nup = -up
        const double tl = wC(L.ws_tl, k)[j], tu = wC(L.ws_tu, k)[j];
        const double rdl = ((lo - v) + tl) * ml;
        const double rdu = ((nup + v) + tu) * mu_;
        const double rml = (wC(L.ws_ll, k)[j] * tl) * ml;
        const double rmu = (wC(L.ws_lu, k)[j] * tu) * mu_;
        wC(L.ws_rdl, k)[j] = rdl;
        wC(L.ws_rdu, k)[j] = rdu;
        wC(L.ws_rml, k)[j] = rml;
        wC(L.ws_rmu, k)[j] = rmu;
        sm += rml;
        sm += rmu;
        nd_ = amax_nan(amax_nan(nd_, rdl), rdu);
        nm_ = amax_nan(amax_nan(nm_, rml), rmu);
      }
      __syncwarp();
    }
    // warp reductions (NaN-safe: a NaN lane makes the sum NaN and we OR a NaN flag into the maxima)
    const double flag = warp_sum((ng_ != ng_ || nb_ != nb_ || nd_ != nd_ || nm_ != nm_) ? 1.0 : 0.0);
    res[0] = warp_max(ng_ == ng_ ? ng_ : 0.0);
    res[1] = warp_max(nb_ == nb_ ? nb_ : 0.0);
    res[2] = warp_max(nd_ == nd_ ? nd_ : 0.0);
    res[3] = warp_max(nm_ == nm_ ? nm_ : 0.0);
    if (flag > 0.0) res[0] = res[0] + __longlong_as_double(0x7ff8000000000000LL);
    sm = warp_sum(sm);
    obj = warp_sum(ob);
    mu = nc_mask > 0 ? sm / (double)nc_mask : 0.0;
  }

  // ------------------------------------------------------------------------------------------------
  // dt, dlam from dz (d_compute_lam_t_qp) fused with the step length (d_compute_alpha_qp)
  // ------------------------------------------------------------------------------------------------
  __device__ void dlam_dt_alpha(double& ap, double& ad) {
    double a_p = 1.0, a_d = 1.0;
    const int nc_m = ncm();
    for (int k = 0; k <= N; ++k) {
      const St s = stage(k);
      if (s.nc == 0) continue;
      const double* dz = wN(L.ws_dz, k);
      const double* mk = gMask(k);
      if (lane < s.n) sx[lane] = dz[lane];
      __syncwarp();
      for (int j = lane; j < s.nc; j += 32) {
        const double v = Jz_row(s, gDCt(k), sx, j);
        const double ml = __ldg(mk + j), mu_ = __ldg(mk + nc_m + j);
        const double tl = wC(L.ws_tl, k)[j], tu = wC(L.ws_tu, k)[j];
        const double ll = wC(L.ws_ll, k)[j], lu = wC(L.ws_lu, k)[j];
        const double dtl = (v - wC(L.ws_rdl, k)[j]) * ml;
        const double dtu = (-v - wC(L.ws_rdu, k)[j]) * mu_;
        const double dll = (-(ll * dtl + wC(L.ws_rml, k)[j]) / tl) * ml;
        const double dlu = (-(lu * dtu + wC(L.ws_rmu, k)[j]) / tu) * mu_;
        wC(L.ws_dtl, k)[j] = dtl;
        wC(L.ws_dtu, k)[j] = dtu;
        wC(L.ws_dll, k)[j] = dll;
        wC(L.ws_dlu, k)[j] = dlu;
        if (dtl < 0.0) a_p = fmin(a_p, -tl / dtl);
        if (dtu < 0.0) a_p = fmin(a_p, -tu / dtu);
        if (dll < 0.0) a_d = fmin(a_d, -ll / dll);
        if (dlu < 0.0) a_d = fmin(a_d, -lu / dlu);
      }
      __syncwarp();
    }
    ap = warp_min(a_p);
    ad = warp_min(a_d);
  }

  // d_compute_alpha_qp on the stored step (after a refinement changed it)
  __device__ void step_lengths(double& ap, double& ad) {
    double a_p = 1.0, a_d = 1.0;
    for (int k = 0; k <= N; ++k) {
      const St s = stage(k);
      for (int j = lane; j < s.nc; j += 32) {
        const double dtl = wC(L.ws_dtl, k)[j], dtu = wC(L.ws_dtu, k)[j], dll = wC(L.ws_dll, k)[j], dlu = wC(L.ws_dlu, k)[j];
        if (dtl < 0.0) a_p = fmin(a_p, -wC(L.ws_tl, k)[j] / dtl);
        if (dtu < 0.0) a_p = fmin(a_p, -wC(L.ws_tu, k)[j] / dtu);
        if (dll < 0.0) a_d = fmin(a_d, -wC(L.ws_ll, k)[j] / dll);
        if (dlu < 0.0) a_d = fmin(a_d, -wC(L.ws_lu, k)[j] / dlu);
      }
    }
    ap = warp_min(a_p);
    ad = warp_min(a_d);
  }

  // d_compute_mu_aff_qp
  __device__ double mu_aff(double alpha, int nc_mask) {
    double sm = 0.0;
    for (int k = 0; k <= N; ++k) {
      const St s = stage(k);
      for (int j = lane; j < s.nc; j += 32) {
        sm += (wC(L.ws_ll, k)[j] + alpha * wC(L.ws_dll, k)[j]) * (wC(L.ws_tl, k)[j] + alpha * wC(L.ws_dtl, k)[j]);
        sm += (wC(L.ws_lu, k)[j] + alpha * wC(L.ws_dlu, k)[j]) * (wC(L.ws_tu, k)[j] + alpha * wC(L.ws_dtu, k)[j]);
      }
    }
    return warp_sum(sm) / (double)nc_mask;
  }

  // ------------------------------------------------------------------------------------------------
  // Iterative refinement (hpipm_d_ocp_qp_ipm.h:74-75 itref_pred_max / itref_corr_max; BALANCE / ROBUST modes).
  // residuals_lin: d_ocp_qp_res_compute_lin -- the residual of the LINEAR KKT system at the current step
  //   lin_g = H dz + rg + G dpi - [0; dpi_{k-1}] + J^T (dlam_u - dlam_l)     lin_b = G^T dz + rb - dx_{k+1}
  //   lin_d_l = (-J dz + dt_l) + rd_l,  lin_d_u = (J dz + dt_u) + rd_u          lin_m = lam dt + t dlam + rm
  // (rg, rb, rd, rm = the right-hand sides of the solve being refined) and its inf-norms.
  // ------------------------------------------------------------------------------------------------
  __device__ void residuals_lin(double lin[4]) {
    double ng_ = 0.0, nb_ = 0.0, nd_ = 0.0, nm_ = 0.0;
    const int nc_m = ncm();
    for (int k = 0; k <= N; ++k) {
      const St s = stage(k);
      const int n = s.n, nu = s.nu;
      const double* dz = wN(L.ws_dz, k);
      const double* rsq = gRSQ(k);
      const double* dct = gDCt(k);
      const double* mk = gMask(k);
      if (lane < n) sx[lane] = dz[lane];
      for (int j = lane; j < s.nc; j += 32) sqx[j] = wC(L.ws_dlu, k)[j] - wC(L.ws_dll, k)[j];
      if (k < N && lane < dm.nx()) sxn[lane] = wX(L.ws_dpi, k)[lane];
      __syncwarp();
      if (lane < n) {
        const int cn = rsq_cn();
        double acc = 0.0;
        for (int j = 0; j < n; ++j) {
          const int a = lane >= j ? lane : j, b = lane >= j ? j : lane;
          acc += __ldg(rsq + pm_index(a, b, cn)) * sx[j];
        }
        double r = acc + wN(L.ws_rg, k)[lane];
        if (k < N) {
          const double* ba = gBAbt(k);
          double a2 = 0.0;
          for (int j = 0; j < s.nxn; ++j) a2 += __ldg(ba + pm_index(lane, j, babt_cn())) * sxn[j];
          r += a2;
        }
        if (k > 0 && lane >= nu) r -= wX(L.ws_dpi, k - 1)[lane - nu];
        r += Jtv_row(s, dct, sqx, lane);
        wN(L.ws_lg, k)[lane] = r;
        ng_ = amax_nan(ng_, r);
      }
      if (k < N && lane < s.nxn) {
        const double* ba = gBAbt(k);
        double acc = 0.0;
        for (int i = 0; i < n; ++i) acc += __ldg(ba + pm_index(i, lane, babt_cn())) * sx[i];
        const double r = (acc + wX(L.ws_rb, k)[lane]) - wN(L.ws_dz, k + 1)[s.nun + lane];
        wX(L.ws_lb, k)[lane] = r;
        nb_ = amax_nan(nb_, r);
      }
      for (int j = lane; j < s.nc; j += 32) {
        const double v = Jz_row(s, dct, sx, j);
        const double ml = __ldg(mk + j), mu_ = __ldg(mk + nc_m + j);
        const double dtl = wC(L.ws_dtl, k)[j], dtu = wC(L.ws_dtu, k)[j];
        const double ldl = ((-v + dtl) + wC(L.ws_rdl, k)[j]) * ml;
        const double ldu = ((v + dtu) + wC(L.ws_rdu, k)[j]) * mu_;
        const double lml = ((wC(L.ws_ll, k)[j] * dtl + wC(L.ws_tl, k)[j] * wC(L.ws_dll, k)[j]) + wC(L.ws_rml, k)[j]) * ml;
        const double lmu = ((wC(L.ws_lu, k)[j] * dtu + wC(L.ws_tu, k)[j] * wC(L.ws_dlu, k)[j]) + wC(L.ws_rmu, k)[j]) * mu_;
        wC(L.ws_ldl, k)[j] = ldl; wC(L.ws_ldu, k)[j] = ldu; wC(L.ws_lml, k)[j] = lml; wC(L.ws_lmu, k)[j] = lmu;
        nd_ = amax_nan(amax_nan(nd_, ldl), ldu);
        nm_ = amax_nan(amax_nan(nm_, lml), lmu);
      }
      __syncwarp();
    }
    lin[0] = warp_max(ng_ == ng_ ? ng_ : 0.0);
    lin[1] = warp_max(nb_ == nb_ ? nb_ : 0.0);
    lin[2] = warp_max(nd_ == nd_ ? nd_ : 0.0);
    lin[3] = warp_max(nm_ == nm_ ? nm_ : 0.0);
    const double flag = warp_sum((ng_ != ng_ || nb_ != nb_ || nd_ != nd_ || nm_ != nm_) ? 1.0 : 0.0);
    if (flag > 0.0) lin[0] = lin[0] + __longlong_as_double(0x7ff8000000000000LL);
  }
  // the refinement loop; returns the number of correction solves, lin[] = the last linear residual norms
  __device__ int refine(int max_it, const double res[4], double lin[4]) {
    const srbd_ipm_args& a = p.a;
    const double tol[4] = {a.tol_stat, a.tol_eq, a.tol_ineq, a.tol_comp};
    const int nc_m = ncm();
    int done = 0;
    for (int it = 0; it < max_it; ++it) {
      residuals_lin(lin);
      bool ok = true;
      for (int i = 0; i < 4; ++i) ok = ok && (lin[i] < a.itref_abs * tol[i] || lin[i] < a.itref_rel * res[i]);
      if (ok) break;
      rhs = Rhs{L.ws_lg, L.ws_lb, L.ws_ldl, L.ws_ldu, L.ws_lml, L.ws_lmu};
      backward(false, true, false);
      forward(L.ws_cz, L.ws_cpi, false);
      rhs = Rhs{L.ws_rg, L.ws_rb, L.ws_rdl, L.ws_rdu, L.ws_rml, L.ws_rmu};
      for (int k = 0; k <= N; ++k) {
        const St s = stage(k);
        const double* cz = wN(L.ws_cz, k);
        const double* mk = gMask(k);
        if (lane < s.n) sx[lane] = cz[lane];
        __syncwarp();
        for (int j = lane; j < s.nc; j += 32) {
          const double v = Jz_row(s, gDCt(k), sx, j);
          const double ml = __ldg(mk + j), mu_ = __ldg(mk + nc_m + j);
          const double ctl = (v - wC(L.ws_ldl, k)[j]) * ml;
          const double ctu = (-v - wC(L.ws_ldu, k)[j]) * mu_;
          const double cll = (-(wC(L.ws_ll, k)[j] * ctl + wC(L.ws_lml, k)[j]) / wC(L.ws_tl, k)[j]) * ml;
          const double clu = (-(wC(L.ws_lu, k)[j] * ctu + wC(L.ws_lmu, k)[j]) / wC(L.ws_tu, k)[j]) * mu_;
          wC(L.ws_dtl, k)[j] += ctl; wC(L.ws_dtu, k)[j] += ctu; wC(L.ws_dll, k)[j] += cll; wC(L.ws_dlu, k)[j] += clu;
        }
        if (lane < s.n) wN(L.ws_dz, k)[lane] += cz[lane];
        if (k < N && lane < dm.nx()) wX(L.ws_dpi, k)[lane] += wX(L.ws_cpi, k)[lane];
        __syncwarp();
      }
      ++done;
    }
    return done;
  }

  // backup res_m (BACKUP_RES_M)
  __device__ void backup_res_m() {
    for (int k = 0; k <= N; ++k) {
      const St s = stage(k);
      for (int j = lane; j < s.nc; j += 32) {
        wC(L.ws_rmlb, k)[j] = wC(L.ws_rml, k)[j];
        wC(L.ws_rmub, k)[j] = wC(L.ws_rmu, k)[j];
      }
    }
  }
  // d_compute_centering_correction_qp (corr = true) / d_compute_centering_qp (corr = false), masked
  __device__ void centering(double sm, bool corr) {
    const int nc_m = ncm();
    for (int k = 0; k <= N; ++k) {
      const St s = stage(k);
      const double* mk = gMask(k);
      for (int j = lane; j < s.nc; j += 32) {
        double a = wC(L.ws_rmlb, k)[j], b = wC(L.ws_rmub, k)[j];
        if (corr) {
          a += wC(L.ws_dtl, k)[j] * wC(L.ws_dll, k)[j];
          b += wC(L.ws_dtu, k)[j] * wC(L.ws_dlu, k)[j];
        }
        wC(L.ws_rml, k)[j] = (a - sm) * __ldg(mk + j);
        wC(L.ws_rmu, k)[j] = (b - sm) * __ldg(mk + nc_m + j);
      }
    }
    __syncwarp();
  }

  // d_update_var_qp
  __device__ void update(double sp, double sd) {
    const int nc_m = ncm();
    for (int k = 0; k <= N; ++k) {
      const St s = stage(k);
      if (lane < s.n) wN(L.ws_z, k)[lane] += sp * wN(L.ws_dz, k)[lane];
      if (k < N && lane < dm.nx()) wX(L.ws_pi, k)[lane] += sd * wX(L.ws_dpi, k)[lane];
      const double* mk = gMask(k);
      for (int j = lane; j < s.nc; j += 32) {
        double tl = wC(L.ws_tl, k)[j] + sp * wC(L.ws_dtl, k)[j];
        double tu = wC(L.ws_tu, k)[j] + sp * wC(L.ws_dtu, k)[j];
        double ll = wC(L.ws_ll, k)[j] + sd * wC(L.ws_dll, k)[j];
        double lu = wC(L.ws_lu, k)[j] + sd * wC(L.ws_dlu, k)[j];
        if (p.a.t_lam_min == 2) {
          if (__ldg(mk + j) != 0.0) {
            tl = tl < p.a.t_min ? p.a.t_min : tl;
            ll = ll < p.a.lam_min ? p.a.lam_min : ll;
          }
          if (__ldg(mk + nc_m + j) != 0.0) {
            tu = tu < p.a.t_min ? p.a.t_min : tu;
            lu = lu < p.a.lam_min ? p.a.lam_min : lu;
          }
        }
        wC(L.ws_tl, k)[j] = tl; wC(L.ws_tu, k)[j] = tu;
        wC(L.ws_ll, k)[j] = ll; wC(L.ws_lu, k)[j] = lu;
      }
    }
    __syncwarp();
  }

  __device__ __forceinline__ double shorten(double alpha) const {
    if (alpha < 1.0) {
      if (p.a.alpha_shorten == 0) return alpha * 0.995;
      return alpha * ((1.0 - alpha) * 0.99 + alpha * 0.9999999);
    }
    return alpha;
  }

  __device__ int count_masks() {
    int c = 0;
    const int nc_m = ncm();
    for (int k = 0; k <= N; ++k) {
      const St s = stage(k);
      const double* mk = gMask(k);
      for (int j = lane; j < s.nc; j += 32) c += (__ldg(mk + j) != 0.0) + (__ldg(mk + nc_m + j) != 0.0);
    }
    return warp_sum_i(c);
  }

  __device__ void stat_row(int row, int col0, const double* v, int n) {
    if (p.stat && row < p.stat_rows && lane == 0) {
      double* r = p.stat + ((size_t)q * p.stat_rows + row) * SRBD_STAT_M;
      for (int i = 0; i < n; ++i) r[col0 + i] = v[i];
    }
  }

  // ------------------------------------------------------------------------------------------------
  // outputs (getters d_ocp_qp_sol_get_*, d_ocp_qp_ipm_get_ric_* and the stage-0 reconstruction of
  // hpipm-cpp/src/ocp_qp_ipm_solver.cpp:337-373)
  // ------------------------------------------------------------------------------------------------
  __device__ void write_outputs(bool unconstrained) {
    const int nx = dm.nx(), nu = dm.nu();
    for (int k = 0; k <= N; ++k) {
      const St s = stage(k);
      const double* z = wN(L.ws_z, k);
      if (k == 0) {
        if (lane < nx) p.sol_x[((size_t)q * (N + 1)) * nx + lane] = p.x0[(size_t)q * nx + lane];
      } else if (lane < nx) {
        p.sol_x[((size_t)q * (N + 1) + k) * nx + lane] = z[s.nu + lane];
        p.sol_pi[((size_t)q * (N + 1) + k) * nx + lane] = wX(L.ws_pi, k - 1)[lane];
      }
      if (k < N && lane < nu) p.sol_u[((size_t)q * N + k) * nu + lane] = z[lane];
    }
    if (p.sol_lam && p.sol_t) {
      size_t o = (size_t)q * L.nct;
      for (int k = 0; k <= N; ++k) {
        const St s = stage(k);
        for (int j = lane; j < s.nc; j += 32) {
          p.sol_lam[o + j] = wC(L.ws_ll, k)[j];
          p.sol_lam[o + s.nc + j] = wC(L.ws_lu, k)[j];
          p.sol_t[o + j] = wC(L.ws_tl, k)[j];
          p.sol_t[o + s.nc + j] = wC(L.ws_tu, k)[j];
        }
        o += 2 * (size_t)s.nc;
      }
    }
    if (!p.ric_P) return;
    // Riccati exports, column-major like Eigen.  K_k = -Linv^T Ls^T  (nu x nx), k>=1
    for (int k = 1; k <= N; ++k) {
      const double* Pk = wP(k);
      double* Po = p.ric_P + ((size_t)q * (N + 1) + k) * nx * nx;
      for (int e = lane; e < nx * nx; e += 32) Po[e] = Pk[e];  // symmetric: row-major == column-major
      const double* xk = wN(L.ws_z, k) + stage(k).nu;
      if (lane < nx) {
        double v;
        if (unconstrained) v = wp(k)[lane];
        else {  // pi_k = P_k x_k + p_k (absolute form by identity)
          double acc = 0.0;
          for (int j = 0; j < nx; ++j) acc += Pk[lane * nx + j] * xk[j];
          v = wX(L.ws_pi, k - 1)[lane] - acc;
        }
        p.ric_p[((size_t)q * (N + 1) + k) * nx + lane] = v;
      }
      if (k < N) {
        const double* Li = wLi(k);
        const double* Ls = wLs(k);
        double* Ko = p.ric_K + ((size_t)q * N + k) * nu * nx;
        for (int e = lane; e < nu * nx; e += 32) {
          const int i = e % nu, c = e / nu;  // K(i,c) = -(Lr^-T Ls^T)(i,c) = -sum_{j>=i} Linv(j,i) Ls(c,j)
          double acc = 0.0;
          for (int j = i; j < nu; ++j) acc += Li[j * nu + i] * Ls[c * nu + j];
          Ko[e] = -acc;
        }
        __syncwarp();
        if (lane < nu) {
          const double* z = wN(L.ws_z, k);
          double v;
          if (unconstrained) {  // -Linv^T lv
            double acc = 0.0;
            for (int j = lane; j < nu; ++j) acc += Li[j * nu + lane] * wlv(k)[j];
            v = -acc;
          } else {  // k = u - K x
            double acc = 0.0;
            for (int c = 0; c < nx; ++c) acc += Ko[lane + nu * c] * z[nu + c];
            v = z[lane] - acc;
          }
          p.ric_k[((size_t)q * N + k) * nu + lane] = v;
        }
      }
      __syncwarp();
    }
    if (p.ric_Lr0) {   // what hpipm-cpp itself reads for its stage-0 reconstruction (ocp_qp_ipm_solver.cpp:352)
      const double* Lr = wLr(0);   // row-major
      double* o = p.ric_Lr0 + (size_t)q * nu * nu;
      for (int e = lane; e < nu * nu; e += 32) {
        const int i = e % nu, j = e / nu;
        o[e] = i >= j ? Lr[i * nu + j] : 0.0;
      }
    }
    // ---- stage 0 (ocp_qp_ipm_solver.cpp:349-373) from the raw stage-0 blocks -------------------------
    {
      const int r_A = 0, r_B = nx * nx, r_b = r_B + nx * nu, r_S = r_b + nx, r_Q = r_S + nu * nx, r_q = r_Q + nx * nx;
      const double* raw = p.raw0 + (size_t)q * (r_q + nx);
      const double *A0 = raw + r_A, *B0 = raw + r_B, *b0 = raw + r_b, *S0 = raw + r_S, *Q0 = raw + r_Q, *q0 = raw + r_q;
      const double* x0 = p.x0 + (size_t)q * nx;
      const double* P1 = wP(1);
      const double* Li0 = wLi(0);
      const double* u0 = wN(L.ws_z, 0);
      // scratch in shared memory: H0 (nu x nx) -> sM, GH (nu x nx) -> sB, AtP (nx x nx) -> sM + offset
      double* H0 = sM;                       // [i*nx + j]
      double* AtP = sM + D::kNU * D::kNX;    // [i*nx + j]
      double* GH = sB;                       // [i*nx + j]
      double* BtP = sB + D::kNU * D::kNX;
      for (int e = lane; e < nu * nx; e += 32) {
        const int i = e / nx, j = e % nx;
        double acc = 0.0;
        for (int l = 0; l < nx; ++l) acc += B0[l + nx * i] * P1[l * nx + j];
        BtP[e] = acc;
      }
      for (int e = lane; e < nx * nx; e += 32) {
        const int i = e / nx, j = e % nx;
        double acc = 0.0;
        for (int l = 0; l < nx; ++l) acc += A0[l + nx * i] * P1[l * nx + j];
        AtP[e] = acc;
      }
      __syncwarp();
      for (int e = lane; e < nu * nx; e += 32) {
        const int i = e / nx, j = e % nx;
        double acc = 0.0;
        for (int l = 0; l < nx; ++l) acc += BtP[i * nx + l] * A0[l + nx * j];
        H0[e] = S0[i + nu * j] + acc;
      }
      __syncwarp();
      for (int e = lane; e < nu * nx; e += 32) {  // GH = Linv^T (Linv H0)
        const int i = e / nx, c = e % nx;
        double acc = 0.0;
        for (int j = i; j < nu; ++j) {
          double y = 0.0;
          for (int l = 0; l <= j; ++l) y += Li0[j * nu + l] * H0[l * nx + c];
          acc += Li0[j * nu + i] * y;
        }
        GH[e] = acc;
      }
      __syncwarp();
      double* K0 = p.ric_K + ((size_t)q * N) * nu * nx;
      for (int e = lane; e < nu * nx; e += 32) K0[(e / nx) + nu * (e % nx)] = -GH[e];
      if (lane < nu) {
        double acc = 0.0;
        for (int j = 0; j < nx; ++j) acc += -GH[lane * nx + j] * x0[j];
        st[lane] = u0[lane] - acc;  // k0
        p.ric_k[((size_t)q * N) * nu + lane] = st[lane];
      }
      // p1 in the exported convention
      if (lane < nx) sxn[lane] = p.ric_p[((size_t)q * (N + 1) + 1) * nx + lane];
      __syncwarp();
      double* P0 = p.ric_P + ((size_t)q * (N + 1)) * nx * nx;
      for (int e = lane; e < nx * nx; e += 32) {
        const int i = e % nx, j = e / nx;  // column-major output
        double s1 = 0.0, s2 = 0.0;
        for (int l = 0; l < nu; ++l) s1 += H0[l * nx + i] * GH[l * nx + j];
        for (int l = 0; l < nx; ++l) s2 += AtP[i * nx + l] * A0[l + nx * j];
        P0[e] = (Q0[i + nx * j] - s1) + s2;
      }
      __syncwarp();
      if (lane < nx) {
        double s1 = 0.0, s2 = 0.0, s3 = 0.0;
        for (int l = 0; l < nx; ++l) s1 += A0[l + nx * lane] * sxn[l];
        for (int l = 0; l < nx; ++l) s2 += AtP[lane * nx + l] * b0[l];
        for (int l = 0; l < nu; ++l) s3 += H0[l * nx + lane] * st[l];
        const double p0 = ((q0[lane] + s1) + s2) + s3;
        p.ric_p[((size_t)q * (N + 1)) * nx + lane] = p0;
        double acc = 0.0;
        for (int j = 0; j < nx; ++j) acc += P0[lane + nx * j] * x0[j];
        p.sol_pi[((size_t)q * (N + 1)) * nx + lane] = p0 + acc;
      }
      __syncwarp();
    }
  }

  // ------------------------------------------------------------------------------------------------
  // d_ocp_qp_ipm_solve for QP q
  // ------------------------------------------------------------------------------------------------
  __device__ void solve_one(int qp) {
    q = qp;
    const srbd_ipm_args& a = p.a;
    const int nc_mask = count_masks();
    double res[4], mu, obj;
    int kk = 0, status = 0;
    if (p.stat && lane == 0) {
      double* r = p.stat + (size_t)q * p.stat_rows * SRBD_STAT_M;
      for (int i = 0; i < p.stat_rows * SRBD_STAT_M; ++i) r[i] = 0.0;
    }
    if (nc_mask == 0) {
      // unconstrained: d_ocp_qp_fact_solve_kkt_unconstr on the QP itself, iter = 0
      backward(true, false, true);
      forward(L.ws_z, L.ws_pi, true);
      for (int k = 0; k <= N; ++k) {
        const St s = stage(k);
        for (int j = lane; j < s.nc; j += 32) {
          wC(L.ws_ll, k)[j] = 0.0; wC(L.ws_lu, k)[j] = 0.0; wC(L.ws_tl, k)[j] = 0.0; wC(L.ws_tu, k)[j] = 0.0;
        }
      }
      __syncwarp();
      residuals(res, mu, obj, 0);
      stat_row(0, 6, res, 4);
      stat_row(0, 10, &obj, 1);
      const bool nan = (res[0] != res[0]);
      status = nan ? 3 : 0;
    } else {
      init_var();
      residuals(res, mu, obj, nc_mask);
      stat_row(0, 5, &mu, 1);
      stat_row(0, 6, res, 4);
      stat_row(0, 10, &obj, 1);
      double alpha = 1.0;
      for (; kk < a.iter_max && alpha > a.alpha_min &&
             (res[0] > a.tol_stat || res[1] > a.tol_eq || res[2] > a.tol_ineq || res[3] > a.tol_comp);
           ++kk) {
        backup_res_m();
        // affine step (factorization)
        backward(true, true, false);
        forward(L.ws_dz, L.ws_dpi, false);
        double ap, ad;
        dlam_dt_alpha(ap, ad);
        double itr[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // itref_pred, itref_corr, lin_res_{stat,eq,ineq,comp}
        if (a.itref_pred_max > 0) {
          double lin[4];
          itr[0] = refine(a.itref_pred_max, res, lin);
          step_lengths(ap, ad);
        }
        const double alpha_aff = fmin(ap, ad);
        double row[5] = {alpha_aff, 0.0, 0.0, 0.0, 0.0};
        if (a.pred_corr == 1) {
          const double mua = mu_aff(alpha_aff, nc_mask);
          const double tmp = mua / mu;
          const double sigma = tmp * tmp * tmp;
          row[1] = mua; row[2] = sigma;
          double sm = sigma * mu;
          sm = sm > a.tau_min ? sm : a.tau_min;
          centering(sm, true);
          backward(false, true, false);
          forward(L.ws_dz, L.ws_dpi, false);
          dlam_dt_alpha(ap, ad);
          if (a.itref_corr_max > 0) {
            itr[1] = refine(a.itref_corr_max, res, itr + 2);
            step_lengths(ap, ad);
          }
          if (a.cond_pred_corr == 1) {
            const double muc = mu_aff(fmin(ap, ad), nc_mask);
            if (muc > a.cond_factor * mua) {
              centering(sigma * mu, false);
              backward(false, true, false);
              forward(L.ws_dz, L.ws_dpi, false);
              dlam_dt_alpha(ap, ad);
            }
          }
        }
        if (!a.split_step) {
          const double al = fmin(ap, ad);
          ap = al; ad = al;
        }
        alpha = fmin(ap, ad);
        row[3] = ap; row[4] = ad;
        update(shorten(ap), shorten(ad));
        residuals(res, mu, obj, nc_mask);
        stat_row(kk + 1, 0, row, 5);
        stat_row(kk + 1, 12, itr, 6);
        stat_row(kk + 1, 5, &mu, 1);
        stat_row(kk + 1, 6, res, 4);
        stat_row(kk + 1, 10, &obj, 1);
      }
      const bool nan = (res[0] != res[0]) || (mu != mu);
      if (kk == a.iter_max) status = 1;
      else if (alpha <= a.alpha_min) status = 2;
      else if (nan) status = 3;
      else status = 0;
    }
    write_outputs(nc_mask == 0);
    if (lane == 0) {
      p.iter[q] = kk;
      p.status[q] = status;
      for (int i = 0; i < 4; ++i) p.res_max[4 * (size_t)q + i] = res[i];
    }
    __syncwarp();
  }
};

template <class D>
__global__ void __launch_bounds__(32) ipm_solve_kernel(const IpmParams p) {
  __shared__ double smem[Solver<D>::kSmemDoubles];
  __shared__ int sidx[3 * kMaxNB];
  __shared__ int s_next;
  extern __shared__ __align__(16) double ws_dyn[];   // the workspace of this CTA's QP when p.ws_in_smem
  if (p.gate && *p.gate != p.gate_value) return;
  if (p.run_gate && *p.run_gate == 0) return;
  if (p.qlist && *p.qcount == 0) return;  // empty rescue list (the usual case): nothing to set up
  Solver<D> S(p, smem, sidx, ws_dyn);
  // per-CTA partial batch statistics (fused epilogue; one set of atomics per CTA at the end)
  long long it_sum = 0, solves = 0;
  int st_cnt[5] = {0, 0, 0, 0, 0};
  double rmax[4] = {0.0, 0.0, 0.0, 0.0};
  const int n_work = p.qlist ? *p.qcount : p.B;
  for (;;) {
    if (threadIdx.x == 0) s_next = atomicAdd(p.counter, 1);
    __syncwarp();
    const int idx = s_next;
    __syncwarp();
    if (idx >= n_work) break;
    const int qp = p.qlist ? p.qlist[idx] : idx;
    if (p.frozen && p.frozen[qp]) continue;
    S.solve_one(qp);
    if (threadIdx.x == 0) {
      const int it = p.iter[qp], st = p.status[qp];
      it_sum += it;
      solves += 1;
      st_cnt[st < 0 || st > 4 ? 4 : st] += 1;
      atomicAdd((unsigned long long*)&p.bstats->iter_hist[it < SRBD_HIST_BINS ? it : SRBD_HIST_BINS - 1], 1ull);
      for (int i = 0; i < 4; ++i) rmax[i] = fmax(rmax[i], p.res_max[4 * (size_t)qp + i]);
    }
  }
  if (threadIdx.x == 0 && solves > 0) {
    atomicAdd((unsigned long long*)&p.bstats->solves, (unsigned long long)solves);
    atomicAdd((unsigned long long*)&p.bstats->iter_sum, (unsigned long long)it_sum);
    for (int i = 0; i < 5; ++i)
      if (st_cnt[i]) atomicAdd((unsigned long long*)&p.bstats->status_count[i], (unsigned long long)st_cnt[i]);
    for (int i = 0; i < 4; ++i) {  // max of non-negative doubles == max of their bit patterns
      atomicMax((unsigned long long*)&p.bstats->res_max[i], (unsigned long long)__double_as_longlong(rmax[i]));
    }
  }
}

}  // namespace srbd
