// srbd_model.cuh — K1 (linearize) and K2 (assemble): stage-parallel kernels over (QP, stage).
//
// K1 replaces SRBDModel::GetShootingDynamic / GetContinuousDynamic and the SO(3) helpers
// (dynamics/SRBD_model.cpp:75-235, dynamics/orientation_tool.h:55-227 of the reference): RK4 defect
// plus Euler Jacobians (only k1's Jacobians reach the outputs, SRBD_model.cpp:180-181, so the 3
// discarded Jacobian evaluations of the reference are not computed).  K2 replaces
// SRBDModel::GetConstrain / Barrier and NMPCSolver::prepareQpStructures
// (SRBD_model.cpp:237-295, NMPC_solver.cpp:276-314).
//
// Both write the packed BLASFEO panel-major stage records K3 consumes (layout.cuh).  One thread
// computes one (QP, stage) in registers and leaves a COMPACT result (the structurally non-zero
// entries) in shared memory; then each warp expands 32 consecutive records into the dense
// panel-major layout with fully coalesced 8-byte stores (the dense records of consecutive
// (QP, stage) pairs are contiguous in HBM).  These kernels are HBM-write bound.
#pragma once
#include <cuda_runtime.h>

#include "layout.cuh"

namespace srbd {

struct ModelDev {
  srbd_model_params m;
  double Ac[24 * 12];  // constraint Jacobian, row-major [g][j] (constant: depends on Rf, mu, Lfx, Lfz only)
};

struct Mat3 {
  double a[9];  // row-major a[3*i+j]
};

__device__ __forceinline__ Mat3 m3_skew(const double v[3]) {
  Mat3 r;
  r.a[0] = 0.0;   r.a[1] = -v[2]; r.a[2] = v[1];
  r.a[3] = v[2];  r.a[4] = 0.0;   r.a[5] = -v[0];
  r.a[6] = -v[1]; r.a[7] = v[0];  r.a[8] = 0.0;
  return r;
}
__device__ __forceinline__ Mat3 m3_mul(const Mat3& A, const Mat3& B) {
  Mat3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) s += A.a[3 * i + k] * B.a[3 * k + j];
      C.a[3 * i + j] = s;
    }
  return C;
}
__device__ __forceinline__ Mat3 m3_mul_bt(const Mat3& A, const Mat3& B) {
  Mat3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) s += A.a[3 * i + k] * B.a[3 * j + k];
      C.a[3 * i + j] = s;
    }
  return C;
}
__device__ __forceinline__ void m3_vec(const Mat3& A, const double v[3], double o[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) s += A.a[3 * i + k] * v[k];
    o[i] = s;
  }
}
__device__ __forceinline__ double clamp_theta(const double r[3]) {  // orientation_tool.h:78-83
  double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  return th < 1e-10 ? 1e-10 : th;
}

// orientation_tool.h:75-86, 144-157 share theta / V / V^2
struct So3 {
  double th, sn, cs;
  Mat3 S, V, VV;  // raw skew, normalized skew, its square
};
__device__ __forceinline__ So3 so3_prepare(const double r[3]) {
  So3 o;
  o.th = clamp_theta(r);
  sincos(o.th, &o.sn, &o.cs);
  o.S = m3_skew(r);
#pragma unroll
  for (int i = 0; i < 9; ++i) o.V.a[i] = o.S.a[i] / o.th;
  o.VV = m3_mul(o.V, o.V);
  return o;
}
__device__ __forceinline__ Mat3 so3_expm(const So3& o) {
  const Mat3 SS = m3_mul(o.S, o.S);
  const double a = o.sn / o.th, b = (1.0 - o.cs) / (o.th * o.th);
  Mat3 R;
#pragma unroll
  for (int i = 0; i < 9; ++i) R.a[i] = (((i % 4) == 0 ? 1.0 : 0.0) + a * o.S.a[i]) + b * SS.a[i];
  return R;
}
__device__ __forceinline__ Mat3 so3_jlt(const So3& o) {  // orientation_tool.h:144-157
  const double cot = 1.0 / tan(0.5 * o.th);
  const double a = 0.5 * cot * o.th;
  Mat3 J;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const double I = (i % 4) == 0 ? 1.0 : 0.0;
    J.a[i] = (a * I + (1.0 - a) * (o.VV.a[i] + I)) - (0.5 * o.th) * o.V.a[i];
  }
  return J;
}
__device__ __forceinline__ Mat3 so3_jl(const So3& o) {  // orientation_tool.h:128-140
  const double s = o.sn / o.th, c = (1.0 - o.cs) / o.th;
  Mat3 J;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const double I = (i % 4) == 0 ? 1.0 : 0.0;
    J.a[i] = (s * I + (1.0 - s) * (o.VV.a[i] + I)) + c * o.V.a[i];
  }
  return J;
}

// xdot = f(x,u)  (SRBD_model.cpp:75-98).  Optionally returns what the Jacobian needs.
__device__ __forceinline__ void srbd_f(const srbd_model_params& m, const double x[12], const double u[12],
                                       double dx[12], So3* so_out, Mat3* Jlt_out, Mat3* RLR_out, double w_out[3]) {
  const So3 so = so3_prepare(x);
  const Mat3 R = so3_expm(so);
  const Mat3 Jlt = so3_jlt(so);
  Mat3 Li;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Li.a[3 * i + j] = m.inertia_inv[i + 3 * j];
  const Mat3 RL = m3_mul(R, Li);
  const Mat3 RLR = m3_mul_bt(RL, R);
  double w[3];
  m3_vec(RLR, x + 3, w);
  m3_vec(Jlt, w, dx);
  double d0[3], d1[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    d0[i] = m.foot_pos[i] - x[6 + i];
    d1[i] = m.foot_pos[3 + i] - x[6 + i];
  }
  const Mat3 S0 = m3_skew(d0), S1 = m3_skew(d1);
  double t0[3], t1[3];
  m3_vec(S0, u + 0, t0);
  m3_vec(S1, u + 6, t1);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    dx[3 + i] = ((u[3 + i] + u[9 + i]) + t0[i]) + t1[i];
    dx[6 + i] = x[9 + i];
    dx[9 + i] = (u[i] + u[6 + i]) / m.mass + m.gravity[i];
  }
  if (so_out) {
    *so_out = so; *Jlt_out = Jlt; *RLR_out = RLR;
    w_out[0] = w[0]; w_out[1] = w[1]; w_out[2] = w[2];
  }
}

// the 3x3 block d(rdot)/dr of j_fx (SRBD_model.cpp:105-118, orientation_tool.h:164-227)
__device__ __forceinline__ Mat3 srbd_drdot_dr(const So3& so, const Mat3& Jlt, const Mat3& RLR, const double l[3],
                                              const double w[3], const double r[3]) {
  const double th = so.th, sn = so.sn, cs = so.cs;
  const double th2 = th * th, th3 = th2 * th;
  const double c1 = (th * sn + (2.0 * (cs - 1.0))) / th3;
  const double c2 = -(2.0 * th - 3.0 * sn + th * cs) / th3;
  const double ca = (th - sn) / th3, cb = (1.0 - cs) / th2;
  Mat3 base;
#pragma unroll
  for (int i = 0; i < 9; ++i) base.a[i] = c1 * so.V.a[i] + c2 * so.VV.a[i];
  Mat3 nJ;
#pragma unroll
  for (int i = 0; i < 9; ++i) nJ.a[i] = -Jlt.a[i];
  Mat3 out;  // column k = (-Jlt dJl_k Jlt) w
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double e[3] = {0.0, 0.0, 0.0};
    e[k] = 1.0;
    const Mat3 E = m3_skew(e);
    const Mat3 ES = m3_mul(E, so.S), SE = m3_mul(so.S, E);
    Mat3 dJ;
#pragma unroll
    for (int i = 0; i < 9; ++i) dJ.a[i] = (ca * (ES.a[i] + SE.a[i]) + cb * E.a[i]) + base.a[i] * r[k];
    const Mat3 T = m3_mul(m3_mul(nJ, dJ), Jlt);
    double col[3];
    m3_vec(T, w, col);
    out.a[0 + k] = col[0]; out.a[3 + k] = col[1]; out.a[6 + k] = col[2];
  }
  const Mat3 Jl = so3_jl(so);
  const Mat3 Sl = m3_skew(l), Sw = m3_skew(w);
  Mat3 X = m3_mul(RLR, Sl);
#pragma unroll
  for (int i = 0; i < 9; ++i) X.a[i] = X.a[i] - Sw.a[i];
  X = m3_mul(m3_mul(Jlt, X), Jl);
#pragma unroll
  for (int i = 0; i < 9; ++i) out.a[i] = out.a[i] + X.a[i];
  return out;
}

// ---------------------------------------------------------------------------------------------------
// K1
// ---------------------------------------------------------------------------------------------------
struct LinParams {
  int B, N;
  const double* x;     // [B][N+1][12]
  const double* u;     // [B][N][12]
  const double* x0;    // [B][12]  (absolute initial state; dx0 = x0 - x[:,0] is embedded into b0)
  double* babt;        // [B][N][28*12] panel-major
  double* gdyn;        // [B][N][kBabtDyn] stage-dependent chunks of the same records (layout.cuh), or null
  double* gconst;      // [28*12] the dense record of (QP 0, stage 1): the source of the model constants of K3's compact BAbt
                       // streaming, or null
  int dense;           // write the dense records (0: only dyn / gconst -- the throughput path of the SRBD K3 variant reads
                       // nothing else; the dense ones are then materialised lazily, capi.cu: ensure_babt)
  // optional work list: linearize only the QPs qlist[0 .. *qcount) (dense records for the rescue list of K3, capi.cu)
  const int* qlist;
  const int* qcount;
  double* defect;      // [B][N][12]
  double* raw0;        // [B][raw0 stride]: A0, B0, b0 (column-major), rest written by K2
  double* dx0;         // [B][12]
  const int* run_gate; // device-side SQP loop (srbd_sqp_solve): return at once if *run_gate == 0 (every QP converged)
};

constexpr int kLinCompact = 40;  // 18 (drdot/d[r,l]) + 3 (F sum) + 6 (d0,d1) + 12 (b) + pad
constexpr int kLinThreads = 128;
constexpr int kRaw0Stride = 144 + 144 + 12 + 144 + 144 + 12;  // A0,B0,b0,S0,Q0,q0

// Element (i,j) of the dense record BAbt = [B^T; A^T; b^T] as ONE term of the compact record c:
//     element = fma(mul, c[src], add)
// c[0..18): jfx rows 0-2 x cols 0-5 (row-major 3x6); c[18..21): F sum; c[21..24): d0; c[24..27): d1; c[27..39): b;
// c[39] = 0 (the source of the constant elements, mul = 0).  336 elements, 48 of them depend on the stage.
struct BabtTerm { int src; double mul, add; };
__host__ __device__ __forceinline__ BabtTerm babt_term(int i, int j, double dt, double minv, bool stage0) {
  BabtTerm t{39, 0.0, 0.0};
  if (i < 12) {  // B^T[i][j] = dt * jfu[j][i]
    const int row = j, col = i;  // jfu(row, col)
    if (row >= 3 && row < 6) {
      const int a = row - 3;
      if (col < 3 || (col >= 6 && col < 9)) {  // skew(d) block: skew(d)(a, b)
        const int b = col < 3 ? col : col - 6;
        if (a != b) {
          const int k3 = 3 - a - b;  // the remaining index
          t.src = (col < 3 ? 21 : 24) + k3;
          t.mul = ((b - a + 3) % 3 == 1) ? -dt : dt;  // skew(a,a+1) = -v[k]
        }
      } else {  // identity blocks at cols 3-5 and 9-11
        const int b = col < 6 ? col - 3 : col - 9;
        t.add = (a == b) ? dt : 0.0;
      }
    } else if (row >= 9) {
      const int a = row - 9;
      if (col < 3) t.add = (a == col) ? dt * minv : 0.0;
      else if (col >= 6 && col < 9) t.add = (a == col - 6) ? dt * minv : 0.0;
    }
    return t;
  }
  if (stage0) {  // nx[0] := 0: stage 0 is [B^T (12 rows); b^T] only
    if (i == 12) { t.src = 27 + j; t.mul = 1.0; t.add = -0.0; }
    return t;
  }
  if (i < 24) {  // A^T[c][j] = A[j][c] = delta + dt * jfx[j][c]
    const int row = j, col = i - 12;
    t.add = row == col ? 1.0 : 0.0;
    if (row < 3) {
      if (col < 6) { t.src = row * 6 + col; t.mul = dt; }
    } else if (row < 6) {
      if (col >= 6 && col < 9) {
        const int a = row - 3, b = col - 6;
        if (a != b) {
          t.src = 18 + (3 - a - b);
          t.mul = ((b - a + 3) % 3 == 1) ? -dt : dt;
        }
      }
    } else if (row < 9) {
      if (col >= 9 && row - 6 == col - 9) t.add = dt;
    }
    return t;
  }
  if (i == 24) { t.src = 27 + j; t.mul = 1.0; t.add = -0.0; }  // (c + -0.0 = c for every c, -0.0 included)
  return t;
}
__host__ __device__ __forceinline__ double babt_elem(const double* c, int i, int j, double dt, double minv, bool stage0) {
  const BabtTerm t = babt_term(i, j, dt, minv, stage0);
  return fma(t.mul, c[t.src], t.add);
}

#ifndef SRBD_K1_MIN_CTAS
#define SRBD_K1_MIN_CTAS 1   // resident CTAs per SM the register allocation of K1 is bounded for (A/B knob)
#endif
__global__ void __launch_bounds__(kLinThreads, SRBD_K1_MIN_CTAS) linearize_kernel(const LinParams p, const ModelDev* __restrict__ md) {
  __shared__ double sc[kLinThreads][kLinCompact + 1];
  __shared__ srbd_model_params sm;
  if (p.run_gate && *p.run_gate == 0) return;
  // items are (QP, stage) pairs; with a work list the QP index is qlist[item / N] and the blocks beyond the list return at once
  const long long total = (long long)(p.qlist ? *p.qcount : p.B) * p.N;
  const long long item0 = (long long)blockIdx.x * kLinThreads;
  if (item0 >= total) return;
  {
    const int nw = sizeof(srbd_model_params) / sizeof(double);
    const double* src = reinterpret_cast<const double*>(&md->m);
    double* dst = reinterpret_cast<double*>(&sm);
    for (int i = threadIdx.x; i < nw; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const long long item = item0 + threadIdx.x;
  if (item < total) {
    const int qi = (int)(item / p.N), k = (int)(item % p.N);
    const int q = p.qlist ? p.qlist[qi] : qi;
    double x[12], xn[12], u[12];
    const double* xp = p.x + ((size_t)q * (p.N + 1) + k) * 12;
#pragma unroll
    for (int i = 0; i < 12; ++i) { x[i] = xp[i]; xn[i] = xp[12 + i]; u[i] = p.u[((size_t)q * p.N + k) * 12 + i]; }
    double k1[12], k2[12], k3[12], k4[12], xt[12], w[3];
    So3 so; Mat3 Jlt, RLR;
    const double dt = sm.dt;
    srbd_f(sm, x, u, k1, &so, &Jlt, &RLR, w);
#pragma unroll
    for (int i = 0; i < 12; ++i) xt[i] = x[i] + (0.5 * dt) * k1[i];
    srbd_f(sm, xt, u, k2, nullptr, nullptr, nullptr, nullptr);
#pragma unroll
    for (int i = 0; i < 12; ++i) xt[i] = x[i] + (0.5 * dt) * k2[i];
    srbd_f(sm, xt, u, k3, nullptr, nullptr, nullptr, nullptr);
#pragma unroll
    for (int i = 0; i < 12; ++i) xt[i] = x[i] + dt * k3[i];
    srbd_f(sm, xt, u, k4, nullptr, nullptr, nullptr, nullptr);
    double* c = sc[threadIdx.x];
    double f[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const double xg = x[i] + (dt / 6.0) * (((k1[i] + 2.0 * k2[i]) + 2.0 * k3[i]) + k4[i]);
      f[i] = xn[i] - xg;
      p.defect[((size_t)q * p.N + k) * 12 + i] = f[i];
    }
    const Mat3 drr = srbd_drdot_dr(so, Jlt, RLR, x + 3, w, x);
    const Mat3 drl = m3_mul(Jlt, RLR);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        c[i * 6 + j] = drr.a[3 * i + j];
        c[i * 6 + 3 + j] = drl.a[3 * i + j];
      }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      c[18 + i] = u[i] + u[6 + i];
      c[21 + i] = sm.foot_pos[i] - x[6 + i];
      c[24 + i] = sm.foot_pos[3 + i] - x[6 + i];
    }
    if (k == 0) {
      // x0 embedding: b0 <- A0 dx0 + b0 (hpipm-cpp/src/ocp_qp_ipm_solver.cpp:225); keep raw A0,B0,b0
      double dx0[12], b0[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        dx0[i] = p.x0[(size_t)q * 12 + i] - x[i];
        p.dx0[(size_t)q * 12 + i] = dx0[i];
      }
      const double minv = 1.0 / sm.mass;
      if (p.raw0) {
      double* raw = p.raw0 + (size_t)q * kRaw0Stride;
      // raw column-major A0 = I + dt jfx, B0 = dt jfu, b0 (the stage-0 reconstruction of the facade needs them)
      for (int e = 0; e < 288; ++e) raw[e] = 0.0;
      double* A0 = raw;
      double* B0 = raw + 144;
      for (int i = 0; i < 12; ++i) A0[i + 12 * i] = 1.0;
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 6; ++j) A0[i + 12 * j] = (i == j ? 1.0 : 0.0) + dt * c[i * 6 + j];
        A0[(6 + i) + 12 * (9 + i)] = 0.0 + dt * 1.0;
        B0[(3 + i) + 12 * (3 + i)] = 0.0 + dt * 1.0;
        B0[(3 + i) + 12 * (9 + i)] = 0.0 + dt * 1.0;
        B0[(9 + i) + 12 * (0 + i)] = 0.0 + dt * minv;
        B0[(9 + i) + 12 * (6 + i)] = 0.0 + dt * minv;
      }
      {
        const Mat3 SF = m3_skew(c + 18), S0 = m3_skew(c + 21), S1 = m3_skew(c + 24);
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) {
            if (i == j) continue;
            A0[(3 + i) + 12 * (6 + j)] = 0.0 + dt * SF.a[3 * i + j];
            B0[(3 + i) + 12 * (0 + j)] = 0.0 + dt * S0.a[3 * i + j];
            B0[(3 + i) + 12 * (6 + j)] = 0.0 + dt * S1.a[3 * i + j];
          }
      }
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        double s = 0.0;
        for (int j = 0; j < 12; ++j) s += A0[i + 12 * j] * dx0[j];
        b0[i] = s + (-f[i]);
        raw[288 + i] = -f[i];
      }
      } else {
        // throughput path (no raw stage-0 blocks: they only serve the Riccati exports / getters): the same A0 entries,
        // element (i, j) = the term of row 12 + j, column i of the dense record, summed in the same order
        c[39] = 0.0;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < 12; ++j) s += babt_elem(c, 12 + j, i, dt, minv, false) * dx0[j];
          b0[i] = s + (-f[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 12; ++i) c[27 + i] = b0[i];
    } else {
#pragma unroll
      for (int i = 0; i < 12; ++i) c[27 + i] = -f[i];
    }
    c[39] = 0.0;
  }
  __syncthreads();
  // expand: each warp writes its 32 records, coalesced over the contiguous dense records.  The per-lane terms of an
  // interior record (11 elements per lane) are resolved once, so an element costs one shared-memory load, one FMA
  // and one store; stage-0 records (1 in N) take the generic path.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double dt = sm.dt, minv = 1.0 / sm.mass;
  constexpr int kRec = 28 * 12, kSlots = (kRec + 31) / 32;
  int tsrc[kSlots];
  double tmul[kSlots], tadd[kSlots];
#pragma unroll
  for (int sl = 0; sl < kSlots; ++sl) {
    const int e = lane + 32 * sl, pnl = e / 48, rem = e - pnl * 48;
    const BabtTerm t = babt_term(4 * pnl + (rem & 3), rem >> 2, dt, minv, false);
    tsrc[sl] = t.src; tmul[sl] = t.mul; tadd[sl] = t.add;
  }
  const long long it0 = item0 + warp * 32;
  // record index of item i: (QP, stage) -> q * N + k (with a work list q = qlist[i / N])
  auto record_of = [&](long long i) -> long long {
    if (!p.qlist) return i;
    const long long qi = i / p.N;
    return (long long)p.qlist[qi] * p.N + (i - qi * p.N);
  };
  int k = (int)(it0 % p.N);
  for (int r = 0; r < 32; ++r, k = (k + 1 == p.N ? 0 : k + 1)) {
    const long long it = it0 + r;
    if (it >= total) break;
    const double* c = sc[warp * 32 + r];
    const long long rec = record_of(it);
    // (the dense record of (QP 0, stage 1) always goes to gconst: K3's compact streaming takes the model constants from it)
    double* dst = p.dense ? p.babt + (size_t)rec * kRec : ((p.gconst && rec == 1 && k == 1) ? p.gconst : nullptr);
    if (!dst) continue;
    if (p.dense && p.gconst && rec == 1 && k == 1) {
#pragma unroll
      for (int sl = 0; sl < kSlots; ++sl) {
        const int e = lane + 32 * sl;
        if (e < kRec) p.gconst[e] = fma(tmul[sl], c[tsrc[sl]], tadd[sl]);
      }
    }
    if (k != 0) {
#pragma unroll
      for (int sl = 0; sl < kSlots; ++sl) {
        const int e = lane + 32 * sl;
        if (e < kRec) dst[e] = fma(tmul[sl], c[tsrc[sl]], tadd[sl]);
      }
    } else {
      for (int e = lane; e < kRec; e += 32) {
        const int pnl = e / 48, rem = e - pnl * 48;
        dst[e] = babt_elem(c, 4 * pnl + (rem & 3), rem >> 2, dt, minv, true);
      }
    }
  }
  // the dyn records (layout.cuh: babt_dyn_off): the same terms, for the 72 doubles of the 36 stage-dependent chunks.  Stage
  // 0 is NOT special here: its record carries the b row (b0 with the x0 embedding) at row 24 and finite A^T entries, which
  // K3 multiplies by dx0 = 0 / masks exactly like the zeros of the dense stage-0 record.
  if (p.gdyn) {
    constexpr int kDS = (kBabtDyn + 31) / 32;
    int dsrc[kDS];
    double dmul[kDS], dadd[kDS];
#pragma unroll
    for (int sl = 0; sl < kDS; ++sl) {
      const int e = lane + 32 * sl;
      const int off = babt_dyn_off(e < kBabtDyn ? e >> 1 : 0) + (e & 1), pnl = off / 48, rem = off - pnl * 48;
      const BabtTerm t = babt_term(4 * pnl + (rem & 3), rem >> 2, dt, minv, false);
      dsrc[sl] = t.src; dmul[sl] = t.mul; dadd[sl] = t.add;
    }
    for (int r = 0; r < 32; ++r) {
      const long long it = it0 + r;
      if (it >= total) break;
      const double* c = sc[warp * 32 + r];
      double* dst = p.gdyn + (size_t)record_of(it) * kBabtDyn;
#pragma unroll
      for (int sl = 0; sl < kDS; ++sl) {
        const int e = lane + 32 * sl;
        if (e < kBabtDyn) dst[e] = fma(dmul[sl], c[dsrc[sl]], dadd[sl]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// K2
// ---------------------------------------------------------------------------------------------------
struct AsmParams {
  int B, N, mode;
  const double* x;      // [B][N+1][12]
  const double* u;      // [B][N][12]
  const double* xref;   // [B][N+1][12]
  const uint8_t* contact;  // [B][N][2] or null
  double* rsq;    // [B][N+1][28*24]
  double* srec;   // [B][N+1][kSrec] compact stage records for the SRBD K3 variant
  double* dct;    // [B][N+1][24*24]
  double* d;      // [B][N+1][48]
  double* dmask;  // [B][N+1][48]
  double* raw0;   // S0, Q0, q0 part
  double* fcon;   // [B][N][24] constraint values (diagnostics, NMPC_solver.cpp:289) or null
  // optional work list: assemble only the QPs qlist[0 .. *qcount) (dense records for the rescue list of K3, capi.cu)
  const int* qlist;
  const int* qcount;
  const int* run_gate; // device-side SQP loop: return at once if *run_gate == 0
};

constexpr int kAsmThreads = 64;
constexpr int kAsmCompact = 64;  // 24 ddb + 12 r + 12 q + 1 flag ... see below

__device__ __forceinline__ void barrier_fn(double v, double mu, double theta, double* db, double* ddb) {
  if (v > theta) {  // SRBD_model.cpp:264-277
    *db = -mu / v;
    *ddb = mu / (v * v);
  } else {          // :279-294
    *db = mu * (v - 2.0 * theta) / (theta * theta);
    *ddb = mu / (theta * theta);
  }
}
__device__ __forceinline__ bool row_soft_in_hard_mode(int g) {
  const int r = g % 12;
  return r == 10 || r == 11;  // the +-x^T tau pair has no strict interior (NMPC_solver.cpp:301 keeps 20 rows)
}

// R-block element (i, j) of RSQrq: R delta_ij + sum_g Ac[g][i] ddb_g Ac[g][j] (NMPC_solver.cpp:308), summed in the
// order g = 0..23.  Ac is two 12x6 blocks (fill_Ac / SRBD_model.cpp:244), so only the 12 rows of the common leg can
// contribute, and in HARD_INEQ mode only its two relaxed rows have ddb != 0 (g_lo = 10, g_cnt = 2): the skipped terms
// are exact zeros, so the sum is bit-identical to the full one.
__device__ __forceinline__ double rblock_elem(const double* sAc, const double* c, int i, int j, double R, int g_lo,
                                              int g_cnt) {
  // (explicit roundings: the hoisted form of the throughput path below must give the same bits)
  double s = 0.0;
  if (i / 6 == j / 6) {
    const int g0 = 12 * (i / 6) + g_lo;
    for (int g = g0; g < g0 + g_cnt; ++g) s = __fma_rn(__dmul_rn(sAc[g * 12 + i], c[g]), sAc[g * 12 + j], s);
  }
  return __dadd_rn(i == j ? R : 0.0, s);
}
// The R tile of the compact stage record in HARD_INEQ mode with everything that does not depend on the stage hoisted out
// of the per-item loop: a lane owns up to four elements (i, j) of the tile (panel prefixes 16 + 32 + 48); only the two
// relaxed rows g0 = 12 leg + 10, g0 + 1 of the common leg contribute, R_ij = R delta_ij + (a_i0 c_g0) a_j0 + (a_i1 c_g1) a_j1
// -- the same operations in the same order as rblock_elem.
struct RTileHard {
  double ai0[4], aj0[4], ai1[4], aj1[4], rd[4];
  int g0[4], off[4];
  bool valid[4];
  __device__ __forceinline__ void init(const double* sAc, double R, int lane) {
#pragma unroll
    for (int sidx = 0; sidx < 4; ++sidx) {
      const int pnl = sidx < 3 ? sidx : 2, sl = sidx < 3 ? 0 : 1;
      const int e = lane + 32 * sl, j = e >> 2, i = 4 * pnl + (e & 3);
      valid[sidx] = j < 4 * pnl + 4;
      off[sidx] = (pnl == 0 ? 0 : (pnl == 1 ? 16 : 48)) + e;
      const bool same = valid[sidx] && (i / 6 == j / 6);
      const int g = same ? 12 * (i / 6) + 10 : 10;
      g0[sidx] = g;
      ai0[sidx] = same ? sAc[g * 12 + i] : 0.0; aj0[sidx] = same ? sAc[g * 12 + j] : 0.0;
      ai1[sidx] = same ? sAc[(g + 1) * 12 + i] : 0.0; aj1[sidx] = same ? sAc[(g + 1) * 12 + j] : 0.0;
      rd[sidx] = (valid[sidx] && i == j) ? R : 0.0;
    }
  }
  __device__ __forceinline__ void write(double* srec, const double* c) const {
#pragma unroll
    for (int sidx = 0; sidx < 4; ++sidx)
      if (valid[sidx]) {
        double t = __fma_rn(__dmul_rn(ai0[sidx], c[g0[sidx]]), aj0[sidx], 0.0);
        t = __fma_rn(__dmul_rn(ai1[sidx], c[g0[sidx] + 1]), aj1[sidx], t);
        srec[off[sidx]] = __dadd_rn(rd[sidx], t);
      }
  }
};
// One RSQrq record (28 x 24 panel-major: 7 panels of 96 doubles, 3 per lane and panel).  TYPE 0: first stage
// (nu = 12, nx = 0), 1: interior, 2: last stage (nu = 0, nx = 12): with the panel and the stage type known at compile
// time most elements fold to a stored zero.
template <int TYPE>
__device__ __forceinline__ double rsq_elem(int i, int j, const double* c, const double* sAc, const double* sQ,
                                           const double* sQf, double R, int g_lo, int g_cnt) {
  constexpr int nu = TYPE == 2 ? 0 : 12, nx = TYPE == 0 ? 0 : 12, n = nu + nx;
  double v = 0.0;
  if (i < n && j < n) {
    if (i < nu && j < nu) v = rblock_elem(sAc, c, i, j, R, g_lo, g_cnt);
    else if (i >= nu && j >= nu && i == j) v = TYPE == 2 ? sQf[i - nu] : sQ[i - nu];
  } else if (i == n && j < n) {
    v = j < nu ? c[24 + j] : c[36 + (j - nu)];
  }
  return v;
}
// Compact stage record for the SRBD K3 variant (ipm_srbd.cuh), kSrec doubles on 128-byte lines:
//   [0, 96)    the lower 12 x 12 block of rows 0..11 of RSQrq (R_k, or Q_N at the last stage) as the prefixes of its
//              three row panels: panel p (rows 4p..4p+3) holds columns 0..4p+3, element (i, j) at 4 j + (i & 3)
//   [108, 132) the gradient row n of RSQrq ([r; q], r only at stage 0, q_N at the last stage)
//   [144, 168) lg (lower bounds of the 24 rows), [168, 192) their masks
// i.e. exactly the shared-memory R tile of the kernel followed by the two per-row vectors every sweep loads: one
// linear cp.async stream and one base pointer instead of panel-prefix / strided-row gathers from three arrays.
constexpr int kSrec = 192;
// DENSE: also the full 28 x 24 record (generic kernel, getters, rescue pass); the throughput path writes the compact
// record only (K3's SRBD variant reads nothing else: 192 instead of 1536 doubles per stage)
template <int TYPE, bool DENSE>
__device__ __forceinline__ void write_rsq(double* dst, double* srec, const double* c, const double* sAc,
                                          const double* sQ, const double* sQf, double R, int lane, int g_lo,
                                          int g_cnt) {
  constexpr int n = (TYPE == 2 ? 0 : 12) + (TYPE == 0 ? 0 : 12);
#pragma unroll
  for (int pnl = 0; pnl < (DENSE ? 7 : 3); ++pnl)
#pragma unroll
    for (int sl = 0; sl < 3; ++sl) {
      const int e = lane + 32 * sl, j = e >> 2, i = 4 * pnl + (e & 3);
      const bool in_tile = pnl < 3 && j < 4 * pnl + 4;   // the R-tile prefix of panels 0..2: columns j <= 4 pnl + 3
      if (!DENSE && 32 * sl >= 16 * (pnl + 1)) continue; // (compile time: no lane of this slot is inside the prefix)
      if (DENSE || in_tile) {
        const double v = rsq_elem<TYPE>(i, j, c, sAc, sQ, sQf, R, g_lo, g_cnt);
        if (DENSE) dst[96 * pnl + e] = v;
        if (in_tile) srec[(pnl == 0 ? 0 : (pnl == 1 ? 16 : 48)) + e] = v;
      }
    }
  if (lane < 24) srec[108 + lane] = lane < n ? rsq_elem<TYPE>(n, lane, c, sAc, sQ, sQf, R, g_lo, g_cnt) : 0.0;
  if (lane < 12) { srec[96 + lane] = 0.0; srec[132 + lane] = 0.0; }
}

// MODE: the assemble mode as a template parameter (p.mode must equal it): the throughput instantiation <false, HARD_INEQ>
// skips the barrier of the hard rows and writes the R tile from hoisted coefficients
template <bool DENSE, int MODE>
__global__ void __launch_bounds__(kAsmThreads) assemble_kernel(const AsmParams p, const ModelDev* __restrict__ md) {
  // compact per item: [0..24) ddb (barrier curvature per row, 0 for hard rows), [24..36) r, [36..48) q,
  // [48..72) lg = -f, stage kind in sk[]
  __shared__ double sc[kAsmThreads][73];
  __shared__ double sAc[24 * 12];
  __shared__ double sQ[12], sQf[12];
  __shared__ double sR;
  if (p.run_gate && *p.run_gate == 0) return;
  for (int i = threadIdx.x; i < 288; i += blockDim.x) sAc[i] = md->Ac[i];
  if (threadIdx.x < 12) { sQ[threadIdx.x] = md->m.Q[threadIdx.x]; sQf[threadIdx.x] = md->m.Qf[threadIdx.x]; }
  if (threadIdx.x == 0) sR = md->m.R;
  __syncthreads();
  const int S = p.N + 1;
  // items are (QP, stage) pairs; with a work list the QP index is qlist[item / S] (the list is short: one grid-stride
  // trip per 64 * gridDim.x items)
  const long long total = (long long)(p.qlist ? *p.qcount : p.B) * S;
  for (long long item0 = (long long)blockIdx.x * kAsmThreads; item0 < total; item0 += (long long)gridDim.x * kAsmThreads) {
  const long long item = item0 + threadIdx.x;
  __syncthreads();  // (second trip: everyone is done reading sc)
  if (item < total) {
    const int qi = (int)(item / S), k = (int)(item % S);
    const int q = p.qlist ? p.qlist[qi] : qi;
    double* c = sc[threadIdx.x];
    const double* xk = p.x + ((size_t)q * S + k) * 12;
    const double* xr = p.xref + ((size_t)q * S + k) * 12;
    if (k < p.N) {
      const srbd_model_params& m = md->m;
      double u[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) u[i] = p.u[((size_t)q * p.N + k) * 12 + i];
      double r[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) r[i] = 0.0;
      for (int g = 0; g < 24; ++g) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 12; ++j) s += sAc[g * 12 + j] * u[j];
        const int leg = g / 12, rr = g % 12;
        double bc = 0.0;
        if (rr == 4) {
          const int st = p.contact ? (int)p.contact[((size_t)q * p.N + k) * 2 + leg] : 1;
          bc = st ? m.fmax : m.swing_fmax;
        } else if (rr == 5) {
          bc = -m.fmin;
        }
        const double f = s + bc;
        if (p.fcon) p.fcon[((size_t)q * p.N + k) * 24 + g] = f;
        double db = 0.0, ddb = 0.0;
        const bool hard_row = MODE == SRBD_HARD_INEQ && !row_soft_in_hard_mode(g);
        if (!hard_row) barrier_fn(f, m.mu_b, m.theta_b, &db, &ddb);
        c[g] = ddb;
        c[48 + g] = -f;
        if (!hard_row) {   // (a hard row adds exact zeros)
#pragma unroll
          for (int j = 0; j < 12; ++j) r[j] += sAc[g * 12 + j] * db;
        }
      }
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        c[24 + i] = sR * u[i] + r[i];                // r = R u + Ac^T db (NMPC_solver.cpp:309)
        c[36 + i] = sQ[i] * (xk[i] - xr[i]);         // q = Q (x - xref)  (:306)
      }
      if (DENSE && k == 0) {
        double* raw = p.raw0 + (size_t)q * kRaw0Stride;
        for (int e = 0; e < 144; ++e) {
          raw[300 + e] = 0.0;                              // S0
          raw[444 + e] = (e % 13 == 0) ? sQ[e / 13] : 0.0; // Q0
        }
        for (int i = 0; i < 12; ++i) raw[588 + i] = c[36 + i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 12; ++i) c[36 + i] = sQf[i] * (xk[i] - xr[i]);  // :313
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int g_lo = MODE == SRBD_HARD_INEQ ? 10 : 0, g_cnt = MODE == SRBD_HARD_INEQ ? 2 : 12;
  constexpr bool kFast = !DENSE && MODE == SRBD_HARD_INEQ;
  RTileHard rt;
  if (kFast) rt.init(sAc, sR, lane);
  const long long it0 = item0 + warp * 32;
  int k = (int)(it0 % S);
  for (int rI = 0; rI < 32; ++rI, k = (k == p.N ? 0 : k + 1)) {
    const long long it_ = it0 + rI;
    if (it_ >= total) break;
    // record index in the [B][N+1] arrays (with a work list: the listed QP's own records)
    const long long it = p.qlist ? (long long)p.qlist[it_ / S] * S + k : it_;
    const double* c = sc[warp * 32 + rI];
    const int nu = k < p.N ? 12 : 0;
    // RSQrq: 28 x 24 panel-major
    double* dst = DENSE ? p.rsq + (size_t)it * (28 * 24) : nullptr;
    double* sr = p.srec + (size_t)it * kSrec;
    if (kFast && k < p.N) {
      rt.write(sr, c);
      const int n = k == 0 ? 12 : 24;
      if (lane < 24) sr[108 + lane] = lane < n ? (lane < 12 ? c[24 + lane] : c[36 + (lane - 12)]) : 0.0;
      if (lane < 12) { sr[96 + lane] = 0.0; sr[132 + lane] = 0.0; }
    } else if (k == 0) write_rsq<0, DENSE>(dst, sr, c, sAc, sQ, sQf, sR, lane, g_lo, g_cnt);
    else if (k < p.N) write_rsq<1, DENSE>(dst, sr, c, sAc, sQ, sQf, sR, lane, g_lo, g_cnt);
    else write_rsq<2, DENSE>(dst, sr, c, sAc, sQ, sQf, sR, lane, g_lo, g_cnt);
    // DCt (n x 24): D^T = Ac^T in the u rows, C = 0; d = [lg | 0 | 0(-ug) | 0], masks
    double* dd = DENSE ? p.dct + (size_t)it * (24 * 24) : nullptr;
    double* dv = DENSE ? p.d + (size_t)it * 48 : nullptr;
    double* dk = DENSE ? p.dmask + (size_t)it * 48 : nullptr;
    if (k < p.N) {
      if (DENSE) {
#pragma unroll
        for (int pnl = 0; pnl < 6; ++pnl)
#pragma unroll
          for (int sl = 0; sl < 3; ++sl) {
            const int e = lane + 32 * sl, g = e >> 2, i = 4 * pnl + (e & 3);
            dd[96 * pnl + e] = (i < nu) ? sAc[g * 12 + i] : 0.0;
          }
      }
      for (int e = lane; e < 48; e += 32) {
        const bool lower = e < 24;
        const int g = lower ? e : e - 24;
        const bool hard = (MODE == SRBD_HARD_INEQ) && !row_soft_in_hard_mode(g);
        const double dvv = (lower && MODE == SRBD_HARD_INEQ) ? c[48 + g] : 0.0;
        const double dkk = (lower && hard) ? 1.0 : 0.0;
        if (DENSE) { dv[e] = dvv; dk[e] = dkk; }
        if (lower) { sr[144 + g] = dvv; sr[168 + g] = dkk; }
      }
    } else {
      for (int e = lane; e < 48; e += 32) {
        if (DENSE) { dv[e] = 0.0; dk[e] = 0.0; }
        sr[144 + e] = 0.0;
      }
    }
  }
  }  // grid-stride trip
}

}  // namespace srbd
