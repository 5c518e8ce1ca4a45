// layout.cuh — data layout of the packed OCP-QP stage records and of the per-QP workspace.
//
// Packed stage records follow HPIPM's d_ocp_qp fields (hpipm_d_ocp_qp.h:57-69 of the reference's
// vendored headers) in BLASFEO panel-major element order (ps = 4, blasfeo_common.h:112,
// blasfeo_block_size.h:76-77): element (i,j) of a matrix with `cn` padded columns lives at
//     (i/4)*4*cn + j*4 + i%4 .
// One array per field (structure of arrays over fields), each [B][stage][record]:
//   BAbt  : (n+1) x nx      rows [B^T (nu); A^T (nx_k); b^T]                 stages 0..N-1
//   RSQrq : (n+1) x n       [[R, .],[S^T.., Q]] lower triangle used, last row [r; q]   stages 0..N
//   DCt   : n x ng          [D^T; C^T]                                        stages 0..N
//   d     : [lo(nb+ng) | pad | -up(nb+ng) | pad]   (upper bounds stored NEGATED like HPIPM)
//   dmask : same layout, 1.0 / 0.0
// Rows/columns of a stage are in the stage's own z = [u_k; x_k] coordinates (stage 0 has no x after
// the x0 embedding, stage N has no u): n_k = nu_k + nx_k, the rq / b row sits at index n_k.
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../include/srbd_b200.h"

namespace srbd {

constexpr int kMaxNX = 12;
constexpr int kMaxNU = 12;
constexpr int kMaxN = kMaxNX + kMaxNU;  // 24
constexpr int kMaxNG = 24;
constexpr int kMaxNB = 24;
constexpr int kMaxNC = kMaxNB + kMaxNG;  // per side

// Stage-dependent part of an SRBD BAbt record (K1 -> K3's SRBD variant).  Of the 336 doubles of a dense record only 48
// depend on the stage (srbd_model.cuh: babt_term); the rest are model constants (0, 1, dt, dt/m) that are the same for
// every stage of every QP.  K1 therefore also writes a 72-double "dyn" record per (QP, stage): the 36 16-byte chunks of
// the dense record that contain a stage-dependent element, verbatim (constant neighbours included), chunk c at record
// offset 2c.  K3 keeps the constants resident in its shared-memory tiles and copies chunk c to tile offset
// babt_dyn_off(c) (the tile is the dense record as it lies in HBM): 576 instead of 2688 bytes per stage and sweep.
//   c  0..11 : the b row (24, j), j = c                 c 21..25 : B^T skew block of leg 0, offsets 12..21
//   c 12..17 : A^T rows 12..15 x columns 0..2           c 26..30 : B^T skew block of leg 1 (rows 6, 7 | row 8)
//   c 18..20 : A^T rows 16, 17 x columns 0..2           c 31..35 : A^T skew block (rows 18, 19 | row 20)
constexpr int kBabtDynChunks = 36;
constexpr int kBabtDyn = 2 * kBabtDynChunks;   // doubles per dyn record
__host__ __device__ inline int babt_dyn_off(int c) {
  if (c < 12) return 288 + 4 * c;
  if (c < 18) return 144 + 2 * (c - 12);
  if (c < 21) return 192 + 4 * (c - 18);
  if (c < 26) return 12 + 2 * (c - 21);
  if (c < 29) return 62 + 4 * (c - 26);
  if (c < 31) return 108 + 4 * (c - 29);
  if (c < 34) return 206 + 4 * (c - 31);
  return 252 + 4 * (c - 34);
}

__host__ __device__ inline int round4(int v) { return (v + 3) & ~3; }
__host__ __device__ inline int pm_index(int i, int j, int cn) { return (i >> 2) * 4 * cn + j * 4 + (i & 3); }

// Everything the kernels need to know about dimensions and strides; filled on the host.
struct QpLayout {
  int N, nx, nu, nbx, nbu, ng, ngN;
  int nm;    // nu + nx
  int ncm;   // max constraints per side over the stages
  int ngm;   // max(ng, ngN)
  // packed record strides (doubles)
  int babt_cn, babt_stride;  // cn = round4(nx), stride = round4(nm+1)*cn
  int rsq_cn, rsq_stride;    // cn = round4(nm)
  int dct_cn, dct_stride;    // cn = round4(ngm), stride = round4(nm)*cn
  int d_stride;              // 2*ncm
  int nct;                   // exported lam/t length per QP (HPIPM order, only real rows)
  int idxb0[kMaxNB], idxb1[kMaxNB], idxbN[kMaxNB];  // box index maps of stage 0 / interior / stage N
  // per-QP workspace offsets (doubles) — see ipm_solve.cuh
  int ws_z, ws_pi, ws_ll, ws_lu, ws_tl, ws_tu;
  int ws_dz, ws_dpi, ws_dll, ws_dlu, ws_dtl, ws_dtu;
  int ws_rg, ws_rb, ws_rdl, ws_rdu, ws_rml, ws_rmu, ws_rmlb, ws_rmub;
  int ws_Li, ws_Ls, ws_lv, ws_P, ws_p, ws_Lr;
  // iterative refinement (ipm_solve.cuh: refine): residual of the linear system and the correction
  int ws_lg, ws_lb, ws_ldl, ws_ldu, ws_lml, ws_lmu, ws_cz, ws_cpi;
  int ws_size;
  int ws_size_core;   // without the refinement vectors (they come last)
};

__host__ __device__ inline int stage_nu(const QpLayout& L, int k) { return k < L.N ? L.nu : 0; }
__host__ __device__ inline int stage_nx(const QpLayout& L, int k) { return k > 0 ? L.nx : 0; }
__host__ __device__ inline int stage_nb(const QpLayout& L, int k) {
  return (k < L.N ? L.nbu : 0) + (k > 0 ? L.nbx : 0);
}
__host__ __device__ inline int stage_ng(const QpLayout& L, int k) { return k < L.N ? L.ng : L.ngN; }

inline int make_layout(const srbd_qp_dims& d, const int* idxbx, const int* idxbu, QpLayout* out) {
  QpLayout L{};
  if (d.N < 1 || d.nx < 1 || d.nu < 1 || d.nx > kMaxNX || d.nu > kMaxNU || d.ng < 0 || d.ng > kMaxNG ||
      d.ngN < 0 || d.ngN > kMaxNG || d.nbx < 0 || d.nbx > d.nx || d.nbu < 0 || d.nbu > d.nu)
    return -1;
  L.N = d.N; L.nx = d.nx; L.nu = d.nu; L.nbx = d.nbx; L.nbu = d.nbu; L.ng = d.ng; L.ngN = d.ngN;
  L.nm = d.nu + d.nx;
  L.ngm = d.ng > d.ngN ? d.ng : d.ngN;
  L.ncm = d.nbu + d.nbx + L.ngm;
  if (L.ncm < 1) L.ncm = 1;
  L.babt_cn = round4(L.nx); L.babt_stride = round4(L.nm + 1) * L.babt_cn;
  L.rsq_cn = round4(L.nm);  L.rsq_stride = round4(L.nm + 1) * L.rsq_cn;
  L.dct_cn = round4(L.ngm > 0 ? L.ngm : 1); L.dct_stride = round4(L.nm) * L.dct_cn;
  L.d_stride = 2 * L.ncm;
  L.nct = 0;
  for (int k = 0; k <= d.N; ++k) L.nct += 2 * (stage_nb(L, k) + stage_ng(L, k));
  for (int j = 0; j < d.nbu; ++j) { L.idxb0[j] = idxbu ? idxbu[j] : j; L.idxb1[j] = L.idxb0[j]; }
  for (int j = 0; j < d.nbx; ++j) {
    const int ix = idxbx ? idxbx[j] : j;
    L.idxb1[d.nbu + j] = d.nu + ix;
    L.idxbN[j] = ix;
  }
  for (int j = 0; j < d.nbu; ++j) if (L.idxb0[j] < 0 || L.idxb0[j] >= d.nu) return -1;
  for (int j = 0; j < d.nbx; ++j) if (L.idxbN[j] < 0 || L.idxbN[j] >= d.nx) return -1;
  const int S = d.N + 1;
  int o = 0;
  auto take = [&](int per_stage) { int r = o; o += S * per_stage; return r; };
  L.ws_z = take(L.nm); L.ws_pi = take(L.nx);
  L.ws_ll = take(L.ncm); L.ws_lu = take(L.ncm); L.ws_tl = take(L.ncm); L.ws_tu = take(L.ncm);
  L.ws_dz = take(L.nm); L.ws_dpi = take(L.nx);
  L.ws_dll = take(L.ncm); L.ws_dlu = take(L.ncm); L.ws_dtl = take(L.ncm); L.ws_dtu = take(L.ncm);
  L.ws_rg = take(L.nm); L.ws_rb = take(L.nx);
  L.ws_rdl = take(L.ncm); L.ws_rdu = take(L.ncm); L.ws_rml = take(L.ncm); L.ws_rmu = take(L.ncm);
  L.ws_rmlb = take(L.ncm); L.ws_rmub = take(L.ncm);
  L.ws_Li = take(L.nu * L.nu); L.ws_Ls = take(L.nx * L.nu); L.ws_lv = take(L.nu);
  L.ws_P = take(L.nx * L.nx); L.ws_p = take(L.nx); L.ws_Lr = take(L.nu * L.nu);
  L.ws_size_core = (o + 15) & ~15;
  L.ws_lg = take(L.nm); L.ws_lb = take(L.nx); L.ws_ldl = take(L.ncm); L.ws_ldu = take(L.ncm);
  L.ws_lml = take(L.ncm); L.ws_lmu = take(L.ncm); L.ws_cz = take(L.nm); L.ws_cpi = take(L.nx);
  L.ws_size = (o + 15) & ~15;
  *out = L;
  return 0;
}

}  // namespace srbd
