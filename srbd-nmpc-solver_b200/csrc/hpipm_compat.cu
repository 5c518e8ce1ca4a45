// hpipm_compat.cu -- the HPIPM C symbols hpipm-cpp links against (include/hpipm_b200_compat.h, SURVEY.md 8(b) "Option A"),
// implemented on top of the batch C-ABI of this library (include/srbd_b200.h) with a batch of one.
//
// Host code only: it never touches a kernel or device memory directly.  The objects HPIPM's callers allocate
// (*_memsize -> malloc -> *_create) hold plain column-major host arrays in exactly the layout srbd_qp_host /
// srbd_sol_host describe, so d_ocp_qp_ipm_solve is: srbd_set_ipm_args, srbd_qp_upload (one H2D copy + pack_kernel +
// structure detection), srbd_qp_solve (K3), srbd_download_solution / _stats / _ric_Lr0.  Device contexts live in a pool
// keyed by the QP dimensions.  Reference call sites: hpipm-cpp/src/ocp_qp_ipm_solver.cpp:103-130,283-407 and
// hpipm-cpp/src/detail/*_wrapper.cpp.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hpipm_b200_compat.h"
#include "../../include/srbd_b200.h"

namespace {

constexpr size_t kAlign = 64;
constexpr uint64_t kMagicQp = 0x5352424451503031ull, kMagicSol = 0x535242444f4c3031ull, kMagicWs = 0x5352424457533031ull;
constexpr int kStatM = 18;   // columns of the statistics table (ocp_qp_ipm_solver.cpp:381)

inline size_t up(size_t n) { return (n + kAlign - 1) & ~(kAlign - 1); }
inline char* aligned(void* p) { return reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + kAlign - 1) & ~(uintptr_t)(kAlign - 1)); }

std::string g_err;

bool same_dims(const srbd_qp_dims& a, const srbd_qp_dims& b) { return std::memcmp(&a, &b, sizeof(a)) == 0; }

// The uniform dimensions the batch C-ABI works with, from HPIPM's per-stage arrays; `why` says what does not fit.
bool shape_of(const d_ocp_qp_dim* dim, srbd_qp_dims* out, std::string* why) {
  auto no = [&](const char* m) { if (why) *why = m; return false; };
  if (!dim || !dim->nx) return no("d_ocp_qp_dim was not created");
  const int N = dim->N;
  if (N < 1) return no("N < 1");
  srbd_qp_dims d{};
  d.N = N; d.nx = dim->nx[N]; d.nu = dim->nu[0]; d.nbx = dim->nbx[N]; d.nbu = dim->nbu[0]; d.ng = dim->ng[0]; d.ngN = dim->ng[N];
  if (dim->nx[0] != 0 || dim->nbx[0] != 0)
    return no("nx[0] != 0: the B200 path expects the initial state embedded (hpipm-cpp sets nx[0] = nbx[0] = 0, ocp_qp_ipm_solver.cpp:128-130)");
  for (int k = 1; k <= N; ++k)
    if (dim->nx[k] != d.nx || dim->nbx[k] != d.nbx) return no("nx / nbx differ between the stages 1..N");
  for (int k = 0; k < N; ++k)
    if (dim->nu[k] != d.nu || dim->nbu[k] != d.nbu || dim->ng[k] != d.ng) return no("nu / nbu / ng differ between the stages 0..N-1");
  if (dim->nu[N] != 0 || dim->nbu[N] != 0) return no("nu[N] != 0");
  for (int k = 0; k <= N; ++k)
    if (dim->ns[k] || dim->nsbx[k] || dim->nsbu[k] || dim->nsg[k]) return no("soft constraints are not supported (hpipm-cpp forces nsg = 0 too)");
  if (d.nx < 1 || d.nu < 1 || d.nx > 12 || d.nu > 12 || d.ng > 24 || d.ngN > 24 || d.nbx > d.nx || d.nbu > d.nu)
    return no("dimensions beyond the compiled maxima (nx, nu <= 12, ng <= 24)");
  *out = d;
  return true;
}

// ---- host stores inside the caller's memory blocks ------------------------------------------------------------------------
struct QpStore {
  uint64_t magic;
  int valid, idx_mismatch;
  srbd_qp_dims d;
  // offsets in doubles from data()
  size_t A, B, b, Q, S, R, q, r, lbx, ubx, lbxm, ubxm, lbu, ubu, lbum, ubum, C, D, lg, ug, lgm, ugm, CN, lgN, ugN, lgNm, ugNm, x0, n_dbl;
  size_t idxbx, idxbu;   // offsets in ints from idata()
  int have_idxbx, have_idxbu;
  double* data() { return reinterpret_cast<double*>(reinterpret_cast<char*>(this) + up(sizeof(QpStore))); }
  int* idata() { return reinterpret_cast<int*>(data() + n_dbl); }
};

void layout_qp(const srbd_qp_dims& d, QpStore* s) {
  const size_t N = d.N, S = N + 1, nx = d.nx, nu = d.nu, nbx = d.nbx, nbu = d.nbu, ng = d.ng, ngN = d.ngN;
  size_t o = 0;
  auto take = [&](size_t n) { const size_t r = o; o += n; return r; };
  s->A = take(N * nx * nx); s->B = take(N * nx * nu); s->b = take(N * nx);
  s->Q = take(S * nx * nx); s->S = take(N * nu * nx); s->R = take(N * nu * nu); s->q = take(S * nx); s->r = take(N * nu);
  s->lbx = take(S * nbx); s->ubx = take(S * nbx); s->lbxm = take(S * nbx); s->ubxm = take(S * nbx);
  s->lbu = take(N * nbu); s->ubu = take(N * nbu); s->lbum = take(N * nbu); s->ubum = take(N * nbu);
  s->C = take(N * ng * nx); s->D = take(N * ng * nu); s->lg = take(N * ng); s->ug = take(N * ng); s->lgm = take(N * ng); s->ugm = take(N * ng);
  s->CN = take(ngN * nx); s->lgN = take(ngN); s->ugN = take(ngN); s->lgNm = take(ngN); s->ugNm = take(ngN);
  s->x0 = take(nx);
  s->n_dbl = (o + 1) & ~(size_t)1;
  s->idxbx = 0; s->idxbu = nbx;
}
size_t qp_bytes(const srbd_qp_dims& d) {
  QpStore t{};
  layout_qp(d, &t);
  return up(sizeof(QpStore)) + t.n_dbl * sizeof(double) + (size_t)(d.nbx + d.nbu + 2) * sizeof(int);
}

struct SolStore {
  uint64_t magic;
  int valid;
  srbd_qp_dims d;
  size_t x, u, pi, lam, t, n_dbl, nct;
  double* data() { return reinterpret_cast<double*>(reinterpret_cast<char*>(this) + up(sizeof(SolStore))); }
};
void layout_sol(const srbd_qp_dims& d, SolStore* s) {
  const size_t N = d.N, S = N + 1;
  size_t o = 0;
  auto take = [&](size_t n) { const size_t r = o; o += n; return r; };
  s->nct = srbd_qp_nct(&d);
  s->x = take(S * d.nx); s->u = take(N * d.nu); s->pi = take(S * d.nx); s->lam = take(s->nct); s->t = take(s->nct);
  s->n_dbl = o;
}

struct WsStore {
  uint64_t magic;
  int valid, stat_rows, solved;
  srbd_qp_dims d;
  size_t P, p, K, k, Lr0, stat, n_dbl;
  double* data() { return reinterpret_cast<double*>(reinterpret_cast<char*>(this) + up(sizeof(WsStore))); }
};
void layout_ws(const srbd_qp_dims& d, int stat_rows, WsStore* s) {
  const size_t N = d.N, S = N + 1;
  size_t o = 0;
  auto take = [&](size_t n) { const size_t r = o; o += n; return r; };
  s->P = take(S * d.nx * d.nx); s->p = take(S * d.nx); s->K = take(N * d.nu * d.nx); s->k = take(N * d.nu);
  s->Lr0 = take((size_t)d.nu * d.nu); s->stat = take((size_t)stat_rows * kStatM);
  s->n_dbl = o;
  s->stat_rows = stat_rows;
}
int stat_rows_for(const d_ocp_qp_ipm_arg* arg) {
  // hpipm-cpp reads iter + 2 rows although only stat_max are promised (a latent over-read upstream when
  // iter >= stat_max): be generous
  int m = arg ? (arg->stat_max > arg->iter_max ? arg->stat_max : arg->iter_max) : 0;
  if (m < 50) m = 50;
  return m + 2;
}

QpStore* qp_store(d_ocp_qp* qp) {
  QpStore* s = qp ? static_cast<QpStore*>(qp->BAbt) : nullptr;
  return s && s->magic == kMagicQp ? s : nullptr;
}
SolStore* sol_store(d_ocp_qp_sol* sol) {
  SolStore* s = sol ? static_cast<SolStore*>(sol->ux) : nullptr;
  return s && s->magic == kMagicSol ? s : nullptr;
}
WsStore* ws_store(d_ocp_qp_ipm_ws* ws) {
  WsStore* s = ws ? static_cast<WsStore*>(ws->core_workspace) : nullptr;
  return s && s->magic == kMagicWs ? s : nullptr;
}

void copy_n(double* dst, const double* src, size_t n) {
  if (n == 0) return;
  if (src) std::memcpy(dst, src, n * sizeof(double));
  else std::memset(dst, 0, n * sizeof(double));
}

// ---- device contexts: one per QP shape, created on first use, destroyed at exit ----------------------------------------------
struct PoolEntry {
  srbd_qp_dims d;
  srbd_ctx* ctx;
  std::vector<double> stat;   // host scratch for the statistics table of the context's stat_rows
};
std::vector<PoolEntry>& pool() {
  static std::vector<PoolEntry>* p = new std::vector<PoolEntry>();   // (never destroyed: release_pool runs from atexit)
  return *p;
}
void release_pool() {
  for (PoolEntry& e : pool())
    if (e.ctx) srbd_ctx_destroy(e.ctx);
  pool().clear();
}
PoolEntry* acquire(const srbd_qp_dims& d) {
  for (PoolEntry& e : pool())
    if (same_dims(e.d, d)) return &e;
  int dev = 0;
  if (const char* s = std::getenv("SRBD_HPIPM_DEVICE")) dev = std::atoi(s);
  srbd_ctx* ctx = nullptr;
  const int rc = srbd_ctx_create(dev, 1, &d, nullptr, &ctx);
  if (rc != SRBD_OK || !ctx) {
    g_err = "srbd_ctx_create failed (rc " + std::to_string(rc) + "): no usable CUDA device or out of memory; there is no CPU fallback";
    return nullptr;
  }
  if (pool().empty()) std::atexit(release_pool);   // registered after the CUDA runtime's own handlers: runs before them
  pool().push_back(PoolEntry{d, ctx, {}});
  return &pool().back();
}

void fail_solve(d_ocp_qp_ipm_ws* ws, const std::string& why) {
  g_err = why;
  if (!ws) return;
  ws->status = 4;   // beyond hpipm_status' 0..3: hpipm-cpp returns HpipmStatus::UnknownFailure (ocp_qp_ipm_solver.cpp:409-414)
  ws->iter = 0;
  for (int i = 0; i < 4; ++i) ws->qp_res[i] = 0.0;
}

}  // namespace

extern "C" {

const char* hpipm_b200_last_error(void) { return g_err.c_str(); }
int hpipm_b200_pool_size(void) { return (int)pool().size(); }

// ---- dimensions -----------------------------------------------------------------------------------------------------------
hpipm_size_t d_ocp_qp_dim_memsize(int N) { return 13 * (size_t)(N + 1) * sizeof(int) + kAlign; }

void d_ocp_qp_dim_create(int N, struct d_ocp_qp_dim* dim, void* memory) {
  int* m = reinterpret_cast<int*>(aligned(memory));
  std::memset(m, 0, 13 * (size_t)(N + 1) * sizeof(int));
  int** f[13] = {&dim->nx, &dim->nu, &dim->nb, &dim->nbx, &dim->nbu, &dim->ng, &dim->ns, &dim->nsbx, &dim->nsbu, &dim->nsg,
                 &dim->nbxe, &dim->nbue, &dim->nge};
  for (int i = 0; i < 13; ++i) *f[i] = m + (size_t)i * (N + 1);
  dim->N = N;
  dim->memsize = d_ocp_qp_dim_memsize(N);
}

void d_ocp_qp_dim_copy_all(struct d_ocp_qp_dim* o, struct d_ocp_qp_dim* d) {
  const int N = o->N < d->N ? o->N : d->N;
  int* const src[13] = {o->nx, o->nu, o->nb, o->nbx, o->nbu, o->ng, o->ns, o->nsbx, o->nsbu, o->nsg, o->nbxe, o->nbue, o->nge};
  int* const dst[13] = {d->nx, d->nu, d->nb, d->nbx, d->nbu, d->ng, d->ns, d->nsbx, d->nsbu, d->nsg, d->nbxe, d->nbue, d->nge};
  for (int i = 0; i < 13; ++i) std::memcpy(dst[i], src[i], (size_t)(N + 1) * sizeof(int));
}

void d_ocp_qp_dim_set_all(int* nx, int* nu, int* nbx, int* nbu, int* ng, int* nsbx, int* nsbu, int* nsg, struct d_ocp_qp_dim* dim) {
  for (int k = 0; k <= dim->N; ++k) {
    dim->nx[k] = nx[k]; dim->nu[k] = nu[k]; dim->nbx[k] = nbx[k]; dim->nbu[k] = nbu[k]; dim->ng[k] = ng[k];
    dim->nb[k] = nbx[k] + nbu[k];
    dim->nsbx[k] = nsbx ? nsbx[k] : 0; dim->nsbu[k] = nsbu ? nsbu[k] : 0; dim->nsg[k] = nsg ? nsg[k] : 0;
    dim->ns[k] = dim->nsbx[k] + dim->nsbu[k] + dim->nsg[k];
  }
}
void d_ocp_qp_dim_set_nx(int stage, int value, struct d_ocp_qp_dim* dim) { dim->nx[stage] = value; }
void d_ocp_qp_dim_set_nbx(int stage, int value, struct d_ocp_qp_dim* dim) {
  dim->nbx[stage] = value;
  dim->nb[stage] = dim->nbx[stage] + dim->nbu[stage];
}
void d_ocp_qp_dim_set_nsbx(int stage, int value, struct d_ocp_qp_dim* dim) {
  dim->nsbx[stage] = value;
  dim->ns[stage] = dim->nsbx[stage] + dim->nsbu[stage] + dim->nsg[stage];
}

// ---- QP data ---------------------------------------------------------------------------------------------------------------
hpipm_size_t d_ocp_qp_memsize(struct d_ocp_qp_dim* dim) {
  srbd_qp_dims d;
  return (shape_of(dim, &d, nullptr) ? qp_bytes(d) : up(sizeof(QpStore))) + kAlign;
}

void d_ocp_qp_create(struct d_ocp_qp_dim* dim, struct d_ocp_qp* qp, void* memory) {
  std::memset(qp, 0, sizeof(*qp));
  qp->dim = dim;
  QpStore* s = reinterpret_cast<QpStore*>(aligned(memory));
  std::memset(s, 0, sizeof(*s));
  s->magic = kMagicQp;
  std::string why;
  s->valid = shape_of(dim, &s->d, &why) ? 1 : 0;
  if (s->valid) {
    layout_qp(s->d, s);
    std::memset(s->data(), 0, s->n_dbl * sizeof(double));
    const size_t S = s->d.N + 1, N = s->d.N;
    double* w = s->data();
    auto ones = [&](size_t off, size_t n) { for (size_t i = 0; i < n; ++i) w[off + i] = 1.0; };   // masks default to "active"
    ones(s->lbxm, S * s->d.nbx); ones(s->ubxm, S * s->d.nbx); ones(s->lbum, N * s->d.nbu); ones(s->ubum, N * s->d.nbu);
    ones(s->lgm, N * s->d.ng); ones(s->ugm, N * s->d.ng); ones(s->lgNm, s->d.ngN); ones(s->ugNm, s->d.ngN);
    std::memset(s->idata(), 0, (size_t)(s->d.nbx + s->d.nbu + 2) * sizeof(int));
  } else {
    g_err = why;
  }
  qp->BAbt = s;
  qp->memsize = d_ocp_qp_memsize(dim);
}

void d_ocp_qp_copy_all(struct d_ocp_qp* o, struct d_ocp_qp* d) {
  QpStore *so = qp_store(o), *sd = qp_store(d);
  if (!so || !sd || !so->valid || !sd->valid || !same_dims(so->d, sd->d)) return;
  std::memcpy(sd->data(), so->data(), so->n_dbl * sizeof(double) + (size_t)(so->d.nbx + so->d.nbu + 2) * sizeof(int));
  sd->idx_mismatch = so->idx_mismatch; sd->have_idxbx = so->have_idxbx; sd->have_idxbu = so->have_idxbu;
}

void d_ocp_qp_set_all(double** A, double** B, double** b, double** Q, double** S, double** R, double** q, double** r,
                      int** idxbx, double** lbx, double** ubx, int** idxbu, double** lbu, double** ubu, double** C,
                      double** D, double** lg, double** ug, double**, double**, double**, double**, int**, double**, double**,
                      struct d_ocp_qp* qp) {
  QpStore* s = qp_store(qp);
  if (!s || !s->valid) return;
  const srbd_qp_dims& d = s->d;
  const size_t nx = d.nx, nu = d.nu, nbx = d.nbx, nbu = d.nbu, ng = d.ng, ngN = d.ngN;
  double* w = s->data();
  int* iw = s->idata();
  s->idx_mismatch = 0; s->have_idxbx = 0; s->have_idxbu = 0;
  auto idx = [&](int* dst, int* have, const int* src, size_t n) {
    if (n == 0 || !src) return;
    if (!*have) { std::memcpy(dst, src, n * sizeof(int)); *have = 1; }
    else if (std::memcmp(dst, src, n * sizeof(int)) != 0) s->idx_mismatch = 1;
  };
  for (int k = 0; k < d.N; ++k) {
    // stage 0 has no state after the x0 embedding: its A, Q, S, q, C blocks have a zero dimension for HPIPM and are not read
    copy_n(w + s->A + k * nx * nx, k > 0 ? A[k] : nullptr, nx * nx);
    copy_n(w + s->B + k * nx * nu, B[k], nx * nu);
    copy_n(w + s->b + k * nx, b[k], nx);
    copy_n(w + s->Q + k * nx * nx, k > 0 ? Q[k] : nullptr, nx * nx);
    copy_n(w + s->S + k * nu * nx, k > 0 && S ? S[k] : nullptr, nu * nx);
    copy_n(w + s->R + k * nu * nu, R[k], nu * nu);
    copy_n(w + s->q + k * nx, k > 0 ? q[k] : nullptr, nx);
    copy_n(w + s->r + k * nu, r[k], nu);
    if (nbu) { copy_n(w + s->lbu + k * nbu, lbu[k], nbu); copy_n(w + s->ubu + k * nbu, ubu[k], nbu); idx(iw + s->idxbu, &s->have_idxbu, idxbu[k], nbu); }
    if (ng) {
      copy_n(w + s->C + k * ng * nx, k > 0 && C ? C[k] : nullptr, ng * nx);
      copy_n(w + s->D + k * ng * nu, D[k], ng * nu);
      copy_n(w + s->lg + k * ng, lg[k], ng); copy_n(w + s->ug + k * ng, ug[k], ng);
    }
  }
  copy_n(w + s->Q + d.N * nx * nx, Q[d.N], nx * nx);
  copy_n(w + s->q + d.N * nx, q[d.N], nx);
  if (nbx)
    for (int k = 1; k <= d.N; ++k) {
      copy_n(w + s->lbx + k * nbx, lbx[k], nbx); copy_n(w + s->ubx + k * nbx, ubx[k], nbx);
      idx(iw + s->idxbx, &s->have_idxbx, idxbx[k], nbx);
    }
  if (ngN) { copy_n(w + s->CN, C[d.N], ngN * nx); copy_n(w + s->lgN, lg[d.N], ngN); copy_n(w + s->ugN, ug[d.N], ngN); }
}

static void set_mask(struct d_ocp_qp* qp, int stage, const double* vec, int which) {
  QpStore* s = qp_store(qp);
  if (!s || !s->valid || !vec || stage < 0 || stage > s->d.N) return;
  const srbd_qp_dims& d = s->d;
  double* w = s->data();
  switch (which) {
    case 0: if (stage >= 1) copy_n(w + s->lbxm + (size_t)stage * d.nbx, vec, d.nbx); break;
    case 1: if (stage >= 1) copy_n(w + s->ubxm + (size_t)stage * d.nbx, vec, d.nbx); break;
    case 2: if (stage < d.N) copy_n(w + s->lbum + (size_t)stage * d.nbu, vec, d.nbu); break;
    case 3: if (stage < d.N) copy_n(w + s->ubum + (size_t)stage * d.nbu, vec, d.nbu); break;
    case 4: if (stage < d.N) copy_n(w + s->lgm + (size_t)stage * d.ng, vec, d.ng); else copy_n(w + s->lgNm, vec, d.ngN); break;
    case 5: if (stage < d.N) copy_n(w + s->ugm + (size_t)stage * d.ng, vec, d.ng); else copy_n(w + s->ugNm, vec, d.ngN); break;
  }
}
void d_ocp_qp_set_lbx_mask(int stage, double* vec, struct d_ocp_qp* qp) { set_mask(qp, stage, vec, 0); }
void d_ocp_qp_set_ubx_mask(int stage, double* vec, struct d_ocp_qp* qp) { set_mask(qp, stage, vec, 1); }
void d_ocp_qp_set_lbu_mask(int stage, double* vec, struct d_ocp_qp* qp) { set_mask(qp, stage, vec, 2); }
void d_ocp_qp_set_ubu_mask(int stage, double* vec, struct d_ocp_qp* qp) { set_mask(qp, stage, vec, 3); }
void d_ocp_qp_set_lg_mask(int stage, double* vec, struct d_ocp_qp* qp) { set_mask(qp, stage, vec, 4); }
void d_ocp_qp_set_ug_mask(int stage, double* vec, struct d_ocp_qp* qp) { set_mask(qp, stage, vec, 5); }

// ---- solution --------------------------------------------------------------------------------------------------------------
hpipm_size_t d_ocp_qp_sol_memsize(struct d_ocp_qp_dim* dim) {
  srbd_qp_dims d;
  size_t n = up(sizeof(SolStore));
  if (shape_of(dim, &d, nullptr)) { SolStore t{}; layout_sol(d, &t); n += t.n_dbl * sizeof(double); }
  return n + kAlign;
}
void d_ocp_qp_sol_create(struct d_ocp_qp_dim* dim, struct d_ocp_qp_sol* sol, void* memory) {
  std::memset(sol, 0, sizeof(*sol));
  sol->dim = dim;
  SolStore* s = reinterpret_cast<SolStore*>(aligned(memory));
  std::memset(s, 0, sizeof(*s));
  s->magic = kMagicSol;
  s->valid = shape_of(dim, &s->d, nullptr) ? 1 : 0;
  if (s->valid) { layout_sol(s->d, s); std::memset(s->data(), 0, s->n_dbl * sizeof(double)); }
  sol->ux = s;
  sol->memsize = d_ocp_qp_sol_memsize(dim);
}
void d_ocp_qp_sol_copy_all(struct d_ocp_qp_sol* o, struct d_ocp_qp_sol* d) {
  SolStore *so = sol_store(o), *sd = sol_store(d);
  if (!so || !sd || !so->valid || !sd->valid || !same_dims(so->d, sd->d)) return;
  std::memcpy(sd->data(), so->data(), so->n_dbl * sizeof(double));
}
void d_ocp_qp_sol_get_x(int stage, struct d_ocp_qp_sol* sol, double* vec) {
  SolStore* s = sol_store(sol);
  if (!s || !s->valid || stage < 1 || stage > s->d.N || !vec) return;   // (stage 0: nx[0] = 0, nothing to copy)
  std::memcpy(vec, s->data() + s->x + (size_t)stage * s->d.nx, (size_t)s->d.nx * sizeof(double));
}
void d_ocp_qp_sol_get_u(int stage, struct d_ocp_qp_sol* sol, double* vec) {
  SolStore* s = sol_store(sol);
  if (!s || !s->valid || stage < 0 || stage >= s->d.N || !vec) return;
  std::memcpy(vec, s->data() + s->u + (size_t)stage * s->d.nu, (size_t)s->d.nu * sizeof(double));
}
void d_ocp_qp_sol_get_pi(int stage, struct d_ocp_qp_sol* sol, double* vec) {
  SolStore* s = sol_store(sol);
  if (!s || !s->valid || stage < 0 || stage >= s->d.N || !vec) return;
  std::memcpy(vec, s->data() + s->pi + (size_t)(stage + 1) * s->d.nx, (size_t)s->d.nx * sizeof(double));
}
void d_ocp_qp_sol_set_x(int stage, double* vec, struct d_ocp_qp_sol* sol) {
  SolStore* s = sol_store(sol);
  if (!s || !s->valid || stage < 1 || stage > s->d.N || !vec) return;
  std::memcpy(s->data() + s->x + (size_t)stage * s->d.nx, vec, (size_t)s->d.nx * sizeof(double));
}
void d_ocp_qp_sol_set_u(int stage, double* vec, struct d_ocp_qp_sol* sol) {
  SolStore* s = sol_store(sol);
  if (!s || !s->valid || stage < 0 || stage >= s->d.N || !vec) return;
  std::memcpy(s->data() + s->u + (size_t)stage * s->d.nu, vec, (size_t)s->d.nu * sizeof(double));
}

// ---- solver arguments --------------------------------------------------------------------------------------------------------
hpipm_size_t d_ocp_qp_ipm_arg_memsize(struct d_ocp_qp_dim*) { return kAlign; }   // (the struct itself holds everything)
void d_ocp_qp_ipm_arg_create(struct d_ocp_qp_dim*, struct d_ocp_qp_ipm_arg* arg, void*) {
  std::memset(arg, 0, sizeof(*arg));
  arg->memsize = kAlign;
}
// d_ocp_qp_ipm_arg_set_default [upstream-recalled, SURVEY.md a18]: HPIPM's per-mode defaults.  The fields this library
// honours are mapped in d_ocp_qp_ipm_solve; abs_form and lq_fact are recorded but not implemented (include/srbd_b200.h:
// srbd_ipm_args_set_mode) -- the delta-form iteration with syrk + potrf runs in every mode.
void d_ocp_qp_ipm_arg_set_default(enum hpipm_mode mode, struct d_ocp_qp_ipm_arg* a) {
  const hpipm_size_t ms = a->memsize;
  std::memset(a, 0, sizeof(*a));
  a->memsize = ms;
  a->mu0 = 1e1; a->alpha_min = 1e-12; a->res_g_max = 1e-6; a->res_b_max = 1e-8; a->res_d_max = 1e-8; a->res_m_max = 1e-8;
  a->reg_prim = 1e-15; a->lam_min = 1e-16; a->t_min = 1e-16; a->tau_min = 1e-16;
  a->iter_max = 15; a->stat_max = 15; a->pred_corr = 1; a->cond_pred_corr = 1; a->square_root_alg = 1;
  a->comp_dual_sol_eq = 1; a->comp_res_exit = 1; a->comp_res_pred = 1; a->split_step = 1; a->t_lam_min = 2;
  switch (mode) {
    case SPEED_ABS:
      a->res_g_max = 1e0; a->res_b_max = 1e0; a->res_d_max = 1e0; a->cond_pred_corr = 0; a->square_root_alg = 0;
      a->abs_form = 1; a->comp_dual_sol_eq = 0; a->comp_res_exit = 0; a->comp_res_pred = 0;
      break;
    case SPEED: break;
    case BALANCE:
      a->iter_max = 30; a->stat_max = 30; a->itref_corr_max = 2; a->lq_fact = 1; a->split_step = 0;
      break;
    case ROBUST:
      a->mu0 = 1e2; a->iter_max = 100; a->stat_max = 100; a->itref_corr_max = 4; a->lq_fact = 2; a->split_step = 0;
      break;
  }
  a->mode = (int)mode;
}
void d_ocp_qp_ipm_arg_set_mu0(double* v, struct d_ocp_qp_ipm_arg* a) { a->mu0 = *v; }
void d_ocp_qp_ipm_arg_set_iter_max(int* v, struct d_ocp_qp_ipm_arg* a) { a->iter_max = *v; }
void d_ocp_qp_ipm_arg_set_alpha_min(double* v, struct d_ocp_qp_ipm_arg* a) { a->alpha_min = *v; }
void d_ocp_qp_ipm_arg_set_tol_stat(double* v, struct d_ocp_qp_ipm_arg* a) { a->res_g_max = *v; }
void d_ocp_qp_ipm_arg_set_tol_eq(double* v, struct d_ocp_qp_ipm_arg* a) { a->res_b_max = *v; }
void d_ocp_qp_ipm_arg_set_tol_ineq(double* v, struct d_ocp_qp_ipm_arg* a) { a->res_d_max = *v; }
void d_ocp_qp_ipm_arg_set_tol_comp(double* v, struct d_ocp_qp_ipm_arg* a) { a->res_m_max = *v; }
void d_ocp_qp_ipm_arg_set_reg_prim(double* v, struct d_ocp_qp_ipm_arg* a) { a->reg_prim = *v; }
void d_ocp_qp_ipm_arg_set_warm_start(int* v, struct d_ocp_qp_ipm_arg* a) { a->warm_start = *v; }
void d_ocp_qp_ipm_arg_set_pred_corr(int* v, struct d_ocp_qp_ipm_arg* a) { a->pred_corr = *v; }
void d_ocp_qp_ipm_arg_set_ric_alg(int* v, struct d_ocp_qp_ipm_arg* a) { a->square_root_alg = *v; }
void d_ocp_qp_ipm_arg_set_split_step(int* v, struct d_ocp_qp_ipm_arg* a) { a->split_step = *v; }

// ---- workspace, solve, getters -------------------------------------------------------------------------------------------------
hpipm_size_t d_ocp_qp_ipm_ws_memsize(struct d_ocp_qp_dim* dim, struct d_ocp_qp_ipm_arg* arg) {
  srbd_qp_dims d;
  size_t n = up(sizeof(WsStore));
  const int rows = stat_rows_for(arg);
  if (shape_of(dim, &d, nullptr)) { WsStore t{}; layout_ws(d, rows, &t); n += t.n_dbl * sizeof(double); }
  else n += (size_t)rows * kStatM * sizeof(double);
  return n + kAlign;
}
void d_ocp_qp_ipm_ws_create(struct d_ocp_qp_dim* dim, struct d_ocp_qp_ipm_arg* arg, struct d_ocp_qp_ipm_ws* ws, void* mem) {
  std::memset(ws, 0, sizeof(*ws));
  ws->dim = dim;
  WsStore* s = reinterpret_cast<WsStore*>(aligned(mem));
  std::memset(s, 0, sizeof(*s));
  s->magic = kMagicWs;
  const int rows = stat_rows_for(arg);
  s->valid = shape_of(dim, &s->d, nullptr) ? 1 : 0;
  if (s->valid) layout_ws(s->d, rows, s);
  else { s->stat = 0; s->n_dbl = (size_t)rows * kStatM; s->stat_rows = rows; }
  std::memset(s->data(), 0, s->n_dbl * sizeof(double));
  ws->core_workspace = s;
  ws->stat = s->data() + s->stat;
  ws->stat_max = arg ? arg->stat_max : 0;
  ws->stat_m = kStatM;
  ws->square_root_alg = arg ? arg->square_root_alg : 0;
  ws->lq_fact = arg ? arg->lq_fact : 0;
  ws->mask_constr = 1;
  ws->memsize = d_ocp_qp_ipm_ws_memsize(dim, arg);
}

void d_ocp_qp_ipm_solve(struct d_ocp_qp* qp, struct d_ocp_qp_sol* sol, struct d_ocp_qp_ipm_arg* arg, struct d_ocp_qp_ipm_ws* ws) {
  g_err.clear();
  QpStore* q = qp_store(qp);
  SolStore* so = sol_store(sol);
  WsStore* w = ws_store(ws);
  if (!q || !so || !w || !arg) return fail_solve(ws, "d_ocp_qp_ipm_solve: an argument was not created by this library's *_create");
  if (!q->valid || !so->valid || !w->valid) {
    std::string why;
    srbd_qp_dims tmp;
    shape_of(qp->dim, &tmp, &why);
    return fail_solve(ws, "unsupported QP shape: " + (why.empty() ? std::string("dimensions changed after *_create") : why));
  }
  if (!same_dims(q->d, so->d) || !same_dims(q->d, w->d)) return fail_solve(ws, "qp, qp_sol and ws were created for different dimensions");
  if (q->idx_mismatch) return fail_solve(ws, "idxbx / idxbu differ between stages: the B200 path needs one index set for stages 1..N / 0..N-1");
  PoolEntry* e = acquire(q->d);
  if (!e) return fail_solve(ws, g_err);
  srbd_ctx* ctx = e->ctx;
  auto ck = [&](int rc, const char* what) {
    if (rc == SRBD_OK) return true;
    fail_solve(ws, std::string(what) + ": " + srbd_last_error(ctx));
    return false;
  };
  // settings: hpipm-cpp's public fields + the hidden constants of the mode (ocp_qp_ipm_solver.cpp:103-116)
  srbd_ipm_args a;
  srbd_ipm_args_default(&a);
  a.iter_max = arg->iter_max; a.alpha_min = arg->alpha_min; a.mu0 = arg->mu0;
  a.tol_stat = arg->res_g_max; a.tol_eq = arg->res_b_max; a.tol_ineq = arg->res_d_max; a.tol_comp = arg->res_m_max;
  a.reg_prim = arg->reg_prim; a.warm_start = arg->warm_start ? 1 : 0; a.pred_corr = arg->pred_corr;
  a.ric_alg = arg->square_root_alg ? 1 : 0; a.split_step = arg->split_step; a.cond_pred_corr = arg->cond_pred_corr;
  a.lam_min = arg->lam_min; a.t_min = arg->t_min; a.tau_min = arg->tau_min; a.t_lam_min = arg->t_lam_min;
  a.itref_pred_max = arg->itref_pred_max; a.itref_corr_max = arg->itref_corr_max;
  if (!ck(srbd_set_ipm_args(ctx, &a), "srbd_set_ipm_args")) return;
  if (!ck(srbd_set_outputs(ctx, 1, 1), "srbd_set_outputs")) return;
  // QP data: the store IS a srbd_qp_host batch of one, contiguous (one H2D copy)
  const srbd_qp_dims& d = q->d;
  double* qd = q->data();
  srbd_qp_host h{};
  h.A = qd + q->A; h.Bm = qd + q->B; h.b = qd + q->b; h.Q = qd + q->Q; h.S = qd + q->S; h.R = qd + q->R; h.q = qd + q->q; h.r = qd + q->r;
  if (d.nbx) { h.idxbx = q->idata() + q->idxbx; h.lbx = qd + q->lbx; h.ubx = qd + q->ubx; h.lbx_mask = qd + q->lbxm; h.ubx_mask = qd + q->ubxm; }
  if (d.nbu) { h.idxbu = q->idata() + q->idxbu; h.lbu = qd + q->lbu; h.ubu = qd + q->ubu; h.lbu_mask = qd + q->lbum; h.ubu_mask = qd + q->ubum; }
  if (d.ng) { h.C = qd + q->C; h.D = qd + q->D; h.lg = qd + q->lg; h.ug = qd + q->ug; h.lg_mask = qd + q->lgm; h.ug_mask = qd + q->ugm; }
  if (d.ngN) { h.CN = qd + q->CN; h.lgN = qd + q->lgN; h.ugN = qd + q->ugN; h.lgN_mask = qd + q->lgNm; h.ugN_mask = qd + q->ugNm; }
  h.x0 = qd + q->x0;   // zeros: hpipm-cpp has already folded x0 into b[0] and r[0] (ocp_qp_ipm_solver.cpp:225,236)
  double* sd = so->data();
  if (a.warm_start) { h.x_init = sd + so->x; h.u_init = sd + so->u; }
  if (!ck(srbd_qp_upload(ctx, &h), "srbd_qp_upload")) return;
  if (!ck(srbd_qp_solve(ctx), "srbd_qp_solve")) return;
  double* wd = w->data();
  srbd_sol_host sh{};
  sh.x = sd + so->x; sh.u = sd + so->u; sh.pi = sd + so->pi; sh.lam = sd + so->lam; sh.t = sd + so->t;
  sh.P = wd + w->P; sh.p = wd + w->p; sh.K = wd + w->K; sh.k = wd + w->k;
  if (!ck(srbd_download_solution(ctx, &sh), "srbd_download_solution")) return;
  if (!ck(srbd_download_ric_lr0(ctx, wd + w->Lr0), "srbd_download_ric_lr0")) return;
  const int rows = srbd_ctx_stat_rows(ctx);
  e->stat.resize((size_t)rows * kStatM);
  int iter = 0, status = 0;
  srbd_stats_host st{};
  st.iter = &iter; st.status = &status; st.res_max = ws->qp_res; st.stat = e->stat.data();
  if (!ck(srbd_download_stats(ctx, &st), "srbd_download_stats")) return;
  const int keep = rows < w->stat_rows ? rows : w->stat_rows;
  std::memcpy(ws->stat, e->stat.data(), (size_t)keep * kStatM * sizeof(double));
  ws->iter = iter;
  ws->status = status;
  ws->valid_ric_vec = 1;
  ws->valid_ric_p = 1;
  w->solved = 1;
}

void d_ocp_qp_ipm_get_iter(struct d_ocp_qp_ipm_ws* ws, int* iter) { *iter = ws->iter; }
void d_ocp_qp_ipm_get_status(struct d_ocp_qp_ipm_ws* ws, int* status) { *status = ws->status; }
void d_ocp_qp_ipm_get_max_res_stat(struct d_ocp_qp_ipm_ws* ws, double* v) { *v = ws->qp_res[0]; }
void d_ocp_qp_ipm_get_max_res_eq(struct d_ocp_qp_ipm_ws* ws, double* v) { *v = ws->qp_res[1]; }
void d_ocp_qp_ipm_get_max_res_ineq(struct d_ocp_qp_ipm_ws* ws, double* v) { *v = ws->qp_res[2]; }
void d_ocp_qp_ipm_get_max_res_comp(struct d_ocp_qp_ipm_ws* ws, double* v) { *v = ws->qp_res[3]; }

void d_ocp_qp_ipm_get_ric_Lr(struct d_ocp_qp*, struct d_ocp_qp_ipm_arg*, struct d_ocp_qp_ipm_ws* ws, int stage, double* Lr) {
  WsStore* w = ws_store(ws);
  if (!w || !w->valid || !w->solved || !Lr) return;
  const size_t n = (size_t)w->d.nu * w->d.nu;
  if (stage == 0) std::memcpy(Lr, w->data() + w->Lr0, n * sizeof(double));
  else if (stage > 0 && stage < w->d.N) {   // only the stage hpipm-cpp reads is exported
    std::memset(Lr, 0, n * sizeof(double));
    g_err = "d_ocp_qp_ipm_get_ric_Lr: only stage 0 is exported by the B200 path";
  }
}
void d_ocp_qp_ipm_get_ric_P(struct d_ocp_qp*, struct d_ocp_qp_ipm_arg*, struct d_ocp_qp_ipm_ws* ws, int stage, double* P) {
  WsStore* w = ws_store(ws);
  if (!w || !w->valid || !w->solved || !P || stage < 1 || stage > w->d.N) return;
  const size_t n = (size_t)w->d.nx * w->d.nx;
  std::memcpy(P, w->data() + w->P + stage * n, n * sizeof(double));
}
void d_ocp_qp_ipm_get_ric_p(struct d_ocp_qp*, struct d_ocp_qp_ipm_arg*, struct d_ocp_qp_ipm_ws* ws, int stage, double* p) {
  WsStore* w = ws_store(ws);
  if (!w || !w->valid || !w->solved || !p || stage < 1 || stage > w->d.N) return;
  std::memcpy(p, w->data() + w->p + (size_t)stage * w->d.nx, (size_t)w->d.nx * sizeof(double));
}
void d_ocp_qp_ipm_get_ric_K(struct d_ocp_qp*, struct d_ocp_qp_ipm_arg*, struct d_ocp_qp_ipm_ws* ws, int stage, double* K) {
  WsStore* w = ws_store(ws);
  if (!w || !w->valid || !w->solved || !K || stage < 1 || stage >= w->d.N) return;   // (stage 0: nu x nx[0] = nu x 0)
  const size_t n = (size_t)w->d.nu * w->d.nx;
  std::memcpy(K, w->data() + w->K + stage * n, n * sizeof(double));
}
void d_ocp_qp_ipm_get_ric_k(struct d_ocp_qp*, struct d_ocp_qp_ipm_arg*, struct d_ocp_qp_ipm_ws* ws, int stage, double* k) {
  WsStore* w = ws_store(ws);
  if (!w || !w->valid || !w->solved || !k || stage < 0 || stage >= w->d.N) return;
  std::memcpy(k, w->data() + w->k + (size_t)stage * w->d.nu, (size_t)w->d.nu * sizeof(double));
}

}  // extern "C"
