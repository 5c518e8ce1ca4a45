/*
 * ocp_qp_ipm.c — CPU ORACLE (test infrastructure): OCP-QP interior-point solve with the semantics of
 * hpipm::OcpQpIpmSolver::solve (hpipm-cpp/src/ocp_qp_ipm_solver.cpp:181-414 of the reference).
 *
 * The arithmetic of that call lives in giaf/hpipm + giaf/blasfeo (unpinned HEAD, NOT vendored in the
 * reference: only headers under hpipm-cpp/include/include).  This file restates the published HPIPM
 * algorithm (the C API it replaces is hpipm_d_ocp_qp_ipm.h:139-242):
 *   d_ocp_qp_set_all / masks           -> embed_qp()            (hpipm_d_ocp_qp.h:57-69,90)
 *   d_ocp_qp_init_var                  -> init_var()            (hpipm_d_ocp_qp_ipm.h:232)
 *   d_ocp_qp_res_compute(+inf_norm)    -> compute_res()         (hpipm_d_ocp_qp_res.h:90,94)
 *   d_ocp_qp_fact_solve_kkt_unconstr   -> kkt_solve(fact=1) on the QP itself  (hpipm_d_ocp_qp_kkt.h:54)
 *   d_ocp_qp_fact_solve_kkt_step       -> kkt_solve(fact=1)     (hpipm_d_ocp_qp_kkt.h:56)
 *   d_ocp_qp_solve_kkt_step            -> kkt_solve(fact=0)     (hpipm_d_ocp_qp_kkt.h:60)
 *   d_compute_{alpha,mu_aff,centering_correction,centering}_qp, d_update_var_qp
 *                                      -> step_length(), mu_aff(), ... (hpipm_d_core_qp_ipm_aux.h:44-62)
 *   d_ocp_qp_ipm_solve                 -> ipm_solve()           (hpipm_d_ocp_qp_ipm.h:238)
 *   getters + stage-0 reconstruction   -> write_outputs()       (ocp_qp_ipm_solver.cpp:337-373)
 * It is pinned against the reference's golden vectors (hpipm-cpp/test/sol0..14.txt) and the analytic
 * Riccati identities of hpipm-cpp/test/ocp_qp_ipm_solver.cpp:60-109; HPIPM's own iteration path is
 * "parity unpinned" (see srbd_oracle.h).
 *
 * Variable order per stage is HPIPM's ux = [u; x]; constraint order per stage is
 * [box-u, box-x, general]; lam/t export order is [lb lg ub ug] (hpipm_d_ocp_qp_sol.h:57-63).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "srbd_oracle.h"

/* Arithmetic type of the restatement.  The default build (double) is THE oracle; the same source compiled with
 * -DORC_REAL=__float128 -DORC_QUAD (oracle/Makefile: _build/libsrbd_oracle_quad.so, entry points suffixed _q) is the
 * higher-precision ARBITER: identical algorithm, constants, operation order and status logic, rounding ~1e-34, inputs
 * and outputs still double.  Where the GPU and the double oracle differ (iterates beyond 1e-9, iteration count, status)
 * the arbiter says what the algorithm does in (nearly) exact arithmetic and which side is closer to it. */
#ifdef ORC_QUAD
#include <quadmath.h>
typedef __float128 real;
#define R_SQRT sqrtq
#define R_FABS fabsq
#define R_ISNAN isnanq
#define ORC_NAME(f) f##_q
#else
typedef double real;
#define R_SQRT sqrt
#define R_FABS fabs
#define R_ISNAN isnan
#define ORC_NAME(f) f
#endif
static inline void cpy_out(double* dst, const real* src, size_t n) { for (size_t i = 0; i < n; ++i) dst[i] = (double)src[i]; }
static inline void cpy_in(real* dst, const double* src, size_t n) { for (size_t i = 0; i < n; ++i) dst[i] = (real)src[i]; }

#ifdef ORC_QUAD
#define orc_qp_nct orc_qp_nct_q
#endif
size_t orc_qp_nct(const srbd_qp_dims* d);

typedef struct {
  int N, nx, nu, nm, ncm, ngm; /* nm = nu+nx (ld of every n-row matrix), ncm = max constraints/side */
  int *nuk, *nxk, *nbk, *ngk;  /* per stage (after the x0 embedding) */
  /* QP data per stage */
  real *H, *g, *G, *bb, *DCt, *lo, *up, *ml, *mu;
  int* idxb;
  /* iterate */
  real *z, *pi, *ll, *lu, *tl, *tu;
  /* step */
  real *dz, *dpi, *dll, *dlu, *dtl, *dtu;
  /* residuals */
  real *rg, *rb, *rdl, *rdu, *rml, *rmu, *rml_bkp, *rmu_bkp;
  /* Gamma / gamma */
  real *Gl, *Gu, *gl, *gu;
  /* iterative refinement: residual of the linear system (rhs of the correction solve) and the correction */
  real *lg_, *lb_, *ldl, *ldu, *lml, *lmu, *cz, *cpi;
  real lin_max[4];
  /* Riccati factors */
  real *Lr, *Ls, *lv, *P, *p, *Lfull, *Pb;
  /* scratch */
  real *M, *AL, *gt, *tmp;
  real res_max[4], mu_res, obj;
  int nc_mask;
} W;

#define HK(w, k) ((w)->H + (size_t)(k) * (w)->nm * (w)->nm)
#define GK(w, k) ((w)->G + (size_t)(k) * (w)->nm * (w)->nx)
#define DK(w, k) ((w)->DCt + (size_t)(k) * (w)->nm * (w)->ngm)
#define VN(w, a, k) ((w)->a + (size_t)(k) * (w)->nm)  /* n-vectors */
#define VX(w, a, k) ((w)->a + (size_t)(k) * (w)->nx)  /* nx-vectors */
#define VC(w, a, k) ((w)->a + (size_t)(k) * (w)->ncm) /* constraint vectors */
#define LRK(w, k) ((w)->Lr + (size_t)(k) * (w)->nu * (w)->nu)
#define LSK(w, k) ((w)->Ls + (size_t)(k) * (w)->nx * (w)->nu)
#define PK(w, k) ((w)->P + (size_t)(k) * (w)->nx * (w)->nx)
#define LFK(w, k) ((w)->Lfull + (size_t)(k) * (w)->nm * (w)->nm)

static real* dalloc(size_t n) { return (real*)calloc(n ? n : 1, sizeof(real)); }

static W* w_create(const srbd_qp_dims* d) {
  W* w = (W*)calloc(1, sizeof(W));
  int N = d->N;
  w->N = N; w->nx = d->nx; w->nu = d->nu; w->nm = d->nu + d->nx;
  w->ngm = d->ng > d->ngN ? d->ng : d->ngN;
  w->ncm = d->nbu + d->nbx + w->ngm;
  if (w->ncm == 0) w->ncm = 1;
  if (w->ngm == 0) w->ngm = 1;
  size_t S = (size_t)(N + 1);
  w->nuk = (int*)calloc(S, sizeof(int)); w->nxk = (int*)calloc(S, sizeof(int));
  w->nbk = (int*)calloc(S, sizeof(int)); w->ngk = (int*)calloc(S, sizeof(int));
  for (int k = 0; k <= N; ++k) {
    w->nuk[k] = k < N ? d->nu : 0;
    w->nxk[k] = k > 0 ? d->nx : 0;                               /* nx[0] := 0 (ocp_qp_ipm_solver.cpp:128) */
    w->nbk[k] = (k < N ? d->nbu : 0) + (k > 0 ? d->nbx : 0);     /* nbx[0] := 0 (:129) */
    w->ngk[k] = k < N ? d->ng : d->ngN;
  }
  size_t nm = w->nm, nx = w->nx ? w->nx : 1, nu = w->nu ? w->nu : 1, nc = w->ncm;
  w->H = dalloc(S * nm * nm); w->g = dalloc(S * nm); w->G = dalloc(S * nm * nx); w->bb = dalloc(S * nx);
  w->DCt = dalloc(S * nm * w->ngm);
  w->lo = dalloc(S * nc); w->up = dalloc(S * nc); w->ml = dalloc(S * nc); w->mu = dalloc(S * nc);
  w->idxb = (int*)calloc(S * nc, sizeof(int));
  w->z = dalloc(S * nm); w->pi = dalloc(S * nx);
  w->ll = dalloc(S * nc); w->lu = dalloc(S * nc); w->tl = dalloc(S * nc); w->tu = dalloc(S * nc);
  w->dz = dalloc(S * nm); w->dpi = dalloc(S * nx);
  w->dll = dalloc(S * nc); w->dlu = dalloc(S * nc); w->dtl = dalloc(S * nc); w->dtu = dalloc(S * nc);
  w->rg = dalloc(S * nm); w->rb = dalloc(S * nx);
  w->rdl = dalloc(S * nc); w->rdu = dalloc(S * nc); w->rml = dalloc(S * nc); w->rmu = dalloc(S * nc);
  w->rml_bkp = dalloc(S * nc); w->rmu_bkp = dalloc(S * nc);
  w->Gl = dalloc(S * nc); w->Gu = dalloc(S * nc); w->gl = dalloc(S * nc); w->gu = dalloc(S * nc);
  w->lg_ = dalloc(S * nm); w->lb_ = dalloc(S * nx); w->ldl = dalloc(S * nc); w->ldu = dalloc(S * nc);
  w->lml = dalloc(S * nc); w->lmu = dalloc(S * nc); w->cz = dalloc(S * nm); w->cpi = dalloc(S * nx);
  w->Lr = dalloc(S * nu * nu); w->Ls = dalloc(S * nx * nu); w->lv = dalloc(S * nm);
  w->P = dalloc(S * nx * nx); w->p = dalloc(S * nx); w->Lfull = dalloc(S * nm * nm); w->Pb = dalloc(S * nx);
  w->M = dalloc(nm * nm); w->AL = dalloc(nm * nx); w->gt = dalloc(nm); w->tmp = dalloc(nm + nx + nc);
  return w;
}

static void w_free(W* w) {
  free(w->nuk); free(w->nxk); free(w->nbk); free(w->ngk);
  free(w->H); free(w->g); free(w->G); free(w->bb); free(w->DCt); free(w->lo); free(w->up); free(w->ml);
  free(w->mu); free(w->idxb); free(w->z); free(w->pi); free(w->ll); free(w->lu); free(w->tl); free(w->tu);
  free(w->dz); free(w->dpi); free(w->dll); free(w->dlu); free(w->dtl); free(w->dtu); free(w->rg);
  free(w->rb); free(w->rdl); free(w->rdu); free(w->rml); free(w->rmu); free(w->rml_bkp); free(w->rmu_bkp);
  free(w->Gl); free(w->Gu); free(w->gl); free(w->gu); free(w->Lr); free(w->Ls); free(w->lv); free(w->P);
  free(w->p); free(w->Lfull); free(w->Pb); free(w->M); free(w->AL); free(w->gt); free(w->tmp);
  free(w->lg_); free(w->lb_); free(w->ldl); free(w->ldu); free(w->lml); free(w->lmu); free(w->cz); free(w->cpi);
  free(w);
}

/* ------------------------------------------------------------------------------------------------ */
/* d_ocp_qp_set_all analog + x0 embedding (ocp_qp_ipm_solver.cpp:225,236; hpipm_d_ocp_qp.h:57-69)     */
/* ------------------------------------------------------------------------------------------------ */
static void embed_qp(W* w, const srbd_qp_dims* d, const srbd_qp_host* qp, int which) {
  const int N = d->N, nx = d->nx, nu = d->nu, nm = w->nm;
  const size_t q = (size_t)which;
  const double* x0 = qp->x0 + q * nx;
  for (int k = 0; k <= N; ++k) {
    const int nuk = w->nuk[k], nxk = w->nxk[k];
    real* H = HK(w, k);
    real* g = VN(w, g, k);
    memset(H, 0, sizeof(real) * nm * nm);
    if (k < N) {
      const double* R = qp->R + (q * N + k) * nu * nu;
      const double* r = qp->r + (q * N + k) * nu;
      const double* S = qp->S ? qp->S + (q * N + k) * nu * nx : NULL;
      for (int j = 0; j < nu; ++j)
        for (int i = 0; i < nu; ++i) H[i + nm * j] = R[i + nu * j];
      for (int i = 0; i < nu; ++i) g[i] = r[i];
      if (k == 0) { /* r0 = S0 x0 + r0 */
        if (S)
          for (int i = 0; i < nu; ++i) {
            real s = 0.0;
            for (int j = 0; j < nx; ++j) s += S[i + nu * j] * x0[j];
            g[i] = s + r[i];
          }
      } else if (S) { /* block (x,u) = S^T and its mirror */
        for (int j = 0; j < nx; ++j)
          for (int i = 0; i < nu; ++i) {
            H[(nuk + j) + nm * i] = S[i + nu * j];
            H[i + nm * (nuk + j)] = S[i + nu * j];
          }
      }
    }
    if (nxk > 0) {
      const double* Q = qp->Q + (q * (N + 1) + k) * nx * nx;
      const double* qq = qp->q + (q * (N + 1) + k) * nx;
      for (int j = 0; j < nx; ++j)
        for (int i = 0; i < nx; ++i) H[(nuk + i) + nm * (nuk + j)] = Q[i + nx * j];
      for (int i = 0; i < nx; ++i) g[nuk + i] = qq[i];
    }
    if (k < N) { /* G = [B^T; A^T], x_{k+1} = G^T z + b */
      const double* A = qp->A + (q * N + k) * nx * nx;
      const double* B = qp->Bm + (q * N + k) * nx * nu;
      const double* b = qp->b + (q * N + k) * nx;
      real* G = GK(w, k);
      real* bb = VX(w, bb, k);
      for (int j = 0; j < nx; ++j) {
        for (int i = 0; i < nu; ++i) G[i + nm * j] = B[j + nx * i];
        if (k > 0)
          for (int i = 0; i < nx; ++i) G[(nu + i) + nm * j] = A[j + nx * i];
      }
      for (int j = 0; j < nx; ++j) bb[j] = b[j];
      if (k == 0) /* b0 = A0 x0 + b0 */
        for (int i = 0; i < nx; ++i) {
          real s = 0.0;
          for (int j = 0; j < nx; ++j) s += A[i + nx * j] * x0[j];
          bb[i] = s + b[i];
        }
    }
    /* constraints: [box-u, box-x, general] */
    int* idxb = w->idxb + (size_t)k * w->ncm;
    real* lo = VC(w, lo, k); real* up = VC(w, up, k);
    real* ml = VC(w, ml, k); real* mu = VC(w, mu, k);
    int c = 0;
    if (k < N)
      for (int j = 0; j < d->nbu; ++j, ++c) {
        size_t o = (q * N + k) * d->nbu + j;
        idxb[c] = qp->idxbu[j];
        lo[c] = qp->lbu[o]; up[c] = qp->ubu[o];
        ml[c] = qp->lbu_mask ? qp->lbu_mask[o] : 1.0;
        mu[c] = qp->ubu_mask ? qp->ubu_mask[o] : 1.0;
      }
    if (k > 0)
      for (int j = 0; j < d->nbx; ++j, ++c) {
        size_t o = (q * (N + 1) + k) * d->nbx + j;
        idxb[c] = nuk + qp->idxbx[j];
        lo[c] = qp->lbx[o]; up[c] = qp->ubx[o];
        ml[c] = qp->lbx_mask ? qp->lbx_mask[o] : 1.0;
        mu[c] = qp->ubx_mask ? qp->ubx_mask[o] : 1.0;
      }
    const int ngk = w->ngk[k];
    real* DCt = DK(w, k);
    memset(DCt, 0, sizeof(real) * nm * w->ngm);
    for (int j = 0; j < ngk; ++j, ++c) {
      if (k < N) {
        size_t o = (q * N + k) * d->ng + j;
        const double* D = qp->D + (q * N + k) * d->ng * nu;
        for (int i = 0; i < nu; ++i) DCt[i + nm * j] = D[j + d->ng * i];
        if (k > 0 && qp->C) { /* C0 is dropped: stage 0 has no state (nx[0] := 0) */
          const double* C = qp->C + (q * N + k) * d->ng * nx;
          for (int i = 0; i < nx; ++i) DCt[(nuk + i) + nm * j] = C[j + d->ng * i];
        }
        lo[c] = qp->lg[o]; up[c] = qp->ug[o];
        ml[c] = qp->lg_mask ? qp->lg_mask[o] : 1.0;
        mu[c] = qp->ug_mask ? qp->ug_mask[o] : 1.0;
      } else {
        size_t o = q * d->ngN + j;
        const double* C = qp->CN + q * d->ngN * nx;
        for (int i = 0; i < nx; ++i) DCt[(nuk + i) + nm * j] = C[j + d->ngN * i];
        lo[c] = qp->lgN[o]; up[c] = qp->ugN[o];
        ml[c] = qp->lgN_mask ? qp->lgN_mask[o] : 1.0;
        mu[c] = qp->ugN_mask ? qp->ugN_mask[o] : 1.0;
      }
    }
    for (int j = 0; j < c; ++j) { /* masks are 0/1 doubles (ocp_qp.hpp:73-109) */
      ml[j] = ml[j] != 0.0 ? 1.0 : 0.0;
      mu[j] = mu[j] != 0.0 ? 1.0 : 0.0;
    }
  }
  w->nc_mask = 0;
  for (int k = 0; k <= N; ++k) {
    int nc = w->nbk[k] + w->ngk[k];
    for (int j = 0; j < nc; ++j) w->nc_mask += (VC(w, ml, k)[j] != 0.0) + (VC(w, mu, k)[j] != 0.0);
  }
}

/* J z for one stage: v[0..nb) box, v[nb..nb+ng) general */
static void apply_J(const W* w, int k, const real* z, real* v) {
  const int n = w->nuk[k] + w->nxk[k], nb = w->nbk[k], ng = w->ngk[k], nm = w->nm;
  const int* idxb = w->idxb + (size_t)k * w->ncm;
  const real* DCt = DK(w, k);
  for (int j = 0; j < nb; ++j) v[j] = z[idxb[j]];
  for (int j = 0; j < ng; ++j) {
    real s = 0.0;
    for (int i = 0; i < n; ++i) s += DCt[i + nm * j] * z[i];
    v[nb + j] = s;
  }
}

/* out += J^T v */
static void apply_Jt_add(const W* w, int k, const real* v, real* out) {
  const int n = w->nuk[k] + w->nxk[k], nb = w->nbk[k], ng = w->ngk[k], nm = w->nm;
  const int* idxb = w->idxb + (size_t)k * w->ncm;
  const real* DCt = DK(w, k);
  for (int j = 0; j < nb; ++j) out[idxb[j]] += v[j];
  for (int j = 0; j < ng; ++j) {
    const real vj = v[nb + j];
    for (int i = 0; i < n; ++i) out[i] += DCt[i + nm * j] * vj;
  }
}

/* d_ocp_qp_init_var (hpipm_d_ocp_qp_ipm.h:232), SURVEY.md Appendix C "INIT" */
static void init_var(W* w, const srbd_ipm_args* a) {
  const real thr0 = a->thr0, mu0 = a->mu0;
  for (int k = 0; k <= w->N; ++k) {
    const int nb = w->nbk[k], ng = w->ngk[k];
    real* z = VN(w, z, k);
    const int* idxb = w->idxb + (size_t)k * w->ncm;
    real *lo = VC(w, lo, k), *up = VC(w, up, k), *tl = VC(w, tl, k), *tu = VC(w, tu, k);
    real *ll = VC(w, ll, k), *lu = VC(w, lu, k), *ml = VC(w, ml, k), *mu = VC(w, mu, k);
    if (k < w->N) memset(VX(w, pi, k), 0, sizeof(real) * w->nx);
    for (int j = 0; j < nb; ++j) {
      const int i = idxb[j];
      tl[j] = -lo[j] + z[i];
      tu[j] = up[j] - z[i];
      if (tl[j] < thr0) {
        if (tu[j] < thr0) {
          z[i] = 0.5 * (lo[j] + up[j]);
          tl[j] = thr0; tu[j] = thr0;
        } else {
          tl[j] = thr0;
          z[i] = lo[j] + thr0;
        }
      } else if (tu[j] < thr0) {
        tu[j] = thr0;
        z[i] = up[j] - thr0;
      }
    }
    real* v = w->tmp;
    apply_J(w, k, z, v);
    for (int j = nb; j < nb + ng; ++j) {
      tl[j] = v[j] - lo[j];
      tu[j] = up[j] - v[j];
      tl[j] = thr0 > tl[j] ? thr0 : tl[j];
      tu[j] = thr0 > tu[j] ? thr0 : tu[j];
    }
    for (int j = 0; j < nb + ng; ++j) { /* lam = mu0/t, masked rows get lam = 0 */
      ll[j] = (mu0 / tl[j]) * ml[j];
      lu[j] = (mu0 / tu[j]) * mu[j];
    }
  }
}

/* d_ocp_qp_res_compute + _compute_inf_norm (hpipm_d_ocp_qp_res.h:90,94), Appendix C "RESIDUALS" */
static void compute_res(W* w) {
  const int nm = w->nm;
  real n_g = 0.0, n_b = 0.0, n_d = 0.0, n_m = 0.0, summ = 0.0, obj = 0.0;
  for (int k = 0; k <= w->N; ++k) {
    const int nuk = w->nuk[k], nxk = w->nxk[k], n = nuk + nxk, nb = w->nbk[k], ng = w->ngk[k], nc = nb + ng;
    const real* H = HK(w, k);
    const real* g = VN(w, g, k);
    const real* z = VN(w, z, k);
    real* rg = VN(w, rg, k);
    real quad = 0.0, lin = 0.0;
    for (int i = 0; i < n; ++i) { /* symv with the lower triangle */
      real s = 0.0;
      for (int j = 0; j < n; ++j) s += (i >= j ? H[i + nm * j] : H[j + nm * i]) * z[j];
      quad += z[i] * s;
      lin += g[i] * z[i];
      rg[i] = s + g[i];
    }
    obj += 0.5 * quad + lin;
    if (k < w->N) {
      const real* G = GK(w, k);
      const real* pi = VX(w, pi, k);
      const real* zn = VN(w, z, k + 1);
      real* rb = VX(w, rb, k);
      for (int i = 0; i < n; ++i) {
        real s = 0.0;
        for (int j = 0; j < w->nx; ++j) s += G[i + nm * j] * pi[j];
        rg[i] += s;
      }
      for (int j = 0; j < w->nx; ++j) {
        real s = 0.0;
        for (int i = 0; i < n; ++i) s += G[i + nm * j] * z[i];
        rb[j] = (s + VX(w, bb, k)[j]) - zn[w->nuk[k + 1] + j];
        if (R_FABS(rb[j]) > n_b) n_b = R_FABS(rb[j]);
      }
    }
    if (k > 0) {
      const real* pim = VX(w, pi, k - 1);
      for (int i = 0; i < nxk; ++i) rg[nuk + i] -= pim[i];
    }
    real *ll = VC(w, ll, k), *lu = VC(w, lu, k), *tl = VC(w, tl, k), *tu = VC(w, tu, k);
    real *ml = VC(w, ml, k), *mu = VC(w, mu, k), *lo = VC(w, lo, k), *up = VC(w, up, k);
    real *rdl = VC(w, rdl, k), *rdu = VC(w, rdu, k), *rml = VC(w, rml, k), *rmu = VC(w, rmu, k);
    real* v = w->tmp;
    for (int j = 0; j < nc; ++j) v[j] = lu[j] - ll[j];
    apply_Jt_add(w, k, v, rg);
    apply_J(w, k, z, v);
    for (int j = 0; j < nc; ++j) {
      rdl[j] = ((lo[j] - v[j]) + tl[j]) * ml[j];
      rdu[j] = ((-up[j] + v[j]) + tu[j]) * mu[j];
      rml[j] = (ll[j] * tl[j]) * ml[j];
      rmu[j] = (lu[j] * tu[j]) * mu[j];
      summ += rml[j];
      summ += rmu[j];
      if (R_FABS(rdl[j]) > n_d) n_d = R_FABS(rdl[j]);
      if (R_FABS(rdu[j]) > n_d) n_d = R_FABS(rdu[j]);
      if (R_FABS(rml[j]) > n_m) n_m = R_FABS(rml[j]);
      if (R_FABS(rmu[j]) > n_m) n_m = R_FABS(rmu[j]);
    }
    for (int i = 0; i < n; ++i)
      if (R_FABS(rg[i]) > n_g) n_g = R_FABS(rg[i]);
  }
  w->res_max[0] = n_g; w->res_max[1] = n_b; w->res_max[2] = n_d; w->res_max[3] = n_m;
  w->mu_res = w->nc_mask > 0 ? summ / (real)w->nc_mask : 0.0;
  w->obj = obj;
}

/* Division by a Cholesky pivot in the triangular solves.  BLASFEO's potrf stores the INVERSE diagonal next to the
 * factor (0 for a non-positive pivot) and its trsv / trsm kernels multiply by it (blasfeo_common.h: dA / use_dA), so a
 * failed pivot zeroes the component instead of producing inf / nan; for a positive pivot this is the plain division
 * (bit-identical to what the golden vectors pin).  At tol 1e-8 this matters for ~1 QP in 1e4..1e5: with mu ~ 1e-11 a
 * pivot of the barrier-augmented Hessian can cancel to <= 0 in the last iteration (profiles/r1_v14_parity_sweep_*). */
static inline real pdiv(real s, real d) { return d > 0.0 ? s / d : 0.0; }

/* dense lower Cholesky of the leading nc columns of an (m x m, ld) symmetric matrix held in its lower
 * triangle, applied right-looking to all m rows (BLASFEO potrf_l_mn semantics).  A non-positive pivot is replaced
 * by 0 and its column is scaled by 0, like BLASFEO does. */
static void potrf_l_mn(int m, int nc, real* A, int ld) {
  for (int j = 0; j < nc; ++j) {
    real dj = A[j + ld * j];
    for (int k = 0; k < j; ++k) dj -= A[j + ld * k] * A[j + ld * k];
    real inv;
    if (dj > 0.0) {
      dj = R_SQRT(dj);
      inv = 1.0 / dj;
    } else {
      dj = 0.0;
      inv = 0.0;
    }
    A[j + ld * j] = dj;
    for (int i = j + 1; i < m; ++i) {
      real s = A[i + ld * j];
      for (int k = 0; k < j; ++k) s -= A[i + ld * k] * A[j + ld * k];
      A[i + ld * j] = s * inv;
    }
  }
}

/*
 * Riccati KKT solve (SURVEY.md Appendix C "KKT_SOLVE").  rhs_g/rhs_b: gradient and dynamics offset of
 * the (step) QP; Gamma/gamma from w->Gl.. when with_constr.  fact=1 factorizes (and stores Lr, Ls, P);
 * fact=0 reuses the factors (vector part only).  Results: oz (N+1 stages of [u;x]), opi.
 */
static void kkt_solve(W* w, const srbd_ipm_args* a, int fact, int with_constr, const real* rhs_g,
                      const real* rhs_b, real* oz, real* opi) {
  const int nm = w->nm, nx = w->nx, nu = w->nu, N = w->N;
  real* M = w->M;
  real* AL = w->AL;
  real* gt = w->gt;
  for (int k = N; k >= 0; --k) {
    const int nuk = w->nuk[k], nxk = w->nxk[k], n = nuk + nxk, nb = w->nbk[k], ng = w->ngk[k], nc = nb + ng;
    const int* idxb = w->idxb + (size_t)k * w->ncm;
    const real* DCt = DK(w, k);
    const real* G = GK(w, k);
    if (fact) {
      const real* H = HK(w, k);
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) M[i + nm * j] = (i >= j) ? H[i + nm * j] : H[j + nm * i];
      if (with_constr) {
        const real *Gl = VC(w, Gl, k), *Gu = VC(w, Gu, k);
        for (int j = 0; j < nb; ++j) M[idxb[j] + nm * idxb[j]] += Gl[j] + Gu[j];
        for (int c = 0; c < n; ++c)
          for (int i = c; i < n; ++i) {
            real s = 0.0;
            for (int j = 0; j < ng; ++j) s += (DCt[i + nm * j] * (Gl[nb + j] + Gu[nb + j])) * DCt[c + nm * j];
            M[i + nm * c] += s;
            if (i != c) M[c + nm * i] = M[i + nm * c];
          }
      }
      if (k < N) {
        if (a->ric_alg == 0) { /* classical: AL = G P_{k+1}; M += AL G^T */
          const real* Pn = PK(w, k + 1);
          for (int j = 0; j < nx; ++j)
            for (int i = 0; i < n; ++i) {
              real s = 0.0;
              for (int l = 0; l < nx; ++l) s += G[i + nm * l] * Pn[l + nx * j];
              AL[i + nm * j] = s;
            }
          for (int c = 0; c < n; ++c)
            for (int i = c; i < n; ++i) {
              real s = 0.0;
              for (int l = 0; l < nx; ++l) s += AL[i + nm * l] * G[c + nm * l];
              M[i + nm * c] += s;
              if (i != c) M[c + nm * i] = M[i + nm * c];
            }
        } else { /* square root: AL = G Lxx_{k+1}; M += AL AL^T */
          const real* Ln = LFK(w, k + 1);
          const int off = w->nuk[k + 1];
          for (int j = 0; j < nx; ++j)
            for (int i = 0; i < n; ++i) {
              real s = 0.0;
              for (int l = j; l < nx; ++l) s += G[i + nm * l] * Ln[(off + l) + nm * (off + j)];
              AL[i + nm * j] = s;
            }
          for (int c = 0; c < n; ++c)
            for (int i = c; i < n; ++i) {
              real s = 0.0;
              for (int l = 0; l < nx; ++l) s += AL[i + nm * l] * AL[c + nm * l];
              M[i + nm * c] += s;
              if (i != c) M[c + nm * i] = M[i + nm * c];
            }
        }
      }
      for (int i = 0; i < n; ++i) M[i + nm * i] += a->reg_prim;
      if (a->ric_alg == 0) {
        potrf_l_mn(n, nuk, M, nm); /* Lr (nu x nu), Ls (nx x nu) */
        real* Lr = LRK(w, k);
        real* Ls = LSK(w, k);
        real* P = PK(w, k);
        for (int j = 0; j < nuk; ++j) {
          for (int i = 0; i < nuk; ++i) Lr[i + nu * j] = i >= j ? M[i + nm * j] : 0.0;
          for (int i = 0; i < nxk; ++i) Ls[i + nx * j] = M[(nuk + i) + nm * j];
        }
        for (int c = 0; c < nxk; ++c) /* P = M_xx - Ls Ls^T, symmetrized from the lower triangle */
          for (int i = c; i < nxk; ++i) {
            real s = M[(nuk + i) + nm * (nuk + c)];
            for (int l = 0; l < nuk; ++l) s -= Ls[i + nx * l] * Ls[c + nx * l];
            P[i + nx * c] = s;
            P[c + nx * i] = s;
          }
      } else {
        potrf_l_mn(n, n, M, nm);
        real* Lf = LFK(w, k);
        real* Lr = LRK(w, k);
        real* Ls = LSK(w, k);
        real* P = PK(w, k);
        for (int j = 0; j < n; ++j)
          for (int i = 0; i < n; ++i) Lf[i + nm * j] = i >= j ? M[i + nm * j] : 0.0;
        for (int j = 0; j < nuk; ++j) {
          for (int i = 0; i < nuk; ++i) Lr[i + nu * j] = Lf[i + nm * j];
          for (int i = 0; i < nxk; ++i) Ls[i + nx * j] = Lf[(nuk + i) + nm * j];
        }
        for (int c = 0; c < nxk; ++c) /* P = Lxx Lxx^T */
          for (int i = c; i < nxk; ++i) {
            real s = 0.0;
            for (int l = 0; l <= c; ++l) s += Lf[(nuk + i) + nm * (nuk + l)] * Lf[(nuk + c) + nm * (nuk + l)];
            P[i + nx * c] = s;
            P[c + nx * i] = s;
          }
      }
    }
    /* gradient */
    for (int i = 0; i < n; ++i) gt[i] = rhs_g[(size_t)k * nm + i];
    if (with_constr) {
      real* v = w->tmp;
      const real *gl = VC(w, gl, k), *gu = VC(w, gu, k);
      for (int j = 0; j < nc; ++j) v[j] = gl[j] - gu[j];
      apply_Jt_add(w, k, v, gt);
    }
    if (k < N) {
      const real* Pn = PK(w, k + 1);
      const real* pn = VX(w, p, k + 1);
      const real* rb = rhs_b + (size_t)k * nx;
      real* t = w->tmp;
      for (int i = 0; i < nx; ++i) {
        real s = 0.0;
        for (int j = 0; j < nx; ++j) s += Pn[i + nx * j] * rb[j];
        t[i] = s + pn[i];
      }
      for (int i = 0; i < n; ++i) {
        real s = 0.0;
        for (int j = 0; j < nx; ++j) s += G[i + nm * j] * t[j];
        gt[i] += s;
      }
    }
    { /* lv = Lr^-1 g_u ; p = g_x - Ls lv */
      const real* Lr = LRK(w, k);
      const real* Ls = LSK(w, k);
      real* lv = VN(w, lv, k);
      real* p = VX(w, p, k);
      for (int i = 0; i < nuk; ++i) {
        real s = gt[i];
        for (int j = 0; j < i; ++j) s -= Lr[i + nu * j] * lv[j];
        lv[i] = pdiv(s, Lr[i + nu * i]);
      }
      for (int i = 0; i < nxk; ++i) {
        real s = gt[nuk + i];
        for (int j = 0; j < nuk; ++j) s -= Ls[i + nx * j] * lv[j];
        p[i] = s;
      }
    }
  }
  /* forward rollout */
  for (int k = 0; k <= N; ++k) {
    const int nuk = w->nuk[k], nxk = w->nxk[k], n = nuk + nxk;
    const real* Lr = LRK(w, k);
    const real* Ls = LSK(w, k);
    const real* lv = VN(w, lv, k);
    real* z = oz + (size_t)k * nm;
    real* t = w->tmp;
    for (int i = 0; i < nuk; ++i) { /* t = Ls^T x + lv */
      real s = 0.0;
      for (int j = 0; j < nxk; ++j) s += Ls[j + nx * i] * z[nuk + j];
      t[i] = s + lv[i];
    }
    for (int i = nuk - 1; i >= 0; --i) { /* u = -Lr^-T t */
      real s = t[i];
      for (int j = i + 1; j < nuk; ++j) s -= Lr[j + nu * i] * t[j];
      t[i] = pdiv(s, Lr[i + nu * i]);
    }
    for (int i = 0; i < nuk; ++i) z[i] = -t[i];
    if (k < N) {
      const real* G = GK(w, k);
      const real* rb = rhs_b + (size_t)k * nx;
      real* zn = oz + (size_t)(k + 1) * nm;
      const int off = w->nuk[k + 1];
      for (int j = 0; j < nx; ++j) {
        real s = 0.0;
        for (int i = 0; i < n; ++i) s += G[i + nm * j] * z[i];
        zn[off + j] = s + rb[j];
      }
      const real* Pn = PK(w, k + 1);
      const real* pn = VX(w, p, k + 1);
      real* pi = opi + (size_t)k * nx;
      for (int i = 0; i < nx; ++i) {
        real s = 0.0;
        for (int j = 0; j < nx; ++j) s += Pn[i + nx * j] * zn[off + j];
        pi[i] = s + pn[i];
      }
    }
  }
}

/* Gamma = lam/t ; gamma = (res_m - lam*res_d)/t   (d_compute_Gamma_gamma_qp / d_compute_gamma_qp) */
static void compute_Gamma_gamma(W* w, int with_Gamma) {
  for (int k = 0; k <= w->N; ++k) {
    const int nc = w->nbk[k] + w->ngk[k];
    real *ll = VC(w, ll, k), *lu = VC(w, lu, k), *tl = VC(w, tl, k), *tu = VC(w, tu, k);
    real *ml = VC(w, ml, k), *mu = VC(w, mu, k);
    for (int j = 0; j < nc; ++j) {
      const real til = 1.0 / tl[j], tiu = 1.0 / tu[j];
      if (with_Gamma) {
        VC(w, Gl, k)[j] = (til * ll[j]) * ml[j];
        VC(w, Gu, k)[j] = (tiu * lu[j]) * mu[j];
      }
      VC(w, gl, k)[j] = (til * (VC(w, rml, k)[j] - ll[j] * VC(w, rdl, k)[j])) * ml[j];
      VC(w, gu, k)[j] = (tiu * (VC(w, rmu, k)[j] - lu[j] * VC(w, rdu, k)[j])) * mu[j];
    }
  }
}

/* Residual of the LINEAR KKT system at the step (dz, dpi, dlam, dt) -- d_ocp_qp_res_compute_lin (hpipm_d_ocp_qp_res.h)
 * with qp_step's right-hand sides (rg, rb, rd, rm of the solve being refined):
 *   lin_g = H dz + rg + G dpi - [0; dpi_{k-1}] + J^T (dlam_u - dlam_l)      lin_b = G^T dz + rb - dx_{k+1}
 *   lin_d_l = (-J dz + dt_l) + rd_l,  lin_d_u = (J dz + dt_u) + rd_u           lin_m = lam dt + t dlam + rm
 * and its inf-norms (stat, eq, ineq, comp) in w->lin_max. */
static void compute_res_lin(W* w) {
  const int nm = w->nm;
  real n_g = 0.0, n_b = 0.0, n_d = 0.0, n_m = 0.0;
  for (int k = 0; k <= w->N; ++k) {
    const int nuk = w->nuk[k], nxk = w->nxk[k], n = nuk + nxk, nb = w->nbk[k], ng = w->ngk[k], nc = nb + ng;
    const real* H = HK(w, k);
    const real* dz = VN(w, dz, k);
    real* lg = VN(w, lg_, k);
    for (int i = 0; i < n; ++i) {
      real s = 0.0;
      for (int j = 0; j < n; ++j) s += (i >= j ? H[i + nm * j] : H[j + nm * i]) * dz[j];
      lg[i] = s + VN(w, rg, k)[i];
    }
    if (k < w->N) {
      const real* G = GK(w, k);
      const real* dpi = VX(w, dpi, k);
      const real* dzn = VN(w, dz, k + 1);
      real* lb = VX(w, lb_, k);
      for (int i = 0; i < n; ++i) {
        real s = 0.0;
        for (int j = 0; j < w->nx; ++j) s += G[i + nm * j] * dpi[j];
        lg[i] += s;
      }
      for (int j = 0; j < w->nx; ++j) {
        real s = 0.0;
        for (int i = 0; i < n; ++i) s += G[i + nm * j] * dz[i];
        lb[j] = (s + VX(w, rb, k)[j]) - dzn[w->nuk[k + 1] + j];
        if (R_FABS(lb[j]) > n_b) n_b = R_FABS(lb[j]);
      }
    }
    if (k > 0) {
      const real* dpim = VX(w, dpi, k - 1);
      for (int i = 0; i < nxk; ++i) lg[nuk + i] -= dpim[i];
    }
    real* v = w->tmp;
    for (int j = 0; j < nc; ++j) v[j] = VC(w, dlu, k)[j] - VC(w, dll, k)[j];
    apply_Jt_add(w, k, v, lg);
    apply_J(w, k, dz, v);
    for (int j = 0; j < nc; ++j) {
      const real ml = VC(w, ml, k)[j], mu = VC(w, mu, k)[j];
      const real ldl = ((-v[j] + VC(w, dtl, k)[j]) + VC(w, rdl, k)[j]) * ml;
      const real ldu = ((v[j] + VC(w, dtu, k)[j]) + VC(w, rdu, k)[j]) * mu;
      const real lml = ((VC(w, ll, k)[j] * VC(w, dtl, k)[j] + VC(w, tl, k)[j] * VC(w, dll, k)[j]) + VC(w, rml, k)[j]) * ml;
      const real lmu = ((VC(w, lu, k)[j] * VC(w, dtu, k)[j] + VC(w, tu, k)[j] * VC(w, dlu, k)[j]) + VC(w, rmu, k)[j]) * mu;
      VC(w, ldl, k)[j] = ldl; VC(w, ldu, k)[j] = ldu; VC(w, lml, k)[j] = lml; VC(w, lmu, k)[j] = lmu;
      if (R_FABS(ldl) > n_d) n_d = R_FABS(ldl);
      if (R_FABS(ldu) > n_d) n_d = R_FABS(ldu);
      if (R_FABS(lml) > n_m) n_m = R_FABS(lml);
      if (R_FABS(lmu) > n_m) n_m = R_FABS(lmu);
    }
    for (int i = 0; i < n; ++i)
      if (R_FABS(lg[i]) > n_g) n_g = R_FABS(lg[i]);
  }
  w->lin_max[0] = n_g; w->lin_max[1] = n_b; w->lin_max[2] = n_d; w->lin_max[3] = n_m;
}

/* Iterative refinement of the current step (hpipm_d_ocp_qp_ipm.h:74-75 itref_pred_max / itref_corr_max): up to `max_it`
 * times: linear residual; stop when every component is below itref_abs * its exit tolerance or itref_rel * the current
 * nonlinear residual; else solve the KKT system for it with the stored factorization (d_ocp_qp_solve_kkt_step on
 * qp_itref) and add the correction to (dz, dpi, dlam, dt).  Returns the number of refinement solves. */
static void kkt_solve(W* w, const srbd_ipm_args* a, int fact, int with_constr, const real* rhs_g,
                      const real* rhs_b, real* oz, real* opi);
static int refine_step(W* w, const srbd_ipm_args* a, int max_it) {
  const int N = w->N;
  const double tol[4] = {a->tol_stat, a->tol_eq, a->tol_ineq, a->tol_comp};
  int done = 0;
  for (int it = 0; it < max_it; ++it) {
    compute_res_lin(w);
    int ok = 1;
    for (int i = 0; i < 4; ++i)
      ok = ok && (w->lin_max[i] < a->itref_abs * tol[i] || w->lin_max[i] < a->itref_rel * w->res_max[i]);
    if (ok) break;
    /* gamma of the correction solve from (lin_m, lin_d); Gamma unchanged */
    for (int k = 0; k <= N; ++k) {
      const int nc = w->nbk[k] + w->ngk[k];
      for (int j = 0; j < nc; ++j) {
        const real til = 1.0 / VC(w, tl, k)[j], tiu = 1.0 / VC(w, tu, k)[j];
        VC(w, gl, k)[j] = (til * (VC(w, lml, k)[j] - VC(w, ll, k)[j] * VC(w, ldl, k)[j])) * VC(w, ml, k)[j];
        VC(w, gu, k)[j] = (tiu * (VC(w, lmu, k)[j] - VC(w, lu, k)[j] * VC(w, ldu, k)[j])) * VC(w, mu, k)[j];
      }
    }
    kkt_solve(w, a, 0, 1, w->lg_, w->lb_, w->cz, w->cpi);
    for (int k = 0; k <= N; ++k) {
      const int n = w->nuk[k] + w->nxk[k], nc = w->nbk[k] + w->ngk[k];
      real* v = w->tmp;
      apply_J(w, k, VN(w, cz, k), v);
      for (int j = 0; j < nc; ++j) {
        const real ml = VC(w, ml, k)[j], mu = VC(w, mu, k)[j];
        const real ctl = (v[j] - VC(w, ldl, k)[j]) * ml;
        const real ctu = (-v[j] - VC(w, ldu, k)[j]) * mu;
        const real cll = (-(VC(w, ll, k)[j] * ctl + VC(w, lml, k)[j]) / VC(w, tl, k)[j]) * ml;
        const real clu = (-(VC(w, lu, k)[j] * ctu + VC(w, lmu, k)[j]) / VC(w, tu, k)[j]) * mu;
        VC(w, dtl, k)[j] += ctl; VC(w, dtu, k)[j] += ctu; VC(w, dll, k)[j] += cll; VC(w, dlu, k)[j] += clu;
      }
      for (int i = 0; i < n; ++i) VN(w, dz, k)[i] += VN(w, cz, k)[i];
      if (k < N)
        for (int i = 0; i < w->nx; ++i) VX(w, dpi, k)[i] += VX(w, cpi, k)[i];
    }
    ++done;
  }
  return done;
}

/* dt = +-J dz - res_d ; dlam = -(lam*dt + res_m)/t   (d_compute_lam_t_qp) */
static void compute_dlam_dt(W* w) {
  for (int k = 0; k <= w->N; ++k) {
    const int nc = w->nbk[k] + w->ngk[k];
    real* v = w->tmp;
    apply_J(w, k, VN(w, dz, k), v);
    for (int j = 0; j < nc; ++j) {
      const real ml = VC(w, ml, k)[j], mu = VC(w, mu, k)[j];
      const real dtl = (v[j] - VC(w, rdl, k)[j]) * ml;
      const real dtu = (-v[j] - VC(w, rdu, k)[j]) * mu;
      VC(w, dtl, k)[j] = dtl;
      VC(w, dtu, k)[j] = dtu;
      VC(w, dll, k)[j] = (-(VC(w, ll, k)[j] * dtl + VC(w, rml, k)[j]) / VC(w, tl, k)[j]) * ml;
      VC(w, dlu, k)[j] = (-(VC(w, lu, k)[j] * dtu + VC(w, rmu, k)[j]) / VC(w, tu, k)[j]) * mu;
    }
  }
}

/* d_compute_alpha_qp: exact min over the ratios (order independent) */
static void step_length(const W* w, real* ap, real* ad) {
  real alpha_p = 1.0, alpha_d = 1.0;
  for (int k = 0; k <= w->N; ++k) {
    const int nc = w->nbk[k] + w->ngk[k];
    for (int j = 0; j < nc; ++j) {
      const real dtl = VC(w, dtl, k)[j], dtu = VC(w, dtu, k)[j], dll = VC(w, dll, k)[j], dlu = VC(w, dlu, k)[j];
      if (dtl < 0.0) { real r = -VC(w, tl, k)[j] / dtl; if (r < alpha_p) alpha_p = r; }
      if (dtu < 0.0) { real r = -VC(w, tu, k)[j] / dtu; if (r < alpha_p) alpha_p = r; }
      if (dll < 0.0) { real r = -VC(w, ll, k)[j] / dll; if (r < alpha_d) alpha_d = r; }
      if (dlu < 0.0) { real r = -VC(w, lu, k)[j] / dlu; if (r < alpha_d) alpha_d = r; }
    }
  }
  *ap = alpha_p; *ad = alpha_d;
}

/* d_compute_mu_aff_qp: sum (lam + alpha dlam)(t + alpha dt) / nc_mask, stage by stage, lower then upper */
static real mu_aff(const W* w, real alpha) {
  real s = 0.0;
  for (int k = 0; k <= w->N; ++k) {
    const int nc = w->nbk[k] + w->ngk[k];
    real sk = 0.0;
    for (int j = 0; j < nc; ++j) {
      sk += (VC(w, ll, k)[j] + alpha * VC(w, dll, k)[j]) * (VC(w, tl, k)[j] + alpha * VC(w, dtl, k)[j]);
      sk += (VC(w, lu, k)[j] + alpha * VC(w, dlu, k)[j]) * (VC(w, tu, k)[j] + alpha * VC(w, dtu, k)[j]);
    }
    s += sk;
  }
  return s / (real)w->nc_mask;
}

static real shorten(const srbd_ipm_args* a, real alpha) {
  if (alpha < 1.0) {
    if (a->alpha_shorten == 0) return alpha * 0.995;
    return alpha * ((1.0 - alpha) * 0.99 + alpha * 0.9999999);
  }
  return alpha;
}

static int any_nan(const W* w) {
  for (int k = 0; k <= w->N; ++k) {
    const int n = w->nuk[k] + w->nxk[k];
    for (int i = 0; i < n; ++i)
      if (R_ISNAN(VN(w, z, k)[i])) return 1;
  }
  return R_ISNAN(w->mu_res) || R_ISNAN(w->res_max[0]) || R_ISNAN(w->res_max[1]) || R_ISNAN(w->res_max[2]) ||
         R_ISNAN(w->res_max[3]);
}

/* d_ocp_qp_ipm_solve (hpipm_d_ocp_qp_ipm.h:238), SURVEY.md Appendix C "MAIN".  Returns status. */
static int ipm_solve(W* w, const srbd_ipm_args* a, int* iter_out, double* stat, int stat_rows) {
  const int N = w->N;
  const size_t S = (size_t)(N + 1);
  if (stat) memset(stat, 0, sizeof(double) * SRBD_STAT_M * stat_rows);
  if (w->nc_mask == 0) { /* unconstrained: one Riccati factor + solve on the QP itself, iter = 0 */
    kkt_solve(w, a, 1, 0, w->g, w->bb, w->z, w->pi);
    for (int k = 0; k <= N; ++k) { /* lam = 0 on every (masked) row, keep t harmless */
      const int nc = w->nbk[k] + w->ngk[k];
      for (int j = 0; j < nc; ++j) {
        VC(w, ll, k)[j] = 0.0; VC(w, lu, k)[j] = 0.0; VC(w, tl, k)[j] = 0.0; VC(w, tu, k)[j] = 0.0;
      }
    }
    compute_res(w);
    if (stat && stat_rows > 0) {
      for (int i = 0; i < 4; ++i) stat[6 + i] = w->res_max[i];
      stat[10] = w->obj;
    }
    *iter_out = 0;
    return any_nan(w) ? 3 : 0;
  }
  init_var(w, a);
  compute_res(w);
  real mu = w->mu_res;
  if (stat && stat_rows > 0) {
    stat[5] = mu;
    for (int i = 0; i < 4; ++i) stat[6 + i] = w->res_max[i];
    stat[10] = w->obj;
  }
  real alpha = 1.0;
  int kk = 0;
  for (; kk < a->iter_max && alpha > a->alpha_min &&
         (w->res_max[0] > a->tol_stat || w->res_max[1] > a->tol_eq || w->res_max[2] > a->tol_ineq ||
          w->res_max[3] > a->tol_comp);
       ++kk) {
    double* row = (stat && kk + 1 < stat_rows) ? stat + SRBD_STAT_M * (kk + 1) : NULL;
    memcpy(w->rml_bkp, w->rml, sizeof(real) * S * w->ncm);
    memcpy(w->rmu_bkp, w->rmu, sizeof(real) * S * w->ncm);
    /* affine (predictor) step */
    compute_Gamma_gamma(w, 1);
    kkt_solve(w, a, 1, 1, w->rg, w->rb, w->dz, w->dpi);
    compute_dlam_dt(w);
    if (a->itref_pred_max > 0) {
      const int nref = refine_step(w, a, a->itref_pred_max);
      if (row) row[12] = nref;
    }
    real ap, ad;
    step_length(w, &ap, &ad);
    real alpha_aff = ap < ad ? ap : ad;
    if (row) row[0] = alpha_aff;
    if (a->pred_corr == 1) {
      const real mua = mu_aff(w, alpha_aff);
      const real tmp = mua / mu;
      const real sigma = tmp * tmp * tmp;
      if (row) { row[1] = mua; row[2] = sigma; }
      /* centering-correction: res_m = res_m_bkp + dt*dlam - max(sigma*mu, tau_min), masked */
      real sm = sigma * mu;
      sm = sm > a->tau_min ? sm : a->tau_min;
      for (int k = 0; k <= N; ++k) {
        const int nc = w->nbk[k] + w->ngk[k];
        for (int j = 0; j < nc; ++j) {
          VC(w, rml, k)[j] = ((VC(w, rml_bkp, k)[j] + VC(w, dtl, k)[j] * VC(w, dll, k)[j]) - sm) * VC(w, ml, k)[j];
          VC(w, rmu, k)[j] = ((VC(w, rmu_bkp, k)[j] + VC(w, dtu, k)[j] * VC(w, dlu, k)[j]) - sm) * VC(w, mu, k)[j];
        }
      }
      compute_Gamma_gamma(w, 0);
      kkt_solve(w, a, 0, 1, w->rg, w->rb, w->dz, w->dpi);
      compute_dlam_dt(w);
      if (a->itref_corr_max > 0) {
        const int nref = refine_step(w, a, a->itref_corr_max);
        if (row) {
          row[13] = nref;
          for (int i = 0; i < 4; ++i) row[14 + i] = (double)w->lin_max[i];
        }
      }
      step_length(w, &ap, &ad);
      if (a->cond_pred_corr == 1) {
        const real al = ap < ad ? ap : ad;
        const real muc = mu_aff(w, al);
        if (muc > a->cond_factor * mua) { /* centering direction only */
          const real sm2 = sigma * mu;
          for (int k = 0; k <= N; ++k) {
            const int nc = w->nbk[k] + w->ngk[k];
            for (int j = 0; j < nc; ++j) {
              VC(w, rml, k)[j] = (VC(w, rml_bkp, k)[j] - sm2) * VC(w, ml, k)[j];
              VC(w, rmu, k)[j] = (VC(w, rmu_bkp, k)[j] - sm2) * VC(w, mu, k)[j];
            }
          }
          compute_Gamma_gamma(w, 0);
          kkt_solve(w, a, 0, 1, w->rg, w->rb, w->dz, w->dpi);
          compute_dlam_dt(w);
          step_length(w, &ap, &ad);
        }
      }
    }
    if (!a->split_step) {
      const real al = ap < ad ? ap : ad;
      ap = al; ad = al;
    }
    alpha = ap < ad ? ap : ad;
    if (row) { row[3] = ap; row[4] = ad; }
    const real sp = shorten(a, ap), sd = shorten(a, ad);
    /* d_update_var_qp: kkt_solve returned the minimizer of the step QP (gradient res_g~, offset res_b),
     * which IS the Newton direction; (z,t) move with alpha_prim, (pi,lam) with alpha_dual. */
    for (int k = 0; k <= N; ++k) {
      const int n = w->nuk[k] + w->nxk[k], nc = w->nbk[k] + w->ngk[k];
      for (int i = 0; i < n; ++i) VN(w, z, k)[i] += sp * VN(w, dz, k)[i];
      if (k < N)
        for (int i = 0; i < w->nx; ++i) VX(w, pi, k)[i] += sd * VX(w, dpi, k)[i];
      for (int j = 0; j < nc; ++j) {
        const real ml = VC(w, ml, k)[j], mu_ = VC(w, mu, k)[j];
        VC(w, tl, k)[j] += sp * VC(w, dtl, k)[j];
        VC(w, tu, k)[j] += sp * VC(w, dtu, k)[j];
        VC(w, ll, k)[j] += sd * VC(w, dll, k)[j];
        VC(w, lu, k)[j] += sd * VC(w, dlu, k)[j];
        if (a->t_lam_min == 2) { /* clip (active rows only; masked rows keep lam = 0 exactly) */
          if (ml != 0.0) {
            if (VC(w, tl, k)[j] < a->t_min) VC(w, tl, k)[j] = a->t_min;
            if (VC(w, ll, k)[j] < a->lam_min) VC(w, ll, k)[j] = a->lam_min;
          }
          if (mu_ != 0.0) {
            if (VC(w, tu, k)[j] < a->t_min) VC(w, tu, k)[j] = a->t_min;
            if (VC(w, lu, k)[j] < a->lam_min) VC(w, lu, k)[j] = a->lam_min;
          }
        }
      }
    }
    compute_res(w);
    mu = w->mu_res;
    if (row) {
      row[5] = mu;
      for (int i = 0; i < 4; ++i) row[6 + i] = w->res_max[i];
      row[10] = w->obj;
    }
  }
  *iter_out = kk;
  /* status order of d_ocp_qp_ipm_solve: MAX_ITER, MIN_STEP, NAN_SOL, SUCCESS (hpipm_common.h:57-64) */
  if (kk == a->iter_max) return 1;
  if (alpha <= a->alpha_min) return 2;
  if (any_nan(w)) return 3;
  return 0;
}

/* getters + stage-0 reconstruction (ocp_qp_ipm_solver.cpp:337-373) */
static void write_outputs(W* w, const srbd_qp_dims* d, const srbd_ipm_args* a, const srbd_qp_host* qp,
                          const srbd_sol_host* sol, int which, int unconstrained) {
  const int N = d->N, nx = d->nx, nu = d->nu, nm = w->nm;
  const size_t q = (size_t)which;
  const double* x0 = qp->x0 + q * nx;
  (void)a;
  if (sol->x) {
    double* X = sol->x + q * (N + 1) * nx;
    memcpy(X, x0, sizeof(double) * nx);
    for (int k = 1; k <= N; ++k) cpy_out(X + (size_t)k * nx, VN(w, z, k) + w->nuk[k], nx);
  }
  if (sol->u)
    for (int k = 0; k < N; ++k) cpy_out(sol->u + (q * N + k) * nu, VN(w, z, k), nu);
  if (sol->lam || sol->t) {
    size_t nct = orc_qp_nct(d), o = 0;
    for (int k = 0; k <= N; ++k) {
      const int nc = w->nbk[k] + w->ngk[k];
      if (sol->lam) {
        cpy_out(sol->lam + q * nct + o, VC(w, ll, k), nc);
        cpy_out(sol->lam + q * nct + o + nc, VC(w, lu, k), nc);
      }
      if (sol->t) {
        cpy_out(sol->t + q * nct + o, VC(w, tl, k), nc);
        cpy_out(sol->t + q * nct + o + nc, VC(w, tu, k), nc);
      }
      o += 2 * (size_t)nc;
    }
  }
  /* K_k = -Lr^-T Ls^T for k>=1 */
  real* Kbuf = dalloc((size_t)nu * nx + 1);
  for (int k = 1; k < N; ++k) {
    const real* Lr = LRK(w, k);
    const real* Ls = LSK(w, k);
    for (int c = 0; c < nx; ++c) { /* column c of K: solve Lr^T y = Ls(c,:)^T */
      for (int i = nu - 1; i >= 0; --i) {
        real s = Ls[c + nx * i];
        for (int j = i + 1; j < nu; ++j) s -= Lr[j + nu * i] * Kbuf[j + nu * c];
        Kbuf[i + nu * c] = pdiv(s, Lr[i + nu * i]);
      }
      for (int i = 0; i < nu; ++i) Kbuf[i + nu * c] = -Kbuf[i + nu * c];
    }
    if (sol->K) cpy_out(sol->K + (q * N + k) * nu * nx, Kbuf, nu * nx);
    if (sol->k) {
      double* kk = sol->k + (q * N + k) * nu;
      const real* z = VN(w, z, k);
      if (unconstrained) { /* true Riccati feed-forward: -Lr^-T lv */
        const real* lv = VN(w, lv, k);
        real t[64];
        for (int i = nu - 1; i >= 0; --i) {
          real s = lv[i];
          for (int j = i + 1; j < nu; ++j) s -= Lr[j + nu * i] * t[j];
          t[i] = pdiv(s, Lr[i + nu * i]);
        }
        for (int i = 0; i < nu; ++i) kk[i] = -t[i];
      } else { /* absolute form by the identity u = K x + k */
        for (int i = 0; i < nu; ++i) {
          real s = 0.0;
          for (int j = 0; j < nx; ++j) s += Kbuf[i + nu * j] * z[nu + j];
          kk[i] = z[i] - s;
        }
      }
    }
  }
  for (int k = 1; k <= N; ++k) {
    if (sol->pi) cpy_out(sol->pi + (q * (N + 1) + k) * nx, VX(w, pi, k - 1), nx);
    if (sol->P) cpy_out(sol->P + (q * (N + 1) + k) * nx * nx, PK(w, k), nx * nx);
    if (sol->p) {
      double* pp = sol->p + (q * (N + 1) + k) * nx;
      if (unconstrained) {
        cpy_out(pp, VX(w, p, k), nx);
      } else { /* pi_k = P_k x_k + p_k */
        const real* P = PK(w, k);
        const real* xk = VN(w, z, k) + w->nuk[k];
        for (int i = 0; i < nx; ++i) {
          real s = 0.0;
          for (int j = 0; j < nx; ++j) s += P[i + nx * j] * xk[j];
          pp[i] = VX(w, pi, k - 1)[i] - s;
        }
      }
    }
  }
  /* stage 0 (ocp_qp_ipm_solver.cpp:349-373) */
  {
    const real* Lr0 = LRK(w, 0);
    const double* A0 = qp->A + (q * N) * nx * nx;
    const double* B0 = qp->Bm + (q * N) * nx * nu;
    const double* b0 = qp->b + (q * N) * nx; /* the ORIGINAL b (:369) */
    const double* S0 = qp->S ? qp->S + (q * N) * nu * nx : NULL;
    const double* Q0 = qp->Q + (q * (N + 1)) * nx * nx;
    const double* q0 = qp->q + (q * (N + 1)) * nx;
    const real* P1 = PK(w, 1);
    /* p1 in the convention exported above */
    real* p1 = dalloc(nx);
    if (unconstrained) memcpy(p1, VX(w, p, 1), sizeof(real) * nx);
    else {
      const real* x1 = VN(w, z, 1) + w->nuk[1];
      for (int i = 0; i < nx; ++i) {
        real s = 0.0;
        for (int j = 0; j < nx; ++j) s += P1[i + nx * j] * x1[j];
        p1[i] = VX(w, pi, 0)[i] - s;
      }
    }
    real* BtP = dalloc((size_t)nu * nx + 1);  /* B0^T P1 */
    real* AtP = dalloc((size_t)nx * nx + 1);  /* A0^T P1 */
    real* H0 = dalloc((size_t)nu * nx + 1);
    real* GH = dalloc((size_t)nu * nx + 1);   /* G0^-1 H0 */
    for (int j = 0; j < nx; ++j)
      for (int i = 0; i < nu; ++i) {
        real s = 0.0;
        for (int l = 0; l < nx; ++l) s += B0[l + nx * i] * P1[l + nx * j];
        BtP[i + nu * j] = s;
      }
    for (int j = 0; j < nx; ++j)
      for (int i = 0; i < nx; ++i) {
        real s = 0.0;
        for (int l = 0; l < nx; ++l) s += A0[l + nx * i] * P1[l + nx * j];
        AtP[i + nx * j] = s;
      }
    for (int j = 0; j < nx; ++j)
      for (int i = 0; i < nu; ++i) {
        real s = 0.0;
        for (int l = 0; l < nx; ++l) s += BtP[i + nu * l] * A0[l + nx * j];
        H0[i + nu * j] = (S0 ? S0[i + nu * j] : 0.0) + s;
      }
    for (int c = 0; c < nx; ++c) { /* GH(:,c) = (Lr0 Lr0^T)^-1 H0(:,c) */
      real y[64];
      for (int i = 0; i < nu; ++i) {
        real s = H0[i + nu * c];
        for (int j = 0; j < i; ++j) s -= Lr0[i + nu * j] * y[j];
        y[i] = pdiv(s, Lr0[i + nu * i]);
      }
      for (int i = nu - 1; i >= 0; --i) {
        real s = y[i];
        for (int j = i + 1; j < nu; ++j) s -= Lr0[j + nu * i] * y[j];
        y[i] = pdiv(s, Lr0[i + nu * i]);
      }
      for (int i = 0; i < nu; ++i) GH[i + nu * c] = y[i];
    }
    real* K0 = dalloc((size_t)nu * nx + 1);
    real* k0 = dalloc(nu + 1);
    for (int i = 0; i < nu * nx; ++i) K0[i] = -GH[i];
    const real* u0 = VN(w, z, 0);
    for (int i = 0; i < nu; ++i) {
      real s = 0.0;
      for (int j = 0; j < nx; ++j) s += K0[i + nu * j] * x0[j];
      k0[i] = u0[i] - s;
    }
    if (sol->K) cpy_out(sol->K + (q * N) * nu * nx, K0, nu * nx);
    if (sol->k) cpy_out(sol->k + (q * N) * nu, k0, nu);
    real* P0 = dalloc((size_t)nx * nx + 1);
    real* p0 = dalloc(nx + 1);
    for (int j = 0; j < nx; ++j)
      for (int i = 0; i < nx; ++i) {
        real s1 = 0.0, s2 = 0.0;
        for (int l = 0; l < nu; ++l) s1 += H0[l + nu * i] * GH[l + nu * j];
        for (int l = 0; l < nx; ++l) s2 += AtP[i + nx * l] * A0[l + nx * j];
        P0[i + nx * j] = (Q0[i + nx * j] - s1) + s2;
      }
    for (int i = 0; i < nx; ++i) {
      real s1 = 0.0, s2 = 0.0, s3 = 0.0;
      for (int l = 0; l < nx; ++l) s1 += A0[l + nx * i] * p1[l];
      for (int l = 0; l < nx; ++l) s2 += AtP[i + nx * l] * b0[l];
      for (int l = 0; l < nu; ++l) s3 += H0[l + nu * i] * k0[l];
      p0[i] = ((q0[i] + s1) + s2) + s3;
    }
    if (sol->P) cpy_out(sol->P + (q * (N + 1)) * nx * nx, P0, nx * nx);
    if (sol->p) cpy_out(sol->p + (q * (N + 1)) * nx, p0, nx);
    if (sol->pi) {
      double* pi0 = sol->pi + (q * (N + 1)) * nx;
      for (int i = 0; i < nx; ++i) {
        real s = 0.0;
        for (int j = 0; j < nx; ++j) s += P0[i + nx * j] * x0[j];
        pi0[i] = p0[i] + s;
      }
    }
    free(p1); free(BtP); free(AtP); free(H0); free(GH); free(K0); free(k0); free(P0); free(p0);
  }
  free(Kbuf);
  (void)nm;
}

size_t orc_qp_nct(const srbd_qp_dims* d) {
  size_t n = 0;
  for (int k = 0; k <= d->N; ++k) {
    int nb = (k < d->N ? d->nbu : 0) + (k > 0 ? d->nbx : 0);
    int ng = k < d->N ? d->ng : d->ngN;
    n += 2 * (size_t)(nb + ng);
  }
  return n;
}

int ORC_NAME(orc_qp_solve_one)(const srbd_qp_dims* d, const srbd_ipm_args* a, const srbd_qp_host* qp,
                     const srbd_sol_host* sol, const srbd_stats_host* st, int stat_rows, int which) {
  W* w = w_create(d);
  embed_qp(w, d, qp, which);
  const int N = d->N, nx = d->nx, nu = d->nu;
  const size_t q = (size_t)which;
  /* primal warm start (ocp_qp_ipm_solver.cpp:328-333): x[i+1], u[i] */
  if (a->warm_start && qp->x_init && qp->u_init) {
    for (int k = 0; k <= N; ++k) {
      real* z = VN(w, z, k);
      if (k < N) cpy_in(z, qp->u_init + (q * N + k) * nu, nu);
      if (k > 0) cpy_in(z + w->nuk[k], qp->x_init + (q * (N + 1) + k) * nx, nx);
    }
  }
  int iter = 0;
  double* stat = (st && st->stat) ? st->stat + q * (size_t)stat_rows * SRBD_STAT_M : NULL;
  const int unconstrained = (w->nc_mask == 0);
  int status = ipm_solve(w, a, &iter, stat, stat_rows);
  write_outputs(w, d, a, qp, sol, which, unconstrained);
  if (st) {
    if (st->iter) st->iter[q] = iter;
    if (st->status) st->status[q] = status;
    if (st->res_max) cpy_out(st->res_max + 4 * q, w->res_max, 4);
  }
  w_free(w);
  return status;
}

#ifndef ORC_QUAD
int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
#endif

int ORC_NAME(orc_qp_solve_batch)(const srbd_qp_dims* d, const srbd_ipm_args* a, const srbd_qp_host* qp,
                       const srbd_sol_host* sol, const srbd_stats_host* st, int stat_rows, int batch,
                       int threads) {
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads)
#endif
  for (int i = 0; i < batch; ++i) ORC_NAME(orc_qp_solve_one)(d, a, qp, sol, st, stat_rows, i);
  (void)threads;
  return 0;
}

#ifndef ORC_QUAD
void orc_ipm_args_default(srbd_ipm_args* a) {
  memset(a, 0, sizeof(*a));
  /* OcpQpIpmSolverSettings defaults (ocp_qp_ipm_solver_settings.hpp:26-86) */
  a->iter_max = 15; a->alpha_min = 1e-8; a->mu0 = 1e2;
  a->tol_stat = a->tol_eq = a->tol_ineq = a->tol_comp = 1e-8;
  a->reg_prim = 1e-12; a->warm_start = 0; a->pred_corr = 1; a->ric_alg = 1; a->split_step = 0;
  /* hidden HPIPM SPEED-mode defaults */
  a->cond_pred_corr = 1; a->cond_factor = 2.0; a->thr0 = 0.1;
  a->lam_min = a->t_min = a->tau_min = 1e-16; a->t_lam_min = 2; a->alpha_shorten = 1;
  a->itref_pred_max = 0; a->itref_corr_max = 0; a->itref_abs = 1.0; a->itref_rel = 1e-3;
}
#endif
