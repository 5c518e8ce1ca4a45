/*
 * srbd_model.c — CPU ORACLE (test infrastructure): SO(3) helpers, SRBD dynamics, shooting
 * linearization, constraint rows, relaxed barrier, SQP assembly and the filter line search.
 * Restates dynamics/orientation_tool.h, dynamics/SRBD_model.cpp and NMPC_solver.cpp:149-314 of the
 * reference; each function cites the lines it follows.  Operation order follows the reference's
 * Eigen expressions (left-associative products) so that rounding differences stay at the ulp level.
 * Parity of this part is UNPINNED by the reference (no tests/fixtures exist for it).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "srbd_oracle.h"

/* ---------------------------------------------------------------------------------------------- */
/* 3x3 helpers, column-major: M[i + 3*j]                                                            */
/* ---------------------------------------------------------------------------------------------- */
static void m3_mul(const double* A, const double* B, double* C) {
  double T[9];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += A[i + 3 * k] * B[k + 3 * j];
      T[i + 3 * j] = s;
    }
  memcpy(C, T, sizeof(T));
}
static void m3_mul_bt(const double* A, const double* B, double* C) { /* C = A * B^T */
  double T[9];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += A[i + 3 * k] * B[j + 3 * k];
      T[i + 3 * j] = s;
    }
  memcpy(C, T, sizeof(T));
}
static void m3_vec(const double* A, const double* v, double* o) {
  double t[3];
  for (int i = 0; i < 3; ++i) {
    double s = 0.0;
    for (int k = 0; k < 3; ++k) s += A[i + 3 * k] * v[k];
    t[i] = s;
  }
  o[0] = t[0]; o[1] = t[1]; o[2] = t[2];
}
static void m3_eye(double* I) {
  memset(I, 0, 9 * sizeof(double));
  I[0] = I[4] = I[8] = 1.0;
}
static double clamp_theta(const double r[3]) { /* orientation_tool.h:78-83 */
  double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  const double h = 1e-10;
  if (th < h) th = h;
  return th;
}

/* orientation_tool.h:55-63 */
void orc_skew(const double v[3], double M[9]) {
  M[0] = 0.0;   M[3] = -v[2]; M[6] = v[1];
  M[1] = v[2];  M[4] = 0.0;   M[7] = -v[0];
  M[2] = -v[1]; M[5] = v[0];  M[8] = 0.0;
}

/* orientation_tool.h:75-86:  R = I + (sin th/th) V + ((1-cos th)/th^2) V V */
void orc_expm(const double r[3], double R[9]) {
  double th = clamp_theta(r), V[9], VV[9];
  orc_skew(r, V);
  m3_mul(V, V, VV);
  double a = sin(th) / th, b = (1.0 - cos(th)) / (th * th);
  m3_eye(R);
  for (int i = 0; i < 9; ++i) R[i] = (R[i] + a * V[i]) + b * VV[i];
}

/* orientation_tool.h:128-140 */
void orc_jl(const double r[3], double J[9]) {
  double th = clamp_theta(r), V[9], VV[9], I[9];
  orc_skew(r, V);
  for (int i = 0; i < 9; ++i) V[i] = V[i] / th;
  m3_mul(V, V, VV);
  m3_eye(I);
  double s = sin(th) / th, c = (1.0 - cos(th)) / th;
  for (int i = 0; i < 9; ++i) J[i] = (s * I[i] + (1.0 - s) * (VV[i] + I[i])) + c * V[i];
}

/* orientation_tool.h:144-157 */
void orc_jlt(const double r[3], double J[9]) {
  double th = clamp_theta(r), V[9], VV[9], I[9];
  orc_skew(r, V);
  for (int i = 0; i < 9; ++i) V[i] = V[i] / th;
  m3_mul(V, V, VV);
  m3_eye(I);
  double cot = 1.0 / tan(0.5 * th);
  double a = 0.5 * cot * th;
  for (int i = 0; i < 9; ++i) J[i] = (a * I[i] + (1.0 - a) * (VV[i] + I[i])) - (0.5 * th) * V[i];
}

/* orientation_tool.h:164-204 */
void orc_djl(const double r[3], double dJ[27]) {
  double th = clamp_theta(r), S[9], V[9], VV[9];
  orc_skew(r, S);
  for (int i = 0; i < 9; ++i) V[i] = S[i] / th;
  m3_mul(V, V, VV);
  double sn = sin(th), cs = cos(th);
  double th2 = th * th, th3 = th2 * th;
  double c1 = (th * sn + (2.0 * (cs - 1.0))) / th3;
  double c2 = -(2.0 * th - 3.0 * sn + th * cs) / th3;
  double base[9];
  for (int i = 0; i < 9; ++i) base[i] = c1 * V[i] + c2 * VV[i];
  double ca = (th - sn) / th3, cb = (1.0 - cs) / th2;
  for (int k = 0; k < 3; ++k) {
    double e[3] = {0.0, 0.0, 0.0}, E[9], ES[9], SE[9];
    e[k] = 1.0;
    orc_skew(e, E);
    m3_mul(E, S, ES);
    m3_mul(S, E, SE);
    for (int i = 0; i < 9; ++i) dJ[9 * k + i] = (ca * (ES[i] + SE[i]) + cb * E[i]) + base[i] * r[k];
  }
}

/* orientation_tool.h:211-227:  d(Jl^-1)/dr_k = -Jl^-1 (dJl/dr_k) Jl^-1 */
void orc_djlt(const double r[3], double dJ[27]) {
  double J[9], nJ[9], d[27], T[9];
  orc_jlt(r, J);
  orc_djl(r, d);
  for (int i = 0; i < 9; ++i) nJ[i] = -J[i];
  for (int k = 0; k < 3; ++k) {
    m3_mul(nJ, d + 9 * k, T);
    m3_mul(T, J, dJ + 9 * k);
  }
}

/* ---------------------------------------------------------------------------------------------- */
/* SRBD_model.cpp:75-141                                                                            */
/* ---------------------------------------------------------------------------------------------- */
void orc_continuous(const srbd_model_params* m, const double* x, const double* u, double* dx,
                    double* jfx, double* jfu) {
  const double* r = x;
  const double* l = x + 3;
  const double* p = x + 6;
  const double* v = x + 9;
  double R[9], Jlt[9], RL[9], RLR[9], w[3];
  orc_expm(r, R);
  orc_jlt(r, Jlt);
  m3_mul(R, m->inertia_inv, RL);  /* (R * Lbody) */
  m3_mul_bt(RL, R, RLR);          /* ... * R^T   */
  m3_vec(RLR, l, w);              /* w = R Lbody R^T l  (:85) */
  double d0[3], d1[3], S0[9], S1[9];
  for (int i = 0; i < 3; ++i) {
    d0[i] = m->foot_pos[i] - p[i];
    d1[i] = m->foot_pos[3 + i] - p[i];
  }
  orc_skew(d0, S0);
  orc_skew(d1, S1);
  if (dx) {
    double t0[3], t1[3];
    m3_vec(Jlt, w, dx);  /* :90 */
    m3_vec(S0, u + 0, t0);
    m3_vec(S1, u + 6, t1);
    for (int i = 0; i < 3; ++i) {
      dx[3 + i] = ((u[3 + i] + u[9 + i]) + t0[i]) + t1[i];                    /* :92-94 */
      dx[6 + i] = v[i];                                                        /* :96 */
      dx[9 + i] = (u[i] + u[6 + i]) / m->mass + m->gravity[i];                 /* :98 */
    }
  }
  if (jfx) {
    double dJ[27], djw[9], Sl[9], Sw[9], X[9], T[9], Jl[9], Fs[3], SF[9];
    orc_djlt(r, dJ);
    for (int k = 0; k < 3; ++k) m3_vec(dJ + 9 * k, w, djw + 3 * k);  /* column k = dJlt_k * w (:111-113) */
    orc_jl(r, Jl);
    orc_skew(l, Sl);
    orc_skew(w, Sw);
    m3_mul(RLR, Sl, X);
    for (int i = 0; i < 9; ++i) X[i] = X[i] - Sw[i];
    m3_mul(Jlt, X, T);
    m3_mul(T, Jl, X);  /* Jlt * (RLR skew(l) - skew(w)) * Jl */
    memset(jfx, 0, 144 * sizeof(double));
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 3; ++i) jfx[i + 12 * j] = djw[i + 3 * j] + X[i + 3 * j];  /* :118 */
    m3_mul(Jlt, RLR, T);
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 3; ++i) jfx[i + 12 * (3 + j)] = T[i + 3 * j];  /* :119 */
    for (int i = 0; i < 3; ++i) Fs[i] = u[i] + u[6 + i];
    orc_skew(Fs, SF);
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 3; ++i) jfx[(3 + i) + 12 * (6 + j)] = SF[i + 3 * j];  /* :120 */
    for (int i = 0; i < 3; ++i) jfx[(6 + i) + 12 * (9 + i)] = 1.0;              /* :121 */
  }
  if (jfu) {
    memset(jfu, 0, 144 * sizeof(double));
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 3; ++i) {
        jfu[(3 + i) + 12 * (0 + j)] = S0[i + 3 * j];  /* :131 */
        jfu[(3 + i) + 12 * (6 + j)] = S1[i + 3 * j];  /* :133 */
      }
    for (int i = 0; i < 3; ++i) {
      jfu[(3 + i) + 12 * (3 + i)] = 1.0;  /* :132 */
      jfu[(3 + i) + 12 * (9 + i)] = 1.0;  /* :134 */
      jfu[(9 + i) + 12 * (0 + i)] = 1.0 / m->mass;  /* :136 */
      jfu[(9 + i) + 12 * (6 + i)] = 1.0 / m->mass;  /* :137 */
    }
  }
}

/* SRBD_model.cpp:143-235.  RK4 defect, Euler Jacobians (only k1_x, k1_u reach the outputs, :180-181) */
void orc_shooting(const srbd_model_params* m, const double* x, const double* xn, const double* u,
                  double* A, double* B, double* b, double* f) {
  double k1[12], k2[12], k3[12], k4[12], xt[12], jfx[144], jfu[144];
  const double dt = m->dt;
  const int need_jac = (A != NULL) || (B != NULL);
  orc_continuous(m, x, u, k1, need_jac ? jfx : NULL, need_jac ? jfu : NULL);
  for (int i = 0; i < 12; ++i) xt[i] = x[i] + (0.5 * dt) * k1[i];
  orc_continuous(m, xt, u, k2, NULL, NULL);
  for (int i = 0; i < 12; ++i) xt[i] = x[i] + (0.5 * dt) * k2[i];
  orc_continuous(m, xt, u, k3, NULL, NULL);
  for (int i = 0; i < 12; ++i) xt[i] = x[i] + dt * k3[i];
  orc_continuous(m, xt, u, k4, NULL, NULL);
  double fl[12];
  for (int i = 0; i < 12; ++i) {
    double xg = x[i] + (dt / 6.0) * (((k1[i] + 2.0 * k2[i]) + 2.0 * k3[i]) + k4[i]);  /* :179 */
    fl[i] = xn[i] - xg;  /* :194-197 (plain subtraction also for the rotation part) */
  }
  if (A) {
    for (int j = 0; j < 12; ++j)
      for (int i = 0; i < 12; ++i) A[i + 12 * j] = (i == j ? 1.0 : 0.0) + dt * jfx[i + 12 * j];  /* :180 */
  }
  if (B) {
    for (int i = 0; i < 144; ++i) B[i] = 0.0 + dt * jfu[i];  /* :181 (j_fx * 0 + j_fu) */
  }
  if (b)
    for (int i = 0; i < 12; ++i) b[i] = -fl[i];  /* :227 */
  if (f)
    for (int i = 0; i < 12; ++i) f[i] = fl[i];
}

/* SRBD_model.cpp:237-260.  Swing contacts (extension): fmax := swing_fmax. */
void orc_constraint(const srbd_model_params* m, const double* u, const uint8_t* stance, double* Ac,
                    double* f) {
  double bc[24];
  memset(Ac, 0, 24 * 12 * sizeof(double));
  memset(bc, 0, sizeof(bc));
  for (int leg = 0; leg < 2; ++leg) {
    const double* Rf = m->foot_rot + 9 * leg; /* column-major 3x3; column c = Rf + 3*c */
    const double* cx = Rf + 0;
    const double* cy = Rf + 3;
    const double* cz = Rf + 6;
    const int r0 = 12 * leg, c0 = 6 * leg;
#define AC(i, j) Ac[(r0 + (i)) + 24 * (c0 + (j))]
    AC(0, 0) = -1.0; AC(0, 2) = m->mu;
    AC(1, 1) = -1.0; AC(1, 2) = m->mu;
    AC(2, 0) = 1.0;  AC(2, 2) = m->mu;
    AC(3, 1) = 1.0;  AC(3, 2) = m->mu;
    AC(4, 2) = -1.0;
    AC(5, 2) = 1.0;
    for (int j = 0; j < 3; ++j) {
      AC(6, j) = m->Lfx * cz[j];  AC(6, 3 + j) = -cy[j];
      AC(7, j) = m->Lfx * cz[j];  AC(7, 3 + j) = cy[j];
      AC(8, j) = m->Lfz * cz[j];  AC(8, 3 + j) = -cz[j];
      AC(9, j) = m->Lfz * cz[j];  AC(9, 3 + j) = cz[j];
      AC(10, 3 + j) = -cx[j];
      AC(11, 3 + j) = cx[j];
    }
#undef AC
    const int st = stance ? (int)stance[leg] : 1;
    bc[r0 + 4] = st ? m->fmax : m->swing_fmax;
    bc[r0 + 5] = -m->fmin;
  }
  if (f) {
    for (int i = 0; i < 24; ++i) {
      double s = 0.0;
      for (int j = 0; j < 12; ++j) s += Ac[i + 24 * j] * u[j];
      f[i] = s + bc[i];  /* :259 */
    }
  }
}

/* SRBD_model.cpp:262-295 */
void orc_barrier(double v, double mu, double theta, double* b, double* db, double* ddb) {
  if (v > theta) {
    if (b) *b = -mu * log(v);
    if (db) *db = -mu / v;
    if (ddb) *ddb = mu / (v * v);
  } else {
    if (b) *b = 0.5 * mu * (((v - 2.0 * theta) / theta) * ((v - 2.0 * theta) / theta) - 1.0) - mu * log(theta);
    if (db) *db = mu * (v - 2.0 * theta) / (theta * theta);
    if (ddb) *ddb = mu / (theta * theta);
  }
}

/* Rows that stay a relaxed barrier in HARD_INEQ mode: the +-x^T tau pair (rows 10,11 of each foot)
 * has an empty strict interior; the reference's commented-out hard variant has 20 = 24-4 rows
 * (NMPC_solver.cpp:301). */
static int row_is_soft_in_hard_mode(int row) {
  int rr = row % 12;
  return rr == 10 || rr == 11;
}

/* NMPC_solver.cpp:276-314 (BARRIER_SOFT) and the hard-inequality variant of :300-304 */
void orc_assemble(const srbd_model_params* m, int N, int mode, const double* x, const double* u,
                  const double* xref, const uint8_t* contact, double* A, double* B, double* b,
                  double* Q, double* S, double* R, double* q, double* r, double* D, double* lg,
                  double* lg_mask, double* defect, double* fcon) {
  double Ac[24 * 12], fc[24], bb[24], db[24], ddb[24];
  for (int k = 0; k < N; ++k) {
    const double* xk = x + 12 * k;
    const double* xn = x + 12 * (k + 1);
    const double* uk = u + 12 * k;
    orc_shooting(m, xk, xn, uk, A ? A + 144 * k : NULL, B ? B + 144 * k : NULL, b ? b + 12 * k : NULL,
                 defect ? defect + 12 * k : NULL);
    orc_constraint(m, uk, contact ? contact + 2 * k : NULL, Ac, fc);
    if (fcon) memcpy(fcon + 24 * k, fc, sizeof(fc));
    for (int i = 0; i < 24; ++i) {
      orc_barrier(fc[i], m->mu_b, m->theta_b, &bb[i], &db[i], &ddb[i]);
      if (mode == SRBD_HARD_INEQ && !row_is_soft_in_hard_mode(i)) {
        db[i] = 0.0;
        ddb[i] = 0.0;
      }
    }
    if (Q) {
      double* Qk = Q + 144 * k;
      memset(Qk, 0, 144 * sizeof(double));
      for (int i = 0; i < 12; ++i) Qk[i + 12 * i] = m->Q[i];
    }
    if (q)
      for (int i = 0; i < 12; ++i) q[12 * k + i] = m->Q[i] * (xk[i] - xref[12 * k + i]);  /* :306 */
    if (S) memset(S + 144 * k, 0, 144 * sizeof(double));
    if (R) { /* R + Ac^T diag(ddb) Ac (:308) */
      double* Rk = R + 144 * k;
      for (int j = 0; j < 12; ++j)
        for (int i = 0; i < 12; ++i) {
          double s = 0.0;
          for (int g = 0; g < 24; ++g) s += (Ac[g + 24 * i] * ddb[g]) * Ac[g + 24 * j];
          Rk[i + 12 * j] = (i == j ? m->R : 0.0) + s;
        }
    }
    if (r) { /* R u + Ac^T db (:309) */
      for (int i = 0; i < 12; ++i) {
        double s = 0.0;
        for (int g = 0; g < 24; ++g) s += Ac[g + 24 * i] * db[g];
        r[12 * k + i] = m->R * uk[i] + s;
      }
    }
    if (mode == SRBD_HARD_INEQ) {
      if (D) memcpy(D + 288 * k, Ac, sizeof(Ac));
      for (int i = 0; i < 24; ++i) {
        if (lg) lg[24 * k + i] = -fc[i]; /* Ac (u+du) + bc >= 0  <=>  Ac du >= -(Ac u + bc) */
        if (lg_mask) lg_mask[24 * k + i] = row_is_soft_in_hard_mode(i) ? 0.0 : 1.0;
      }
    }
  }
  if (Q) {
    double* QN = Q + 144 * N;
    memset(QN, 0, 144 * sizeof(double));
    for (int i = 0; i < 12; ++i) QN[i + 12 * i] = m->Qf[i];  /* :312 */
  }
  if (q)
    for (int i = 0; i < 12; ++i) q[12 * N + i] = m->Qf[i] * (x[12 * N + i] - xref[12 * N + i]);  /* :313 */
}

/* cost of one stage as the line search evaluates it (NMPC_solver.cpp:166-187, 213-231) */
static void stage_merit(const srbd_model_params* m, int N, int mode, int k, const double* x, const double* u,
                        const double* xref, const uint8_t* contact, double* phi, double* theta,
                        double* Jx, double* Ju) {
  const double* xk = x + 12 * k;
  if (k == N) {
    double s = 0.0;
    for (int i = 0; i < 12; ++i) {
      double e = xk[i] - xref[12 * k + i];
      s += e * (m->Qf[i] * e);
      if (Jx) Jx[i] = m->Qf[i] * e;
    }
    *phi += 0.5 * s;
    return;
  }
  const double* uk = u + 12 * k;
  double f[12], Ac[288], fc[24];
  orc_shooting(m, xk, x + 12 * (k + 1), uk, NULL, NULL, NULL, f);
  double ss = 0.0;
  for (int i = 0; i < 12; ++i) ss += f[i] * f[i];
  *theta += 0.5 * ss;  /* :175 */
  double s = 0.0;
  for (int i = 0; i < 12; ++i) {
    double e = xk[i] - xref[12 * k + i];
    s += e * (m->Q[i] * e);
    if (Jx) Jx[i] = m->Q[i] * e;
  }
  *phi += 0.5 * s;  /* :177 */
  orc_constraint(m, uk, contact ? contact + 2 * k : NULL, Ac, fc);
  double bsum = 0.0, db[24];
  for (int g = 0; g < 24; ++g) {
    double bv;
    db[g] = 0.0;
    /* HARD_INEQ extension: the hard rows are constraints of the QP, not cost terms (only the rows the assembly keeps
     * as a relaxed barrier enter the merit function); mode 0 = the reference, all 24 rows */
    if (mode == SRBD_HARD_INEQ && !row_is_soft_in_hard_mode(g)) continue;
    orc_barrier(fc[g], m->mu_b, m->theta_b, &bv, &db[g], NULL);
    bsum += bv;
  }
  double uu = 0.0;
  for (int i = 0; i < 12; ++i) uu += uk[i] * (m->R * uk[i]);
  *phi += bsum + 0.5 * uu;  /* :186 */
  if (Ju)
    for (int i = 0; i < 12; ++i) {
      double a = 0.0;
      for (int g = 0; g < 24; ++g) a += Ac[g + 24 * i] * db[g];
      Ju[i] = a + m->R * uk[i];  /* :187 */
    }
}

/* NMPC_solver.cpp:149-274.  The reference's out-of-bounds read of u at k==N (:163,:210) has no effect
 * on the result and is not reproduced; `alpha` is carried across calls exactly like the member
 * alpha_ (NMPC_solver.h:104). */
int orc_line_search(const srbd_model_params* m, int N, double* x, double* u, const double* xref,
                    const uint8_t* contact, const double* dx, const double* du, double* alpha,
                    double* merit) {
  return orc_line_search_mode(m, N, SRBD_BARRIER_SOFT, x, u, xref, contact, dx, du, alpha, merit);
}

int orc_line_search_mode(const srbd_model_params* m, int N, int mode, double* x, double* u, const double* xref,
                         const uint8_t* contact, const double* dx, const double* du, double* alpha,
                         double* merit) {
  const double theta_max = 1e-6, theta_min = 5e-10, eta = 1e-4, byta_phi = 1e-6, byta_theta = 1e-6,
               byta_alpha = 0.5, alpha_min = 1e-4; /* NMPC_solver.h:97-103 */
  double theta = 0.0, phi = 0.0, dphi = 0.0;
  double* Jx = (double*)malloc(sizeof(double) * 12 * (N + 1));
  double* Ju = (double*)malloc(sizeof(double) * 12 * (N > 0 ? N : 1));
  double* xa = (double*)malloc(sizeof(double) * 12 * (N + 1));
  double* ua = (double*)malloc(sizeof(double) * 12 * (N > 0 ? N : 1));
  for (int k = 0; k <= N; ++k) stage_merit(m, N, mode, k, x, u, xref, contact, &phi, &theta, Jx + 12 * k, k < N ? Ju + 12 * k : NULL);
  for (int k = 0; k <= N; ++k) { /* :191-198 */
    double s = 0.0;
    for (int i = 0; i < 12; ++i) s += dx[12 * k + i] * Jx[12 * k + i];
    dphi += s;
    if (k < N) {
      double t = 0.0;
      for (int i = 0; i < 12; ++i) t += du[12 * k + i] * Ju[12 * k + i];
      dphi += t;
    }
  }
  while (*alpha > alpha_min) { /* :200-264 */
    double theta_a = 0.0, phi_a = 0.0;
    for (int i = 0; i < 12 * (N + 1); ++i) xa[i] = x[i] + *alpha * dx[i];
    for (int i = 0; i < 12 * N; ++i) ua[i] = u[i] + *alpha * du[i];
    for (int k = 0; k <= N; ++k) stage_merit(m, N, mode, k, xa, ua, xref, contact, &phi_a, &theta_a, NULL, NULL);
    int accept = 0;
    if (theta_a > theta_max) {
      if (theta_a < (1.0 - byta_theta) * theta) accept = 1;
    } else if ((fmax(theta_a, theta) < theta_min) && (dphi < 0.0)) {
      if (phi_a < phi + eta * (*alpha) * dphi) accept = 1;
    } else {
      if ((phi_a < phi - byta_phi * theta) || (theta_a < (1.0 - byta_theta) * theta)) accept = 1;
    }
    if (accept) {
      memcpy(x, xa, sizeof(double) * 12 * (N + 1));
      memcpy(u, ua, sizeof(double) * 12 * N);
      break;
    }
    *alpha = byta_alpha * (*alpha);
  }
  free(Jx); free(Ju); free(xa); free(ua);
  if (merit) { merit[0] = phi; merit[1] = dphi; merit[2] = theta; }
  return (dphi > -1e-3 && theta < 1e-6) ? 1 : 0; /* :267 */
}

/* defaults: SRBD_model.cpp:12-23, NMPC_solver.cpp:56-58,334-338, config/mpc_option.yaml */
void orc_model_params_default(srbd_model_params* p, int horizon) {
  memset(p, 0, sizeof(*p));
  p->mass = 15.0;
  p->dt = 0.015;
  const double Ib[3] = {0.541667, 0.516667, 1.0416667};
  for (int i = 0; i < 3; ++i) p->inertia_inv[i + 3 * i] = 1.0 / Ib[i];
  p->foot_pos[1] = -0.1;
  p->foot_pos[4] = 0.1;
  for (int leg = 0; leg < 2; ++leg)
    for (int i = 0; i < 3; ++i) p->foot_rot[9 * leg + i + 3 * i] = 1.0;
  p->mu = 0.5; p->Lfx = 0.05; p->Lfz = 0.05; p->fmax = 1000.0; p->fmin = 0.0;
  p->gravity[2] = -9.8;
  p->Q[11] = 10.0;
  const double Qf[12] = {0.5, 0.5, 0.5, 0.01, 0.01, 0.01, 100, 100, 100, 0.0, 0.0, 100};
  for (int i = 0; i < 12; ++i) p->Qf[i] = (double)horizon * Qf[i];
  p->R = 1e-4;
  p->mu_b = 0.1; p->theta_b = 5.0;
  p->swing_fmax = 1.0;
}
