// Host-side check (no GPU needed) of babt_term() -- the one-term-per-element form K1 expands the compact
// linearization record with -- against the dense definition A = I + dt jfx, B = dt jfu built from the blocks of
// SRBD_model.cpp:105-141 (jfx: d(rdot)/d[r,l], skew(F sum), I; jfu: skew(d_leg), I, I/m).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "../../srbd-nmpc-solver_b200/csrc/srbd_model.cuh"

static void skew(const double* v, double S[3][3]) {
  S[0][0] = 0; S[0][1] = -v[2]; S[0][2] = v[1];
  S[1][0] = v[2]; S[1][1] = 0; S[1][2] = -v[0];
  S[2][0] = -v[1]; S[2][1] = v[0]; S[2][2] = 0;
}

int main() {
  using namespace srbd;
  const double dt = 0.015, minv = 1.0 / 15.0;
  int bad = 0, varying = 0;
  srand(11);
  for (int trial = 0; trial < 200; ++trial) {
    double c[40];
    for (int i = 0; i < 39; ++i) c[i] = (rand() / (double)RAND_MAX - 0.5) * 20.0;
    c[39] = 0.0;
    double jfx[12][12] = {{0}}, jfu[12][12] = {{0}}, SF[3][3], S0[3][3], S1[3][3];
    skew(c + 18, SF); skew(c + 21, S0); skew(c + 24, S1);
    for (int a = 0; a < 3; ++a) {
      for (int b = 0; b < 6; ++b) jfx[a][b] = c[a * 6 + b];
      for (int b = 0; b < 3; ++b) {
        jfx[3 + a][6 + b] = SF[a][b];
        jfu[3 + a][b] = S0[a][b];
        jfu[3 + a][6 + b] = S1[a][b];
      }
      jfx[6 + a][9 + a] = 1.0;
      jfu[3 + a][3 + a] = 1.0; jfu[3 + a][9 + a] = 1.0;
      jfu[9 + a][a] = minv; jfu[9 + a][6 + a] = minv;
    }
    for (int st0 = 0; st0 < 2; ++st0)
      for (int i = 0; i < 28; ++i)
        for (int j = 0; j < 12; ++j) {
          double ref = 0.0;
          if (i < 12) ref = dt * jfu[j][i];
          else if (st0) ref = (i == 12) ? c[27 + j] : 0.0;
          else if (i < 24) ref = fma(dt, jfx[j][i - 12], j == i - 12 ? 1.0 : 0.0);
          else if (i == 24) ref = c[27 + j];
          const double got = babt_elem(c, i, j, dt, minv, st0 != 0);
          if (!(ref == got)) {
            if (bad < 10) printf("mismatch st0=%d (%d,%d): ref %.17g got %.17g\n", st0, i, j, ref, got);
            ++bad;
          }
          if (trial == 0 && !st0 && babt_term(i, j, dt, minv, false).mul != 0.0) ++varying;
        }
  }
  if (varying != 48) { printf("expected 48 stage-dependent elements, found %d\n", varying); ++bad; }
  // the dyn chunks of layout.cuh (what K3's compact BAbt streaming copies per stage): 36 disjoint, 16-byte aligned pairs of
  // the dense record that together contain every stage-dependent element
  {
    bool covered[336] = {false};
    for (int c = 0; c < kBabtDynChunks; ++c) {
      const int o = babt_dyn_off(c);
      if (o < 0 || o + 1 >= 336 || (o & 1) || covered[o] || covered[o + 1]) { printf("bad dyn chunk %d at %d\n", c, o); ++bad; continue; }
      covered[o] = covered[o + 1] = true;
    }
    for (int e = 0; e < 336; ++e) {
      const int pnl = e / 48, rem = e - pnl * 48;
      const bool dep = babt_term(4 * pnl + (rem & 3), rem >> 2, dt, minv, false).mul != 0.0;
      if (dep && !covered[e]) { printf("stage-dependent element %d is in no dyn chunk\n", e); ++bad; }
    }
  }
  if (bad) { printf("FAIL %d\n", bad); return 1; }
  printf("OK\n");
  return 0;
}
