// rsqrt_seed.cu — accuracy of the FP64 reciprocal-square-root seed (rsqrt.approx.ftz.f64 = MUFU.RSQ64H) and of the
// refinements K3's pivot inverse builds on it: third-order step, + one Newton step, CUDA's rsqrt(), 1/sqrt().
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/rsqrt_seed scripts/microbench/rsqrt_seed.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(const double* x, double* o, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = x[i], y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));
  o[4 * i] = y;
  const double tt = v * y, e = fma(-tt, y, 1.0), q = e * fma(0.375, e, 0.5);
  double y1 = fma(y, q, y);
  o[4 * i + 1] = y1;
  const double t2 = v * y1, e2 = fma(-t2, y1, 1.0);
  o[4 * i + 2] = fma(y1 * 0.5, e2, y1);
  o[4 * i + 3] = rsqrt(v);
}
int main() {
  const int n = 1 << 22;
  double *hx = new double[n], *ho = new double[4 * n], *dx, *dout;
  unsigned long long s = 88172645463325252ull;
  for (int i = 0; i < n; ++i) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    const double u = (double)(s >> 11) * (1.0 / 9007199254740992.0);
    hx[i] = exp(-40.0 + 100.0 * u) * (1.0 + 1e-3 * (i % 977));
  }
  cudaMalloc(&dx, n * 8); cudaMalloc(&dout, 4 * n * 8);
  cudaMemcpy(dx, hx, n * 8, cudaMemcpyHostToDevice);
  k<<<n / 256, 256>>>(dx, dout, n);
  cudaMemcpy(ho, dout, 4 * n * 8, cudaMemcpyDeviceToHost);
  const char* names[4] = {"seed rsqrt.approx.ftz.f64", "third-order step (round-1 pivot inverse)", "+ one Newton step", "CUDA rsqrt()"};
  for (int c = 0; c < 4; ++c) {
    long double worst = 0, sum = 0;
    for (int i = 0; i < n; ++i) {
      const long double t = 1.0L / sqrtl((long double)hx[i]);
      const long double e = fabsl(((long double)ho[4 * i + c] - t) / t);
      if (e > worst) worst = e;
      sum += e;
    }
    printf("%-42s max rel. error %.3Le (%.2Lf ulp)  mean %.3Le\n", names[c], worst, worst / 1.1102230246251565e-16L, sum / n);
  }
  return 0;
}
