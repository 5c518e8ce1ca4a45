// fp64_pipes.cu — microbenchmark of the FP64 paths on sm_100a: DFMA vs DMMA (mma.sync m8n8k4 / m16n8k8 f64)
// throughput and dependent-issue latency, DFMA+DMMA overlap, LDS / SHFL latency.  Design input for K3.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_pipes scripts/microbench/fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int ILP>
__global__ void k_dfma(double* out, int iters, double x) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], x, 1.0);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_dmma(double* out, int iters, double x) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
  double a = x, b = x * 0.5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_dmma1688(double* out, int iters, double x) {
  double c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  double a[4] = {x, x * 0.5, x * 0.25, x * 2}, b[2] = {x * 0.5, x};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma1688(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: per loop trip MM dmma884 + FF dfma (independent chains)
template <int MM, int FF>
__global__ void k_mix(double* out, int iters, double x) {
  double c0[MM], c1[MM], acc[FF];
#pragma unroll
  for (int i = 0; i < MM; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
#pragma unroll
  for (int i = 0; i < FF; ++i) acc[i] = threadIdx.x + i;
  double a = x, b = x * 0.5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < (MM > FF ? MM : FF); ++i) {
      if (i < MM) dmma884(c0[i], c1[i], a, b);
      if (i < FF) acc[i] = fma(acc[i], x, 1.0);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < MM; ++i) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < FF; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// latency: single warp, dependent chain; clock64
__global__ void k_lat(long long* res, double* out, double x, int iters) {
  __shared__ double sh[64];
  sh[threadIdx.x] = threadIdx.x; sh[threadIdx.x + 32] = 1.0;
  __syncwarp();
  double acc = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) acc = fma(acc, x, 1.0);
  long long t1 = clock64();
  double c0 = acc, c1 = 1.0;
  for (int it = 0; it < iters; ++it) dmma884(c0, c1, x, x);
  long long t2 = clock64();
  double s = c0 + c1;
  for (int it = 0; it < iters; ++it) s = __shfl_sync(0xffffffffu, s, (threadIdx.x + 1) & 31);
  long long t3 = clock64();
  int idx = threadIdx.x;
  for (int it = 0; it < iters; ++it) idx = (int)sh[idx & 31] ;
  long long t4 = clock64();
  double cc[4] = {s, 1, 2, 3}; double a4[4] = {x, x, x, x}, b2[2] = {x, x};
  for (int it = 0; it < iters; ++it) dmma1688(cc, a4, b2);
  long long t5 = clock64();
  // dependent DADD / DMUL
  double q = s;
  for (int it = 0; it < iters; ++it) q = q + x;
  long long t6 = clock64();
  // sqrt / rsqrt / div
  double r = fabs(q) + 2.0;
  for (int it = 0; it < iters; ++it) r = rsqrt(r) + 2.0;
  long long t7 = clock64();
  double r2 = r;
  for (int it = 0; it < iters; ++it) r2 = 1.0 / sqrt(r2) + 2.0;
  long long t8 = clock64();
  double r3 = r2;
  for (int it = 0; it < iters; ++it) r3 = 1.0 / r3 + 2.0;
  long long t9 = clock64();
  if (threadIdx.x == 0) {
    res[0] = t1 - t0; res[1] = t2 - t1; res[2] = t3 - t2; res[3] = t4 - t3; res[4] = t5 - t4; res[5] = t6 - t5;
    res[6] = t7 - t6; res[7] = t8 - t7; res[8] = t9 - t8;
  }
  out[threadIdx.x] = s + idx + cc[0] + cc[1] + cc[2] + cc[3] + q + r + r2 + r3;
}

template <typename F>
float time_kernel(F launch, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 1024));
  long long* res; CK(cudaMalloc(&res, 16 * sizeof(long long)));
  const int iters = 4096;
  // throughput, vary warps per SM
  for (int wps : {4, 8, 16, 32}) {
    const int threads = 128, blocks = sms * (wps / 4);
    const double nthr = (double)blocks * threads;
    float ms;
    ms = time_kernel([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001); }, 5);
    printf("warps/SM %2d  DFMA ilp8      : %7.2f TFLOP/s\n", wps, 2.0 * nthr * iters * 8 / ms * 1e-9);
    ms = time_kernel([&] { k_dmma<8><<<blocks, threads>>>(out, iters, 1.0000001); }, 5);
    printf("warps/SM %2d  DMMA884 ilp8   : %7.2f TFLOP/s\n", wps, 2.0 * 256 * (nthr / 32) * iters * 8 / ms * 1e-9);
    ms = time_kernel([&] { k_dmma<2><<<blocks, threads>>>(out, iters, 1.0000001); }, 5);
    printf("warps/SM %2d  DMMA884 ilp2   : %7.2f TFLOP/s\n", wps, 2.0 * 256 * (nthr / 32) * iters * 2 / ms * 1e-9);
    ms = time_kernel([&] { k_dmma1688<4><<<blocks, threads>>>(out, iters, 1.0000001); }, 5);
    printf("warps/SM %2d  DMMA1688 ilp4  : %7.2f TFLOP/s\n", wps, 2.0 * 1024 * (nthr / 32) * iters * 4 / ms * 1e-9);
    ms = time_kernel([&] { k_mix<4, 4><<<blocks, threads>>>(out, iters, 1.0000001); }, 5);
    printf("warps/SM %2d  mix 4 DMMA + 4 DFMA : %7.2f TFLOP/s total (DMMA part %.2f, DFMA part %.2f)\n", wps,
           (2.0 * 256 * (nthr / 32) * 4 + 2.0 * nthr * 4) * iters / ms * 1e-9, 2.0 * 256 * (nthr / 32) * 4 * iters / ms * 1e-9,
           2.0 * nthr * 4 * iters / ms * 1e-9);
    ms = time_kernel([&] { k_mix<1, 8><<<blocks, threads>>>(out, iters, 1.0000001); }, 5);
    printf("warps/SM %2d  mix 1 DMMA + 8 DFMA : %7.2f TFLOP/s total\n", wps,
           (2.0 * 256 * (nthr / 32) * 1 + 2.0 * nthr * 8) * iters / ms * 1e-9);
  }
  k_lat<<<1, 32>>>(res, out, 1.0000001, iters);
  CK(cudaDeviceSynchronize());
  long long h[16]; CK(cudaMemcpy(h, res, sizeof(h), cudaMemcpyDeviceToHost));
  const char* names[] = {"DFMA", "DMMA884", "SHFL", "LDS(dep, +cvt)", "DMMA1688", "DADD", "rsqrt()+DADD", "1/sqrt()+DADD", "1/x+DADD"};
  for (int i = 0; i < 9; ++i) printf("latency %-16s %7.1f cycles\n", names[i], (double)h[i] / iters);
  return 0;
}
