// dmma_power.cu — does the FP64 tensor pipe burn less power when 7 of the 8 A-operand rows are zero?
// (K3's vector sweeps issue their gemvs as DMMA m8n8k4 with the vector in row 0 of the A operand; the other rows
// replicate it.)  Runs a DMMA-saturating kernel for <seconds> in one of four operand modes while the caller samples
// `nvidia-smi --query-gpu=power.draw,clocks.sm`:
//   0: all 8 rows distinct   1: rows replicate row 0   2: rows 1..7 zero   3: DFMA instead of DMMA
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dmma_power scripts/microbench/dmma_power.cu
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void k_dmma(double* out, int iters, int mode, double x) {
  const int lane = threadIdx.x & 31, r = lane >> 2, t = lane & 3;
  double c0[8], c1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c0[i] = 0.001 * (lane + i); c1[i] = 0.002 * i; }
  // A operand: lane (r, t) holds A[r][t]
  double a = x * (1.0 + 0.37 * t);
  if (mode == 0) a *= (1.0 + 0.11 * r);
  if (mode == 2 && r != 0) a = 0.0;
  const double b = 0.5 * x * (1.0 + 0.013 * lane);
  if (mode == 2 && r != 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { c0[i] = 0.0; c1[i] = 0.0; }
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c0[i], c1[i], a, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) { c0[i] *= 0.5; c1[i] *= 0.5; }  // keep the values bounded (DMUL: same in every mode)
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma(double* out, int iters, double x) {
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], x, 1.0);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const double seconds = argc > 2 ? atof(argv[2]) : 3.0;
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int blocks = prop.multiProcessorCount * 4, threads = 128;  // 16 warps / SM
  double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  const int iters = 20000;
  const auto t0 = std::chrono::steady_clock::now();
  long long launches = 0;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms_sum = 0;
  while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < seconds) {
    cudaEventRecord(e0);
    if (mode == 3) k_dfma<<<blocks, threads>>>(out, iters * 8, 0.999999);
    else k_dmma<<<blocks, threads>>>(out, iters, mode, 0.7);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms_sum += ms; ++launches;
  }
  const double flops = mode == 3 ? 2.0 * 8 * iters * 8 * (double)blocks * threads
                                 : 2.0 * 256 * 8 * iters * (double)blocks * (threads / 32);
  printf("mode %d: %lld launches, %.2f ms each, %.1f TFLOP/s issued\n", mode, launches, ms_sum / launches,
         flops / (ms_sum / launches * 1e-3) / 1e12);
  return 0;
}
