// Measures the FP64 peaks the K2/K3 rooflines are quoted against (not part of the product):
//   DFMA  : register-resident fused multiply-add chains
//   DMMA  : mma.sync.aligned.m8n8k4.f64 chains (the FP64 tensor path on sm_100a)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, int iters) {
  double a[ILP], b = 1.0000001, c = 1e-9;
  for (int k = 0; k < ILP; ++k) a[k] = threadIdx.x + k;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int k = 0; k < ILP; ++k) a[k] = fma(a[k], b, c);
  double s = 0;
  for (int k = 0; k < ILP; ++k) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma_kernel(double* out, int iters) {
  double c0[ILP], c1[ILP];
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int k = 0; k < ILP; ++k) c0[k] = c1[k] = 0.0;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int k = 0; k < ILP; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[k]), "+d"(c1[k]) : "d"(a), "d"(b));
  double s = 0;
  for (int k = 0; k < ILP; ++k) s += c0[k] + c1[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    for (int bps : {1, 2}) {
      if (threads * bps > 2048) continue;
      const int grid = sms * bps;
      float ms = time_ms([&] { dfma_kernel<8><<<grid, threads>>>(out, iters); });
      double fl = 2.0 * 8 * iters * (double)threads * grid;
      printf("DFMA  threads %4d x %d CTA/SM: %.2f TFLOP/s\n", threads, bps, fl / ms / 1e9);
      ms = time_ms([&] { dmma_kernel<8><<<grid, threads>>>(out, iters); });
      fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)(threads / 32) * grid;
      printf("DMMA  threads %4d x %d CTA/SM: %.2f TFLOP/s\n", threads, bps, fl / ms / 1e9);
    }
  }
  return 0;
}
