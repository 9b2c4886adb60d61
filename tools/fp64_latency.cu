// Dependent-issue latencies of the FP64 instructions the K3 diagonal-tile chain is made of (not part of the
// product): one warp, one long dependent chain per instruction, clock64 around it.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 512;

__global__ void lat_kernel(double* out, long long* cyc, double seed) {
  const int lane = threadIdx.x;
  double x = seed + lane * 1e-9, b = 1.0000001, c = 1e-9;
  long long t0, t1;
  // DFMA
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, b, c);
  t1 = clock64();
  if (lane == 0) cyc[0] = t1 - t0;
  // DMUL
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x * b;
  t1 = clock64();
  if (lane == 0) cyc[1] = t1 - t0;
  // DADD
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x + c;
  t1 = clock64();
  if (lane == 0) cyc[2] = t1 - t0;
  // rsqrt (library sequence)
  x = fabs(x) + 1.0;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) x = rsqrt(x) + 1.0;
  t1 = clock64();
  if (lane == 0) cyc[3] = t1 - t0;   // includes one DADD per step
  // 1/x (library)
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) x = 1.0 / x + 1.0;
  t1 = clock64();
  if (lane == 0) cyc[4] = t1 - t0;
  // sqrt
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) x = sqrt(x) + 1.0;
  t1 = clock64();
  if (lane == 0) cyc[5] = t1 - t0;
  // DMMA dependent chain (accumulator feeds the next)
  double c0 = x, c1 = x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(b), "d"(c));
  t1 = clock64();
  if (lane == 0) cyc[6] = t1 - t0;
  // DMMA where the A operand depends on the previous result
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    double d0 = 0.0, d1 = 0.0;
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(c0), "d"(c));
    c0 = d0;
  }
  t1 = clock64();
  if (lane == 0) cyc[7] = t1 - t0;
  // shuffle of a double (two 32-bit shuffles)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
  t1 = clock64();
  if (lane == 0) cyc[8] = t1 - t0;
  // float FFMA for comparison
  float f = (float)x, fb = 1.0000001f, fc = 1e-9f;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) f = fmaf(f, fb, fc);
  t1 = clock64();
  if (lane == 0) cyc[9] = t1 - t0;
  // independent DFMAs (8 chains): issue cost per warp instruction
  double y[8];
  for (int k = 0; k < 8; ++k) y[k] = x + k;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) y[k] = fma(y[k], b, c);
  t1 = clock64();
  if (lane == 0) cyc[10] = t1 - t0;
  double s = x + c0 + c1 + f;
  for (int k = 0; k < 8; ++k) s += y[k];
  out[lane] = s;
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 32 * sizeof(double));
  cudaMalloc(&cyc, 16 * sizeof(long long));
  for (int rep = 0; rep < 2; ++rep) lat_kernel<<<1, 32>>>(out, cyc, 1.5);
  long long h[16];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
  const char* names[] = {"DFMA dependent", "DMUL dependent", "DADD dependent", "rsqrt(double)+DADD", "1/x (double)+DADD",
                         "sqrt(double)+DADD", "DMMA accumulate chain", "DMMA operand chain", "shfl double",
                         "FFMA dependent", "DFMA 8 independent chains (per instr)"};
  for (int k = 0; k < 11; ++k)
    printf("%-40s %8.1f cycles per op\n", names[k], (double)h[k] / (k == 10 ? N * 8 : N));
  return 0;
}
