// Throughput half of the Ozaki sizing (tests/test_ozaki_emulation.py is the error half):
// what the int8 tensor path reaches from plain mma.sync on sm_100a -- the legacy warp-level path, i.e. a LOWER bound
// for the tcgen05.mma kind::i8 route (nominal 4.5 POP/s dense on B200) -- next to the FP64 DMMA the Schur SYRK uses.
//   IMMA : mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 chains, register-resident operands
//   DMMA : mma.sync.aligned.m8n8k4.f64 chains
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_peak imma_peak.cu      (not part of the product)
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void imma_kernel(int* out, int iters) {
  int c[ILP][4];
  const unsigned a0 = 0x01020304u * (threadIdx.x + 1), a1 = a0 ^ 0x11111111u, a2 = a0 + 7, a3 = a1 + 3;
  const unsigned b0 = 0x04030201u + threadIdx.x, b1 = b0 ^ 0x22222222u;
  for (int k = 0; k < ILP; ++k) c[k][0] = c[k][1] = c[k][2] = c[k][3] = 0;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int k = 0; k < ILP; ++k)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(c[k][0]), "+r"(c[k][1]), "+r"(c[k][2]), "+r"(c[k][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  int s = 0;
  for (int k = 0; k < ILP; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
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
  void* out;
  cudaMalloc(&out, sizeof(double) * sms * 2 * 1024);
  const int iters = 20000;
  double best_i = 0, best_d = 0;
  for (int threads : {128, 256, 512, 1024}) {
    for (int bps : {1, 2}) {
      if (threads * bps > 2048) continue;
      const int grid = sms * bps;
      float ms = time_ms([&] { imma_kernel<8><<<grid, threads>>>((int*)out, iters); });
      const double ops = 2.0 * 16 * 8 * 32 * 8 * iters * (double)(threads / 32) * grid;
      printf("IMMA m16n8k32 s8  threads %4d x %d CTA/SM: %8.1f TOP/s\n", threads, bps, ops / ms / 1e9);
      if (ops / ms / 1e9 > best_i) best_i = ops / ms / 1e9;
      ms = time_ms([&] { dmma_kernel<8><<<grid, threads>>>((double*)out, iters); });
      const double fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)(threads / 32) * grid;
      printf("DMMA m8n8k4 f64   threads %4d x %d CTA/SM: %8.2f TFLOP/s\n", threads, bps, fl / ms / 1e9);
      if (fl / ms / 1e9 > best_d) best_d = fl / ms / 1e9;
    }
  }
  printf("best: IMMA (mma.sync) %.1f TOP/s, DMMA %.2f TFLOP/s, ratio %.1f\n", best_i, best_d, best_i / best_d);
  printf("FP64-equivalent of an Ozaki split on this path: 15 GEMMs (5 slices) %.1f, 21 (6) %.1f, 28 (7) %.1f TFLOP/s "
         "at full int8 rate, before quantisation and recombination\n", best_i / 15, best_i / 21, best_i / 28);
  return 0;
}
