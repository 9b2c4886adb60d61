// How fast can ONE warp issue independent DMMA.8x8x4 (mma.sync.m8n8k4.f64), and how does that change with more warps
// on the same SM sub-partition?  (not part of the product; sizing for K3's FACTOR trailing update)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_issue tools/dmma_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int N>
__global__ void bench(double* out, long long* cycles, int reps, int active_mask) {
  const int warp = threadIdx.x >> 5;
  if (!((active_mask >> warp) & 1)) return;
  double c0[N], c1[N];
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
#pragma unroll
  for (int i = 0; i < N; ++i) c0[i] = c1[i] = 0.0;
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < N; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) s += c0[i] + c1[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int N>
void run(double* out, long long* cyc, int threads, int mask, const char* what) {
  const int reps = 4000;
  long long h;
  bench<N><<<1, threads>>>(out, cyc, reps, mask);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %d independent accumulators: %6.1f cycles per DMMA (warp 0)\n", what, N, (double)h / reps / N);
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8192); cudaMalloc(&cyc, 64);
  run<1>(out, cyc, 32, 1, "1 warp");
  run<2>(out, cyc, 32, 1, "1 warp");
  run<4>(out, cyc, 32, 1, "1 warp");
  run<8>(out, cyc, 32, 1, "1 warp");
  run<4>(out, cyc, 256, 0x11, "2 warps on one sub-partition (0, 4)");
  run<4>(out, cyc, 512, 0x1111, "4 warps on one sub-partition");
  run<4>(out, cyc, 128, 0xf, "4 warps, one per sub-partition");
  run<4>(out, cyc, 256, 0xff, "8 warps, two per sub-partition");
  run<1>(out, cyc, 256, 0xff, "8 warps, two per sub-partition");
  run<1>(out, cyc, 512, 0xffff, "16 warps, four per sub-partition");
  run<2>(out, cyc, 512, 0xffff, "16 warps, four per sub-partition");
  return 0;
}
