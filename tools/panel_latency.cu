// How long does the 8-column register panel of K3's FACTOR task take by itself?  (not part of the product)
// One warp (optionally with idle or DMMA-busy neighbours) runs the diagonal-block arithmetic of factor_panel
// (k3_dag.cu) on register data, R times back to back with a true dependency between repetitions, timed by clock64.
//   full  : the 8 x 8 block, every lane redundantly (36 values), as factor_panel does
//   chain : only what is on the pivot chain (rsqrt -> column scale -> next pivot), no trailing update of the block
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o panel_latency panel_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double rsqrt_pivot(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double a = y * y;
  const double e = fma(d, -a, 1.0);
  const double p = fma(e, 0.375, 0.5);
  const double q = y * e;
  return fma(p, q, y);
}

template <bool FULL>
__device__ __forceinline__ double panel(double (&D)[36]) {
  double acc = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int jj = j * (j + 1) / 2 + j;
    const double d = D[jj];
    const double rs = rsqrt_pivot(d);
    D[jj] = d * rs;
#pragma unroll
    for (int i = j + 1; i < 8; ++i) D[i * (i + 1) / 2 + j] *= rs;
    if (FULL) {
#pragma unroll
      for (int c = j + 1; c < 8; ++c)
#pragma unroll
        for (int i = c; i < 8; ++i) D[i * (i + 1) / 2 + c] -= D[i * (i + 1) / 2 + j] * D[c * (c + 1) / 2 + j];
    } else if (j + 1 < 8) {
      const int c = j + 1;
      D[c * (c + 1) / 2 + c] -= D[c * (c + 1) / 2 + j] * D[c * (c + 1) / 2 + j];
    }
    acc += D[jj];
  }
  return acc;
}

template <bool FULL>
__global__ void bench(double* out, long long* cycles, int reps, int busy_mode) {
  const int warp = threadIdx.x >> 5;
  if (warp > 0 && !(busy_mode >= 2 && warp < busy_mode)) {   // neighbours: 0 = exit at once, 1 = DMMA chains on every sub-partition while warp 0 measures;
                                                              // n >= 2: warps 1 .. n-1 run the same panel (FACTOR's redundant panel warps)
    if (busy_mode == 1) {
      double c0 = 0, c1 = 0, a = threadIdx.x * 1e-3, b = 1.0;
      for (int i = 0; i < reps * 64; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
      out[threadIdx.x] = c0 + c1;
    }
    return;
  }
  double D[36];
  double total = 0.0;
  const long long t0 = clock64();
  double carry = 0.0;
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) D[i * (i + 1) / 2 + j] = (i == j ? 9.0 + i : 0.1 * (i + 1) / (j + 2)) + carry * 1e-30;
    carry = panel<FULL>(D);
    total += carry;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
  out[threadIdx.x] = total;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8192); cudaMalloc(&cyc, 64);
  const int reps = 2000;
  for (int threads : {32, 256}) for (int busy : {0, 1, 2, 3, 4, 8}) {
    if (threads == 32 && busy) continue;
    long long h;
    bench<true><<<1, threads>>>(out, cyc, reps, busy); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("full  panel, %3d threads, neighbours %s (%d): %6.0f cycles per 8 columns\n", threads, busy == 1 ? "DMMA" : (busy ? "panel warps" : "idle"), busy, (double)h / reps);
    bench<false><<<1, threads>>>(out, cyc, reps, busy); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("chain only,  %3d threads, neighbours %s (%d): %6.0f cycles per 8 columns\n", threads, busy == 1 ? "DMMA" : (busy ? "panel warps" : "idle"), busy, (double)h / reps);
  }
  return 0;
}
