// DMMA.8x8x4 fed from shared memory the way K3's FACTOR update does it (fragments from a row-major tile, ld = 100):
// cycles per k-panel (6 LDS.64 + 4 DMMA in four independent chains) for one warp, two warps on one SM sub-partition,
// and FACTOR's five update warps (3 .. 7).  (not part of the product)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_smem tools/dmma_smem.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kLd = 100;

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>   // 0: loads then DMMAs in one iteration; 1: operands one iteration ahead; 2: DMMAs only (registers); 3: loads only
__global__ void bench(double* out, long long* cycles, int reps, int active_mask) {
  extern __shared__ double A[];
  for (int e = threadIdx.x; e < 96 * kLd; e += blockDim.x) A[e] = 1e-3 * (e % 97);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (!((active_mask >> warp) & 1)) return;
  const int fr = lane >> 2, fc = lane & 3;
  const double* b_ = A + (16 + fr) * kLd + fc;
  const double* a0_ = A + (24 + 8 * (warp & 3) + fr) * kLd + fc;
  const double* a1_ = A + (56 + 8 * (warp & 3) + fr) * kLd + fc;
  double c00 = 0, c01 = 0, c10 = 0, c11 = 0, d00 = 0, d01 = 0, d10 = 0, d11 = 0;
  double b0 = b_[0], b1 = b_[4], x00 = a0_[0], x01 = a0_[4], x10 = a1_[0], x11 = a1_[4];
  double sink = 0;
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll 1
    for (int k = 8; k < 88; k += 8) {
      if (MODE == 0) {
        b0 = b_[k]; b1 = b_[k + 4]; x00 = a0_[k]; x01 = a0_[k + 4]; x10 = a1_[k]; x11 = a1_[k + 4];
        dmma(c00, c01, x00, b0); dmma(c10, c11, x10, b0); dmma(d00, d01, x01, b1); dmma(d10, d11, x11, b1);
      } else if (MODE == 1) {
        const double nb0 = b_[k], nb1 = b_[k + 4], nx00 = a0_[k], nx01 = a0_[k + 4], nx10 = a1_[k], nx11 = a1_[k + 4];
        dmma(c00, c01, x00, b0); dmma(c10, c11, x10, b0); dmma(d00, d01, x01, b1); dmma(d10, d11, x11, b1);
        b0 = nb0; b1 = nb1; x00 = nx00; x01 = nx01; x10 = nx10; x11 = nx11;
      } else if (MODE == 2) {
        dmma(c00, c01, x00, b0); dmma(c10, c11, x10, b0); dmma(d00, d01, x01, b1); dmma(d10, d11, x11, b1);
      } else {
        sink += b_[k] + b_[k + 4] + a0_[k] + a0_[k + 4] + a1_[k] + a1_[k + 4];
      }
    }
  }
  const long long t1 = clock64();
  out[threadIdx.x] = c00 + c01 + c10 + c11 + d00 + d01 + d10 + d11 + sink;
  if (threadIdx.x == 32 * (31 - __clz(active_mask))) cycles[0] = t1 - t0;   // the highest active warp reports
}

template <int MODE>
void run(double* out, long long* cyc, int mask, const char* what) {
  const int reps = 400;
  long long h;
  cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * kLd * 8);
  bench<MODE><<<1, 256, 96 * kLd * 8>>>(out, cyc, reps, mask);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const char* names[] = {"loads then DMMAs", "operands one panel ahead", "DMMAs only", "loads only"};
  printf("%-34s %-26s %6.1f cycles per k-panel (6 LDS.64 + 4 DMMA)\n", what, names[MODE], (double)h / reps / 10);
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8192); cudaMalloc(&cyc, 64);
  const struct { int mask; const char* what; } cases[] = {{0x80, "one warp (7)"}, {0x88, "warps 3 + 7 (one sub-partition)"}, {0xf8, "warps 3 .. 7"}, {0xff, "all 8 warps"}};
  for (auto& c : cases) {
    run<0>(out, cyc, c.mask, c.what);
    run<1>(out, cyc, c.mask, c.what);
    run<2>(out, cyc, c.mask, c.what);
    run<3>(out, cyc, c.mask, c.what);
  }
  return 0;
}
