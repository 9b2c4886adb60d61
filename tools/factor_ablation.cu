// What does a block step of K3's FACTOR task consist of?  (not part of the product)
// One CTA runs factor_tile's loop (k3_dag.cu, included as source so the panel and DMMA helpers are the product's own) on a
// synthetic SPD 96 x 96 tile with pieces switched off, timed by clock64 per block step.
//   bit 0: no panel        bit 1: no phase-2 update      bit 2: no phase-1 update     bit 3: panel on warp 0 only (32 rows)
//   bit 5: phase-2 update with four blocks' operands requested before the first DMMA
//   bit 7: look-ahead order (panels 0 .. I applied to block column I + 2 during panel I + 1)
//   bit 4: update warps start at warp 4 (none beside panel warps 0-2 on their sub-partitions except warp 4-6)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I rsba_b200/csrc -o tools/factor_ablation
//        tools/factor_ablation.cu -L rsba_b200/lib -lrsba_cuda -Xlinker -rpath=$PWD/rsba_b200/lib
#include "../rsba_b200/csrc/k3_dag.cu"

#include <cmath>
#include <cstdio>
#include <vector>

namespace rsba {
namespace {

template <int V>
__device__ __forceinline__ void factor_tile_v(double* A, double* rdiag, int* info, long long* steps) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int fr = lane >> 2, fc = lane & 3;
  if (!(V & 1) && warp < 3) factor_panel(A, rdiag, 0, warp, (V & 8) ? 1 : 3, lane, 0, info);
  __syncthreads();
  for (int I = 0; I + 1 < kTile / 8; ++I) {
    if (tid == 0) steps[I] = clock64();
    const int o = 8 * I;
    const int nb = kTile / 8 - 1 - I;
    const double* Pm = A + (o + 8) * kLd + o;
    double* C22 = A + (o + 8) * kLd + o + 8;
    auto update_block = [&](int bi, int bj) {
      double c0 = 0.0, c1 = 0.0;
      const double* ap = Pm + (8 * bi + fr) * kLd + fc;
      const double* bp = Pm + (8 * bj + fr) * kLd + fc;
      dmma(c0, c1, ap[0], bp[0]);
      dmma(c0, c1, ap[4], bp[4]);
      double* cp = C22 + (8 * bi + fr) * kLd + 8 * bj + 2 * fc;
      cp[0] -= c0;
      cp[1] -= c1;
    };
    if (!(V & 4))
      for (int bi = warp; bi < nb; bi += 8) update_block(bi, 0);
    __syncthreads();
    if (tid == 0) steps[16 + I] = clock64();
    const int rows_next = kTile - (o + 8) - 8;
    int np = rows_next > 64 ? 3 : (rows_next > 32 ? 2 : 1);
    if (V & 8) np = 1;
    const int first_update = (V & 16) ? 4 : np;
    if (warp < np) {
      if (!(V & 1)) factor_panel(A, rdiag, o + 8, warp, np, lane, 0, info);
      if (tid == 0) steps[32 + I] = clock64();
    } else if ((V & 128) && !(V & 2)) {
      // look-ahead order: this step applies panels 0 .. I to block column I + 2 only (K = 8 (I + 1), at most two
      // blocks per warp, one read-modify-write of C per block); phase 1 then only has panel I + 1 left to apply
      const int u = warp - np, nu = 8 - np;
      const int o2 = o + 16;
      const int nblk = nb - 1;                       // block rows o2 / 8 .. 11 of column I + 2
      if (u < nblk) {
        const int bi0 = u, bi1 = u + nu;
        const bool two = bi1 < nblk;
        const double* b_ = A + (o2 + fr) * kLd + fc;
        const double* a0_ = A + (o2 + 8 * bi0 + fr) * kLd + fc;
        const double* a1_ = A + (o2 + 8 * (two ? bi1 : bi0) + fr) * kLd + fc;
        // four independent accumulator chains (2 blocks x the two K halves of a panel), operands one panel ahead
        double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0, d00 = 0.0, d01 = 0.0, d10 = 0.0, d11 = 0.0;
        const int K = o + 8;
        double b0 = b_[0], b1 = b_[4], x00 = a0_[0], x01 = a0_[4], x10 = a1_[0], x11 = a1_[4];
#pragma unroll 1
        for (int k = 8; k < K; k += 8) {
          const double nb0 = b_[k], nb1 = b_[k + 4], nx00 = a0_[k], nx01 = a0_[k + 4], nx10 = a1_[k], nx11 = a1_[k + 4];
          dmma(c00, c01, x00, b0);
          dmma(c10, c11, x10, b0);
          dmma(d00, d01, x01, b1);
          dmma(d10, d11, x11, b1);
          b0 = nb0; b1 = nb1; x00 = nx00; x01 = nx01; x10 = nx10; x11 = nx11;
        }
        dmma(c00, c01, x00, b0);
        dmma(c10, c11, x10, b0);
        dmma(d00, d01, x01, b1);
        dmma(d10, d11, x11, b1);
        c00 += d00; c01 += d01; c10 += d10; c11 += d11;
        double2* p0 = reinterpret_cast<double2*>(A + (o2 + 8 * bi0 + fr) * kLd + o2 + 2 * fc);
        const double2 v0 = *p0;
        *p0 = make_double2(v0.x - c00, v0.y - c01);
        if (two) {
          double2* p1 = reinterpret_cast<double2*>(A + (o2 + 8 * bi1 + fr) * kLd + o2 + 2 * fc);
          const double2 v1 = *p1;
          *p1 = make_double2(v1.x - c10, v1.y - c11);
        }
      }
      if (warp == 7 && lane == 0) steps[48 + I] = clock64();
    } else if (!(V & 2) && warp >= first_update) {
      const int u = warp - first_update, nu = 8 - first_update;
#pragma unroll 1
      for (int q = u; u >= 0 && 1 + q <= nb - 1 - q; q += nu)
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
          const int r1 = 1 + q, r2 = nb - 1 - q;
          const int bi = pass == 0 ? r1 : r2;
          if (bi >= nb || (pass == 1 && r2 <= r1) || (pass == 0 && r1 > r2)) continue;
          const double* ap = Pm + (8 * bi + fr) * kLd + fc;
          const double a0 = ap[0], a1 = ap[4];
          double* crow = C22 + (8 * bi + fr) * kLd + 2 * fc;
          if (V & 32) {
            int stamp = 64 + 8 * pass;
            if ((V & 64) && I == 0 && warp == 7 && lane == 0) steps[stamp++] = clock64();
#pragma unroll 1
            for (int bj0 = 1; bj0 <= bi; bj0 += 4) {
              double b0[4], b1[4], c0[4], c1[4];
              double2 old[4];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const int bj = min(bj0 + t, bi);
                const double* bp = Pm + (8 * bj + fr) * kLd + fc;
                b0[t] = bp[0];
                b1[t] = bp[4];
                old[t] = *reinterpret_cast<const double2*>(crow + 8 * bj);
                c0[t] = c1[t] = 0.0;
              }
#pragma unroll
              for (int t = 0; t < 4; ++t) dmma(c0[t], c1[t], a0, b0[t]);
#pragma unroll
              for (int t = 0; t < 4; ++t) dmma(c0[t], c1[t], a1, b1[t]);
#pragma unroll
              for (int t = 0; t < 4; ++t)
                if (bj0 + t <= bi) *reinterpret_cast<double2*>(crow + 8 * (bj0 + t)) = make_double2(old[t].x - c0[t], old[t].y - c1[t]);
              if ((V & 64) && I == 0 && warp == 7 && lane == 0) steps[stamp++] = clock64();
            }
            continue;
          }
#pragma unroll 4
          for (int bj = 1; bj <= bi; ++bj) {
            const double* bp = Pm + (8 * bj + fr) * kLd + fc;
            double c0 = 0.0, c1 = 0.0;
            dmma(c0, c1, a0, bp[0]);
            dmma(c0, c1, a1, bp[4]);
            double2* cp = reinterpret_cast<double2*>(crow + 8 * bj);
            const double2 old = *cp;
            *cp = make_double2(old.x - c0, old.y - c1);
          }
        }
      if (warp == 7 && lane == 0) steps[48 + I] = clock64();
    }
    __syncthreads();
  }
  if (tid == 0) steps[11] = clock64();
}

template <int V>
__global__ void __launch_bounds__(kDagThreads, 1) ablate(const double* tile, double* out, long long* steps, int* info) {
  extern __shared__ __align__(16) double smem[];
  double* A = smem;
  double* rdiag = A + 3 * kTile * kLd;
  for (int e = threadIdx.x; e < kTile * kTile; e += kDagThreads) A[(e / kTile) * kLd + e % kTile] = tile[e];
  __syncthreads();
  factor_tile_v<V>(A, rdiag, info, steps);
  __syncthreads();
  for (int e = threadIdx.x; e < kTile * kTile; e += kDagThreads) out[e] = A[(e / kTile) * kLd + e % kTile];
}

std::vector<double> g_reference;

template <int V>
void run(const double* d_tile, double* d_out, long long* d_steps, int* d_info, const char* what) {
  cudaFuncSetAttribute(ablate<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDagSmem);
  long long h[96];
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(d_steps, 0, sizeof(h));
    ablate<V><<<1, kDagThreads, kDagSmem>>>(d_tile, d_out, d_steps, d_info);
    cudaMemcpy(h, d_steps, sizeof(h), cudaMemcpyDeviceToHost);
  }
  cudaError_t e = cudaGetLastError();
  {
    std::vector<double> out(kTile * kTile);
    cudaMemcpy(out.data(), d_out, out.size() * 8, cudaMemcpyDeviceToHost);
    if (V == 0) g_reference = out;
    double worst = 0.0;
    for (int i = 0; i < kTile; ++i)
      for (int j = 0; j <= i; ++j) worst = fmax(worst, fabs(out[i * kTile + j] - g_reference[i * kTile + j]));
    if (!(V & 7)) printf("[max |L - L(as shipped)| = %.2e] ", worst);
  }
  printf("%-58s total %6lld |", what, h[11] - h[0]);
  for (int I = 0; I < 11; ++I) printf(" %4lld", (I < 10 ? h[I + 1] : h[11]) - h[I]);
  printf(" | phase1");
  for (int I = 0; I < 11; ++I) printf(" %4lld", h[16 + I] - h[I]);
  printf(" | panel(w0)");
  for (int I = 0; I < 11; ++I) printf(" %4lld", h[32 + I] ? h[32 + I] - h[16 + I] : 0);
  printf(" | update(w7)");
  for (int I = 0; I < 11; ++I) printf(" %4lld", h[48 + I] ? h[48 + I] - h[16 + I] : 0);
  if (V & 64) {
    printf(" | warp 7, step 0, stamps relative to the phase-1 barrier:");
    for (int i = 64; i < 80; ++i) printf(" %lld", h[i] ? h[i] - h[16] : 0);
  }
  printf("%s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
}

}  // namespace
}  // namespace rsba

int main() {
  using namespace rsba;
  std::vector<double> t(kTile * kTile);
  for (int i = 0; i < kTile; ++i)
    for (int j = 0; j < kTile; ++j) t[i * kTile + j] = (i == j ? 200.0 : 0.0) + 1.0 / (1 + (i > j ? i - j : j - i));
  double *d_tile, *d_out; long long* d_steps; int* d_info;
  cudaMalloc(&d_tile, t.size() * 8); cudaMalloc(&d_out, t.size() * 8); cudaMalloc(&d_steps, 96 * 8); cudaMalloc(&d_info, 4);
  cudaMemcpy(d_tile, t.data(), t.size() * 8, cudaMemcpyHostToDevice);
  cudaMemset(d_info, 0, 4);
  run<0>(d_tile, d_out, d_steps, d_info, "as shipped");
  run<1>(d_tile, d_out, d_steps, d_info, "no panel");
  run<2>(d_tile, d_out, d_steps, d_info, "no phase-2 update");
  run<4>(d_tile, d_out, d_steps, d_info, "no phase-1 update");
  run<6>(d_tile, d_out, d_steps, d_info, "panel only");
  run<7>(d_tile, d_out, d_steps, d_info, "barriers only");
  run<8 + 6>(d_tile, d_out, d_steps, d_info, "panel only, one warp");
  run<16>(d_tile, d_out, d_steps, d_info, "update on warps 4-7 only");
  run<3>(d_tile, d_out, d_steps, d_info, "phase 1 only");
  run<5>(d_tile, d_out, d_steps, d_info, "phase-2 update only");
  run<32>(d_tile, d_out, d_steps, d_info, "batched update");
  run<32 + 5>(d_tile, d_out, d_steps, d_info, "batched phase-2 update only");
  run<128>(d_tile, d_out, d_steps, d_info, "look-ahead update");
  run<128 + 5>(d_tile, d_out, d_steps, d_info, "look-ahead update only");
  run<64 + 32 + 5>(d_tile, d_out, d_steps, d_info, "batched phase-2 update only, stamps");
  run<32 + 16>(d_tile, d_out, d_steps, d_info, "batched update on warps 4-7 only");
  return 0;
}
