// Issue-to-completion time of short tcgen05.mma sequences (kind::f16, M = 128, K = 16 per instruction) on one SM:
// R back-to-back MMAs accumulating into one TMEM tile, then tcgen05.commit -> mbarrier; cycles from the first issue to the
// barrier flip, for N = 64 / 128 / 256, operands smem x smem (SS) and TMEM x smem (TS).  This is what bounds an attention
// tile (4-8 Q K^T MMAs + 8 P V MMAs).  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../videollama2_b200/csrc -I../include
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace vl2;

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) k(long long* out, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];   // A: 128 x 64 (16 KB, SW128 K-major), B: 256 x 64 (32 KB)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    constexpr uint32_t hi = umma_desc_sw128_hi(1024);
    const uint32_t a_lo = umma_desc_sw128_lo(smem_u32(smem), 16), b_lo = umma_desc_sw128_lo(smem_u32(smem + 16384), 16);
    if (leader) {
      for (int rep = 0; rep < 6; ++rep) {
        const int R = 1 << rep;   // 1, 2, 4, 8, 16, 32 MMAs
        long long best = 1 << 30;
        for (int t = 0; t < reps; ++t) {
          const uint32_t parity = (rep * reps + t) & 1;
          const long long t0 = clock64();
          for (int i = 0; i < R; ++i) {
            const uint32_t off = ((i & 3) * 32) >> 4;
            if (TS) umma_bf16_ts_lohi(tmem, tmem + 256 + (i & 3) * 8, b_lo + off, hi, idesc, i != 0);
            else umma_bf16_ss_lohi(tmem, a_lo + off, hi, b_lo + off, hi, idesc, i != 0);
          }
          umma_commit(&bar);
          const long long t1 = clock64();
          mbar_wait(&bar, parity);
          const long long t2 = clock64();
          if (t2 - t0 < best) { best = t2 - t0; out[rep * 2 + 1] = t1 - t0; }
        }
        out[rep * 2] = best;
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tmem, 512); }
}

template <int N, bool TS>
static void run(const char* name) {
  long long* d; cudaMalloc(&d, 16 * 8);
  cudaFuncSetAttribute(k<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  k<N, TS><<<1, 128, 64 * 1024>>>(d, 20);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h[12]; cudaMemcpy(h, d, 96, cudaMemcpyDeviceToHost);
  printf("%-22s rated %3d clk/MMA | issue->done (issue only) for R = 1,2,4,8,16,32 MMAs:", name, N / 2);
  for (int i = 0; i < 6; ++i) printf("  %lld (%lld)", h[2 * i], h[2 * i + 1]);
  printf("  => marginal %.0f clk/MMA\n", (double)(h[10] - h[6]) / 24.0);
}
int main() {
  run<64, false>("SS 128x64x16");
  run<128, false>("SS 128x128x16");
  run<256, false>("SS 128x256x16");
  run<64, true>("TS 128x64x16");
  run<128, true>("TS 128x128x16");
  return 0;
}
