#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
// per-SM throughput of exp2 variants: all 4 sub-partitions busy, 8 warps per SM like the attention kernel's softmax role
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(float* out, float seed, int iters) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = seed * (i + 1) + threadIdx.x * 1e-6f;
  float acc = 0.f;
  // loop-carried state for the "pure" modes (ptxas would hoist a loop-invariant ex2 chain out of the loop)
  unsigned hs[8];
  float ys[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (MODE == 4) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hs[i]) : "f"(x[2 * i + 1]), "f"(x[2 * i]));
    else asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hs[i]) : "f"(x[2 * i + 1]), "f"(x[2 * i]));
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) ys[i] = x[i];
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {          // 16 x ex2.approx.ftz.f32
#pragma unroll
      for (int i = 0; i < 16; ++i) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[i])); acc += y; x[i] = x[i] * 0.999f; }
    } else if (MODE == 1) {   // 8 x (cvt.f16x2.f32 + ex2.approx.f16x2) for the same 16 values, sum in f16x2 then f32
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        unsigned h, e;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x[i + 1]), "f"(x[i]));
        asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(e) : "r"(h));
        __half2 hh = *reinterpret_cast<__half2*>(&e);
        float2 f = __half22float2(hh);
        acc += f.x + f.y; x[i] *= 0.999f; x[i + 1] *= 0.999f;
      }
    } else if (MODE == 2) {   // ex2 only, f16x2, minimal other work (pure MUFU rate)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(hs[i]) : "r"(hs[i]));
      hs[it & 7] ^= 0x04000400u;   // keep the chain data-dependent on the iteration
    } else if (MODE == 3) {   // pure f32 MUFU rate: 4 rounds of 16 dependent-free ex2
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(ys[i]) : "f"(ys[i]));
      ys[it & 15] = -ys[it & 15];
    } else if (MODE == 4) {   // bf16x2 MUFU
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(hs[i]) : "r"(hs[i]));
      hs[it & 7] ^= 0x00800080u;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += __uint_as_float(hs[i]);
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += ys[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE>
static void run(const char* name, int per_iter) {
  float* d; cudaMalloc(&d, 148 * 256 * 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 20000;
  k<MODE><<<148, 256>>>(d, -0.01f, 100);
  cudaEventRecord(a);
  k<MODE><<<148, 256>>>(d, -0.01f, iters);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double elems = (double)iters * per_iter * 256;   // per SM
  printf("%-28s %8.3f ms  %.2f exp/ns/SM  (=%.1f exp/clk/SM at %d MHz nominal)\n", name, ms, elems / (ms * 1e6), elems / (ms * 1e-3 * clk * 1e3), clk / 1000);
}
int main() {
  run<3>("f32 ex2 pure", 64);
  run<2>("f16x2 ex2 pure", 64);
  run<4>("bf16x2 ex2 pure", 64);
  run<0>("f32 ex2 + fadd + fmul", 16);
  run<1>("f16x2 ex2 + cvt + sum", 16);
  return 0;
}
