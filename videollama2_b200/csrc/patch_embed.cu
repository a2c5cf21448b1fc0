// ViT patch embedding as ONE implicit-GEMM kernel on tcgen05 (HF CLIPVisionEmbeddings.forward, HF:clip/modeling_clip.py:202-218,
// followed by CLIPVisionTransformer.pre_layrnorm :739-741; HF SiglipVisionEmbeddings for the VideoLLaMA2.1 tower):
//
//   tok[f, 1+p, :] = LN( conv_PxP_strideP(pixels[f])[p, :] + pos[1+p, :] ),   tok[f, 0, :] = LN( cls + pos[0, :] )     (CLIP)
//   tok[f, p, :]   =     conv_PxP_strideP(pixels[f])[p, :] + bias + pos[p, :]                                          (SigLIP)
//
// No im2col matrix exists: the A operand (128 patches x K = 3 P^2 pixel values) is gathered from the NCHW frames by 128
// producer threads (one patch each; the 28-byte pixel runs of a 14 x 14 patch break TMA's 16-byte rule, so this is plain
// LDG) straight into the 128-byte-swizzled shared-memory layout tcgen05.mma reads, ONCE per 128-patch tile and for the
// whole K (10 k-blocks = 160 KB).  The conv weight [C, Kpad] streams through a 2-stage TMA ring, one 256-column N tile at a
// time, with two TMEM accumulator stages so the epilogue of N tile j overlaps the MMAs of N tile j+1.
// Epilogue (thread = patch row): + position row (+ bias); CLIP: every row's sum / sum of squares accumulate in the row's
// thread across the N tiles while the un-normalised fp32 row goes to an L2-resident scratch; after the last N tile each warp
// normalises its 32 rows (lanes across columns, coalesced) and writes the bf16 tokens.  One launch replaces
// vl2_patch_im2col + vl2_gemm_bf16 + vl2_clip_embed_finish.
#include "host_common.h"
#include "ptx.cuh"
#include "row_common.cuh"

namespace vl2 {

static constexpr int kPeBM = 128, kPeBN = 256, kPeBK = 64;
static constexpr int kPeMaxKB = 10;                           // K padded to at most 640 (3 * 14 * 14 = 588)
static constexpr int kPeThreads = 320;                        // 4 gather warps, TMA warp, MMA warp, 4 epilogue warps
static constexpr int kPeStageB = kPeBN * kPeBK * 2;           // 32 KB
static constexpr int kPeSmem = kPeMaxKB * kPeBM * kPeBK * 2 + 2 * kPeStageB + 512;

struct PatchEmbedParams {
  const __nv_bfloat16* px;      // [F, 3, H, W]
  const __nv_bfloat16* pos;     // [np + has_cls, C]
  const __nv_bfloat16* cls;     // [C] or NULL
  const __nv_bfloat16* gamma;   // pre-LN (CLIP) or NULL
  const __nv_bfloat16* beta;
  const float* bias;            // conv bias (SigLIP) or NULL
  __nv_bfloat16* out;           // [F * (np + has_cls), C]
  float* scratch;               // [F * np, C] fp32 (CLIP mode)
  int F, H, W, P, G, np, C, K, nkb, n_tiles, M;
  float eps;
};

__global__ void __launch_bounds__(kPeThreads, 1)
patch_embed_kernel(const __grid_constant__ CUtensorMap tmap_w, const PatchEmbedParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) { asm volatile("trap;"); }
  uint8_t* sA = smem;                                              // [nkb][128 rows][128 B], swizzled
  uint8_t* sB = smem + kPeMaxKB * kPeBM * kPeBK * 2;               // [2][256 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 2 * kPeStageB);
  uint64_t* a_full = bars;            // [10]  gather -> MMA (128 arrivals)
  uint64_t* b_full = bars + 10;       // [2]   TMA -> MMA
  uint64_t* b_empty = bars + 12;      // [2]   MMA -> TMA
  uint64_t* t_full = bars + 14;       // [2]   MMA -> epilogue
  uint64_t* t_empty = bars + 16;      // [2]   epilogue -> MMA (128 arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kPeBM;
  const bool clip = p.gamma != nullptr;
  const int has_cls = p.cls != nullptr ? 1 : 0;

  pdl_launch_dependents();
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    for (int i = 0; i < kPeMaxKB; ++i) mbar_init(&a_full[i], 128);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 128);
    }
    fence_barrier_init();
  }
  pdl_wait();
  if (warp == 5) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < 4) {
    // ===================== A gather: thread = patch row, k = (c, i, j) runs of P contiguous pixels =====================
    const int r = threadIdx.x;
    const int m = m0 + r;
    const bool valid = m < p.M;
    const int f = valid ? m / p.np : 0, pp = valid ? m % p.np : 0;
    const int ph = pp / p.G, pw = pp % p.G;
    const __nv_bfloat16* src0 = p.px + ((int64_t)f * 3 * p.H + (int64_t)ph * p.P) * p.W + pw * p.P;   // (c = 0, i = 0, j = 0)
    const int64_t cstride = (int64_t)p.H * p.W;
    int c = 0, i = 0, j = 0;                       // decomposition of the running k (advances by 2: P is even)
    const uint32_t row_base = smem_u32(sA) + r * 128;
    for (int kb = 0; kb < p.nkb; ++kb) {
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint32_t w4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = kb * 64 + ch * 8 + e * 2;
          uint32_t v = 0;
          if (valid && k < p.K) v = *reinterpret_cast<const uint32_t*>(src0 + c * cstride + (int64_t)i * p.W + j);
          w4[e] = v;
          j += 2;
          if (j == p.P) { j = 0; if (++i == p.P) { i = 0; ++c; } }
        }
        sts128(row_base + kb * (kPeBM * 128) + ((ch ^ (r & 7)) << 4), make_uint4(w4[0], w4[1], w4[2], w4[3]));
      }
      fence_proxy_async_smem();       // generic-proxy smem writes -> visible to the tensor core's async proxy
      mbar_arrive(&a_full[kb]);
    }
  } else if (warp == 4) {
    if (lane == 0) {
      // ===================== TMA: conv weight tiles [256 output channels x 64 k] =====================
      int g = 0;
      for (int jn = 0; jn < p.n_tiles; ++jn)
        for (int kb = 0; kb < p.nkb; ++kb, ++g) {
          const int st = g & 1;
          mbar_wait(&b_empty[st], ((g >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&b_full[st], kPeStageB);
          tma_load_2d(sB + st * kPeStageB, &tmap_w, &b_full[st], kb * kPeBK, jn * kPeBN);
        }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = umma_idesc_bf16(kPeBM, kPeBN, 0, 0);
      int g = 0;
      for (int jn = 0; jn < p.n_tiles; ++jn) {
        const int as = jn & 1;
        mbar_wait(&t_empty[as], ((jn >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * kPeBN;
        for (int kb = 0; kb < p.nkb; ++kb, ++g) {
          const int st = g & 1;
          if (jn == 0) mbar_wait(&a_full[kb], 0);
          mbar_wait(&b_full[st], (g >> 1) & 1);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(sA) + kb * (kPeBM * 128);
          const uint32_t b_addr = smem_u32(sB) + st * kPeStageB;
#pragma unroll
          for (int k = 0; k < kPeBK / 16; ++k)
            umma_bf16_ss(d_tmem, umma_desc_sw128(a_addr + k * 32, 16, 1024), umma_desc_sw128(b_addr + k * 32, 16, 1024), idesc,
                         (kb | k) != 0);
          umma_commit(&b_empty[st]);
        }
        umma_commit(&t_full[as]);
      }
    }
  } else {
    // ===================== epilogue: thread = patch row (TMEM lane) =====================
    const int q = warp & 3;                            // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;
    const int m = m0 + r;
    const bool valid = m < p.M;
    const int f = valid ? m / p.np : 0, pp = valid ? m % p.np : 0;
    const int64_t out_row = (int64_t)f * (p.np + has_cls) + has_cls + pp;
    const __nv_bfloat16* pos_row = p.pos + (int64_t)(has_cls + pp) * p.C;
    float s = 0.f, sq = 0.f;
    for (int jn = 0; jn < p.n_tiles; ++jn) {
      const int as = jn & 1;
      mbar_wait(&t_full[as], (jn >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + as * kPeBN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int ck = 0; ck < kPeBN / 32; ++ck) {
        const int col0 = jn * kPeBN + ck * 32;
        if (col0 >= p.C) break;                        // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32(taddr + ck * 32, v);
        uint4 pv[4];
        if (valid) {
#pragma unroll
          for (int g = 0; g < 4; ++g) pv[g] = __ldg(reinterpret_cast<const uint4*>(pos_row + col0) + g);
        }
        tmem_ld_wait();
        if (!valid) continue;
        float x[32];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float pf[8];
          unpack8(pv[g], pf);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float a = __uint_as_float(v[g * 8 + e]);
            // CLIP: the conv output is rounded to bf16 (HF's conv returns bf16) before the fp32 position add + LayerNorm;
            // SigLIP: conv + bias + position in fp32, one rounding at the store
            x[g * 8 + e] = clip ? __bfloat162float(__float2bfloat16_rn(a)) + pf[e]
                                : a + (p.bias != nullptr ? __ldg(p.bias + col0 + g * 8 + e) : 0.f) + pf[e];
          }
        }
        if (clip) {
          float* srow = p.scratch + (int64_t)m * p.C + col0;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            *reinterpret_cast<float4*>(srow + g * 4) = make_float4(x[g * 4], x[g * 4 + 1], x[g * 4 + 2], x[g * 4 + 3]);
            s += (x[g * 4] + x[g * 4 + 1]) + (x[g * 4 + 2] + x[g * 4 + 3]);
            sq = fmaf(x[g * 4], x[g * 4], fmaf(x[g * 4 + 1], x[g * 4 + 1], fmaf(x[g * 4 + 2], x[g * 4 + 2], fmaf(x[g * 4 + 3], x[g * 4 + 3], sq))));
          }
        } else {
          __nv_bfloat16* orow = p.out + out_row * p.C + col0;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float o8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o8[e] = x[g * 8 + e];
            *reinterpret_cast<uint4*>(orow + g * 8) = pack8(o8);
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&t_empty[as]);
    }
    if (clip) {
      // pre-LN: this thread holds its row's statistics; the warp now walks its 32 rows with lanes across columns
      const float mean = s / (float)p.C;
      const float rstd = rsqrtf(fmaxf(sq / (float)p.C - mean * mean, 0.f) + p.eps);
      __syncwarp();                                    // the rows' scratch writes of every lane are visible to the warp
      for (int rr = 0; rr < 32; ++rr) {
        const float mu = __shfl_sync(0xffffffffu, mean, rr), rs = __shfl_sync(0xffffffffu, rstd, rr);
        const int mm = m0 + q * 32 + rr;
        if (mm >= p.M) break;                          // warp-uniform
        const int ff = mm / p.np, pq = mm % p.np;
        const float* srow = p.scratch + (int64_t)mm * p.C;
        __nv_bfloat16* orow = p.out + ((int64_t)ff * (p.np + 1) + 1 + pq) * p.C;
        for (int c0 = lane * 8; c0 < p.C; c0 += 256) {
          const float4 a = __ldcg(reinterpret_cast<const float4*>(srow + c0));
          const float4 b = __ldcg(reinterpret_cast<const float4*>(srow + c0 + 4));
          float gm[8], bt[8], o8[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(p.gamma + c0)), gm);
          unpack8(__ldg(reinterpret_cast<const uint4*>(p.beta + c0)), bt);
          const float xs[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) o8[e] = (xs[e] - mu) * rs * gm[e] + bt[e];
          *reinterpret_cast<uint4*>(orow + c0) = pack8(o8);
        }
      }
      // the class-token row is the same for every frame: LN(cls + pos[0]); block 0's first epilogue warp writes all F copies
      if (blockIdx.x == 0 && q == 0) {
        float cs = 0.f, cq = 0.f;
        for (int c0 = lane * 8; c0 < p.C; c0 += 256) {
          float a[8], b[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(p.cls + c0)), a);
          unpack8(__ldg(reinterpret_cast<const uint4*>(p.pos + c0)), b);
#pragma unroll
          for (int e = 0; e < 8; ++e) { const float xv = a[e] + b[e]; cs += xv; cq = fmaf(xv, xv, cq); }
        }
        cs = warp_sum(cs);
        cq = warp_sum(cq);
        const float mu = cs / (float)p.C;
        // two-pass variance for the single class row (matches the stand-alone kernel's centred form closely enough: fp32)
        const float rs = rsqrtf(fmaxf(cq / (float)p.C - mu * mu, 0.f) + p.eps);
        for (int c0 = lane * 8; c0 < p.C; c0 += 256) {
          float a[8], b[8], gm[8], bt[8], o8[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(p.cls + c0)), a);
          unpack8(__ldg(reinterpret_cast<const uint4*>(p.pos + c0)), b);
          unpack8(__ldg(reinterpret_cast<const uint4*>(p.gamma + c0)), gm);
          unpack8(__ldg(reinterpret_cast<const uint4*>(p.beta + c0)), bt);
#pragma unroll
          for (int e = 0; e < 8; ++e) o8[e] = (a[e] + b[e] - mu) * rs * gm[e] + bt[e];
          const uint4 pk = pack8(o8);
          for (int ff = 0; ff < p.F; ++ff) *reinterpret_cast<uint4*>(p.out + (int64_t)ff * (p.np + 1) * p.C + c0) = pk;
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace vl2

using namespace vl2;

extern "C" int vl2_patch_embed(const vl2_patch_embed_args* a, void* stream) {
  VL2_REQUIRE(a != nullptr, VL2_E_BADSHAPE, "vl2_patch_embed: null args");
  VL2_REQUIRE(a->F > 0 && a->P > 0 && a->P % 2 == 0 && a->H >= a->P && a->W >= a->P && a->W % 2 == 0, VL2_E_BADSHAPE,
              "vl2_patch_embed: need F > 0, even patch size, even W, H/W >= P (F=%d P=%d H=%d W=%d)", a->F, a->P, a->H, a->W);
  const int K = 3 * a->P * a->P;
  VL2_REQUIRE(a->Kpad % 64 == 0 && a->Kpad >= K && a->Kpad <= kPeMaxKB * 64, VL2_E_UNSUPPORTED,
              "vl2_patch_embed: Kpad must be a multiple of 64 in [3*P*P, %d] (Kpad=%d)", kPeMaxKB * 64, a->Kpad);
  VL2_REQUIRE(a->C > 0 && a->C % 32 == 0, VL2_E_BADSHAPE, "vl2_patch_embed: C %% 32 == 0 required (C=%d)", a->C);
  VL2_REQUIRE(a->pixels && a->weight && a->pos && a->out, VL2_E_BADSHAPE, "vl2_patch_embed: null pointer");
  const bool clip = a->gamma != nullptr;
  VL2_REQUIRE(!clip || (a->beta != nullptr && a->cls != nullptr && a->scratch != nullptr), VL2_E_BADSHAPE,
              "vl2_patch_embed: the CLIP form needs cls, gamma, beta and an fp32 scratch [F*np, C]");
  VL2_REQUIRE(clip || a->cls == nullptr, VL2_E_UNSUPPORTED, "vl2_patch_embed: a class token without pre-LN is not a tower of the path");
  VL2_REQUIRE(aligned16(a->pixels) && aligned16(a->weight) && aligned16(a->pos) && aligned16(a->out) && aligned16(a->scratch) &&
                  aligned16(a->gamma) && aligned16(a->beta) && aligned16(a->cls),
              VL2_E_BADALIGN, "vl2_patch_embed: pointers must be 16-byte aligned");
  CUtensorMap tw;
  {
    uint64_t dims[2] = {(uint64_t)a->Kpad, (uint64_t)a->C};
    uint64_t str[1] = {(uint64_t)a->Kpad * 2};
    uint32_t box[2] = {kPeBK, kPeBN};
    int rc = make_tmap_bf16(&tw, a->weight, 2, dims, str, box);
    if (rc) return rc;
  }
  PatchEmbedParams p;
  p.px = (const __nv_bfloat16*)a->pixels; p.pos = (const __nv_bfloat16*)a->pos; p.cls = (const __nv_bfloat16*)a->cls;
  p.gamma = (const __nv_bfloat16*)a->gamma; p.beta = (const __nv_bfloat16*)a->beta; p.bias = a->bias;
  p.out = (__nv_bfloat16*)a->out; p.scratch = a->scratch;
  p.F = a->F; p.H = a->H; p.W = a->W; p.P = a->P; p.G = a->W / a->P; p.np = (a->H / a->P) * (a->W / a->P); p.C = a->C;
  p.K = K; p.nkb = a->Kpad / 64; p.n_tiles = (a->C + kPeBN - 1) / kPeBN; p.M = a->F * p.np; p.eps = a->eps;
  VL2_SMEM_OPT_IN(patch_embed_kernel, kPeSmem);
  const int blocks = (p.M + kPeBM - 1) / kPeBM;
  VL2_CHECK_CUDA(launch_kernel(patch_embed_kernel, dim3(blocks), dim3(kPeThreads), kPeSmem, (cudaStream_t)stream, 1, tw, p));
  VL2_CHECK_LAUNCH("patch_embed_kernel");
  return VL2_OK;
}
