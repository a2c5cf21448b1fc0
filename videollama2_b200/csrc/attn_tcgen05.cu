// FlashAttention-style fused attention for sm_100a (non-causal ViT heads and causal GQA decoder heads).
//
// Work item = one 128-row query tile of one (batch, head).  320 threads:
//   warps 0..7  softmax / output warps: TWO threads per query row (warps w and w+4 share TMEM lane quarter w & 3 and
//               take key columns [0,64) / [64,128) of every S tile), so each SM sub-partition has two warps to overlap
//               TMEM loads, MUFU exp2 and packing; the half-row maxima are exchanged through smem once per tile
//   warp 8      TMA producers: lane 0 loads Q and the K ring (2 stages, freed right after Q K^T), lane 1 the V ring
//               (2 stages, freed after P V)
//   warp 9      TMEM allocator + MMA issuer (one ELECTED lane, descriptors as (lo, hi) words: back-to-back UTCHMMAs):
//                 S_j  = Q K_j^T      tcgen05.mma 128x128x16, A/B K-major SW128 from smem, accumulator in TMEM (2 buffers)
//                 O   += P_j V_j      tcgen05.mma 128xDx16, B = V tile as MN-major SW128 operand (no transpose pass)
// Online softmax state (m, l) lives in registers of the row's thread; O accumulates in TMEM over all key tiles and is
// rescaled in place only when a row maximum grows by more than 2^8 (lazy rescale).  S is read from TMEM once per tile.
//
// Two kernels:
//   attn_fwd_persistent_kernel (default): one CTA per SM loops over work items.  P_j goes to TENSOR MEMORY (tcgen05.st,
//     own columns) and P V reads its A operand from there - no shared-memory write / read for P; Q K^T(j+2) is issued
//     ahead of P V(j); the item's output tile is staged in smem and leaves through ONE TMA store per 64-column atom;
//     the two threads of a row synchronise through a 64-thread pair barrier; unmasked and masked key tiles are separate
//     loops.  Cycle traces (vl2_debug_attn_trace / _timeline, tools/attn_trace.py) are their own instantiation.
//   attn_fwd_kernel (VL2_ATTN_PERSISTENT=0, the round-1 form kept UNCHANGED for A/B runs): one CTA per item, `lane == 0`
//     MMA issuer with descriptors rebuilt per MMA, P through a double-buffered smem tile, per-thread 16-byte output stores.
#include <stdlib.h>

#include <type_traits>

#include "host_common.h"
#include "ptx.cuh"

namespace vl2 {

static constexpr int kAttnThreads = 320;
static constexpr int kSoftmaxThreads = 256;
static constexpr int BQ = 128;
static constexpr int BKV = 128;

template <int D>
struct AttnCfg {
  static constexpr int kAtoms = D / 64;                 // 64-column (128-byte) swizzle atoms per row
  static constexpr int kTileBytes = BKV * D * 2;        // one Q / K / V tile
  static constexpr int kAtomBytes = 128 * 128;          // 128 rows x 128 B
  static constexpr int kPBytes = BQ * BKV * 2;          // 32 KB
  static constexpr int kOffQ = 0;
  static constexpr int kOffK = kTileBytes;              // 2 stages, released as soon as Q K^T of the tile has run
  static constexpr int kOffV = 3 * kTileBytes;          // 2 stages, released after P V of the tile
  static constexpr int kOffP = 5 * kTileBytes;          // 2 buffers: softmax(j) never waits for P V(j-1)
  static constexpr int kOffBar = kOffP + 2 * kPBytes;
  static constexpr int kOffMax = kOffBar + 256;         // float [2 parities][2 halves][128 rows]
  static constexpr int kSmemUsed = kOffMax + 2 * 2 * 128 * 4;
  // at least 116 KB so that only one CTA is resident per SM (each CTA allocates all 512 TMEM columns)
  static constexpr int kSmemBytes = (kSmemUsed > 116 * 1024) ? kSmemUsed : 116 * 1024;
  static constexpr int kTmemCols = 512;
  static constexpr int kColS0 = 0, kColS1 = 128, kColO = 256;
  static constexpr int kColP0 = 384, kColP1 = 448;      // persistent kernel: P_j as 64 columns of 16-bit pairs (A operand of P V)
};

// Debug cycle trace (vl2_attn_args.reserved == 777): block (0,0,0), softmax thread 0 accumulates the cycles it spends
// in each phase of the key-tile loop; read back with vl2_debug_attn_trace().
__device__ long long g_attn_trace[16];
// Timeline trace (vl2_attn_args.reserved == 779, persistent kernel): clock64 stamps of CTA 0's SECOND work item, per key
// tile j: softmax thread 0 at [8j + 0..6] (tile start, S arrived, S in registers, maxima exchanged, P buffer / O free,
// P stored, p_full arrive) and [120..123] (item epilogue: start, l exchanged, last P V complete, O stored);
// MMA thread at [128 + 8j + 0..5] (before Q K^T(j+1), its K arrived, issued, P(j) arrived, V(j) arrived, P V(j) issued),
// [250..251] (Q arrived, Q K^T(0) issued); K loader at [256 + j] and V loader at [288 + j] (stage free, load issued).
__device__ long long g_attn_tl[320];
#define VL2_TL(cond, slot) do { if (cond) g_attn_tl[slot] = clock64(); } while (0)
static constexpr int kTlItem = 1;

struct AttnParams {
  int trace;
  void* out;
  int64_t ldo;
  int S, Hq, group;  // group = Hq / Hkv
  int d_true;        // real head width (<= D); columns beyond it are TMA zero fill
  int n_batch;       // B (persistent variant: the grid no longer carries it)
  int causal;
  float scale_log2;  // softmax scale * log2(e)
};

template <int D>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  using Cfg = AttnCfg<D>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) { asm volatile("trap;"); }
  uint8_t* sQ = smem + Cfg::kOffQ;
  uint8_t* sK = smem + Cfg::kOffK;
  uint8_t* sV = smem + Cfg::kOffV;
  uint8_t* sP = smem + Cfg::kOffP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* p_full = bars + 11;   // [2]
  uint64_t* o_full = bars + 13;   // [2]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 15);
  float* smax = reinterpret_cast<float*>(smem + Cfg::kOffMax);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // heavy (late) causal tiles first
  // Non-causal: x = query tile, y = head (consecutive CTAs share a head's K/V in L2).  Causal: x = head, y = query-tile
  // rank, heaviest (latest) tile first: the hardware hands CTAs out in x-fastest order, so every head's 14-tile CTA is
  // scheduled before any 13-tile one - longest-processing-time-first list scheduling over the SMs instead of one head
  // after the other (and the 4 query heads of a GQA group, adjacent in x, read the same K/V tiles at the same time).
  const int qt = p.causal ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.x;
  const int head = p.causal ? (int)blockIdx.x : (int)blockIdx.y;
  const int b = blockIdx.z;
  const int kvh = head / p.group;
  const int q0 = qt * BQ;
  const int n_kv = p.causal ? (qt + 1) : (p.S + BKV - 1) / BKV;

  pdl_launch_dependents();
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], kSoftmaxThreads);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  pdl_wait();
  if (warp == 9) {
    tmem_alloc(tmem_base_ptr, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 8) {
    // ===================== TMA producers: lane 0 = Q + K ring, lane 1 = V ring (independent, never block each other)
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, Cfg::kTileBytes);
#pragma unroll
      for (int a = 0; a < Cfg::kAtoms; ++a)
        tma_load_4d(sQ + a * Cfg::kAtomBytes, &tmap_q, q_full, a * 64, head, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait(&k_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[st], Cfg::kTileBytes);
#pragma unroll
        for (int a = 0; a < Cfg::kAtoms; ++a)
          tma_load_4d(sK + st * Cfg::kTileBytes + a * Cfg::kAtomBytes, &tmap_k, &k_full[st], a * 64, kvh, j * BKV, b);
      }
    } else if (lane == 1) {
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait(&v_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[st], Cfg::kTileBytes);
#pragma unroll
        for (int a = 0; a < Cfg::kAtoms; ++a)
          tma_load_4d(sV + st * Cfg::kTileBytes + a * Cfg::kAtomBytes, &tmap_v, &v_full[st], a * 64, kvh, j * BKV, b);
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_qk = umma_idesc_bf16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, D, 0, 1);  // B (= V tile) is MN-major
      const uint32_t q_addr = smem_u32(sQ);
      auto issue_qk = [&](int j) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t k_addr = smem_u32(sK + st * Cfg::kTileBytes);
        const uint32_t d_tmem = tmem_base + (st ? Cfg::kColS1 : Cfg::kColS0);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk >> 2) * Cfg::kAtomBytes + (kk & 3) * 32;
          umma_bf16_ss(d_tmem, umma_desc_sw128(q_addr + off, 16, 1024), umma_desc_sw128(k_addr + off, 16, 1024),
                       idesc_qk, kk != 0);
        }
        umma_commit(&s_full[st]);
        umma_commit(&k_empty[st]);   // K_j is free as soon as Q K_j^T has run (the loader can fetch K_{j+2} early)
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        if (j + 1 < n_kv) issue_qk(j + 1);
        mbar_wait(&p_full[st], (j >> 1) & 1);
        mbar_wait(&v_full[st], (j >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t v_addr = smem_u32(sV + st * Cfg::kTileBytes);
        const uint32_t p_addr = smem_u32(sP + st * Cfg::kPBytes);
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          const uint64_t da = umma_desc_sw128(p_addr + (kk >> 2) * Cfg::kAtomBytes + (kk & 3) * 32, 16, 1024);
          // MN-major B: 16 keys = 2 groups of 8 rows (1024 B each); next 64-wide d atom is kAtomBytes away (LBO)
          const uint64_t db = umma_desc_sw128(v_addr + kk * 2048, Cfg::kAtomBytes, 1024);
          umma_bf16_ss(tmem_base + Cfg::kColO, da, db, idesc_pv, (j | kk) != 0);
        }
        umma_commit(&o_full[st]);
        umma_commit(&v_empty[st]);
      }
    }
  } else {
    // ===================== softmax / output warps (two threads per query row) =====================
    // One TMEM pass over S per tile; O accumulates in TMEM across tiles (tcgen05.mma accumulate) and is rescaled
    // in place only when a row maximum grows by more than 2^8 (lazy rescale: probabilities stay <= 256, exact after
    // the final division by l).  Thread (r, hf) owns columns [64 hf, 64 hf + 64) of row r's scores and columns
    // [hf D/2, (hf+1) D/2) of its output; both threads of a row make identical rescale decisions from the exchanged
    // row maximum, so their partial sums l share one scale and are added once at the end.
    const int hf = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    const int qi = q0 + r;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t o_taddr = tmem_base + lane_sel + Cfg::kColO + hf * (D / 2);
    float m = -INFINITY, l = 0.f;
    constexpr float kRescaleThreshold = 8.f;

    const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0;
    long long t0 = 0, acc_t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define VL2_TR(i) do { if (tr) { const long long t1 = clock64(); acc_t[i] += t1 - t0; t0 = t1; } } while (0)
    for (int j = 0; j < n_kv; ++j) {
      if (tr) t0 = clock64();
      const int st = j & 1;
      const uint32_t s_taddr = tmem_base + lane_sel + (st ? Cfg::kColS1 : Cfg::kColS0) + hf * 64;
      const int kv0 = j * BKV + hf * 64;
      const bool need_mask = (j * BKV + BKV > p.S) || (p.causal && (j * BKV + BKV - 1 > q0));
      mbar_wait(&s_full[st], (j >> 1) & 1);
      tc_fence_after_sync();
      VL2_TR(0);   // waiting for S_j
      uint32_t sv[2][32];
#pragma unroll
      for (int c = 0; c < 2; ++c) tmem_ld_32x32(s_taddr + c * 32, sv[c]);
      tmem_ld_wait();
      VL2_TR(1);   // TMEM load
      if (need_mask) {  // warp-uniform; predicated selects, no per-element branches
        const int lim = p.causal ? min(p.S - 1, qi) : (p.S - 1);   // last visible key index for this row
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            sv[c][i] = (kv0 + c * 32 + i <= lim) ? sv[c][i] : 0xff800000u;  // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        mx0 = fmaxf(mx0, __uint_as_float(sv[0][i]));
        mx1 = fmaxf(mx1, __uint_as_float(sv[1][i]));
      }
      // exchange the half-row maxima (double-buffered by tile parity: one named barrier per tile)
      float* sm = smax + st * 256;
      sm[hf * 128 + r] = fmaxf(mx0, mx1);
      asm volatile("bar.sync 2, %0;" ::"n"(kSoftmaxThreads) : "memory");
      VL2_TR(2);   // mask + max + exchange barrier
      const float m_tile = fmaxf(sm[r], sm[128 + r]) * p.scale_log2;
      // reference maximum for this tile: keep the old one unless it is too stale
      float m_use = m, alpha = 1.f;
      const bool grow = (m_tile > m + kRescaleThreshold) || (m == -INFINITY);
      if (grow) {
        m_use = (m_tile == -INFINITY) ? 0.f : m_tile;   // fully masked rows stay finite
        alpha = (m == -INFINITY) ? 0.f : fast_exp2(m - m_use);
      }
      // P is double buffered: buffer (j & 1) was last read by P V(j-2); O (TMEM) may only be rescaled once P V(j-1)
      // is complete
      if (j > 1) mbar_wait(&o_full[st], ((j >> 1) & 1) ^ 1);
      if (j > 0) {
        if (__any_sync(0xffffffffu, grow)) {
          mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1);
          tc_fence_after_sync();
#pragma unroll
          for (int c = 0; c < D / 64; ++c) {
            uint32_t ov[32];
            tmem_ld_32x32(o_taddr + c * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st_32x32(o_taddr + c * 32, ov);
          }
          tmem_st_wait();
        }
      }
      VL2_TR(3);   // waiting for P V(j-2) / rescale
      // probabilities -> smem (bf16, K-major SW128 A operand): this thread's 64 columns are exactly atom `hf`
      float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float pr[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) pr[i] = fast_exp2(fmaf(__uint_as_float(sv[c][i]), p.scale_log2, -m_use));
#pragma unroll
        for (int i = 0; i < 32; i += 4) { rs0 += pr[i]; rs1 += pr[i + 1]; rs2 += pr[i + 2]; rs3 += pr[i + 3]; }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int chunk = c * 4 + g;
          uint8_t* dst = sP + st * Cfg::kPBytes + hf * Cfg::kAtomBytes + r * 128 + ((chunk ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(dst) =
              make_uint4(pack_bf16(pr[g * 8 + 0], pr[g * 8 + 1]), pack_bf16(pr[g * 8 + 2], pr[g * 8 + 3]),
                         pack_bf16(pr[g * 8 + 4], pr[g * 8 + 5]), pack_bf16(pr[g * 8 + 6], pr[g * 8 + 7]));
        }
      }
      l = l * alpha + ((rs0 + rs1) + (rs2 + rs3));
      m = m_use;
      VL2_TR(4);   // exp2 + pack + st.shared
      // make the generic-proxy smem writes visible to the tensor core (async proxy), then signal
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(&p_full[st]);
      VL2_TR(5);   // proxy fence + arrive
    }
    if (tr) {
      for (int i = 0; i < 6; ++i) g_attn_trace[i] = acc_t[i];
      g_attn_trace[6] = 0;
      g_attn_trace[7] = n_kv;
      g_attn_trace[8] = 1;
      g_attn_trace[9] = 0;
    }
#undef VL2_TR
    // combine the two half-row sums, then each thread normalises and stores its half of the output columns
    float* sl = smax;   // reuse: all max exchanges are complete once every thread passed its last bar.sync
    asm volatile("bar.sync 2, %0;" ::"n"(kSoftmaxThreads) : "memory");
    sl[hf * 128 + r] = l;
    asm volatile("bar.sync 2, %0;" ::"n"(kSoftmaxThreads) : "memory");
    const float inv = 1.f / (sl[r] + sl[128 + r]);
    mbar_wait(&o_full[(n_kv - 1) & 1], ((n_kv - 1) >> 1) & 1);
    tc_fence_after_sync();
    __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + ((int64_t)b * p.S + qi) * p.ldo + head * p.d_true + hf * (D / 2);
#pragma unroll
    for (int c = 0; c < D / 64; ++c) {
      uint32_t ov[32];
      tmem_ld_32x32(o_taddr + c * 32, ov);
      tmem_ld_wait();
      if (qi < p.S) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (hf * (D / 2) + c * 32 + g * 8 >= p.d_true) break;   // head_dim < D: the zero-padded columns are not stored
          *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = make_uint4(
              pack_bf16(__uint_as_float(ov[g * 8 + 0]) * inv, __uint_as_float(ov[g * 8 + 1]) * inv),
              pack_bf16(__uint_as_float(ov[g * 8 + 2]) * inv, __uint_as_float(ov[g * 8 + 3]) * inv),
              pack_bf16(__uint_as_float(ov[g * 8 + 4]) * inv, __uint_as_float(ov[g * 8 + 5]) * inv),
              pack_bf16(__uint_as_float(ov[g * 8 + 6]) * inv, __uint_as_float(ov[g * 8 + 7]) * inv));
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Persistent variant (the default; VL2_ATTN_PERSISTENT=0 selects the one-CTA-per-item kernel above): one CTA per SM loops
// over work items, so barrier / TMEM set-up happens once per CTA, the K / V rings keep streaming across items and causal
// items are balanced by a closed-form schedule.  Extra barriers: q_empty (the Q tile may be reloaded), o_free (O has been
// read out of TMEM); every ring stage / phase is driven by a tile counter that runs across the items.
// ---------------------------------------------------------------------------------------------------------
// O *= alpha for this thread's half of a row's output columns (lazy rescale, rare).  Out of line on purpose: the key-tile
// loop of the persistent kernel has to stay small and straight (see there).
template <int D>
__device__ __noinline__ void attn_rescale_o(uint32_t o_taddr, float alpha) {
#pragma unroll
  for (int c = 0; c < D / 64; ++c) {
    uint32_t ov[32];
    tmem_ld_32x32(o_taddr + c * 32, ov);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
    tmem_st_32x32(o_taddr + c * 32, ov);
  }
  tmem_st_wait();
}

// TRACE: the cycle-trace instantiation (vl2_attn_args.reserved = 777 / 778 / 779); the production one carries no trace
// registers (the D = 128 softmax loop has none to spare).
template <int D, bool TRACE>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_persistent_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o,
                const AttnParams p) {
  using Cfg = AttnCfg<D>;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) { asm volatile("trap;"); }
  uint8_t* sQ = smem + Cfg::kOffQ;
  uint8_t* sK = smem + Cfg::kOffK;
  uint8_t* sV = smem + Cfg::kOffV;
  uint8_t* sP = smem + Cfg::kOffP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* p_full = bars + 11;   // [2]
  uint64_t* o_full = bars + 13;   // [2]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 15);
  uint64_t* q_empty = bars + 16;  // MMA -> Q loader: the previous item's Q K^T MMAs have read Q
  uint64_t* o_free = bars + 17;   // softmax -> MMA: the previous item's O has been read out of TMEM
  uint64_t* stage_free = bars + 18;   // thread 0 -> softmax: the previous item's output store has finished reading its P buffer
  float* smax = reinterpret_cast<float*>(smem + Cfg::kOffMax);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // heavy (late) causal tiles first
  // Persistent CTA: work items = (query tile, head, batch), sorted heaviest first (causal: latest query tile first) and
  // dealt to the CTAs boustrophedon-wise (round 0: cta, round 1: G-1-cta, ...): a closed-form schedule within a few
  // per cent of longest-processing-time-first, identical in every role of the CTA.
  const int n_qt = (p.S + BQ - 1) / BQ;
  const int hb_count = p.Hq * p.n_batch;
  const int n_items = n_qt * hb_count;
  const int G = (int)gridDim.x, cta = (int)blockIdx.x;
  struct Item { int head, b, kvh, q0, n_kv; };
  auto item_of = [&](int it, Item& w) -> bool {
    const int idx = it * G + ((it & 1) ? (G - 1 - cta) : cta);
    if (idx >= n_items) return false;
    const int r = idx / hb_count, hb = idx - r * hb_count;
    const int qt = p.causal ? (n_qt - 1 - r) : r;
    w.head = hb % p.Hq;
    w.b = hb / p.Hq;
    w.kvh = w.head / p.group;
    w.q0 = qt * BQ;
    w.n_kv = p.causal ? (qt + 1) : (p.S + BKV - 1) / BKV;
    return true;
  };

  pdl_launch_dependents();
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_o);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(stage_free, 1);
    mbar_init(o_free, kSoftmaxThreads);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], kSoftmaxThreads);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  pdl_wait();
  if (warp == 9) {
    tmem_alloc(tmem_base_ptr, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 8) {
    // ===================== TMA producers: lane 0 = Q + K ring, lane 1 = V ring (independent, never block each other)
    Item w;
    if (lane == 0) {
      int gk = 0;   // K tiles loaded so far (ring stage / phase)
      for (int it = 0; item_of(it, w); ++it) {
        const bool tl = TRACE && p.trace == 3 && blockIdx.x == 0 && it == kTlItem;
        if (it > 0) mbar_wait_parked(q_empty, (it - 1) & 1);
        mbar_arrive_expect_tx(q_full, Cfg::kTileBytes);
#pragma unroll
        for (int a = 0; a < Cfg::kAtoms; ++a)
          tma_load_4d(sQ + a * Cfg::kAtomBytes, &tmap_q, q_full, a * 64, w.head, w.q0, w.b);
        for (int j = 0; j < w.n_kv; ++j, ++gk) {
          const int st = gk & 1;
          mbar_wait_parked(&k_empty[st], ((gk >> 1) & 1) ^ 1);
          VL2_TL(tl, 256 + j);
          mbar_arrive_expect_tx(&k_full[st], Cfg::kTileBytes);
#pragma unroll
          for (int a = 0; a < Cfg::kAtoms; ++a)
            tma_load_4d(sK + st * Cfg::kTileBytes + a * Cfg::kAtomBytes, &tmap_k, &k_full[st], a * 64, w.kvh, j * BKV, w.b);
        }
      }
    } else if (lane == 1) {
      int gv = 0;
      for (int it = 0; item_of(it, w); ++it) {
        const bool tl = TRACE && p.trace == 3 && blockIdx.x == 0 && it == kTlItem;
        for (int j = 0; j < w.n_kv; ++j, ++gv) {
          const int st = gv & 1;
          mbar_wait_parked(&v_empty[st], ((gv >> 1) & 1) ^ 1);
          VL2_TL(tl, 288 + j);
          mbar_arrive_expect_tx(&v_full[st], Cfg::kTileBytes);
#pragma unroll
          for (int a = 0; a < Cfg::kAtoms; ++a)
            tma_load_4d(sV + st * Cfg::kTileBytes + a * Cfg::kAtomBytes, &tmap_v, &v_full[st], a * 64, w.kvh, j * BKV, w.b);
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer (one elected lane, (lo, hi) descriptor words: see attn_fwd_kernel) =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc_qk = umma_idesc_bf16(BQ, BKV, 0, 0);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, D, 0, 1);  // B (= V tile) is MN-major
    constexpr uint32_t hi_kmaj = umma_desc_sw128_hi(1024);
    const uint32_t q_lo = umma_desc_sw128_lo(smem_u32(sQ), 16);
    const uint32_t k_lo0 = umma_desc_sw128_lo(smem_u32(sK), 16);
    const uint32_t v_lo0 = umma_desc_sw128_lo(smem_u32(sV), Cfg::kAtomBytes);   // MN-major: LBO = next 64-wide d atom
    if (leader) {
      bool tl = false;
      int tl_slot = 0;
      int gq = 0, gp = 0;   // S tiles issued / P V tiles issued so far (ring stages and phases run across items)
      auto issue_qk = [&]() {
        const int st = gq & 1;
        mbar_wait_parked(&k_full[st], (gq >> 1) & 1);
        VL2_TL(tl, tl_slot + 1);
        tc_fence_after_sync();
        const uint32_t k_lo = k_lo0 + st * (Cfg::kTileBytes >> 4);
        const uint32_t d_tmem = tmem_base + (st ? Cfg::kColS1 : Cfg::kColS0);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = ((kk >> 2) * Cfg::kAtomBytes + (kk & 3) * 32) >> 4;
          umma_bf16_ss_lohi(d_tmem, q_lo + off, hi_kmaj, k_lo + off, hi_kmaj, idesc_qk, kk != 0);
        }
        umma_commit(&s_full[st]);
        umma_commit(&k_empty[st]);   // K_j is free as soon as Q K_j^T has run (the loader can fetch K_{j+2} early)
        ++gq;
      };
      auto issue_pv = [&](int j, int it) {
        const int st = gp & 1;
        VL2_TL(tl, 128 + 8 * j + 3);
        if (j == 0 && it > 0) mbar_wait_parked(o_free, (it - 1) & 1);   // the previous item's O has left TMEM
        tc_fence_after_sync();
        const uint32_t v_lo = v_lo0 + st * (Cfg::kTileBytes >> 4);
        const uint32_t p_tmem = tmem_base + (st ? Cfg::kColP1 : Cfg::kColP0);
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          // A = P_j in TENSOR MEMORY: 128 keys as 64 columns of 16-bit pairs, 16 keys = 8 columns.
          // B = V, MN-major: 16 keys = 2 groups of 8 rows (1024 B each)
          umma_bf16_ts_lohi(tmem_base + Cfg::kColO, p_tmem + kk * 8, v_lo + ((kk * 2048) >> 4), hi_kmaj, idesc_pv, (j | kk) != 0);
        }
        umma_commit(&o_full[st]);
        umma_commit(&v_empty[st]);
        VL2_TL(tl, 128 + 8 * j + 5);
        ++gp;
      };
      // Issue order.  P_j has its own TMEM columns, so when it arrives the S buffer it was computed from is free and
      // Q K^T(j+2) goes into the tensor pipe AHEAD of P V(j): S_{j+2} is ready a full tile before the softmax warps need it
      // (with Q K^T queued behind P V the chain P_j -> P V(j) -> Q K^T(j+2) -> S_{j+2} took as long as a softmax tile, and
      // the softmax warps waited for S on every tile).
      Item w;
      for (int it = 0; item_of(it, w); ++it) {
        tl = TRACE && p.trace == 3 && blockIdx.x == 0 && it == kTlItem;
        mbar_wait_parked(q_full, it & 1);
        VL2_TL(tl, 250);
        tl_slot = 240;
        issue_qk();
        VL2_TL(tl, 251);
        if (w.n_kv > 1) issue_qk();
        if (w.n_kv <= 2) umma_commit(q_empty);   // last Q K^T of the item issued: Q may be overwritten once they ran
        for (int j = 0; j < w.n_kv; ++j) {
          tl_slot = 128 + 8 * j;
          VL2_TL(tl, tl_slot + 0);
          const int st = gp & 1;
          mbar_wait_parked(&p_full[st], (gp >> 1) & 1);   // (S_j consumed: its buffer is free)
          mbar_wait_parked(&v_full[st], (gp >> 1) & 1);
          if (j + 2 < w.n_kv) {
            issue_qk();
            if (j + 3 == w.n_kv) umma_commit(q_empty);
          }
          VL2_TL(tl, tl_slot + 2);
          issue_pv(j, it);
        }
      }
    }
  } else {
    // ===================== softmax / output warps (two threads per query row) =====================
    // One TMEM pass over S per tile; O accumulates in TMEM across tiles (tcgen05.mma accumulate) and is rescaled
    // in place only when a row maximum grows by more than 2^8 (lazy rescale: probabilities stay <= 256, exact after
    // the final division by l).  Thread (r, hf) owns columns [64 hf, 64 hf + 64) of row r's scores and columns
    // [hf D/2, (hf+1) D/2) of its output; both threads of a row make identical rescale decisions from the exchanged
    // row maximum, so their partial sums l share one scale and are added once at the end.
    // The two threads of a row live in warps w and w + 4: their exchanges go through a 64-thread named barrier per warp
    // pair, so a pair never waits for the other three.  The item's output leaves through shared memory: each thread writes
    // its normalised row chunk into the (free) P buffer of the item's last tile in the 128-byte-swizzle box layout and one
    // thread issues a TMA store per 64-column atom (full lines, rows beyond S clipped by the tensor map) - the per-thread
    // 16-byte row stores it replaces cost 1.3 k (D = 64) to 4.2 k (D = 128) cycles per item in the timeline trace.
    const int hf = warp >> 2;
    const int pair_bar = 2 + (warp & 3);
#define VL2_PAIR_SYNC() asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory")
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t o_taddr = tmem_base + lane_sel + Cfg::kColO + hf * (D / 2);
    constexpr float kRescaleThreshold = 8.f;
    int gt = 0;   // key tiles processed so far by this CTA: TMEM / smem stage and barrier phases run across items

    const bool tr0 = TRACE && p.trace && blockIdx.x == 0 && threadIdx.x == 0;
    // 32-bit cycle counters (the low clock word; a launch is far shorter than 2^32 cycles): the trace must not cost the
    // D = 128 instantiation registers it does not have
    unsigned t0 = 0, acc_t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned tr_tiles = 0, tr_items = 0;
    const unsigned tr_begin = tr0 ? (unsigned)clock64() : 0u;
#define VL2_TR(i) do { if (tr) { const unsigned t1 = (unsigned)clock64(); acc_t[i] += t1 - t0; t0 = t1; } } while (0)
    Item w;
    for (int it = 0; item_of(it, w); ++it) {
    const bool tr = tr0 && (p.trace == 2 || (p.trace == 1 && it == 0));   // 1: the CTA's first (cold) item; 2: every item
    const bool tl = tr0 && p.trace == 3 && it == kTlItem;                  // 3: timeline of one item (g_attn_tl)
#define VL2_TRJ(i) do { VL2_TR(i); VL2_TL(tl, 8 * j + (i) + 1); } while (0)
    const int n_kv = w.n_kv, q0 = w.q0;
    const int qi = q0 + r;
    float m = -INFINITY, l = 0.f;
    // One key tile.  MASK is a compile-time flag and the loop below runs the unmasked tiles first: only an item's last
    // tile can need a mask (sequence end / causal diagonal), and with the 200-instruction masking block and the rescale
    // block inside ONE loop body every tile paid for jumping over them - the body (~18 KB) is three times the 6 KB L0
    // instruction cache and the whole kernel as large as the 32 KB L1.5, so each taken branch landed on a line that
    // had to come from L2 (a fixed ~250 cycles at the top of every tile in the timeline trace).
    auto tile = [&](const int j, auto mask_tag) {
      constexpr bool MASK = decltype(mask_tag)::value;
      if (tr) t0 = (unsigned)clock64();
      VL2_TL(tl, 8 * j);
      const int st = gt & 1;
      const uint32_t s_taddr = tmem_base + lane_sel + (st ? Cfg::kColS1 : Cfg::kColS0) + hf * 64;
      const uint32_t p_taddr = tmem_base + lane_sel + (st ? Cfg::kColP1 : Cfg::kColP0) + hf * 32;
      const int kv0 = j * BKV + hf * 64;
      // S_j has arrived; P V(j-2) has read this tile's P buffer (both polls in flight together)
      if (gt > 1) mbar_wait2(&s_full[st], (gt >> 1) & 1, &o_full[st], ((gt >> 1) & 1) ^ 1);
      else mbar_wait(&s_full[st], (gt >> 1) & 1);
      tc_fence_after_sync();
      if (it > 0 && threadIdx.x == 0 && j == 0) {   // the previous item's output store has read its staging buffer
        bulk_wait_read_all();
        mbar_arrive(stage_free);
      }
      VL2_TRJ(0);   // waiting for S_j
      if (TRACE && tr0 && p.trace == 2) {   // the same poll again, now certainly satisfied: what a poll costs by itself
        mbar_wait(&s_full[st], (gt >> 1) & 1);
        if (tr) { const unsigned t1 = (unsigned)clock64(); acc_t[7] += t1 - t0; t0 = t1; }
      }
      uint32_t sv[2][32];
#pragma unroll
      for (int c = 0; c < 2; ++c) tmem_ld_32x32(s_taddr + c * 32, sv[c]);
      tmem_ld_wait();
      VL2_TRJ(1);   // TMEM load
      if constexpr (MASK) {  // predicated selects, no per-element branches
        const int lim = p.causal ? min(p.S - 1, qi) : (p.S - 1);   // last visible key index for this row
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            sv[c][i] = (kv0 + c * 32 + i <= lim) ? sv[c][i] : 0xff800000u;  // -inf
      }
      // row maximum of the 64 scores: 8 independent chains of three-input maxima (a single chain of 32 dependent
      // FMNMX was ~130 cycles of pure latency per tile)
      float mxc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t* e = &sv[q >> 2][(q & 3) * 8];
        mxc[q] = fmax3(fmax3(__uint_as_float(e[0]), __uint_as_float(e[1]), __uint_as_float(e[2])),
                       fmax3(__uint_as_float(e[3]), __uint_as_float(e[4]), __uint_as_float(e[5])),
                       fmaxf(__uint_as_float(e[6]), __uint_as_float(e[7])));
      }
      const float mx_half = fmax3(fmax3(mxc[0], mxc[1], mxc[2]), fmax3(mxc[3], mxc[4], mxc[5]), fmaxf(mxc[6], mxc[7]));
      // exchange the half-row maxima (double-buffered by tile parity: one named barrier per tile)
      float* sm = smax + st * 256;
      sm[hf * 128 + r] = mx_half;
      VL2_PAIR_SYNC();
      VL2_TRJ(2);   // mask + max + exchange barrier
      const float m_tile = fmaxf(sm[r], sm[128 + r]) * p.scale_log2;
      // reference maximum for this tile: keep the old one unless it is too stale
      float m_use = m, alpha = 1.f;
      const bool grow = (m_tile > m + kRescaleThreshold) || (m == -INFINITY);
      if (grow) {
        m_use = (m_tile == -INFINITY) ? 0.f : m_tile;   // fully masked rows stay finite
        alpha = (m == -INFINITY) ? 0.f : fast_exp2(m - m_use);
      }
      // O (TMEM) may only be rescaled once P V(j-1) is complete
      if (j > 0) {
        if (__any_sync(0xffffffffu, grow)) {
          mbar_wait(&o_full[(gt - 1) & 1], ((gt - 1) >> 1) & 1);
          tc_fence_after_sync();
          attn_rescale_o<D>(o_taddr, alpha);
        }
      }
      VL2_TRJ(3);   // waiting for P V(j-2) / rescale
      // probabilities -> TENSOR MEMORY (16-bit pairs, this thread's 64 keys = 32 columns of the tile's P buffer): the A
      // operand of P V comes from TMEM, so P costs no shared-memory write here and no shared-memory read in the MMA
      // (with P in smem the kernel moved 224 KB per D = 128 tile through a 128 B/clk shared memory: 1750 cycles)
      float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float pr[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) pr[i] = fast_exp2(fmaf(__uint_as_float(sv[c][i]), p.scale_log2, -m_use));
#pragma unroll
        for (int i = 0; i < 32; i += 4) { rs0 += pr[i]; rs1 += pr[i + 1]; rs2 += pr[i + 2]; rs3 += pr[i + 3]; }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(pr[2 * i], pr[2 * i + 1]);
        tmem_st_32x32_x16(p_taddr + c * 16, pk);
      }
      l = l * alpha + ((rs0 + rs1) + (rs2 + rs3));
      m = m_use;
      tmem_st_wait();
      VL2_TRJ(4);   // exp2 + pack + tcgen05.st
      tc_fence_before_sync();
      mbar_arrive(&p_full[st]);
      VL2_TRJ(5);   // proxy fence + arrive
    };
    {
      const bool last_masked = p.causal || (n_kv * BKV > p.S);
      const int n_plain = n_kv - (last_masked ? 1 : 0);
      int j = 0;
      for (; j < n_plain; ++j, ++gt) tile(j, std::false_type{});
      for (; j < n_kv; ++j, ++gt) tile(j, std::true_type{});
    }
    if (tr) { tr_tiles += n_kv; ++tr_items; t0 = (unsigned)clock64(); }
    VL2_TL(tl, 120);
    // combine the two half-row sums, then each thread normalises and stores its half of the output columns.  The
    // exchange reuses the max buffer of the item's LAST tile: the barrier below orders it after every read of that
    // tile's maxima, the next item's first tile uses the other buffer, and the buffer is only rewritten two tiles later.
    float* sl = smax + ((gt - 1) & 1) * 256;
    VL2_PAIR_SYNC();
    sl[hf * 128 + r] = l;
    VL2_PAIR_SYNC();
    const float inv = 1.f / (sl[r] + sl[128 + r]);
    VL2_TL(tl, 121);
    // the last P V is complete and the previous item's output store has read the staging tile (polls in flight together)
    if (it > 0) mbar_wait2(&o_full[(gt - 1) & 1], ((gt - 1) >> 1) & 1, stage_free, (it - 1) & 1);
    else mbar_wait(&o_full[(gt - 1) & 1], ((gt - 1) >> 1) & 1);
    tc_fence_after_sync();
    VL2_TL(tl, 122);
    uint8_t* sO = sP;   // output staging tile [128 rows x D], 128-byte-swizzle box layout
#pragma unroll
    for (int c = 0; c < D / 64; ++c) {
      uint32_t ov[32];
      tmem_ld_32x32(o_taddr + c * 32, ov);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = hf * (D / 2) + c * 32 + g * 8;   // first of 8 output columns (within the head)
        uint8_t* dst = sO + (col >> 6) * Cfg::kAtomBytes + r * 128 + ((((col & 63) >> 3) ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = make_uint4(
            pack_bf16(__uint_as_float(ov[g * 8 + 0]) * inv, __uint_as_float(ov[g * 8 + 1]) * inv),
            pack_bf16(__uint_as_float(ov[g * 8 + 2]) * inv, __uint_as_float(ov[g * 8 + 3]) * inv),
            pack_bf16(__uint_as_float(ov[g * 8 + 4]) * inv, __uint_as_float(ov[g * 8 + 5]) * inv),
            pack_bf16(__uint_as_float(ov[g * 8 + 6]) * inv, __uint_as_float(ov[g * 8 + 7]) * inv));
      }
    }
    // O has left TMEM (tcgen05.wait::ld above): the MMA warp may start the next item's first P V (accumulate = 0)
    tc_fence_before_sync();
    mbar_arrive(o_free);
    fence_proxy_async_smem();
    asm volatile("bar.sync 6, %0;" ::"n"(kSoftmaxThreads) : "memory");   // every row chunk of the item is staged
    if (threadIdx.x == 0) {
#pragma unroll
      for (int a = 0; a < Cfg::kAtoms; ++a)
        if (a * 64 < p.d_true) tma_store_4d(&tmap_o, sO + a * Cfg::kAtomBytes, a * 64, w.head, q0, w.b);
      bulk_commit_group();
    }
    VL2_TR(6);   // item epilogue: l exchange, wait for the last P V, O out of TMEM, normalise, store
    VL2_TL(tl, 123);
    }   // items
    if (tr0) {
      for (int i = 0; i < 7; ++i) g_attn_trace[i] = acc_t[i];
      g_attn_trace[7] = tr_tiles;
      g_attn_trace[8] = tr_items;
      g_attn_trace[10] = acc_t[7];
      g_attn_trace[9] = (unsigned)clock64() - tr_begin;   // whole item loop of this CTA (all items, traced or not)
    }
    if (threadIdx.x == 0) bulk_wait_read_all();   // the last store still reads this CTA's shared memory
#undef VL2_PAIR_SYNC
#undef VL2_TRJ
#undef VL2_TR
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

static bool attn_persistent_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VL2_ATTN_PERSISTENT");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;   // default on: -15..-19 % on the towers, -6 % on the decoder (profiles/)
  }
  return v == 1;
}

template <int D>
static int launch_attn(const vl2_attn_args* a, cudaStream_t stream) {
  using Cfg = AttnCfg<D>;
  CUtensorMap tq, tk, tv;
  // 4-D maps (d, head, row, batch): a head narrower than the kernel's D (e.g. SigLIP's 72) is zero-filled by the TMA
  // unit beyond its true width, so QK^T and PV simply see zero columns.
  const uint32_t box[4] = {64, 1, 128, 1};
  const uint64_t dt = (uint64_t)a->D;
  {
    uint64_t dims[4] = {dt, (uint64_t)a->Hq, (uint64_t)a->S, (uint64_t)a->B};
    uint64_t str[3] = {dt * 2, (uint64_t)a->ldq * 2, (uint64_t)a->ldq * 2 * a->S};
    int rc = make_tmap_bf16(&tq, a->q, 4, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {dt, (uint64_t)a->Hkv, (uint64_t)a->S, (uint64_t)a->B};
    uint64_t str[3] = {dt * 2, (uint64_t)a->ldk * 2, (uint64_t)a->ldk * 2 * a->S};
    int rc = make_tmap_bf16(&tk, a->k, 4, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {dt, (uint64_t)a->Hkv, (uint64_t)a->S, (uint64_t)a->B};
    uint64_t str[3] = {dt * 2, (uint64_t)a->ldv * 2, (uint64_t)a->ldv * 2 * a->S};
    int rc = make_tmap_bf16(&tv, a->v, 4, dims, str, box);
    if (rc) return rc;
  }
  CUtensorMap to;   // output rows through the same 4-D form (persistent kernel: TMA store of the staged tile)
  {
    uint64_t dims[4] = {dt, (uint64_t)a->Hq, (uint64_t)a->S, (uint64_t)a->B};
    uint64_t str[3] = {dt * 2, (uint64_t)a->ldo * 2, (uint64_t)a->ldo * 2 * a->S};
    int rc = make_tmap_bf16(&to, a->out, 4, dims, str, box);
    if (rc) return rc;
  }
  AttnParams p;
  p.trace = (a->reserved == 777) ? 1 : (a->reserved == 778) ? 2 : (a->reserved == 779) ? 3 : 0;
  p.out = a->out; p.ldo = a->ldo; p.d_true = a->D; p.n_batch = a->B; p.S = a->S; p.Hq = a->Hq; p.group = a->Hq / a->Hkv; p.causal = a->causal;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  VL2_SMEM_OPT_IN(attn_fwd_kernel<D>, Cfg::kSmemBytes);
  const int n_qt = (a->S + BQ - 1) / BQ;
  if (attn_persistent_enabled()) {
    const int n_items = n_qt * a->Hq * a->B;
    const int ctas = n_items < sm_count() ? n_items : sm_count();
    if (p.trace) {
      VL2_SMEM_OPT_IN((attn_fwd_persistent_kernel<D, true>), Cfg::kSmemBytes);
      VL2_CHECK_CUDA(launch_kernel(attn_fwd_persistent_kernel<D, true>, dim3(ctas), dim3(kAttnThreads), Cfg::kSmemBytes, stream, 1, tq, tk, tv, to, p));
    } else {
      VL2_SMEM_OPT_IN((attn_fwd_persistent_kernel<D, false>), Cfg::kSmemBytes);
      VL2_CHECK_CUDA(launch_kernel(attn_fwd_persistent_kernel<D, false>, dim3(ctas), dim3(kAttnThreads), Cfg::kSmemBytes, stream, 1, tq, tk, tv, to, p));
    }
    VL2_CHECK_LAUNCH("attn_fwd_persistent_kernel");
    return VL2_OK;
  }
  dim3 grid(a->causal ? a->Hq : n_qt, a->causal ? n_qt : a->Hq, a->B);
  VL2_CHECK_CUDA(launch_kernel(attn_fwd_kernel<D>, grid, dim3(kAttnThreads), Cfg::kSmemBytes, stream, 1, tq, tk, tv, p));
  VL2_CHECK_LAUNCH("attn_fwd_kernel");
  return VL2_OK;
}

}  // namespace vl2

extern "C" int vl2_attention(const vl2_attn_args* a, void* stream) {
  using namespace vl2;
  VL2_REQUIRE(a != nullptr, VL2_E_BADSHAPE, "vl2_attention: null args");
  VL2_REQUIRE(a->B > 0 && a->S > 0 && a->Hq > 0 && a->Hkv > 0 && a->Hq % a->Hkv == 0, VL2_E_BADSHAPE,
              "vl2_attention: bad B/S/heads (%d,%d,%d,%d)", a->B, a->S, a->Hq, a->Hkv);
  VL2_REQUIRE(a->D >= 8 && a->D <= 128 && a->D % 8 == 0, VL2_E_UNSUPPORTED,
              "vl2_attention: head_dim %d unsupported (multiple of 8, <= 128)", a->D);
  VL2_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0, VL2_E_BADALIGN,
              "vl2_attention: row strides must be multiples of 8 elements");
  VL2_REQUIRE(aligned16(a->q) && aligned16(a->k) && aligned16(a->v) && aligned16(a->out), VL2_E_BADALIGN,
              "vl2_attention: pointers must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->D <= 64) return launch_attn<64>(a, st);
  return launch_attn<128>(a, st);
}

// Debug: copy the timeline stamps of the last vl2_attention launch with reserved == 779 (see g_attn_tl) to host memory.
extern "C" int vl2_debug_attn_timeline(long long* host_out320) {
  VL2_CHECK_CUDA(cudaDeviceSynchronize());
  VL2_CHECK_CUDA(cudaMemcpyFromSymbol(host_out320, vl2::g_attn_tl, 320 * sizeof(long long)));
  return VL2_OK;
}

// Debug: copy the cycle trace of the last traced vl2_attention launch (see g_attn_trace) to host memory.
extern "C" int vl2_debug_attn_trace(long long* host_out16) {
  VL2_CHECK_CUDA(cudaDeviceSynchronize());
  VL2_CHECK_CUDA(cudaMemcpyFromSymbol(host_out16, vl2::g_attn_trace, 16 * sizeof(long long)));
  return VL2_OK;
}
