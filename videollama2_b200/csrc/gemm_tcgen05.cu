// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,Nout] = epi(A[M,K] * W[N,K]^T).
//
//   warp 0      TMA producer   (one elected lane): cp.async.bulk.tensor A/B tiles -> 128B-swizzled smem ring
//   warp 1      MMA issuer     (one elected lane, (lo, hi) descriptor words: the four UTCHMMAs of a k-block are issued back
//               to back): tcgen05.mma.cta_group::1/2.kind::f16, 128 x BN x 16, accumulators in TMEM
//   warp 2      TMEM allocator (2 x BN fp32 columns: the epilogue of tile i overlaps the main loop of tile i+1)
//   warps 4..11 epilogue: tcgen05.ld (thread == accumulator row; two warps per TMEM lane quarter take alternate
//               32-column chunks, next chunk + residual prefetched while the current one is processed, bias staged in
//               smem) -> row_scale/bias/activation/SwiGLU/residual -> global
//
// Tiles are 128 x BN, BN in {64,...,256} step 32 chosen per launch by a wave/smem-bandwidth cost model (choose_bn);
// BLOCK_K = 64 bf16 = one 128-byte swizzle atom; the grid is persistent
// (<= #SMs CTAs, static round-robin over tiles, M fastest so that a wave shares W tiles through L2).
// M/N/K tails: TMA zero-fills out-of-bounds reads, stores are predicated.
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"

namespace vl2 {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int kEpiWarps = 8;                       // 2 per TMEM lane quarter: they split the 32-column chunks
static constexpr int kEpiThreads = kEpiWarps * 32;
static constexpr int kGemmThreads = 128 + kEpiThreads;    // warps 0..3: TMA / MMA / TMEM alloc / spare
static constexpr int kConvLine = 16;                      // Conv3d front end: output lines and columns per time step are padded to 16

// PAIR = true: two CTAs of a cluster (one TPC) run tcgen05.mma.cta_group::2 on a 256 x BN tile; each CTA stages its own
// 128 rows of A and HALF of the B tile, so a k-block costs 16 KB + BN*64 B of smem/L2 traffic per CTA instead of
// 16 KB + BN*128 B, and the ring is deep enough (6 stages at BN=256) to cover the L2/HBM latency at the full MMA rate.
// BN2 > 0 (the "wide" tile, lean pair kernel only): the tile is BN + BN2 columns wide, held as TWO accumulators (TMEM columns
// [0, BN) and [256, 256 + BN2)) that share every A k-block: a k-block then costs 16 KB of A for 416 columns instead of
// 224 or 256 - the narrow pair tiles are bound by operand ingest (~58 B/clk/SM), not by the tensor pipe.  All 512
// TMEM columns hold ONE accumulator stage, so the epilogue is not overlapped with the next tile: the tile choice uses it
// only where a launch then needs a single round of the persistent grid (o_proj / down_proj at S = 1776).
template <int BN, bool PAIR, int BN2 = 0>
struct GemmCfg {
  static_assert(BN % 32 == 0 && BN >= 64 && BN <= 256, "BN must be a multiple of 32 in [64,256]");
  static_assert(BN2 == 0 || (PAIR && BN2 % 32 == 0 && BN2 >= 64 && BN2 <= 256), "a wide tile is a cta_group::2 tile");
  static constexpr int kTileN = BN + BN2;
  static constexpr int kRowsB = PAIR ? BN / 2 : BN;   // B rows staged by one CTA
  static constexpr int kRowsB2 = BN2 / 2;             // ... of the second accumulator
  static constexpr int kStageBytesA = BM * BK * 2;
  static constexpr int kStageBytesB1 = kRowsB * BK * 2;
  static constexpr int kStageBytesB = (kRowsB + kRowsB2) * BK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStagingBytes = kEpiWarps * 4096;  // per epilogue warp: 32 rows x 128 B transpose buffer
  static constexpr int kExtraBytes = kStagingBytes + 2 * 256 * 4 /*bias (+ colsum) staging*/ + 512 /*barriers*/;
  static constexpr int kMaxStages = (227 * 1024 - kExtraBytes) / kStageBytes;
  static constexpr int kStages = kMaxStages > 8 ? 8 : kMaxStages;
  static constexpr int kAccStages = BN2 > 0 ? 1 : 2;
  static constexpr int kAccStride = BN > 128 ? 256 : (BN > 64 ? 128 : 64);  // TMEM columns between accumulator stages
  static constexpr int kTmemCols = 2 * kAccStride;                           // power of two >= 32 (wide: 512 = both accumulators)
  static constexpr int kSmemBytes = kStages * kStageBytes + kExtraBytes;
  static constexpr int kChunks = BN / 32;
  static_assert(BN2 == 0 || kAccStride == 256, "wide tile: the second accumulator lives at TMEM column 256");
  static_assert(kStageBytesB1 % 1024 == 0 || BN2 == 0, "second B box must start on a swizzle-atom boundary");
};

struct GemmParams {
  void* C;
  const float* bias;
  const void* residual;
  const float* row_scale;
  int64_t ldc, ldr;
  int M, N, K;
  int act;
  int out_f32;
  int num_m_tiles, num_n_tiles;
  // epilogue-fused all-gather: peers' copies of C (NVLink P2P stores) or one NVSwitch multicast address
  __nv_bfloat16* bcast[8];
  __nv_bfloat16* mc;
  int n_bcast;
  // folded RMSNorm / LayerNorm (see vl2.h)
  const float* rms_sumsq_in;
  const float* ln_sum_in;
  const float* ln_colsum;
  float* rowsum_out;
  float* sumsq_out;
  int rms_nparts;
  float rms_inv_dim, rms_eps;
  // split-K of the last, partial round (see launch_gemm): items >= split_first are K-slices of the remaining tiles
  int split_first, split_s, num_items;
  float* ws_partial;      // [tiles past split_first][split_s - 1][tile rows][BN] fp32 partial accumulators
  unsigned* ws_flags;     // [tiles past split_first][2] arrival counters (self-resetting)
  // implicit-GEMM Conv3d(k = s = 2) front end (see vl2.h: conv_C > 0).  A rows are the output positions in a padded
  // enumeration (to, ho16, wo16): 16 x 16 rows per output time step, so that one 128-row tile is 8 whole output lines of
  // one time step and a k-block (one tap x 64 channels) of the tile is ONE 5-D TMA box over x viewed as
  // [T, H/2, 2, W/2, 2C] (the stride-2 walks along H and W become unit walks over the coarse indices at a fixed parity);
  // the spatial zero padding and the padded rows are TMA out-of-bounds fill.
  int conv_C, conv_pad, conv_Ho, conv_Wo, conv_To;
  // RoPE in the QKV epilogue (see vl2.h): adjacent output columns (2i, 2i+1) of a head are a rotation pair
  const uint32_t* rope_tab;
  int rope_cols, rope_D, rope_pos0;
  int trace;   // debug: CTA 0 records clock64() at tile boundaries of its MMA and epilogue roles (vl2_debug_gemm_trace)
};

// [0] tiles traced; per tile t (<= 7), at 8*t + 1: MMA waits for the accumulator stage, MMA starts issuing, MMA issued
// the last commit, epilogue warp 4 starts the tile, its accumulator is complete, its last span is stored.
__device__ long long g_gemm_trace[128];   // [64 + 8 t ..]: phases of warp 4's first span of tile t (see the epilogue)

// one 16-byte store replicated by the NVSwitch to every GPU of the multicast group
__device__ __forceinline__ void multimem_st128(void* mc_addr, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(__uint_as_float(v.x)),
               "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
               : "memory");
}

__device__ __forceinline__ float fast_sigmoid_mul(float x, float k_log2e) {
  // x * sigmoid(k x) = x / (1 + 2^(-k*log2e*x)); approximate reciprocal on the MUFU pipe
  return __fdividef(x, 1.f + fast_exp2(-k_log2e * x));
}
__device__ __forceinline__ float gelu_tanh(float x) {
  // 0.5 x (1 + tanh(u)) = x * sigmoid(2u), u = sqrt(2/pi) (x + 0.044715 x^3)      (HF "gelu_pytorch_tanh", SigLIP MLP)
  const float u2 = x * fmaf(0.0356774081f * x, x, 0.7978845608f) * 2.885390082f;   // 2 u log2(e)
  return __fdividef(x, 1.f + fast_exp2(-u2));
}
__device__ __forceinline__ float act_apply(float x, int act) {
  switch (act) {
    case VL2_ACT_QUICK_GELU: return fast_sigmoid_mul(x, 1.702f * 1.4426950408889634f);
    case VL2_ACT_SILU: return fast_sigmoid_mul(x, 1.4426950408889634f);
    case VL2_ACT_GELU_ERF: return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
    case VL2_ACT_GELU_TANH: return gelu_tanh(x);
    default: return x;
  }
}

// One unit of the persistent loop: a whole output tile (part = -1), or - for the tiles of the last, partial round - one
// K-slice of a tile: part 0 owns the tile (adds the other slices' partials and runs the epilogue), parts > 0 dump their
// fp32 accumulators to the workspace.  Every role (producer, MMA, epilogue) decodes the same sequence.
constexpr size_t kSplitFlagBytes = 65536;  // arrival counters at the head of the split-K workspace, one 128-byte line each
constexpr int kSplitFlagStride = 32;       // (zero at first use, self-resetting)
constexpr double kSplitOverheadCycles = 16000.0;   // dump + fence + flag + owner's partial reads (trace: llm_qkv, 3 slices)
struct WorkItem { int tile, kb0, kb1, part; };
__device__ __forceinline__ WorkItem decode_item(int item, const GemmParams& p, int nkb) {
  WorkItem w;
  if (item < p.split_first) {
    w.tile = item; w.kb0 = 0; w.kb1 = nkb; w.part = -1;
  } else {
    const int j = item - p.split_first;
    w.tile = p.split_first + j / p.split_s;
    w.part = j % p.split_s;
    w.kb0 = (int)((long long)w.part * nkb / p.split_s);
    w.kb1 = (int)((long long)(w.part + 1) * nkb / p.split_s);
  }
  return w;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}


// ---------------------------------------------------------------------------------------------------------------------
// Lean epilogue (LEAN = true instantiations): everything the hot launches of the path use - row scale / folded RMSNorm,
// bias, activation, SwiGLU, RoPE, residual, row statistics, bf16 output - and nothing else (the Conv3d front end, peer /
// multicast stores, split-K, folded LayerNorm and fp32 output stay in the general epilogue below).  Differences that
// matter for speed (profiles/r02_gemm_epilogue_trace.txt: the general epilogue spends 2.2 k cycles per 64-column span in
// its smem -> global store loop and 2.7-3 k staging the residual, against ~1.3 k per 32-column half of arithmetic):
//   * the unit of work is 32 OUTPUT columns: thread = accumulator row writes its 64 bytes into a [32 rows x 64 B]
//     staging block (64-byte swizzle, conflict-free), ONE lane issues a TMA store of the block (cp.async.bulk.tensor
//     shared -> global, clipped at the matrix edge by the tensor map) - no per-thread address arithmetic or stores;
//   * the residual block of a unit arrives by TMA load into the same staging block (the sum is formed in place), issued
//     one unit ahead - also across tiles - into the other of the warp's two blocks, so its latency is never exposed;
//   * feature tests read one register of flags instead of the kernel parameters in constant memory.
// ---------------------------------------------------------------------------------------------------------------------
enum : uint32_t { kFBias = 1, kFRes = 2, kFSwiglu = 4, kFScale = 8, kFSumsq = 16, kFRowsum = 32, kFRope = 64, kFTrace = 128 };

// PLAIN = true: the instantiation for launches without activation, SwiGLU and RoPE (tower qkv / out_proj / fc2, decoder
// o_proj / down_proj, most connector convolutions): those bodies are not compiled in.  Not a micro-optimisation of the
// epilogue alone - with the bodies present the 256-wide pair kernel is 11.3 k SASS instructions, the epilogue warps stream
// through them while the single MMA-issuing thread and the TMA producer run their short loops, and the k-loop of the K = 1024
// GEMMs ran 4-8 % slower (instruction-cache misses in the issuing thread; no effect on the wide tile, whose epilogue does not
// overlap its mainloop): profiles/r02_gemm_epilogue_trace.txt, last part.
template <int BN, bool PAIR, int BN2, bool PLAIN>
__device__ __forceinline__ void epilogue_lean(const CUtensorMap* tmap_c, const CUtensorMap* tmap_r, const GemmParams& p,
                                              uint8_t* smem_stage, float* sbias, uint64_t* tmem_full, uint64_t* tmem_empty,
                                              uint64_t* res_bars, uint32_t tmem_base, uint32_t rank, int tile0,
                                              int tile_stride) {
  using Cfg = GemmCfg<BN, PAIR, BN2>;
  constexpr int kTileM = PAIR ? 2 * BM : BM;
  constexpr int kTileN = Cfg::kTileN;
  constexpr uint32_t kUnitBytes = 32 * 64;
  constexpr uint32_t kExpFeat = PLAIN ? ~(kFSwiglu | kFRope) : ~0u;   // features compiled into this instantiation
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ew = warp & 3, grp = (warp - 4) >> 2, etid = threadIdx.x - 128;
  const int row_in_tile = ew * 32 + lane;
  uint32_t feat = (p.bias != nullptr ? kFBias : 0u) | (p.residual != nullptr ? kFRes : 0u) |
                  (p.act == VL2_ACT_SWIGLU ? kFSwiglu : 0u) |
                  ((p.row_scale != nullptr || p.rms_sumsq_in != nullptr) ? kFScale : 0u) |
                  (p.sumsq_out != nullptr ? kFSumsq : 0u) | (p.rowsum_out != nullptr ? kFRowsum : 0u) |
                  (p.rope_tab != nullptr ? kFRope : 0u) | ((p.trace && blockIdx.x == 0 && etid == 0) ? kFTrace : 0u);
  asm volatile("" : "+r"(feat));   // keep the flags in a register (do not re-derive them from constant memory)
  int act = PLAIN ? (int)VL2_ACT_NONE : p.act;
  if (!PLAIN) asm volatile("" : "+r"(act));
  // accumulator columns per output unit: 32, or 64 for SwiGLU ((gate, up) pairs -> 32 outputs).  The two warps of a TMEM
  // lane quarter take alternate units.
  const int ustep = (kExpFeat & feat & kFSwiglu) ? 64 : 32;
  uint8_t* stg = smem_stage + (warp - 4) * 4096;                 // two [32 x 64 B] blocks
  const uint32_t my_row = smem_u32(stg) + lane * 64;
  const uint32_t swz = (uint32_t)(lane >> 1) & 3u;               // 64-byte swizzle key of this thread's row
  uint64_t* rbar = res_bars + (warp - 4) * 2;
  uint32_t uc = 0;          // units this warp has processed: staging block uc & 1, residual barrier phase (uc >> 1) & 1
  bool res_ahead = false;   // (lane 0) the first residual block of the coming tile is already in flight
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  // TMEM column of tile column c: the second accumulator of a wide tile starts at column 256
  auto tcol = [](int c) { return (BN2 > 0 && c >= BN) ? 256 + (c - BN) : c; };
  int it = 0;
  for (int tile = tile0; tile < num_tiles; tile += tile_stride, ++it) {
    const int as = Cfg::kAccStages == 2 ? (it & 1) : 0;
    const uint32_t aphase = Cfg::kAccStages == 2 ? ((it >> 1) & 1) : (it & 1);
    const bool etr = (feat & kFTrace) && it < 7;
    if (etr) g_gemm_trace[8 * it + 4] = clock64();
    const int m0 = (tile % p.num_m_tiles) * kTileM + (int)rank * BM;
    const int n0 = (tile / p.num_m_tiles) * kTileN;
    const int row = m0 + row_in_tile;
    const bool row_ok = row < p.M;
    const int rbase = m0 + ew * 32;
    const int c_first = grp * ustep;
    const bool first_ok = c_first < kTileN && n0 + c_first < p.N;   // this warp has work in the tile
    if ((feat & kFRes) && lane == 0 && first_ok && !res_ahead) {
      bulk_wait_read_all();
      mbar_arrive_expect_tx(&rbar[uc & 1], kUnitBytes);
      tma_load_2d(stg + (uc & 1) * kUnitBytes, tmap_r, &rbar[uc & 1], n0 + c_first, rbase);
    }
    res_ahead = false;
    if (feat & kFBias) {
      // stage this tile's bias slice once; double buffered by accumulator stage (a wide tile has one stage and uses the
      // whole area: one more barrier keeps its writes behind the previous tile's reads)
      if (Cfg::kAccStages == 1) asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      for (int i = etid; i < kTileN; i += kEpiThreads) sbias[as * 256 + i] = (n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
    }
    float rs = 1.f;
    if (feat & kFScale) {
      if (p.row_scale != nullptr && row_ok) rs = p.row_scale[row];
      if (p.rms_sumsq_in != nullptr && row_ok) {
        // sum the producer's per-32-column partials in a fixed order (deterministic: no atomics anywhere)
        const float* pp = p.rms_sumsq_in + (int64_t)row * p.rms_nparts;
        const int nvec = (p.rms_nparts % 4 == 0) ? p.rms_nparts : 0;   // rows are 16-byte aligned only then
        float q = 0.f;
        int i = 0;
        for (; i + 4 <= nvec; i += 4) {
          const float4 t4 = *reinterpret_cast<const float4*>(pp + i);
          q += (t4.x + t4.y) + (t4.z + t4.w);
        }
        for (; i < p.rms_nparts; ++i) q += pp[i];
        rs *= rsqrtf(q * p.rms_inv_dim + p.rms_eps);
      }
    }
    if (feat & kFBias) asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
    const uint32_t sb = smem_u32(sbias + as * 256);
    mbar_wait(&tmem_full[as], aphase);
    tc_fence_after_sync();
    if (etr) g_gemm_trace[8 * it + 5] = clock64();
    const uint32_t taddr = tmem_base + as * Cfg::kAccStride + ((uint32_t)(ew * 32) << 16);
    uint32_t v[32];
    if (first_ok) tmem_ld_32x32(taddr + tcol(c_first), v);

#pragma unroll 1
    for (int c0 = c_first; c0 < kTileN; c0 += 2 * ustep) {
      const int col0 = n0 + c0;                     // first accumulator column of the unit
      if (col0 >= p.N) break;                       // warp-uniform
      const int ucols = min(min(ustep, kTileN - c0), p.N - col0);   // its valid accumulator columns (multiple of 8)
      const bool next_unit_ok = c0 + 2 * ustep < kTileN && col0 + 2 * ustep < p.N;
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {              // 32 accumulator columns at a time (two passes only for SwiGLU)
        if (hf * 32 >= ucols) break;                // warp-uniform
        const int cc = c0 + hf * 32;                // tile column of these 32 accumulator columns
        const bool ptr_ = etr && c0 == c_first && hf == 0;
        if (ptr_) g_gemm_trace[64 + 8 * it] = clock64();
        tmem_ld_wait();
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
        // this warp's next 32 accumulator columns: second half of a SwiGLU unit, else the first of its next unit
        if (hf == 0 && ucols > 32) tmem_ld_32x32(taddr + tcol(cc + 32), v);
        else if (next_unit_ok) tmem_ld_32x32(taddr + tcol(c0 + 2 * ustep), v);
        if (ptr_) g_gemm_trace[64 + 8 * it + 1] = clock64();
        if (feat & kFScale) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] *= rs;
        }
        if (feat & kFBias) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 bq = lds_f4(sb + (cc + g * 4) * 4);
            x[g * 4 + 0] += bq.x; x[g * 4 + 1] += bq.y; x[g * 4 + 2] += bq.z; x[g * 4 + 3] += bq.w;
          }
        }
        if (!(kExpFeat & feat & kFSwiglu)) {
          // one branch per activation (warp-uniform), each with a fully unrolled body; a launch without activation takes ONE
          // jump over all of them (every skipped 200-instruction body is an instruction-cache miss at the landing site:
          // compiling the bodies out made the unit 300-450 cycles shorter, profiles/r02_gemm_epilogue_trace.txt)
          if (act == VL2_ACT_NONE) {
          } else if (act == VL2_ACT_QUICK_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = fast_sigmoid_mul(x[j], 1.702f * 1.4426950408889634f);
          } else if (act == VL2_ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = fast_sigmoid_mul(x[j], 1.4426950408889634f);
          } else if (act == VL2_ACT_GELU_ERF) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = 0.5f * x[j] * (1.f + erff(x[j] * 0.70710678118654752f));
          } else if (act == VL2_ACT_GELU_TANH) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = gelu_tanh(x[j]);
          }
          if ((kExpFeat & feat & kFRope) && col0 < p.rope_cols && row_ok) {
            // 32 accumulator columns = 16 rotation pairs of one head, frequencies i0 .. i0 + 15, angle of this row's position
            const int i0 = (col0 % p.rope_D) >> 1;
            const uint4* tr = reinterpret_cast<const uint4*>(p.rope_tab + (int64_t)(p.rope_pos0 + row) * (p.rope_D >> 1) + i0);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 cs = __ldg(tr + g);
              const uint32_t e[4] = {cs.x, cs.y, cs.z, cs.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float c = bf16_lo(e[j]), sn = bf16_hi(e[j]);
                const float a = x[g * 8 + 2 * j], b = x[g * 8 + 2 * j + 1];
                x[g * 8 + 2 * j] = a * c - b * sn;
                x[g * 8 + 2 * j + 1] = b * c + a * sn;
              }
            }
          }
        }
        if (ptr_) g_gemm_trace[64 + 8 * it + 2] = clock64();
        // ---- staging block of this output unit
        const uint32_t blk = my_row + (uc & 1) * kUnitBytes;
        if (hf == 0) {
          if (lane == 0) {
            bulk_wait_read_all();     // the stores issued so far have read their blocks (the last one >= half a unit ago)
            if (feat & kFRes) {
              // residual of this warp's NEXT unit -> the other block: the next unit of the tile, or the next tile's first
              int nc = -1, nr = rbase;
              if (next_unit_ok) nc = col0 + 2 * ustep;
              else if (tile + tile_stride < num_tiles) {
                const int tn = tile + tile_stride;
                const int n0n = (tn / p.num_m_tiles) * kTileN;
                if (c_first < kTileN && n0n + c_first < p.N) {
                  nc = n0n + c_first;
                  nr = (tn % p.num_m_tiles) * kTileM + (int)rank * BM + ew * 32;
                  res_ahead = true;
                }
              }
              if (nc >= 0) {
                mbar_arrive_expect_tx(&rbar[(uc + 1) & 1], kUnitBytes);
                tma_load_2d(stg + ((uc + 1) & 1) * kUnitBytes, tmap_r, &rbar[(uc + 1) & 1], nc, nr);
              }
            }
          }
          if (feat & kFRes) mbar_wait(&rbar[uc & 1], (uc >> 1) & 1);
          else __syncwarp();
        }
        if (ptr_) g_gemm_trace[64 + 8 * it + 3] = clock64();
        if (kExpFeat & feat & kFSwiglu) {
          // accumulator columns interleave (gate, up): 32 columns -> 16 outputs = chunks 2*hf, 2*hf+1 of the unit
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float g0 = x[g * 16 + 4 * j + 0], u0 = x[g * 16 + 4 * j + 1];
              const float g1 = x[g * 16 + 4 * j + 2], u1 = x[g * 16 + 4 * j + 3];
              o[j] = pack_bf16(fast_sigmoid_mul(g0, 1.4426950408889634f) * u0, fast_sigmoid_mul(g1, 1.4426950408889634f) * u1);
            }
            sts128(blk + ((uint32_t)((hf * 2 + g) ^ swz) << 4), make_uint4(o[0], o[1], o[2], o[3]));
          }
        } else {
          float ssq = 0.f, ssum = 0.f;
          if (feat & kFRes) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 r = lds128(blk + ((uint32_t)(g ^ swz) << 4));
              x[g * 8 + 0] += bf16_lo(r.x); x[g * 8 + 1] += bf16_hi(r.x);
              x[g * 8 + 2] += bf16_lo(r.y); x[g * 8 + 3] += bf16_hi(r.y);
              x[g * 8 + 4] += bf16_lo(r.z); x[g * 8 + 5] += bf16_hi(r.z);
              x[g * 8 + 6] += bf16_lo(r.w); x[g * 8 + 7] += bf16_hi(r.w);
            }
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint4 pk = make_uint4(pack_bf16(x[g * 8], x[g * 8 + 1]), pack_bf16(x[g * 8 + 2], x[g * 8 + 3]),
                                        pack_bf16(x[g * 8 + 4], x[g * 8 + 5]), pack_bf16(x[g * 8 + 6], x[g * 8 + 7]));
            sts128(blk + ((uint32_t)(g ^ swz) << 4), pk);
            if ((feat & kFSumsq) && g * 8 < ucols) {   // statistics of what the consumer will read
              const float a0 = bf16_lo(pk.x), a1 = bf16_hi(pk.x), a2 = bf16_lo(pk.y), a3 = bf16_hi(pk.y);
              const float a4 = bf16_lo(pk.z), a5 = bf16_hi(pk.z), a6 = bf16_lo(pk.w), a7 = bf16_hi(pk.w);
              ssq += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3 + a4 * a4 + a5 * a5 + a6 * a6 + a7 * a7;
              if (feat & kFRowsum) ssum += ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
            }
          }
          if ((feat & kFSumsq) && row_ok) {   // one slot per 32 output columns
            const int64_t slot = (int64_t)row * (p.N >> 5) + (col0 >> 5);
            p.sumsq_out[slot] = ssq;
            if (feat & kFRowsum) p.rowsum_out[slot] = ssum;
          }
        }
        if (ptr_) g_gemm_trace[64 + 8 * it + 4] = clock64();
        if (hf * 32 + 32 >= ucols) {   // the unit is complete
          fence_proxy_async_smem();   // this thread's smem writes -> visible to the TMA unit
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(tmap_c, stg + (uc & 1) * kUnitBytes, (kExpFeat & feat & kFSwiglu) ? (col0 >> 1) : col0, rbase);
            bulk_commit_group();
          }
          ++uc;
        }
        if (ptr_) g_gemm_trace[64 + 8 * it + 5] = clock64();
      }
    }
    if (etr) g_gemm_trace[8 * it + 6] = clock64();
    // all tcgen05.ld of this thread have completed (wait::ld above) -> hand the accumulator stage back
    tmem_ld_wait();
    tc_fence_before_sync();
    if (PAIR) mbar_arrive_cluster(&tmem_empty[as], 0); else mbar_arrive(&tmem_empty[as]);
  }
  if (lane == 0) bulk_wait_read_all();   // the staging blocks must outlive the last stores' reads
}

template <int BN, bool PAIR, bool LEAN, int BN2 = 0, bool PLAIN = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_b2, const __grid_constant__ CUtensorMap tmap_c,
                         const __grid_constant__ CUtensorMap tmap_r, const GemmParams p) {
  static_assert(BN2 == 0 || LEAN, "the wide tile exists for the lean epilogue only");
  using Cfg = GemmCfg<BN, PAIR, BN2>;
  constexpr int kTileN = Cfg::kTileN;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;   // 0 = leader (issues the MMAs), 1 = peer
  const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;        // persistent tile loop start / stride
  const int tile_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int kTileM = PAIR ? 2 * BM : BM;
  // SWIZZLE_128B operands need 1024-byte aligned stage bases: the kernel has no static smem, so the dynamic window
  // starts at offset 0 of the CTA's shared memory (checked below; a misaligned base traps instead of corrupting).
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) { asm volatile("trap;"); }
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kStageBytesA;
  uint8_t* smem_stage = smem + Cfg::kStages * Cfg::kStageBytes;                       // [kEpiWarps][4096]
  float* sbias = reinterpret_cast<float*>(smem_stage + Cfg::kStagingBytes);           // [2][256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + 512);
  uint64_t* full_bar = bars;                        // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + Cfg::kStages;        // [kStages]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;    // [2]        MMA -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;             // [2]        epilogue -> MMA
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* res_bars = tmem_empty + 3;              // [kEpiWarps][2]  lean epilogue: residual block landed (TMA)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_k_blocks = (p.K + BK - 1) / BK;

  pdl_launch_dependents();   // let the next kernel's CTAs take the SMs this kernel's last wave leaves idle
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (BN2 > 0) tma_prefetch_desc(&tmap_b2);
    if (LEAN) {
      tma_prefetch_desc(&tmap_c);
      if (p.residual != nullptr) tma_prefetch_desc(&tmap_r);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], PAIR ? 2 * kEpiThreads : kEpiThreads);  // pair: both CTAs' epilogues arrive on the leader's
    }
    if (LEAN) {
      for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(&res_bars[i], 1);
    }
    fence_barrier_init();
  }
  pdl_wait();                // predecessors complete + visible: nothing above touched global memory or TMEM
  if (warp == 2) {
    if (PAIR) { tmem_alloc_pair(tmem_base_ptr, Cfg::kTmemCols); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_base_ptr, Cfg::kTmemCols); tmem_relinquish(); }
  }
  tc_fence_before_sync();
  if (PAIR) cluster_sync_all(); else __syncthreads();   // barrier inits + TMEM address visible (cluster-wide for a pair)
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 0) {
    if (elect_one()) {
      // ===================== TMA producer (one elected lane: see the MMA issuer) =====================
      int stage = 0;
      uint32_t phase = 0;
      const bool conv = !LEAN && p.conv_C > 0;
      const int slabs = conv ? p.conv_C / BK : 1;
      for (int item = tile0; item < p.num_items; item += tile_stride) {
        const WorkItem w = decode_item(item, p, num_k_blocks);
        const int m0 = (w.tile % p.num_m_tiles) * kTileM + (int)rank * BM;
        const int nt0 = (w.tile / p.num_m_tiles) * kTileN;                    // first column of the tile
        const int n0 = nt0 + (int)rank * (PAIR ? BN / 2 : 0);                 // first B row this CTA stages
        // conv: this 128-row tile = output lines ho_base .. ho_base + 7 of time step `to`
        const int line0 = m0 / kConvLine, to = line0 / kConvLine, ho_base = line0 % kConvLine;
        int tap = w.kb0 / slabs, slab = w.kb0 - tap * slabs;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (PAIR) {
            // both CTAs' bytes land on the leader's full barrier
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          }
          uint8_t* dst_a = smem_a + stage * Cfg::kStageBytesA;
          if (conv) {
            // implicit im2col: k-block = (tap, 64-channel slab).  Input index along each axis = 2 * out + e, e = d - pad in
            // {-1, 0, 1}: parity e & 1, coarse index out + floor(e / 2)  (-1 = the zero-padded border -> out of bounds)
            const int et = (tap >> 2) - p.conv_pad, eh = ((tap >> 1) & 1) - p.conv_pad, ew = (tap & 1) - p.conv_pad;
            const int ph = eh & 1, pw = ew & 1;
            const int c0 = pw * p.conv_C + slab * BK, wc0 = (ew - pw) / 2, hc0 = ho_base + (eh - ph) / 2, t = 2 * to + et;
            if (PAIR) tma_load_5d_pair(dst_a, &tmap_a, &full_bar[stage], c0, wc0, ph, hc0, t);
            else tma_load_5d(dst_a, &tmap_a, &full_bar[stage], c0, wc0, ph, hc0, t);
            if (++slab == slabs) { slab = 0; ++tap; }
          } else if (PAIR) {
            tma_load_2d_pair(dst_a, &tmap_a, &full_bar[stage], kb * BK, m0);
          } else {
            tma_load_2d(dst_a, &tmap_a, &full_bar[stage], kb * BK, m0);
          }
          if (PAIR) tma_load_2d_pair(smem_b + stage * Cfg::kStageBytesB, &tmap_b, &full_bar[stage], kb * BK, n0);
          else tma_load_2d(smem_b + stage * Cfg::kStageBytesB, &tmap_b, &full_bar[stage], kb * BK, n0);
          if (BN2 > 0)   // second accumulator's half of B (wide tile: always a pair)
            tma_load_2d_pair(smem_b + stage * Cfg::kStageBytesB + Cfg::kStageBytesB1, &tmap_b2, &full_bar[stage], kb * BK,
                             nt0 + BN + (int)rank * (BN2 / 2));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only for a pair): ONE ELECTED lane - the compiler then knows the branch
    // is single-threaded and emits the UTCHMMAs back to back; under `lane == 0` it wrapped every MMA in an elect loop with
    // five register -> uniform-register moves (~20 instructions and ~100 cycles per MMA, more than a 128 x 128 x 16 MMA
    // takes to execute).  Descriptors as (lo, hi) words: one uniform add per operand per MMA (ptx.cuh). =====================
    const bool leader = (rank == 0) && elect_one();
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN, 0, 0);
      constexpr uint32_t idesc2 = umma_idesc_bf16(kTileM, BN2 > 0 ? BN2 : BN, 0, 0);
      constexpr uint32_t dhi = umma_desc_sw128_hi(1024);
      const uint32_t a_lo0 = umma_desc_sw128_lo(smem_u32(smem_a), 16);
      const uint32_t b_lo0 = umma_desc_sw128_lo(smem_u32(smem_b), 16);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = tile0; item < p.num_items; item += tile_stride, ++it) {
        const WorkItem w = decode_item(item, p, num_k_blocks);
        const int as = Cfg::kAccStages == 2 ? (it & 1) : 0;
        const uint32_t aphase = Cfg::kAccStages == 2 ? ((it >> 1) & 1) : (it & 1);
        const bool tr = p.trace && blockIdx.x == 0 && it < 7;
        if (tr) g_gemm_trace[8 * it + 1] = clock64();
        mbar_wait(&tmem_empty[as], aphase ^ 1);  // epilogue drained this accumulator stage
        tc_fence_after_sync();
        if (tr) g_gemm_trace[8 * it + 2] = clock64();
        const uint32_t d_tmem = tmem_base + as * Cfg::kAccStride;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t a_lo = a_lo0 + stage * (Cfg::kStageBytesA >> 4);
          const uint32_t b_lo = b_lo0 + stage * (Cfg::kStageBytesB >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint32_t acc = (kb != w.kb0) || (k != 0);
            if (PAIR) umma_bf16_ss_pair_lohi(d_tmem, a_lo + 2 * k, dhi, b_lo + 2 * k, dhi, idesc, acc);
            else umma_bf16_ss_lohi(d_tmem, a_lo + 2 * k, dhi, b_lo + 2 * k, dhi, idesc, acc);
            if (BN2 > 0)   // the same A k-slice into the second accumulator (TMEM column 256)
              umma_bf16_ss_pair_lohi(d_tmem + 256, a_lo + 2 * k, dhi, b_lo + (Cfg::kStageBytesB1 >> 4) + 2 * k, dhi, idesc2, acc);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs have read it
          if (PAIR) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (PAIR) umma_commit_pair(&tmem_full[as]); else umma_commit(&tmem_full[as]);
        if (tr) { g_gemm_trace[8 * it + 3] = clock64(); g_gemm_trace[0] = it + 1; }
      }
    }
  } else if (warp >= 4 && LEAN) {
    epilogue_lean<BN, PAIR, BN2, PLAIN>(&tmap_c, &tmap_r, p, smem_stage, sbias, tmem_full, tmem_empty, res_bars, tmem_base, rank,
                                        tile0, tile_stride);
  } else if (warp >= 4) {
    // ===================== general epilogue (8 warps) =====================
    // warp w: TMEM lane quarter (w & 3); the two warps of a quarter take alternate 64-column spans of the tile.
    // Accumulator rows live one per thread (TMEM lane == row), which is the wrong shape for global memory, so every
    // 32-row x 64-column block goes through a per-warp swizzled smem buffer: residual rows come in and output rows go
    // out as full 128-byte lines (8 lanes x 16 B per row) instead of 32 different rows per instruction.
    const int ew = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int etid = threadIdx.x - 128;
    const int row_in_tile = ew * 32 + lane;
    const bool swiglu = (p.act == VL2_ACT_SWIGLU);
    const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(p.residual);
    const uint32_t stg = smem_u32(smem_stage + (warp - 4) * 4096);
    const uint32_t my_row = stg + lane * 128;             // this thread's row in the staging buffer
    const int sw = lane & 7;                              // its swizzle key
    const int t_row = lane >> 3, t_chunk = lane & 7;      // transposed role: 8 lanes per row, 4 rows per instruction
    int it = 0;
    for (int item = tile0; item < p.num_items; item += tile_stride, ++it) {
      const WorkItem w = decode_item(item, p, num_k_blocks);
      const int tile = w.tile;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const bool etr = p.trace && blockIdx.x == 0 && it < 7 && etid == 0;
      if (etr) g_gemm_trace[8 * it + 4] = clock64();
      const int m0 = (tile % p.num_m_tiles) * kTileM + (int)rank * BM;
      const int n0 = (tile / p.num_m_tiles) * BN;
      if (w.part > 0) {
        // ---- K-slice of a split tile: dump the fp32 accumulators (thread = row, 128-byte pieces), then signal the owner
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after_sync();
        const uint32_t taddr_p = tmem_base + as * Cfg::kAccStride + ((uint32_t)(ew * 32) << 16);
        // layout: [tile slot][slice][CTA half][epilogue warp quarter][32-column chunk][8 groups][lane][4 floats]: every
        // warp-level store / load below is one contiguous 512-byte piece (thread = row would touch 32 lines per instruction)
        float* dst = p.ws_partial +
                     ((size_t)(tile - p.split_first) * (p.split_s - 1) + (w.part - 1)) * ((size_t)kTileM * BN) +
                     (size_t)((int)rank * 4 + ew) * (32 * BN) + lane * 4;
#pragma unroll 1
        for (int c = grp * 32; c < BN; c += 64) {
          uint32_t v[32];
          tmem_ld_32x32(taddr_p + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 8; ++g)
            __stcg(reinterpret_cast<float4*>(dst + (c / 32) * 1024 + g * 128),
                   make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]),
                               __uint_as_float(v[4 * g + 3])));
        }
        __threadfence();
        tc_fence_before_sync();
        if (PAIR) mbar_arrive_cluster(&tmem_empty[as], 0); else mbar_arrive(&tmem_empty[as]);
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");   // every thread's partials are fenced
        if (etid == 0) atomicAdd(p.ws_flags + ((tile - p.split_first) * 2 + (int)rank) * kSplitFlagStride, 1u);
        continue;
      }
      const int row = m0 + row_in_tile;
      const bool row_ok = row < p.M;
      // bias staging is double buffered by accumulator stage; with a folded LayerNorm the second buffer holds the column
      // sums instead, so the pair is single buffered and one more barrier keeps this tile's writes behind the previous
      // tile's reads
      const bool ln = p.ln_sum_in != nullptr;
      const int bsel = ln ? 0 : as;
      const uint32_t sb = smem_u32(sbias + bsel * 256);
      if (ln) asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      if (p.bias != nullptr) {  // stage this tile's bias slice once (broadcast LDS later instead of exposed LDG latency)
        for (int i = etid; i < BN; i += kEpiThreads) sbias[bsel * 256 + i] = (n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
      }
      if (ln) {
        for (int i = etid; i < BN; i += kEpiThreads) sbias[256 + i] = (n0 + i < p.N) ? __ldg(p.ln_colsum + n0 + i) : 0.f;
      }
      float rs = (p.row_scale != nullptr && row_ok) ? p.row_scale[row] : 1.f;
      float mu = 0.f;
      if (p.rms_sumsq_in != nullptr && row_ok) {
        // sum the producer's per-32-column partials in a fixed order (deterministic: no atomics anywhere)
        const int nvec = (p.rms_nparts % 4 == 0) ? p.rms_nparts : 0;   // rows are 16-byte aligned only then
        auto sum_parts = [&](const float* pp) {
          float q = 0.f;
          int i = 0;
          for (; i + 4 <= nvec; i += 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(pp + i);
            q += (t4.x + t4.y) + (t4.z + t4.w);
          }
          for (; i < p.rms_nparts; ++i) q += pp[i];
          return q;
        };
        float q = sum_parts(p.rms_sumsq_in + (int64_t)row * p.rms_nparts) * p.rms_inv_dim;   // E[x^2]
        if (ln) {
          mu = sum_parts(p.ln_sum_in + (int64_t)row * p.rms_nparts) * p.rms_inv_dim;
          q = fmaxf(q - mu * mu, 0.f);                                                        // variance
        }
        rs *= rsqrtf(q + p.rms_eps);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after_sync();
      const float* part_src = nullptr;
      if (w.part == 0) {
        // owner of a split tile: the other K-slices run concurrently on other CTAs of this (last) round
        // ONE thread polls (a line of its own per counter): hundreds of pollers on one L2 line starve the arriving atomics
        if (etid == 0) {
          const unsigned* flag = p.ws_flags + ((tile - p.split_first) * 2 + (int)rank) * kSplitFlagStride;
          unsigned spins = 0;
          while (ld_acquire_gpu(flag) < (unsigned)(p.split_s - 1)) {
            if (++spins > (1u << 24)) { asm volatile("trap;"); }
            __nanosleep(200);
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        part_src = p.ws_partial + (size_t)(tile - p.split_first) * (p.split_s - 1) * ((size_t)kTileM * BN) +
                   (size_t)((int)rank * 4 + ew) * (32 * BN) + lane * 4;
      }
      if (etr) g_gemm_trace[8 * it + 5] = clock64();
      const uint32_t taddr = tmem_base + as * Cfg::kAccStride + ((uint32_t)(ew * 32) << 16);
      const int rbase = m0 + ew * 32;  // first row of this warp's 32-row block

#pragma unroll 1
      for (int sp = grp; sp * 64 < BN; sp += 2) {
        const int c0 = sp * 64;                       // first accumulator column of the span inside the tile
        const int col0 = n0 + c0;
        if (col0 >= p.N) break;                       // warp-uniform
        const int span = min(min(64, BN - c0), p.N - col0);   // 8..64 valid accumulator columns (multiple of 8)
        uint32_t v[32];
        float ssq = 0.f, ssum = 0.f;                  // sum of squares / sum of this thread's (row's) outputs in the span
        const bool ptr_ = etr && sp == 0;             // phase trace of this warp's first span
        if (ptr_) g_gemm_trace[64 + 8 * it] = clock64();
        tmem_ld_32x32(taddr + c0, v);
        // residual block -> smem (coalesced: 8 lanes x 16 B per row)
        if (res != nullptr) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = t_row + 4 * i;
            const int gr = rbase + rr, gc = col0 + t_chunk * 8;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (gr < p.M && t_chunk * 8 < span) val = *reinterpret_cast<const uint4*>(res + (int64_t)gr * p.ldr + gc);
            sts128(stg + rr * 128 + ((t_chunk ^ (rr & 7)) << 4), val);
          }
          __syncwarp();
        }
        if (ptr_) g_gemm_trace[64 + 8 * it + 1] = clock64();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {              // two 32-column halves of the span
          if (hf * 32 >= span) break;                 // warp-uniform
          tmem_ld_wait();
          if (ptr_ && hf == 0) g_gemm_trace[64 + 8 * it + 2] = clock64();
          float x[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
          if (hf == 0 && span > 32) tmem_ld_32x32(taddr + c0 + 32, v);   // prefetch the second half
          if (part_src != nullptr) {     // split tile: add the other K-slices in a fixed order (deterministic)
            for (int pp = 0; pp < p.split_s - 1; ++pp) {
              const float* src = part_src + (size_t)pp * ((size_t)kTileM * BN) + ((c0 + hf * 32) / 32) * 1024;
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 t4 = __ldcg(reinterpret_cast<const float4*>(src + g * 128));
                x[4 * g] += t4.x; x[4 * g + 1] += t4.y; x[4 * g + 2] += t4.z; x[4 * g + 3] += t4.w;
              }
            }
          }
          if (ln) {   // folded LayerNorm: remove the row mean's contribution mu * sum_k W'[n,k] before the 1/std scale
            const uint32_t sc = smem_u32(sbias + 256);
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 cq = lds_f4(sc + (c0 + hf * 32 + g * 4) * 4);
              x[g * 4 + 0] = fmaf(-mu, cq.x, x[g * 4 + 0]); x[g * 4 + 1] = fmaf(-mu, cq.y, x[g * 4 + 1]);
              x[g * 4 + 2] = fmaf(-mu, cq.z, x[g * 4 + 2]); x[g * 4 + 3] = fmaf(-mu, cq.w, x[g * 4 + 3]);
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] *= rs;
          if (p.bias != nullptr) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 bq = lds_f4(sb + (c0 + hf * 32 + g * 4) * 4);
              x[g * 4 + 0] += bq.x; x[g * 4 + 1] += bq.y; x[g * 4 + 2] += bq.z; x[g * 4 + 3] += bq.w;
            }
          }
          if (swiglu) {
            // accumulator columns interleave (gate, up): 32 columns -> 16 outputs = staging chunks 2*hf, 2*hf+1
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              uint32_t o[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float g0 = x[g * 16 + 4 * j + 0], u0 = x[g * 16 + 4 * j + 1];
                const float g1 = x[g * 16 + 4 * j + 2], u1 = x[g * 16 + 4 * j + 3];
                o[j] = pack_bf16(fast_sigmoid_mul(g0, 1.4426950408889634f) * u0, fast_sigmoid_mul(g1, 1.4426950408889634f) * u1);
              }
              sts128(my_row + (((hf * 2 + g) ^ sw) << 4), make_uint4(o[0], o[1], o[2], o[3]));
            }
            continue;
          }
          // one branch per activation (warp-uniform), each with a fully unrolled body
          if (p.act == VL2_ACT_QUICK_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = fast_sigmoid_mul(x[j], 1.702f * 1.4426950408889634f);
          } else if (p.act == VL2_ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = fast_sigmoid_mul(x[j], 1.4426950408889634f);
          } else if (p.act == VL2_ACT_GELU_ERF) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = 0.5f * x[j] * (1.f + erff(x[j] * 0.70710678118654752f));
          } else if (p.act == VL2_ACT_GELU_TANH) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = gelu_tanh(x[j]);
          }
          if (p.rope_tab != nullptr && col0 + hf * 32 < p.rope_cols && row_ok) {
            // 32 accumulator columns = 16 rotation pairs of one head, frequencies i0 .. i0 + 15, angle of this row's position
            const int i0 = ((col0 + hf * 32) % p.rope_D) >> 1;
            const uint4* tr = reinterpret_cast<const uint4*>(p.rope_tab + (int64_t)(p.rope_pos0 + row) * (p.rope_D >> 1) + i0);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 cs = __ldg(tr + g);
              const uint32_t e[4] = {cs.x, cs.y, cs.z, cs.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float c = bf16_lo(e[j]), sn = bf16_hi(e[j]);
                const float a = x[g * 8 + 2 * j], b = x[g * 8 + 2 * j + 1];
                x[g * 8 + 2 * j] = a * c - b * sn;
                x[g * 8 + 2 * j + 1] = b * c + a * sn;
              }
            }
          }
          if (res != nullptr) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 r = lds128(my_row + (((hf * 4 + g) ^ sw) << 4));
              x[g * 8 + 0] += bf16_lo(r.x); x[g * 8 + 1] += bf16_hi(r.x);
              x[g * 8 + 2] += bf16_lo(r.y); x[g * 8 + 3] += bf16_hi(r.y);
              x[g * 8 + 4] += bf16_lo(r.z); x[g * 8 + 5] += bf16_hi(r.z);
              x[g * 8 + 6] += bf16_lo(r.w); x[g * 8 + 7] += bf16_hi(r.w);
            }
          }
          if (p.out_f32) {   // fp32 logits: rare, written directly (row-per-thread)
            if (row_ok) {
              float* crow = reinterpret_cast<float*>(p.C) + (int64_t)row * p.ldc + col0 + hf * 32;
#pragma unroll
              for (int g = 0; g < 8; ++g)
                if (hf * 32 + g * 4 < span)
                  *reinterpret_cast<float4*>(crow + g * 4) = make_float4(x[g * 4], x[g * 4 + 1], x[g * 4 + 2], x[g * 4 + 3]);
            }
          } else {
            // own row -> smem (each thread overwrites only the chunks it just read its residual from)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 pk = make_uint4(pack_bf16(x[g * 8], x[g * 8 + 1]), pack_bf16(x[g * 8 + 2], x[g * 8 + 3]),
                                          pack_bf16(x[g * 8 + 4], x[g * 8 + 5]), pack_bf16(x[g * 8 + 6], x[g * 8 + 7]));
              sts128(my_row + (((hf * 4 + g) ^ sw) << 4), pk);
              if (p.sumsq_out != nullptr && hf * 32 + g * 8 < span) {   // statistics of what the consumer will read
                const float a0 = bf16_lo(pk.x), a1 = bf16_hi(pk.x), a2 = bf16_lo(pk.y), a3 = bf16_hi(pk.y);
                const float a4 = bf16_lo(pk.z), a5 = bf16_hi(pk.z), a6 = bf16_lo(pk.w), a7 = bf16_hi(pk.w);
                ssq += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3 + a4 * a4 + a5 * a5 + a6 * a6 + a7 * a7;
                if (p.rowsum_out != nullptr) ssum += ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
              }
            }
          }
          if (ptr_) g_gemm_trace[64 + 8 * it + 3 + hf] = clock64();
        }
        if (p.sumsq_out != nullptr && row_ok && !swiglu && !p.out_f32) {
          // one slot per 32 output columns: this span owns slots col0/32 (its sum) and col0/32 + 1 (zero)
          float* slot = p.sumsq_out + (int64_t)row * (p.N >> 5) + (col0 >> 5);
          slot[0] = ssq;
          if (span > 32) slot[1] = 0.f;
          if (p.rowsum_out != nullptr) {
            float* slot2 = p.rowsum_out + (int64_t)row * (p.N >> 5) + (col0 >> 5);
            slot2[0] = ssum;
            if (span > 32) slot2[1] = 0.f;
          }
        }
        if (!p.out_f32) {
          __syncwarp();
          // smem -> global, full lines.  Output span: `span` columns (or span/2 for SwiGLU).
          const int out_span = swiglu ? (span >> 1) : span;
          const int out_col0 = swiglu ? (col0 >> 1) : col0;
          __nv_bfloat16* cbase = reinterpret_cast<__nv_bfloat16*>(p.C);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = t_row + 4 * i;
            int gr = rbase + rr;
            if (p.conv_C > 0) {   // padded (to, ho16, wo16) row -> real output row, or nothing for the padding positions
              const int wo = gr % kConvLine, ho = (gr / kConvLine) % kConvLine, to = gr / (kConvLine * kConvLine);
              gr = (wo < p.conv_Wo && ho < p.conv_Ho && to < p.conv_To) ? (to * p.conv_Ho + ho) * p.conv_Wo + wo : p.M;
            }
            if (gr < p.M && t_chunk * 8 < out_span) {
              const uint4 val = lds128(stg + rr * 128 + ((t_chunk ^ (rr & 7)) << 4));
              const int64_t off = (int64_t)gr * p.ldc + out_col0 + t_chunk * 8;
              *reinterpret_cast<uint4*>(cbase + off) = val;
              if (p.mc != nullptr) {
                multimem_st128(p.mc + off, val);
              } else {
                for (int pi = 0; pi < p.n_bcast; ++pi) *reinterpret_cast<uint4*>(p.bcast[pi] + off) = val;
              }
            }
          }
          __syncwarp();  // the buffer is reused by the next span
        }
        if (ptr_) g_gemm_trace[64 + 8 * it + 5] = clock64();
      }
      if (etr) g_gemm_trace[8 * it + 6] = clock64();
      if (w.part == 0) {   // every thread has consumed the partials: re-arm the counter for the next launch
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        if (etid == 0) p.ws_flags[((tile - p.split_first) * 2 + (int)rank) * kSplitFlagStride] = 0u;
      }
      // all tcgen05.ld of this thread have completed (wait::ld above) -> hand the accumulator stage back
      tmem_ld_wait();
      tc_fence_before_sync();
      if (PAIR) mbar_arrive_cluster(&tmem_empty[as], 0); else mbar_arrive(&tmem_empty[as]);
    }
  }

  tc_fence_before_sync();
  if (PAIR) cluster_sync_all(); else __syncthreads();   // nobody may still be using the pair's TMEM / barriers
  if (warp == 2) {
    tc_fence_after_sync();
    if (PAIR) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

static bool splitk_enabled();
static bool lean_enabled();
static bool wide_enabled();

static bool lean_eligible(const vl2_gemm_args* a) {   // launches the lean epilogue can serve (modulo split-K / SwiGLU tile width)
  return lean_enabled() && a->reserved4 != 1 && a->conv_C == 0 && a->n_bcast == 0 && a->mc_out == nullptr && !a->out_f32 &&
         a->ln_sum_in == nullptr;
}

template <int BN, bool PAIR, int BN2 = 0>
static int launch_gemm(const vl2_gemm_args* a, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, PAIR, BN2>;
  CUtensorMap ta, tb, tb2;
  const bool conv = a->conv_C > 0;
  int conv_To = 0, conv_Ho = 0, conv_Wo = 0;
  if (conv) {
    // x [T,H,W,C] viewed as [T, H/2, 2, W/2, 2C]: a tap's stride-2 walks along H and W are unit walks over the coarse
    // indices at fixed parities (the W parity is folded into the channel coordinate: offset parity*C of the merged 2C axis).
    // Box = 64 channels x 16 coarse columns (one padded output line) x 8 coarse rows (the tile's 8 output lines).
    const int p_ = a->conv_pad;
    conv_To = (a->conv_T + 2 * p_ - 2) / 2 + 1;
    conv_Ho = (a->conv_H + 2 * p_ - 2) / 2 + 1;
    conv_Wo = (a->conv_W + 2 * p_ - 2) / 2 + 1;
    // odd H / W without padding (SigLIP 27 x 27, stc_connector_v35): the last coarse index would pair the last pixel with
    // the first one of the next row / frame; no stored output needs it, so it is declared out of bounds (zero fill)
    const uint64_t wcols = (a->conv_W % 2 == 1 && p_ == 0) ? (uint64_t)a->conv_W / 2 : (uint64_t)(a->conv_W + 1) / 2;
    const uint64_t hrows = (a->conv_H % 2 == 1 && p_ == 0) ? (uint64_t)a->conv_H / 2 : (uint64_t)(a->conv_H + 1) / 2;
    const uint64_t C_ = (uint64_t)a->conv_C, W_ = (uint64_t)a->conv_W, H_ = (uint64_t)a->conv_H;
    uint64_t dims[5] = {2 * C_, wcols, 2, hrows, (uint64_t)a->conv_T};
    uint64_t str[4] = {2 * C_ * 2, W_ * C_ * 2, 2 * W_ * C_ * 2, H_ * W_ * C_ * 2};
    uint32_t box[5] = {BK, (uint32_t)kConvLine, 1, (uint32_t)(BM / kConvLine), 1};
    int rc = make_tmap_bf16(&ta, a->A, 5, dims, str, box);
    if (rc) return rc;
  } else {
    uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->M};
    uint64_t str[1] = {(uint64_t)a->lda * 2};
    uint32_t box[2] = {BK, BM};
    int rc = make_tmap_bf16(&ta, a->A, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->N};
    uint64_t str[1] = {(uint64_t)a->ldw * 2};
    uint32_t box[2] = {BK, (uint32_t)Cfg::kRowsB};
    int rc = make_tmap_bf16(&tb, a->W, 2, dims, str, box);
    if (rc) return rc;
    tb2 = tb;
    if (BN2 > 0) {
      uint32_t box2[2] = {BK, (uint32_t)Cfg::kRowsB2};
      rc = make_tmap_bf16(&tb2, a->W, 2, dims, str, box2);
      if (rc) return rc;
    }
  }
  GemmParams p;
  p.C = a->C; p.bias = a->bias; p.residual = a->residual; p.row_scale = a->row_scale;
  p.ldc = a->ldc; p.ldr = a->ldr; p.M = a->M; p.N = a->N; p.K = a->K; p.act = a->act; p.out_f32 = a->out_f32;
  p.rms_sumsq_in = a->rms_sumsq_in; p.sumsq_out = a->sumsq_out; p.rms_nparts = a->rms_nparts;
  p.ln_sum_in = a->ln_sum_in; p.ln_colsum = a->ln_colsum; p.rowsum_out = a->rowsum_out;
  p.rms_inv_dim = a->rms_inv_dim; p.rms_eps = a->rms_eps;
  p.trace = (a->reserved2 == 777) ? 1 : 0;
  p.n_bcast = a->n_bcast;
  p.mc = reinterpret_cast<__nv_bfloat16*>(a->mc_out);
  for (int i = 0; i < 8; ++i) p.bcast[i] = reinterpret_cast<__nv_bfloat16*>(i < a->n_bcast ? a->bcast_out[i] : nullptr);
  const int tile_m = PAIR ? 2 * BM : BM;
  p.rope_tab = a->rope_tab; p.rope_cols = a->rope_cols; p.rope_D = a->rope_D; p.rope_pos0 = a->rope_pos0;
  p.conv_C = conv ? a->conv_C : 0; p.conv_pad = a->conv_pad; p.conv_Ho = conv_Ho; p.conv_Wo = conv_Wo; p.conv_To = conv_To;
  const int m_rows = conv ? conv_To * kConvLine * kConvLine : a->M;     // rows of the (padded) A enumeration
  p.num_m_tiles = (m_rows + tile_m - 1) / tile_m;
  p.num_n_tiles = (a->N + Cfg::kTileN - 1) / Cfg::kTileN;
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int slots = PAIR ? sm_count() / 2 : sm_count();
  // Split-K of the last, partial round (wave quantisation): the `rem` tiles left for `slots` CTAs (pairs) are cut into s
  // K-slices each, so the round costs a fraction of a tile time.  Slice 0 owns the tile: it adds the other slices' fp32
  // partials (workspace, fixed order -> deterministic) and runs the epilogue.  Needs the caller's workspace (splitk_ws).
  p.split_first = tiles;
  p.split_s = 1;
  p.ws_flags = nullptr;
  p.ws_partial = nullptr;
  {
    const int num_kb = (a->K + BK - 1) / BK;
    const int rem = tiles % slots;
    // s slices per remaining tile; the rem * s items may need more than one sub-round: pick the s with the shortest tail
    // ceil(rem * s / slots) / s (in tile times; 1.0 = unsplit).  Items are dealt out in index order, an owner only waits for
    // higher-numbered items and an item is only queued behind a lower-numbered one, so the waits cannot form a cycle.
    int s = 1;
    if (rem > 0) {
      // cost of the tail round in cycles: ~550 per k-block (profiles/r01_gemm_tile_trace.txt) + the measured cost of the
      // partials' round trip through L2, the owner's wait and its longer epilogue; a split must beat the unsplit round by 10 %
      const double tile_cycles = 550.0 * num_kb;
      double best = 0.9 * tile_cycles;
      for (int c = 2; c <= 4; ++c) {
        if (num_kb / c < 8) break;
        const double tail = (double)((rem * c + slots - 1) / slots) / c;
        const double cost = tail * tile_cycles + kSplitOverheadCycles;
        if (cost < best) { best = cost; s = c; }
      }
    }
    const bool forced = rem > 0 && a->reserved3 >= 2 && a->reserved3 <= 4 && num_kb >= a->reserved3;   // test hook
    if (forced) s = a->reserved3;
    const size_t need = kSplitFlagBytes + (size_t)rem * (s > 1 ? s - 1 : 0) * tile_m * BN * sizeof(float);
    if (BN2 == 0 && s > 1 && (splitk_enabled() || forced) && a->splitk_ws != nullptr && (size_t)a->splitk_ws_bytes >= need && (size_t)2 * rem * kSplitFlagStride * 4 <= kSplitFlagBytes &&
        aligned16(a->splitk_ws)) {
      p.split_first = tiles - rem;
      p.split_s = s;
      p.ws_flags = reinterpret_cast<unsigned*>(a->splitk_ws);
      p.ws_partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a->splitk_ws) + kSplitFlagBytes);
    }
  }
  p.num_items = p.split_first + (tiles - p.split_first) * p.split_s;
  const int units = p.num_items < slots ? p.num_items : slots;
  // The lean epilogue (TMA stores / TMA residual loads, see epilogue_lean) serves every launch that does not need the
  // general one; reserved4 == 1 forces the general kernel (test hook: both must agree).
  const bool swiglu = a->act == VL2_ACT_SWIGLU;
  const bool lean = lean_eligible(a) && p.split_s == 1 && !(swiglu && BN % 64 != 0);
  CUtensorMap tc = ta, tr = ta;   // placeholders when unused (never dereferenced)
  if (lean) {
    // output / residual as [32 rows x 32 columns] boxes, 64-byte swizzle: one box = one staging block of an epilogue warp
    uint64_t dims[2] = {(uint64_t)(swiglu ? a->N / 2 : a->N), (uint64_t)a->M};
    uint64_t str[1] = {(uint64_t)a->ldc * 2};
    uint32_t box[2] = {32, 32};
    int rc = make_tmap_bf16(&tc, a->C, 2, dims, str, box, 64);
    if (rc) return rc;
    if (a->residual != nullptr) {
      uint64_t strr[1] = {(uint64_t)a->ldr * 2};
      rc = make_tmap_bf16(&tr, a->residual, 2, dims, strr, box, 64);
      if (rc) return rc;
    }
  }
  const bool plain = a->act == VL2_ACT_NONE && a->rope_tab == nullptr;   // epilogue without activation / SwiGLU / RoPE bodies
  const dim3 grid(PAIR ? 2 * units : units), block(kGemmThreads);
  const int cl = PAIR ? 2 : 1;
#define VL2_LAUNCH_GEMM(...)                                                                                        \
  do {                                                                                                              \
    VL2_SMEM_OPT_IN((gemm_bf16_tcgen05_kernel<__VA_ARGS__>), Cfg::kSmemBytes);                                      \
    VL2_CHECK_CUDA(launch_kernel(gemm_bf16_tcgen05_kernel<__VA_ARGS__>, grid, block, Cfg::kSmemBytes, stream, cl, ta, tb, tb2, \
                                 tc, tr, p));                                                                       \
  } while (0)
  if constexpr (BN2 > 0) {
    VL2_REQUIRE(lean && !swiglu, VL2_E_UNSUPPORTED, "vl2_gemm_bf16: the wide tile serves lean, non-SwiGLU launches only");
    if (plain) VL2_LAUNCH_GEMM(BN, PAIR, true, BN2, true);
    else VL2_LAUNCH_GEMM(BN, PAIR, true, BN2, false);
  } else if (lean) {
    if (plain) VL2_LAUNCH_GEMM(BN, PAIR, true, 0, true);
    else VL2_LAUNCH_GEMM(BN, PAIR, true, 0, false);
  } else {
    VL2_LAUNCH_GEMM(BN, PAIR, false, 0, false);
  }
#undef VL2_LAUNCH_GEMM
  VL2_CHECK_LAUNCH("gemm_bf16_tcgen05_kernel");
  return VL2_OK;
}

// Tile choice.  Cost model fitted to ncu launch times (profiles/r01_launches_v4_summary.txt):
//   * one 128 x BN x 16 MMA occupies the tensor pipe for ~BN/2 cycles (same per CTA for a cta_group::2 pair);
//   * every k-block moves A (16 KB) + the staged part of B into smem and out again at ~190 B/cycle combined;
//   * the smem ring must cover the L2/HBM latency (~3000 cycles): a k-block cannot retire faster than L / stages
//     (this is what holds the 4-stage 128x256 single-CTA tile at ~750 cycles per k-block = 83 % tensor-active);
//   * a launch costs waves x tile time + one un-overlapped epilogue.
struct TileChoice { int bn; bool pair; int bn2; };
static constexpr int kWideBN = 224, kWideBN2 = 192;   // the wide tile: 416 columns in two accumulators
static constexpr double kIngestBytesPerClk = 58.0;    // L2 -> smem per SM, measured (scripts/vit_shard_probe.py --sweep)

static TileChoice choose_tile(int M, int N, int K, int sms, bool allow_pair, bool allow_split, bool allow_wide) {
  const int kb = (K + BK - 1) / BK;
  const double L = 3000.0;
  const int extra = 8 * 4096 + 2048 + 512;
  static const int cands[7] = {256, 224, 192, 160, 128, 96, 64};
  TileChoice best = {256, false, 0};
  double best_cost = -1;
  for (int pair = (allow_pair ? 1 : 0); pair >= 0; --pair) {
    for (int i = 0; i < 7; ++i) {
      const int bn = cands[i];
      if (pair && bn < 128) continue;
      if (bn > 64 && N <= bn - 32) continue;  // do not pick a tile much wider than the matrix
      const int tile_m = pair ? 256 : 128;
      const long tiles = (long)((M + tile_m - 1) / tile_m) * ((N + bn - 1) / bn);
      const long slots = pair ? sms / 2 : sms;
      const long full = tiles / slots, rem = tiles % slots;
      const double stage_bytes = 16384.0 + (pair ? bn / 2 : bn) * 128.0;
      int stages = (int)((227 * 1024 - extra) / stage_bytes);
      if (stages > 8) stages = 8;
      const double mma = 2.0 * bn;
      const double smem = 2.0 * stage_bytes / 190.0;
      const double lat = L / stages;
      double per_kb = mma > smem ? mma : smem;
      if (lat > per_kb) per_kb = lat;
      // the last, partial round costs a whole tile time, or a fraction of it when launch_gemm can split it along K
      double tail = rem > 0 ? 1.0 : 0.0;
      if (allow_split && rem > 0) {
        for (int c = 2; c <= 4 && kb / c >= 8; ++c) {
          const double t = (double)((rem * c + slots - 1) / slots) / c + kSplitOverheadCycles / (550.0 * kb);
          if (t < tail * 0.9) tail = t;
        }
      }
      const double cost = ((double)full + tail) * (kb * per_kb + 600.0) + 2000.0 + bn * 16.0;
      if (best_cost < 0 || cost < best_cost * 0.98) { best_cost = cost; best.bn = bn; best.pair = pair != 0; }
    }
  }
  // The wide tile against the winner.  Both sides are costed with the operand-ingest bound here (a k-block cannot retire
  // faster than its bytes arrive: 16 KB of A + the CTA's share of B at ~58 B/clk - what holds the 256 x 224 / 256 x 256
  // pair tiles at 530 / 565 cycles per k-block although their MMAs need 448 / 512), which is the whole point of the wide
  // tile: 832 cycles of MMA per k-block for 416 columns on the same 16 KB of A.  Its epilogue is exposed on every tile.
  if (allow_wide && allow_pair && best.pair && N >= kWideBN + kWideBN2 / 2) {
    const long slots = sms / 2;
    const long mt = (M + 255) / 256;
    const long tiles_n = mt * ((N + best.bn - 1) / best.bn);
    const double ing_n = (16384.0 + best.bn * 64.0) / kIngestBytesPerClk;
    const double per_n = ing_n > 2.0 * best.bn ? ing_n : 2.0 * best.bn;
    const double cost_n = (double)((tiles_n + slots - 1) / slots) * (kb * per_n + 600.0) + 2000.0 + best.bn * 16.0;
    const int wn = kWideBN + kWideBN2;
    const long tiles_w = mt * ((N + wn - 1) / wn);
    const double per_w = 2.0 * wn;                       // MMA-bound (ingest: (16384 + 64 * 416) / 58 = 741)
    const double epi_w = 12000.0;                        // 13 units of 32 columns over two warps per lane quarter
    const double cost_w = (double)((tiles_w + slots - 1) / slots) * (kb * per_w + 600.0 + epi_w) + 2000.0;
    if (cost_w < 0.97 * cost_n) { best.bn = kWideBN; best.bn2 = kWideBN2; best.pair = true; }
  }
  return best;
}

// VL2_GEMM_SPLITK=1 enables the split-K tail round (needs the caller's workspace).  Default OFF: measured on the 7B shapes
// it only pays for K = 14336 (down_proj 162 -> 157 us; the partials' round trip costs ~16 k cycles, half a K = 4096 tile),
// i.e. 0.4 % of a step, and it makes a GEMM's rounding depend on M (how many tiles land in the last round), which would
// end the bit-exact frame-sharding / frame-independence properties of the vision tower (tests/test_fullsize_gpu.py).
// (An L2 prefetch cursor in the TMA producer - cp.async.bulk.prefetch.tensor 6/12/24 k-blocks ahead - was measured 20-30 %
// slower on every shape and removed: profiles/experiments/gemm_l2_prefetch.txt.)
static bool splitk_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VL2_GEMM_SPLITK");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// VL2_GEMM_LEAN=0 routes every launch through the general epilogue (A/B measurements; both are parity-tested).
static bool lean_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VL2_GEMM_LEAN");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}

// VL2_GEMM_WIDE=0 keeps the cost model away from the wide (two-accumulator) tile (A/B measurements)
static bool wide_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VL2_GEMM_WIDE");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}

static bool pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VL2_GEMM_PAIR");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}

}  // namespace vl2

// Host-side planning only (no launch, no device work): which tile the cost model picks for a problem and how the
// persistent grid will be filled.  Lets tests pin the scheduler's decisions for the shapes of the path on a CPU box.
extern "C" int vl2_gemm_plan(int M, int N, int K, int with_splitk_ws, int32_t* out6) {
  using namespace vl2;
  VL2_REQUIRE(M > 0 && N > 0 && K > 0 && out6 != nullptr, VL2_E_BADSHAPE, "vl2_gemm_plan: bad arguments");
  const int sms = sm_count();
  // (planned for a launch the lean epilogue can serve: bias / activation / residual / statistics, bf16 output)
  const TileChoice t = choose_tile(M, N, K, sms, pair_enabled(), splitk_enabled() && with_splitk_ws != 0, wide_enabled());
  const int tile_m = t.pair ? 2 * BM : BM;
  const int tiles = ((M + tile_m - 1) / tile_m) * ((N + t.bn + t.bn2 - 1) / (t.bn + t.bn2));
  const int slots = t.pair ? sms / 2 : sms;
  out6[0] = t.bn + t.bn2;
  out6[1] = t.pair ? 1 : 0;
  out6[2] = tiles;
  out6[3] = slots;
  out6[4] = (tiles + slots - 1) / slots;   // rounds of the persistent loop
  out6[5] = sms;
  return VL2_OK;
}

// Debug: copy the tile-boundary cycle trace of the last traced launch (args->reserved2 == 777) to host memory.
extern "C" int vl2_debug_gemm_trace(long long* host_out128) {
  VL2_REQUIRE(host_out128 != nullptr, VL2_E_BADSHAPE, "vl2_debug_gemm_trace: null output");
  VL2_CHECK_CUDA(cudaDeviceSynchronize());
  VL2_CHECK_CUDA(cudaMemcpyFromSymbol(host_out128, vl2::g_gemm_trace, 128 * sizeof(long long)));
  return VL2_OK;
}

extern "C" int vl2_gemm_bf16(const vl2_gemm_args* a, void* stream) {
  using namespace vl2;
  VL2_REQUIRE(a != nullptr, VL2_E_BADSHAPE, "vl2_gemm_bf16: null args");
  VL2_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, VL2_E_BADSHAPE, "vl2_gemm_bf16: M,N,K must be positive (%d,%d,%d)",
              a->M, a->N, a->K);
  VL2_REQUIRE(a->K % 8 == 0 && a->N % 8 == 0, VL2_E_BADSHAPE, "vl2_gemm_bf16: K and N must be multiples of 8 (%d,%d)",
              a->K, a->N);
  VL2_REQUIRE(a->lda % 8 == 0 && a->ldw % 8 == 0 && a->ldc % 8 == 0 && (a->residual == nullptr || a->ldr % 8 == 0),
              VL2_E_BADALIGN, "vl2_gemm_bf16: leading dimensions must be multiples of 8 elements");
  VL2_REQUIRE(a->lda >= a->K && a->ldw >= a->K, VL2_E_BADSHAPE, "vl2_gemm_bf16: lda/ldw smaller than K");
  VL2_REQUIRE(aligned16(a->A) && aligned16(a->W) && aligned16(a->C) && aligned16(a->residual) && aligned16(a->bias),
              VL2_E_BADALIGN, "vl2_gemm_bf16: pointers must be 16-byte aligned");
  VL2_REQUIRE(a->act >= VL2_ACT_NONE && a->act <= VL2_ACT_GELU_TANH, VL2_E_UNSUPPORTED, "vl2_gemm_bf16: unknown act %d",
              a->act);
  VL2_REQUIRE(a->sumsq_out == nullptr || (!a->out_f32 && a->act != VL2_ACT_SWIGLU && a->N % 32 == 0), VL2_E_UNSUPPORTED,
              "vl2_gemm_bf16: sumsq_out supports bf16, non-SwiGLU outputs with N %% 32 == 0 only");
  VL2_REQUIRE(a->rms_sumsq_in == nullptr || a->rms_nparts > 0, VL2_E_BADSHAPE, "vl2_gemm_bf16: rms_nparts must be positive");
  VL2_REQUIRE(a->ln_sum_in == nullptr || (a->rms_sumsq_in != nullptr && a->ln_colsum != nullptr), VL2_E_BADSHAPE,
              "vl2_gemm_bf16: ln_sum_in needs rms_sumsq_in (sums of squares) and ln_colsum");
  VL2_REQUIRE(a->rowsum_out == nullptr || a->sumsq_out != nullptr, VL2_E_UNSUPPORTED,
              "vl2_gemm_bf16: rowsum_out is written together with sumsq_out");
  VL2_REQUIRE(a->n_bcast >= 0 && a->n_bcast <= 8, VL2_E_BADSHAPE, "vl2_gemm_bf16: n_bcast must be in [0,8]");
  if (a->n_bcast > 0 || a->mc_out != nullptr) {
    VL2_REQUIRE(!a->out_f32 && a->act != VL2_ACT_SWIGLU, VL2_E_UNSUPPORTED,
                "vl2_gemm_bf16: the broadcast epilogue supports bf16, non-SwiGLU outputs only");
    for (int i = 0; i < a->n_bcast; ++i)
      VL2_REQUIRE(a->bcast_out[i] != nullptr && aligned16(a->bcast_out[i]), VL2_E_BADALIGN,
                  "vl2_gemm_bf16: broadcast target %d must be a 16-byte aligned device pointer", i);
  }
  if (a->act == VL2_ACT_SWIGLU) {
    VL2_REQUIRE(a->residual == nullptr && !a->out_f32 && a->N % 16 == 0, VL2_E_UNSUPPORTED,
                "vl2_gemm_bf16: SWIGLU epilogue needs N %% 16 == 0, bf16 output and no residual");
  }
  if (a->rope_tab != nullptr) {
    VL2_REQUIRE(a->rope_D >= 32 && a->rope_D % 32 == 0 && a->rope_cols > 0 && a->rope_cols % a->rope_D == 0 &&
                    a->rope_cols <= a->N && a->rope_pos0 >= 0 && aligned16(a->rope_tab),
                VL2_E_BADSHAPE, "vl2_gemm_bf16: rope epilogue needs head width %% 32 == 0, rope_cols a multiple of it (<= N)");
    VL2_REQUIRE(!a->out_f32 && a->act == VL2_ACT_NONE && a->residual == nullptr, VL2_E_UNSUPPORTED,
                "vl2_gemm_bf16: the rope epilogue supports bf16 output without activation / residual");
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int m_plan = a->M;
  if (a->conv_C > 0) {
    const int p_ = a->conv_pad;
    VL2_REQUIRE(a->conv_T > 0 && a->conv_H > 0 && a->conv_W > 0 && (p_ == 0 || p_ == 1) && a->conv_C % 64 == 0,
                VL2_E_BADSHAPE, "vl2_gemm_bf16: conv front end needs T,H,W > 0, pad in {0,1}, C %% 64 == 0");
    const int To = (a->conv_T + 2 * p_ - 2) / 2 + 1, Ho = (a->conv_H + 2 * p_ - 2) / 2 + 1, Wo = (a->conv_W + 2 * p_ - 2) / 2 + 1;
    VL2_REQUIRE((a->conv_W % 2 == 0 && a->conv_H % 2 == 0) || p_ == 0, VL2_E_UNSUPPORTED,
                "vl2_gemm_bf16: conv front end: odd H / W needs pad 0");
    VL2_REQUIRE(To > 0 && Ho > 0 && Wo > 0 && Wo <= kConvLine && Ho <= kConvLine, VL2_E_UNSUPPORTED,
                "vl2_gemm_bf16: conv front end supports up to %d x %d output positions per time step (got %d x %d)", kConvLine,
                kConvLine, Ho, Wo);
    VL2_REQUIRE(a->K == 8 * a->conv_C && a->M == To * Ho * Wo, VL2_E_BADSHAPE,
                "vl2_gemm_bf16: conv front end needs K == 8*C and M == To*Ho*Wo (M=%d, expected %d)", a->M, To * Ho * Wo);
    VL2_REQUIRE(a->residual == nullptr && a->row_scale == nullptr && a->rms_sumsq_in == nullptr && a->sumsq_out == nullptr &&
                    !a->out_f32 && a->act != VL2_ACT_SWIGLU && a->n_bcast == 0 && a->mc_out == nullptr,
                VL2_E_UNSUPPORTED, "vl2_gemm_bf16: the conv front end supports bias + activation epilogues, bf16 output");
    m_plan = To * kConvLine * kConvLine;
  }
  const bool split_ok = splitk_enabled() && a->splitk_ws != nullptr && a->conv_C == 0;
  const bool wide_ok = wide_enabled() && lean_eligible(a) && a->act != VL2_ACT_SWIGLU && !split_ok && a->reserved3 == 0;
  TileChoice t = choose_tile(m_plan, a->N, a->K, sm_count(), pair_enabled(), split_ok, wide_ok);
  // test hook: reserved = BN forces a single-CTA tile width, 1000 + BN forces the cta_group::2 pair kernel
  // (2416 forces the wide 224 + 192 pair tile)
  if (a->reserved >= 64 && a->reserved <= 256 && a->reserved % 32 == 0) { t.bn = a->reserved; t.pair = false; t.bn2 = 0; }
  if (a->reserved >= 1128 && a->reserved <= 1256 && (a->reserved - 1000) % 32 == 0) { t.bn = a->reserved - 1000; t.pair = true; t.bn2 = 0; }
  if (a->reserved == 2000 + kWideBN + kWideBN2) {
    VL2_REQUIRE(lean_eligible(a) && a->act != VL2_ACT_SWIGLU && pair_enabled(), VL2_E_UNSUPPORTED,
                "vl2_gemm_bf16: the wide tile serves lean, non-SwiGLU launches only");
    t.bn = kWideBN; t.bn2 = kWideBN2; t.pair = true;
  }
  if (t.bn2 > 0) return launch_gemm<kWideBN, true, kWideBN2>(a, st);
  if (t.pair) {
    switch (t.bn) {
      case 256: return launch_gemm<256, true>(a, st);
      case 224: return launch_gemm<224, true>(a, st);
      case 192: return launch_gemm<192, true>(a, st);
      case 160: return launch_gemm<160, true>(a, st);
      default: return launch_gemm<128, true>(a, st);
    }
  }
  switch (t.bn) {
    case 256: return launch_gemm<256, false>(a, st);
    case 224: return launch_gemm<224, false>(a, st);
    case 192: return launch_gemm<192, false>(a, st);
    case 160: return launch_gemm<160, false>(a, st);
    case 128: return launch_gemm<128, false>(a, st);
    case 96: return launch_gemm<96, false>(a, st);
    default: return launch_gemm<64, false>(a, st);
  }
}
