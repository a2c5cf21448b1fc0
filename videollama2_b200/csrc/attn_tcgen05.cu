// FlashAttention-style fused attention for sm_100a (non-causal ViT heads and causal GQA decoder heads).
//
// One CTA = TWO 128-row query tiles (A, B) of one (batch, head), sharing every K/V tile.  320 threads:
//   warps 0..3  softmax group A, warps 4..7 softmax group B: thread r of a group owns query row r of its tile
//               (TMEM lane r) -> no cross-thread reductions; two warps per SM sub-partition hide each other's latency
//   warp 8      TMA producer (one lane): Q_A, Q_B once, then K tiles (2-stage ring) and V tiles (1 or 2 stages)
//   warp 9      TMEM allocator + MMA issuer (one lane), per key tile j and group g:
//                 S_g = Q_g K_j^T     tcgen05.mma 128x128x16, A/B K-major SW128, accumulator in TMEM
//                 O_g += P_g V_j      tcgen05.mma 128xDx16,   A = P_g (bf16, written to smem by the group's warps),
//                                     B = V tile as MN-major SW128 operand (no transpose pass); accumulates in TMEM
// The issue order  PV_A(j), QK_A(j+1), PV_B(j), QK_B(j+1)  staggers the groups: while group A runs its softmax the
// tensor core works for group B and vice versa.  Online-softmax state (m, l) lives in registers; O is rescaled in place
// in TMEM only when a row maximum grows by more than 2^8 (lazy rescale).  S is read from TMEM once per tile.
#include "host_common.h"
#include "ptx.cuh"

namespace vl2 {

static constexpr int kAttnThreads = 320;
static constexpr int BQ = 128;   // rows per query tile (two tiles per CTA)
static constexpr int BKV = 128;

template <int D>
struct AttnCfg {
  static constexpr int kAtoms = D / 64;                 // 64-column (128-byte) swizzle atoms per row
  static constexpr int kTileBytes = BKV * D * 2;        // one Q / K / V tile
  static constexpr int kAtomBytes = 128 * 128;          // 128 rows x 128 B
  static constexpr int kPBytes = BQ * BKV * 2;          // 32 KB per group
  static constexpr int kVStages = (D == 64) ? 2 : 1;    // D=128: 2*Q + 2*K + 1*V + 2*P = 224 KB
  static constexpr int kOffQ = 0;                       // [2] tiles
  static constexpr int kOffK = 2 * kTileBytes;          // [2] stages
  static constexpr int kOffV = 4 * kTileBytes;          // [kVStages]
  static constexpr int kOffP = (4 + kVStages) * kTileBytes;   // [2] groups
  static constexpr int kOffBar = kOffP + 2 * kPBytes;
  static constexpr int kSmemUsed = kOffBar + 256;
  // at least 116 KB so that only one CTA is resident per SM (each CTA allocates all 512 TMEM columns)
  static constexpr int kSmemBytes = kSmemUsed > 116 * 1024 ? kSmemUsed : 116 * 1024;
  static constexpr int kTmemCols = 512;
  static constexpr int kColS = 0;     // S_A at 0, S_B at 128
  static constexpr int kColO = 256;   // O_A at 256, O_B at 256 + D
};

struct AttnParams {
  void* out;
  int64_t ldo;
  int S, Hq, group;  // group = Hq / Hkv
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
  uint8_t* sPall = smem + Cfg::kOffP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2] per group
  uint64_t* p_full = bars + 11;   // [2] per group
  uint64_t* o_full = bars + 13;   // [2] per group
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // heavy (late) causal tile pairs first
  const int qp = p.causal ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int kvh = head / p.group;
  const int q0 = qp * 2 * BQ;
  const int n_tiles_total = (p.S + BKV - 1) / BKV;
  // key tiles each group needs (0 = the group's query tile lies entirely past the sequence)
  int n_kv[2];
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int qt = 2 * qp + g;
    n_kv[g] = (qt * BQ >= p.S) ? 0 : (p.causal ? (qt + 1) : n_tiles_total);
  }
  const int n_max = n_kv[0] > n_kv[1] ? n_kv[0] : n_kv[1];

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
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_base_ptr, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 8) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      mbar_arrive_expect_tx(q_full, 2 * Cfg::kTileBytes);
#pragma unroll
      for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int a = 0; a < Cfg::kAtoms; ++a)
          tma_load_3d(sQ + g * Cfg::kTileBytes + a * Cfg::kAtomBytes, &tmap_q, q_full, head * D + a * 64, q0 + g * BQ, b);
      for (int j = 0; j < n_max; ++j) {
        const int ks = j & 1;
        mbar_wait(&k_empty[ks], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[ks], Cfg::kTileBytes);
#pragma unroll
        for (int a = 0; a < Cfg::kAtoms; ++a)
          tma_load_3d(sK + ks * Cfg::kTileBytes + a * Cfg::kAtomBytes, &tmap_k, &k_full[ks], kvh * D + a * 64, j * BKV, b);
        const int vs = j % Cfg::kVStages;
        const uint32_t vph = (j / Cfg::kVStages) & 1;
        mbar_wait(&v_empty[vs], vph ^ 1);
        mbar_arrive_expect_tx(&v_full[vs], Cfg::kTileBytes);
#pragma unroll
        for (int a = 0; a < Cfg::kAtoms; ++a)
          tma_load_3d(sV + vs * Cfg::kTileBytes + a * Cfg::kAtomBytes, &tmap_v, &v_full[vs], kvh * D + a * 64, j * BKV, b);
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_qk = umma_idesc_bf16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, D, 0, 1);  // B (= V tile) is MN-major
      auto issue_qk = [&](int g, int j) {  // S_g = Q_g K_j^T ; K_j must have landed (caller waited)
        const uint32_t q_addr = smem_u32(sQ + g * Cfg::kTileBytes);
        const uint32_t k_addr = smem_u32(sK + (j & 1) * Cfg::kTileBytes);
        const uint32_t d_tmem = tmem_base + Cfg::kColS + g * 128;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk >> 2) * Cfg::kAtomBytes + (kk & 3) * 32;
          umma_bf16_ss(d_tmem, umma_desc_sw128(q_addr + off, 16, 1024), umma_desc_sw128(k_addr + off, 16, 1024),
                       idesc_qk, kk != 0);
        }
        umma_commit(&s_full[g]);
      };
      auto issue_pv = [&](int g, int j) {  // O_g += P_g V_j
        const uint32_t p_addr = smem_u32(sPall + g * Cfg::kPBytes);
        const uint32_t v_addr = smem_u32(sV + (j % Cfg::kVStages) * Cfg::kTileBytes);
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          const uint64_t da = umma_desc_sw128(p_addr + (kk >> 2) * Cfg::kAtomBytes + (kk & 3) * 32, 16, 1024);
          // MN-major B: 16 keys = 2 groups of 8 rows (1024 B each); next 64-wide d atom is kAtomBytes away (LBO)
          const uint64_t db = umma_desc_sw128(v_addr + kk * 2048, Cfg::kAtomBytes, 1024);
          umma_bf16_ss(tmem_base + Cfg::kColO + g * D, da, db, idesc_pv, (j | kk) != 0);
        }
        umma_commit(&o_full[g]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after_sync();
      if (n_kv[0] > 0) issue_qk(0, 0);
      if (n_kv[1] > 0) issue_qk(1, 0);
      umma_commit(&k_empty[0]);   // (K_0 is re-read by nobody else; stage reusable once these MMAs finish)
      for (int j = 0; j < n_max; ++j) {
        bool v_ready = false, k_ready = false;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (j < n_kv[g]) {
            mbar_wait(&p_full[g], j & 1);
            if (!v_ready) { mbar_wait(&v_full[j % Cfg::kVStages], (j / Cfg::kVStages) & 1); v_ready = true; }
            tc_fence_after_sync();
            issue_pv(g, j);
            if (j + 1 < n_kv[g]) {
              if (!k_ready) { mbar_wait(&k_full[(j + 1) & 1], ((j + 1) >> 1) & 1); k_ready = true; }
              tc_fence_after_sync();
              issue_qk(g, j + 1);
            }
          }
        }
        umma_commit(&v_empty[j % Cfg::kVStages]);           // V_j free once both groups' P V have read it
        if (j + 1 < n_max) umma_commit(&k_empty[(j + 1) & 1]);  // K_{j+1} free once both groups' Q K^T have read it
      }
    }
  } else {
    // ===================== softmax / output warps (thread == query row of its group's tile) =====================
    const int g = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    const int qi = q0 + g * BQ + r;
    const int nk = n_kv[g];
    uint8_t* sP = sPall + g * Cfg::kPBytes;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t s_taddr = tmem_base + lane_sel + Cfg::kColS + g * 128;
    const uint32_t o_taddr = tmem_base + lane_sel + Cfg::kColO + g * D;
    float m = -INFINITY, l = 0.f;
    constexpr float kRescaleThreshold = 8.f;

    for (int j = 0; j < nk; ++j) {
      const int kv0 = j * BKV;
      const bool need_mask = (kv0 + BKV > p.S) || (p.causal && (kv0 + BKV - 1 > q0 + g * BQ));
      mbar_wait(&s_full[g], j & 1);
      tc_fence_after_sync();
      uint32_t sv[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32(s_taddr + c * 32, sv[c]);
      tmem_ld_wait();
      if (need_mask) {  // warp-uniform; predicated selects, no per-element branches
        const int lim = p.causal ? min(p.S - 1, qi) : (p.S - 1);   // last visible key index for this row
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            sv[c][i] = (kv0 + c * 32 + i <= lim) ? sv[c][i] : 0xff800000u;  // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        mx0 = fmaxf(mx0, __uint_as_float(sv[0][i]));
        mx1 = fmaxf(mx1, __uint_as_float(sv[1][i]));
        mx2 = fmaxf(mx2, __uint_as_float(sv[2][i]));
        mx3 = fmaxf(mx3, __uint_as_float(sv[3][i]));
      }
      const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;
      // reference maximum for this tile: keep the old one unless it is too stale
      float m_use = m, alpha = 1.f;
      const bool grow = (m_tile > m + kRescaleThreshold) || (m == -INFINITY);
      if (grow) {
        m_use = (m_tile == -INFINITY) ? 0.f : m_tile;   // fully masked rows stay finite
        alpha = (m == -INFINITY) ? 0.f : fast_exp2(m - m_use);
      }
      // the previous P V must be complete before P (smem) is overwritten or O (TMEM) is rescaled
      if (j > 0) {
        mbar_wait(&o_full[g], (j - 1) & 1);
        tc_fence_after_sync();
        if (__any_sync(0xffffffffu, grow)) {
#pragma unroll
          for (int c = 0; c < D / 32; ++c) {
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
      // probabilities -> smem (bf16, K-major SW128 A operand), row sum
      float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float pr[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) pr[i] = fast_exp2(fmaf(__uint_as_float(sv[c][i]), p.scale_log2, -m_use));
#pragma unroll
        for (int i = 0; i < 32; i += 4) { rs0 += pr[i]; rs1 += pr[i + 1]; rs2 += pr[i + 2]; rs3 += pr[i + 3]; }
        // 32 columns = 4 chunks of 16 B inside atom (c >> 1), chunk index (c & 1) * 4 + q
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (c & 1) * 4 + q;
          sts128(smem_u32(sP) + (c >> 1) * Cfg::kAtomBytes + r * 128 + ((chunk ^ (r & 7)) << 4),
                 make_uint4(pack_bf16(pr[q * 8 + 0], pr[q * 8 + 1]), pack_bf16(pr[q * 8 + 2], pr[q * 8 + 3]),
                            pack_bf16(pr[q * 8 + 4], pr[q * 8 + 5]), pack_bf16(pr[q * 8 + 6], pr[q * 8 + 7])));
        }
      }
      l = l * alpha + ((rs0 + rs1) + (rs2 + rs3));
      m = m_use;
      // make the generic-proxy smem writes visible to the tensor core (async proxy), then signal
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(&p_full[g]);
    }
    if (nk > 0) {
      mbar_wait(&o_full[g], (nk - 1) & 1);
      tc_fence_after_sync();
      const float inv = 1.f / l;
      __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + ((int64_t)b * p.S + qi) * p.ldo + head * D;
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint32_t ov[32];
        tmem_ld_32x32(o_taddr + c * 32, ov);
        tmem_ld_wait();
        if (qi < p.S) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            *reinterpret_cast<uint4*>(orow + c * 32 + q * 8) = make_uint4(
                pack_bf16(__uint_as_float(ov[q * 8 + 0]) * inv, __uint_as_float(ov[q * 8 + 1]) * inv),
                pack_bf16(__uint_as_float(ov[q * 8 + 2]) * inv, __uint_as_float(ov[q * 8 + 3]) * inv),
                pack_bf16(__uint_as_float(ov[q * 8 + 4]) * inv, __uint_as_float(ov[q * 8 + 5]) * inv),
                pack_bf16(__uint_as_float(ov[q * 8 + 6]) * inv, __uint_as_float(ov[q * 8 + 7]) * inv));
          }
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

template <int D>
static int launch_attn(const vl2_attn_args* a, cudaStream_t stream) {
  using Cfg = AttnCfg<D>;
  CUtensorMap tq, tk, tv;
  const uint32_t box[3] = {64, 128, 1};
  {
    uint64_t dims[3] = {(uint64_t)a->Hq * D, (uint64_t)a->S, (uint64_t)a->B};
    uint64_t str[2] = {(uint64_t)a->ldq * 2, (uint64_t)a->ldq * 2 * a->S};
    int rc = make_tmap_bf16(&tq, a->q, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)a->Hkv * D, (uint64_t)a->S, (uint64_t)a->B};
    uint64_t str[2] = {(uint64_t)a->ldk * 2, (uint64_t)a->ldk * 2 * a->S};
    int rc = make_tmap_bf16(&tk, a->k, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)a->Hkv * D, (uint64_t)a->S, (uint64_t)a->B};
    uint64_t str[2] = {(uint64_t)a->ldv * 2, (uint64_t)a->ldv * 2 * a->S};
    int rc = make_tmap_bf16(&tv, a->v, 3, dims, str, box);
    if (rc) return rc;
  }
  AttnParams p;
  p.out = a->out; p.ldo = a->ldo; p.S = a->S; p.Hq = a->Hq; p.group = a->Hq / a->Hkv; p.causal = a->causal;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  static bool attr_set = false;
  if (!attr_set) {
    VL2_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  dim3 grid((a->S + 2 * BQ - 1) / (2 * BQ), a->Hq, a->B);
  attn_fwd_kernel<D><<<grid, kAttnThreads, Cfg::kSmemBytes, stream>>>(tq, tk, tv, p);
  VL2_CHECK_LAUNCH("attn_fwd_kernel");
  return VL2_OK;
}

}  // namespace vl2

extern "C" int vl2_attention(const vl2_attn_args* a, void* stream) {
  using namespace vl2;
  VL2_REQUIRE(a != nullptr, VL2_E_BADSHAPE, "vl2_attention: null args");
  VL2_REQUIRE(a->B > 0 && a->S > 0 && a->Hq > 0 && a->Hkv > 0 && a->Hq % a->Hkv == 0, VL2_E_BADSHAPE,
              "vl2_attention: bad B/S/heads (%d,%d,%d,%d)", a->B, a->S, a->Hq, a->Hkv);
  VL2_REQUIRE(a->D == 64 || a->D == 128, VL2_E_UNSUPPORTED, "vl2_attention: head_dim %d unsupported (64 or 128)", a->D);
  VL2_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0, VL2_E_BADALIGN,
              "vl2_attention: row strides must be multiples of 8 elements");
  VL2_REQUIRE(aligned16(a->q) && aligned16(a->k) && aligned16(a->v) && aligned16(a->out), VL2_E_BADALIGN,
              "vl2_attention: pointers must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->D == 64) return launch_attn<64>(a, st);
  return launch_attn<128>(a, st);
}
