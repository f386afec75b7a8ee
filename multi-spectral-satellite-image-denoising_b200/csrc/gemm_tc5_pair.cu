// CTA-pair (tcgen05 cta_group::2) variant of the plain-mode GEMM / tap-major conv kernel of gemm_tc5.cuh.
//
// Why: the single-CTA kernel moves (128 + BN) x 64 bf16 through L2 -> shared memory per k-block for a 128 x BN tile, and
// ncu puts its wide GEMMs at the TMA / L2 delivery rate (~50 B/clk/SM, tensor pipe 58-61 % active).  Two CTAs of a cluster
// (same TPC) compute one 256 x BN tile with ONE instruction stream: each CTA loads its own 128 rows of A and only HALF of
// the W tile (BN/2 rows); `tcgen05.mma.cta_group::2` (M = 256) reads the other half from the peer's shared memory.  Operand
// bytes per CTA per k-block drop from 48 KB to 32 KB (BN = 256) and the ring gets deeper.
//
// Roles per CTA: warp 0 TMA producer (both CTAs; transaction bytes land on the LEADER's full barrier), warps 1 / 3 MMA
// issuers (leader CTA only; even / odd tiles, one TMEM accumulator buffer each), warp 2 TMEM allocator (cta_group::2 in both
// CTAs), warps 4-19 epilogue (both CTAs, own 128 rows; the fused epilogues of gemm_tc5.cuh unchanged).  `tcgen05.commit ...
// multicast::cluster` frees a stage / publishes an accumulator in both CTAs at once; the epilogue warps of both CTAs hand a
// TMEM buffer back by arriving on the leader's barrier through the cluster shared-memory window.
#include "gemm_tc5.cuh"

namespace gemm {
namespace tc5 {

// MERGED = the dx-merged 3x3 conv mode of gemm_tc5.cuh with the packed weight matrix RESIDENT in shared memory, split over the
// pair: each CTA keeps its BN/2 output channels of all taps (96 channels: 2 x 120 KB), so a stage is one 136-row A box only.
template <int BN, bool MERGED>
struct PairCfg {
  static constexpr uint32_t A_BOX_BYTES = (MERGED ? MERGE_ROWS : BLOCK_M) * BLOCK_K * 2;   // own 128 (+8) rows
  static constexpr uint32_t W_BOX_BYTES = (BN / 2) * BLOCK_K * 2;                           // own half of one W k-block
  static constexpr uint32_t STAGE_BYTES = A_BOX_BYTES + (MERGED ? 0u : W_BOX_BYTES);
  static constexpr int MAX_RES_KB = 20;                                                     // 10 taps x 2 k-blocks
  static constexpr uint32_t RESIDENT_BYTES = MERGED ? MAX_RES_KB * W_BOX_BYTES : 0u;
  using E = Cfg<BN, (MERGED ? MODE_MERGE3 : MODE_PLAIN)>;                                   // epilogue slab geometry
  // merged (narrow) tiles: 8 epilogue warps, each owning all slabs of its rows (the shared memory goes to the resident weights
  // and a deeper A ring instead of 16 staging buffers), and FOUR TMEM accumulator buffers: with one instruction stream feeding
  // two SMs the issue work per tile halves, and what bounds a narrow tile is the commit -> drain -> remote-arrive round trip of
  // its accumulator buffer, which two extra buffers hide.
  static constexpr int EPI_WARPS = MERGED ? 8 : E::EPI_WARPS;
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;
  static constexpr int SLAB_SPLIT = MERGED ? E::NSLAB : E::SLAB_SPLIT;
  static constexpr uint32_t STG_WARP_BYTES = 32 * E::STG_PITCH + SLAB_SPLIT * E::SC * 4;
  static constexpr uint32_t STG_BYTES = EPI_WARPS * STG_WARP_BYTES;
  static constexpr int NBUF = (4 * BN <= 512) ? 4 : 2;
  static constexpr uint32_t BAR_BYTES = 512;
  static constexpr int kStagesRaw = (int)((227u * 1024u - 1024u - BAR_BYTES - STG_BYTES - RESIDENT_BYTES) / STAGE_BYTES);
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr uint32_t SMEM_BYTES = kStages * STAGE_BYTES + RESIDENT_BYTES + STG_BYTES + 1024 + BAR_BYTES;
  static constexpr uint32_t TMEM_COLS = 512;
  static_assert(BN % 16 == 0 && 2 * BN <= 512 && W_BOX_BYTES % 1024 == 0 && A_BOX_BYTES % 1024 == 0, "pair tile geometry");
  static_assert(kStages >= 3, "pipeline too shallow");
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// 2-SM TMA load: the data lands in THIS CTA's shared memory, the transaction bytes on the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int BN, bool MERGED>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((PairCfg<BN, MERGED>::THREADS), 1)
gemm_tc5_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmS, const Problem p, int tiles_m2, int tiles_n, int epi) {
  using C = PairCfg<BN, MERGED>;
  using E = typename C::E;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* smem_a = smem;
  unsigned char* smem_b = smem + C::kStages * C::A_BOX_BYTES;                  // per-stage W halves (plain mode)
  unsigned char* smem_res = smem + C::kStages * C::STAGE_BYTES;                // resident W half (merged mode)
  unsigned char* smem_stg = smem_res + C::RESIDENT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + C::STG_BYTES);
  uint64_t* empty = bars + 2 * C::kStages;        // bars[0 .. 2*kStages): full[issuer][stage] (used in the leader CTA)
  uint64_t* tmem_full = bars + 3 * C::kStages;    // [NBUF]
  uint64_t* tmem_empty = tmem_full + C::NBUF;     // [NBUF], used in the leader CTA
  uint64_t* res_full = tmem_empty + C::NBUF;      // used in the leader CTA
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::kStages; ++i) { mbar_init(&bars[i], 1); mbar_init(&bars[C::kStages + i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < C::NBUF; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], C::EPI_WARPS); }   // both CTAs' warps of one group
    mbar_init(res_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_ptr)), "r"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();                              // barrier inits + TMEM of both CTAs visible before any cross-CTA traffic
  tcgen05_fence_after();
  msd_pdl_enter();                                 // the prologue above ran under the previous kernel's tail; no global memory touched yet
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  const int num_tiles = tiles_m2 * tiles_n;        // 256-row pair tiles
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int kbpt = p.kb_per_tap;
  const int nk_last = min(BLOCK_K / 16, (p.kvalid - (kbpt - 1) * BLOCK_K + 15) / 16);
  const int n_dy = 3, has_skip = p.skip_tap >= 0 ? 1 : 0;
  const int steps_per_tile = MERGED ? (n_dy + has_skip) * kbpt : p.num_kb;

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    int stage = 0; uint32_t phase = 0;
    uint32_t it = 0;
    if constexpr (MERGED) {
      // this CTA's half (BN/2 output channels) of the whole packed weight matrix stays resident; both halves are counted on the
      // leader's barrier, which its MMA issuers wait for once
      if (elect_one()) {
        if (leader) mbar_arrive_expect_tx(res_full, 2u * (uint32_t)p.num_kb * C::W_BOX_BYTES);
        for (int kb = 0; kb < p.num_kb; ++kb) tma_load_2d_pair(&tmB, res_full, smem_res + kb * C::W_BOX_BYTES, kb * BLOCK_K, (int)rank * (BN / 2));
      }
      __syncwarp();
    }
    const int tap0 = p.tap_off[0], tap3 = p.tap_off[3], tap6 = p.tap_off[6];
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const int tm = tile / tiles_n;
      const int m0 = tm * (2 * BLOCK_M) + (int)rank * BLOCK_M, n0 = (tile - tm * tiles_n) * BN + (int)rank * (BN / 2);
      uint64_t* full = bars + (it & 1u) * C::kStages;
      if constexpr (MERGED) {
        // rows m0 + off(dy, dx=-1) .. +135: windows at +0 / +1 / +2 rows are the dx taps; the skip tensor is read through the middle one
        for (int dy = 0; dy < n_dy + has_skip; ++dy) {
          const bool is_skip = dy == n_dy;
          const int row = is_skip ? m0 - 1 : m0 + (dy == 0 ? tap0 : dy == 1 ? tap3 : tap6);
          for (int kc = 0; kc < kbpt; ++kc) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one()) {
              if (leader) mbar_arrive_expect_tx(&full[stage], 2 * C::A_BOX_BYTES);
              tma_load_2d_pair(is_skip ? &tmS : &tmA, &full[stage], smem_a + stage * C::A_BOX_BYTES, kc * BLOCK_K, row);
            }
            __syncwarp();
            if (++stage == C::kStages) { stage = 0; phase ^= 1; }
          }
        }
        continue;
      }
      int tap = 0, kc = 0;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(&full[stage], 2 * C::STAGE_BYTES);      // both CTAs' boxes of this stage
          tma_load_2d_pair(tap == p.skip_tap ? &tmS : &tmA, &full[stage], smem_a + stage * C::A_BOX_BYTES, kc * BLOCK_K, m0 + p.tap_off[tap]);
          tma_load_2d_pair(&tmB, &full[stage], smem_b + stage * C::W_BOX_BYTES, kb * BLOCK_K, n0);
        }
        __syncwarp();
        if (++kc == kbpt) { kc = 0; ++tap; }
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if ((warp == 1 || warp == 3) && leader) {
    // ===== MMA issuers (leader CTA only) =====
    constexpr uint32_t idesc = make_idesc(2 * BLOCK_M, BN);
    constexpr uint32_t HI = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t par = warp == 1 ? 0u : 1u;
    const uint32_t a_lo0 = (smem_u32(smem_a) & 0x3FFFFu) >> 4, b_lo0 = (smem_u32(smem_b) & 0x3FFFFu) >> 4;
    const uint32_t r_lo0 = (smem_u32(smem_res) & 0x3FFFFu) >> 4;
    uint64_t* full = bars + par * C::kStages;
    uint32_t pmask = 0, it = 0;
    int gstep = 0;
    if constexpr (MERGED) { mbar_wait(res_full, 0); tcgen05_fence_after(); }
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it, gstep += steps_per_tile) {
      if ((it & 1u) != par) continue;
      int stage = gstep % C::kStages;
      // this issuer's n-th tile goes to accumulator buffer par (+2 on odd n when there are four buffers)
      const uint32_t n = it >> 1;
      const uint32_t buf = C::NBUF == 4 ? par + 2u * (n & 1u) : par;
      const uint32_t uses = C::NBUF == 4 ? n >> 1 : n;
      mbar_wait(&tmem_empty[buf], (uses & 1u) ^ 1u);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + buf * BN;
      uint32_t acc = 0;
      if constexpr (MERGED) {
        for (int dy = 0; dy < n_dy + has_skip; ++dy) {
          const bool is_skip = dy == n_dy;
          for (int kc = 0; kc < kbpt; ++kc) {
            mbar_wait(&full[stage], (pmask >> stage) & 1u);
            pmask ^= 1u << stage;
            tcgen05_fence_after();
            const int nk = (kc == kbpt - 1) ? nk_last : BLOCK_K / 16;
            const bool last = (dy == n_dy + has_skip - 1) && (kc == kbpt - 1);
            if (elect_one()) {
              const uint32_t a_lo = a_lo0 + stage * (C::A_BOX_BYTES >> 4);
              const int j0 = is_skip ? 1 : 0, j1 = is_skip ? 2 : 3;                       // window j = rows +j = tap dx = j - 1
              for (int j = j0; j < j1; ++j) {
                const int tap = is_skip ? p.skip_tap : 3 * dy + j;
                const uint64_t da = ((uint64_t)HI << 32) | (a_lo + 8u * j);
                const uint64_t db = ((uint64_t)HI << 32) | (r_lo0 + (uint32_t)(tap * kbpt + kc) * (C::W_BOX_BYTES >> 4));
                for (int k = 0; k < nk; ++k) { tcgen05_mma_bf16_pair(tmem_d, da + 2u * k, db + 2u * k, idesc, acc); acc = 1; }
              }
              tcgen05_commit_pair(&empty[stage]);
              if (last) tcgen05_commit_pair(&tmem_full[buf]);
            }
            __syncwarp();
            acc = 1;
            if (++stage == C::kStages) stage = 0;
          }
        }
        continue;
      }
      int kc = 0;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full[stage], (pmask >> stage) & 1u);
        pmask ^= 1u << stage;
        tcgen05_fence_after();
        const int nk = (kc == kbpt - 1) ? nk_last : BLOCK_K / 16;
        if (elect_one()) {
          const uint64_t da = ((uint64_t)HI << 32) | (a_lo0 + stage * (C::A_BOX_BYTES >> 4));
          const uint64_t db = ((uint64_t)HI << 32) | (b_lo0 + stage * (C::W_BOX_BYTES >> 4));
          for (int k = 0; k < nk; ++k) tcgen05_mma_bf16_pair(tmem_d, da + 2u * k, db + 2u * k, idesc, k == 0 ? acc : 1u);
          tcgen05_commit_pair(&empty[stage]);
          if (kb == p.num_kb - 1) tcgen05_commit_pair(&tmem_full[buf]);
        }
        __syncwarp();
        acc = 1;
        if (++kc == kbpt) kc = 0;
        if (++stage == C::kStages) stage = 0;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs, own 128 rows) =====
    const int e = warp - 4;
    if (p.mode == EPI_CONVT && p.zero_out_halo) zero_convt_out_halo(p, (int)threadIdx.x - 128, C::EPI_WARPS * 32);
    const int q = warp & 3, grp = (e >> 2) & 1, half = e >> 3;
    const uint32_t stg_s = smem_u32(smem_stg + e * C::STG_WARP_BYTES);
    const uint32_t bias_s = stg_s + 32 * E::STG_PITCH;
    const int slab_begin = half ? C::SLAB_SPLIT : 0, slab_end = half ? E::NSLAB : C::SLAB_SPLIT;
    const int ncols = (slab_end - slab_begin) * E::SC;
    // column sums of the stored tensor (training BatchNorm statistics): one slot per epilogue warp behind the nine taps' weights
    uint32_t sums_s = 0u;
    if constexpr (MERGED) {
      static_assert((uint32_t)C::EPI_WARPS * 2u * BN * 4u <= 2u * C::W_BOX_BYTES, "column-sum slots must fit the skip tap's two k-blocks");
      if (p.bn_sums != nullptr && epi == E_CONV_RAW) {
        sums_s = smem_u32(smem_res + (C::MAX_RES_KB - 2) * C::W_BOX_BYTES) + (uint32_t)e * (uint32_t)(2 * BN * 4);
        bn_sums_zero(sums_s, BN, lane);
      }
    }
    int it = 0, bias_n0 = -1;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      if ((it & 1) != grp) continue;
      const uint32_t n = (uint32_t)it >> 1;
      const uint32_t as = C::NBUF == 4 ? (uint32_t)grp + 2u * (n & 1u) : (uint32_t)grp;
      const uint32_t aphase = (C::NBUF == 4 ? n >> 1 : n) & 1u;
      const int tm = tile / tiles_n;
      const int m0 = tm * (2 * BLOCK_M) + (int)rank * BLOCK_M, n0 = (tile - tm * tiles_n) * BN;
      if (n0 != bias_n0) {
        bias_n0 = n0;
        float b[(C::SLAB_SPLIT * E::SC + 31) / 32];
#pragma unroll
        for (int k = 0; k < (C::SLAB_SPLIT * E::SC + 31) / 32; ++k) {
          const int i = k * 32 + lane;
          b[k] = (p.bias != nullptr && i < ncols) ? __ldg(p.bias + n0 + slab_begin * E::SC + i) : 0.0f;
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < (C::SLAB_SPLIT * E::SC + 31) / 32; ++k) {
          const int i = k * 32 + lane;
          if (i < ncols) asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(bias_s + 4u * i), "f"(b[k]) : "memory");
        }
        __syncwarp();
        if constexpr (!MERGED) {
          if (p.ln_s != nullptr)      // folded LayerNorm: the column sums of the weights, behind the bf16-pitched staging rows
            stage_column_slice<(C::SLAB_SPLIT * E::SC + 31) / 32>(p.ln_s + n0 + slab_begin * E::SC, stg_s + 32u * (E::SC * 2 + 16), ncols, lane);
        }
      }
      float ln_a = 0.0f, ln_c = 0.0f;
      if constexpr (!MERGED) {
        if (epi == E_RM_BF16_LN || epi == E_RM_BF16_GELU_LN) ln_row_coeffs(p, m0 + q * 32 + lane, ln_a, ln_c);
        if (epi == E_RM_F32_RESID || epi == E_RM_F32_RESID_STATS)
          prefetch_rows_l2(p.resid, (size_t)p.ldo * 4, m0 + q * 32, p.M - (m0 + q * 32), (uint32_t)(n0 + slab_begin * E::SC) * 4u, (uint32_t)ncols * 4u, lane);
      }
      if (epi == E_CONV_SKIPADD)
        prefetch_rows_l2(p.skip, (size_t)p.ldo * 2, m0 + q * 32, p.M - (m0 + q * 32), (uint32_t)(n0 + slab_begin * E::SC) * 2u, (uint32_t)ncols * 2u, lane);
      if constexpr (!MERGED) {
        if (epi == E_RM_BF16_DGELU)
          prefetch_rows_l2(p.gelu_pre, (size_t)p.ldo * 2, m0 + q * 32, p.M - (m0 + q * 32), (uint32_t)(n0 + slab_begin * E::SC) * 2u, (uint32_t)ncols * 2u, lane);
      }
      mbar_wait(&tmem_full[as], aphase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
#define MSD_EPI_CASE(EE) case EE: epilogue_tile<BN, EE, true>(p, stg_s, bias_s, taddr, &tmem_empty[as], m0 + q * 32, n0, lane, slab_begin, slab_end); break;
      switch (epi) {
        MSD_EPI_CASE(E_RM_F32)
        MSD_EPI_CASE(E_RM_BF16)
        MSD_EPI_CASE(E_RM_BF16_GELU)
        MSD_EPI_CASE(E_RM_F32_RESID)
        MSD_EPI_CASE(E_CONV)
        case E_CONV_RAW: epilogue_tile<BN, E_CONV_RAW, true>(p, stg_s, bias_s, taddr, &tmem_empty[as], m0 + q * 32, n0, lane, slab_begin, slab_end, nullptr, 0.0f, 0.0f, sums_s); break;
        MSD_EPI_CASE(E_CONVT)
        MSD_EPI_CASE(E_CONVT_RAW)
        MSD_EPI_CASE(E_CONV_SKIPADD)
        MSD_EPI_CASE(E_RM_F32_RESID_STATS)
        MSD_EPI_CASE(E_RM_BF16_GELU_PRE)
        MSD_EPI_CASE(E_RM_BF16_DGELU)
        case E_RM_BF16_LN: epilogue_tile<BN, E_RM_BF16_LN, true>(p, stg_s, bias_s, taddr, &tmem_empty[as], m0 + q * 32, n0, lane, slab_begin, slab_end, nullptr, ln_a, ln_c); break;
        case E_RM_BF16_GELU_LN: epilogue_tile<BN, E_RM_BF16_GELU_LN, true>(p, stg_s, bias_s, taddr, &tmem_empty[as], m0 + q * 32, n0, lane, slab_begin, slab_end, nullptr, ln_a, ln_c); break;
        default: epilogue_tile<BN, E_GENERIC, true>(p, stg_s, bias_s, taddr, &tmem_empty[as], m0 + q * 32, n0, lane, slab_begin, slab_end); break;
      }
#undef MSD_EPI_CASE
    }
    if constexpr (MERGED) {
      if (p.bn_sums != nullptr && epi == E_CONV_RAW)
        bn_sums_flush(smem_u32(smem_res + (C::MAX_RES_KB - 2) * C::W_BOX_BYTES), p.bn_sums, BN, (int)threadIdx.x - 128, C::EPI_WARPS);
    }
  }
  tcgen05_fence_before();
  cluster_sync_all();                              // no CTA may free TMEM / exit while its peer still reads its shared memory
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

template <int BN, bool MERGED>
int launch_pair_mode(const Problem& p, cudaStream_t s) {
  using C = PairCfg<BN, MERGED>;
  CUtensorMap tmA, tmB;
  const uint32_t a_box_rows = MERGED ? MERGE_ROWS : BLOCK_M;
  int rc = make_map(&tmA, p.A, (uint64_t)p.kvalid, (uint64_t)p.a_rows, (uint64_t)p.ldA, a_box_rows);
  if (rc) return rc;
  rc = make_map(&tmB, p.Wt, (uint64_t)p.num_kb * BLOCK_K, (uint64_t)p.N, (uint64_t)(p.w_ld_kb ? p.w_ld_kb : p.num_kb) * BLOCK_K, BN / 2);
  if (rc) return rc;
  CUtensorMap tmS = tmA;
  if (p.skip_tap >= 0) {
    rc = make_map(&tmS, p.skip, (uint64_t)p.kvalid, (uint64_t)p.a_rows, (uint64_t)p.ldA, a_box_rows);
    if (rc) return rc;
  }
  static PerDeviceOnce attr_set;
  if (attr_set.first()) {
    MSD_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc5_pair_kernel<BN, MERGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
  }
  const int tiles_m2 = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M), tiles_n = p.N / BN;
  const int tiles = tiles_m2 * tiles_n;
  const int clusters = tiles < num_sms() / 2 ? tiles : num_sms() / 2;
  MSD_CHECK_CUDA(msd_launch_pdl(gemm_tc5_pair_kernel<BN, MERGED>, dim3(2 * clusters), dim3(C::THREADS), C::SMEM_BYTES, s, tmA, tmB, tmS, p, tiles_m2, tiles_n, classify(p)));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

template <int BN>
int launch_pair(const Problem& p, cudaStream_t s) {
  if constexpr (BN == 96 || BN == 48) return launch_pair_mode<BN, true>(p, s);
  else return launch_pair_mode<BN, false>(p, s);
}

template int launch_pair<256>(const Problem&, cudaStream_t);
template int launch_pair<192>(const Problem&, cudaStream_t);
template int launch_pair<96>(const Problem&, cudaStream_t);
template int launch_pair<48>(const Problem&, cudaStream_t);

}  // namespace tc5
}  // namespace gemm
