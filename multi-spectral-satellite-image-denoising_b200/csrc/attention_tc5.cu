// tcgen05 / TMEM / TMA self-attention for the 144-token ViT-B blocks (reference: nn.MultiheadAttention in the fallback
// CustomBlock, src/models/encoder.py:158-169, or timm Attention -> SDPA, encoder.py:126-134):
//   out = softmax(q k^T / sqrt(64)) v per (image, head); no mask, no dropout.
//
// Persistent, warp-specialised, one CTA per SM; a CTA walks (image, head) problems p = blockIdx.x, + gridDim.x, ...
//   warp 4   : TMA producer -- three boxes [144 rows x 64] of the qkv matrix (q, k, v of one head) per problem into a
//              3-stage 128B-swizzled ring
//   warp 5   : MMA issuer   -- S = Q K^T (K-major A and B, N = 144, four K=16 steps) and O = P V (A = P from shared
//              memory, B = V as an MN-major operand straight from its TMA box, nine K=16 steps).  144 query rows are
//              one M=128 tile (rows 0..127) plus a second M=128 tile whose A descriptor starts 16 rows later
//              (rows 16..143): its last 16 accumulator lanes are rows 128..143, the other 112 are recomputed and ignored.
//   warp 6   : TMEM allocator (S tiles at columns 0 / 144, O tiles at 288 / 352)
//   warps 0-3: softmax + epilogue of rows 0..127, thread = one row: tcgen05.ld the whole 144-column score row into
//              registers (S is released to the tensor core for the next problem right after), max / exp2 / sum, P (bf16)
//              written as the K-major swizzled A operand of P V; the row's 1/sum is applied to O in the epilogue.
//   warp 7   : the same for rows 128..143 (TMEM lanes 112..127 of the shifted tile).
//              (Rotating the left-over rows over four extra warps, one per lane quadrant, balances the sub-partitions but needs
//              10 warps = 168 registers per thread; the 144-value score row then spills and the kernel is slower: 88 vs 66 us.)
// Software pipeline per softmax warp: softmax(i+1) runs before the epilogue of problem i, so P V (i) and the S MMAs
// of problem i+2 execute under it; the tensor core never waits on a warp that is waiting on the tensor core.
#include "gemm_tc5.cuh"

namespace {

using namespace gemm::tc5;

constexpr int T = 144, HD = 64;
constexpr int kStagesA = 3;
constexpr uint32_t TILE_BYTES = T * HD * 2;            // 18,432: one [144 x 64] bf16 box, 18 swizzle atoms
constexpr uint32_t STAGE_BYTES_A = 3 * TILE_BYTES;     // q, k, v
constexpr uint32_t P_BYTES = 3 * TILE_BYTES;           // P[144 x 144] as three 64-key column blocks
constexpr uint32_t SMEM_A = kStagesA * STAGE_BYTES_A + P_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int kThreadsA = 256;
constexpr uint32_t COL_S0 = 0, COL_S1 = 144, COL_O0 = 288, COL_O1 = 352;
constexpr int kSoftmaxWarps = 5;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// instruction descriptor with A K-major and B MN-major (bit 16)
__host__ __device__ constexpr uint32_t make_idesc_bmn(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__global__ void __launch_bounds__(kThreadsA, 1)
attention_tc5_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ out, int heads, int problems) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* smem_p = smem;
  unsigned char* smem_ring = smem + P_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_ring + kStagesA * STAGE_BYTES_A);
  uint64_t* qkv_full = bars;                   // [kStagesA]
  uint64_t* qkv_empty = bars + kStagesA;       // [kStagesA]
  uint64_t* s_full = bars + 2 * kStagesA;
  uint64_t* s_free = s_full + 1;
  uint64_t* p_full = s_full + 2;
  uint64_t* p_free = s_full + 3;
  uint64_t* o_full = s_full + 4;
  uint64_t* o_free = s_full + 5;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(s_full + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 4 && lane == 0) asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmQKV)) : "memory");
  if (warp == 5 && lane == 0) {
    for (int i = 0; i < kStagesA; ++i) { mbar_init(&qkv_full[i], 1); mbar_init(&qkv_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(s_free, kSoftmaxWarps);
    mbar_init(p_full, kSoftmaxWarps); mbar_init(p_free, 1);
    mbar_init(o_full, 1); mbar_init(o_free, kSoftmaxWarps);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 6) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  msd_pdl_enter();      // the prologue above ran under the previous kernel's tail; no global memory touched yet
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  const int n_local = ((int)blockIdx.x < problems) ? (problems - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int E = heads * HD;

  if (warp == 4) {
    // ===== TMA producer =====
    int stage = 0; uint32_t phase = 0;
    for (int j = 0; j < n_local; ++j) {
      const int problem = blockIdx.x + j * gridDim.x;
      const int b = problem / heads, h = problem - b * heads;
      mbar_wait(&qkv_empty[stage], phase ^ 1);
      if (elect_one()) {
        unsigned char* dst = smem_ring + stage * STAGE_BYTES_A;
        mbar_arrive_expect_tx(&qkv_full[stage], STAGE_BYTES_A);
        tma_load_2d(&tmQKV, &qkv_full[stage], dst, h * HD, b * T);
        tma_load_2d(&tmQKV, &qkv_full[stage], dst + TILE_BYTES, E + h * HD, b * T);
        tma_load_2d(&tmQKV, &qkv_full[stage], dst + 2 * TILE_BYTES, 2 * E + h * HD, b * T);
      }
      __syncwarp();
      if (++stage == kStagesA) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 5) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc_s = make_idesc(128, T);
    constexpr uint32_t idesc_o = make_idesc_bmn(128, HD);
    constexpr uint32_t HI = (1024u >> 4) | (1u << 14) | (2u << 29);                 // SBO 1024 B, version 1, SWIZZLE_128B
    const uint32_t stage_lo0 = (smem_u32(smem_ring) & 0x3FFFFu) >> 4;
    const uint32_t p_lo0 = (smem_u32(smem_p) & 0x3FFFFu) >> 4;
    constexpr uint32_t shift = (16u * 128u) >> 4;                                    // second tile starts 16 rows further down
    auto issue_s = [&](int j) {
      const int stage = j % kStagesA;
      mbar_wait(&qkv_full[stage], (uint32_t)(j / kStagesA) & 1u);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t q_lo = stage_lo0 + stage * (STAGE_BYTES_A >> 4), k_lo = q_lo + (TILE_BYTES >> 4);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint64_t db = ((uint64_t)HI << 32) | (k_lo + 2u * kk);
          tcgen05_mma_bf16(tmem_base + COL_S0, ((uint64_t)HI << 32) | (q_lo + 2u * kk), db, idesc_s, kk ? 1u : 0u);
          tcgen05_mma_bf16(tmem_base + COL_S1, ((uint64_t)HI << 32) | (q_lo + shift + 2u * kk), db, idesc_s, kk ? 1u : 0u);
        }
        tcgen05_commit(s_full);
      }
      __syncwarp();
    };
    if (n_local > 0) issue_s(0);
    for (int it = 0; it < n_local; ++it) {
      const uint32_t par = (uint32_t)it & 1u;
      if (it + 1 < n_local) {
        mbar_wait(s_free, par);                        // every softmax warp holds S(it) in registers
        issue_s(it + 1);
      }
      mbar_wait(p_full, par);                          // P(it) is in shared memory
      mbar_wait(o_free, par ^ 1u);                     // O(it-1) has been read out (passes at once for it = 0)
      tcgen05_fence_after();
      if (elect_one()) {
        const int stage = it % kStagesA;
        const uint32_t v_lo = stage_lo0 + stage * (STAGE_BYTES_A >> 4) + 2u * (TILE_BYTES >> 4);
#pragma unroll
        for (int kk = 0; kk < T / 16; ++kk) {
          // B = V rows 16kk .. 16kk+15 (two 8-row groups of the MN-major atom stack); LBO (one 64-wide atom only) unused
          const uint64_t db = ((uint64_t)HI << 32) | ((uint64_t)(TILE_BYTES >> 4) << 16) | (v_lo + kk * (2048u >> 4));
          const uint32_t pa = p_lo0 + (uint32_t)(kk >> 2) * (TILE_BYTES >> 4) + 2u * (kk & 3);
          tcgen05_mma_bf16(tmem_base + COL_O0, ((uint64_t)HI << 32) | pa, db, idesc_o, kk ? 1u : 0u);
          tcgen05_mma_bf16(tmem_base + COL_O1, ((uint64_t)HI << 32) | (pa + shift), db, idesc_o, kk ? 1u : 0u);
        }
        tcgen05_commit(o_full);
        tcgen05_commit(&qkv_empty[stage]);
        tcgen05_commit(p_free);
      }
      __syncwarp();
    }
  } else if (warp < 4 || warp == 7) {
    // ===== softmax + epilogue: thread = one query row =====
    const bool rem = warp == 7;
    const int quad = warp & 3;
    const int row = rem ? 16 + 96 + lane : 32 * quad + lane;          // query row this lane owns
    const bool active = !rem || lane >= 16;                           // lanes 0..15 of the left-over warp duplicate rows 112..127
    const uint32_t t_s = tmem_base + ((uint32_t)(32 * quad) << 16) + (rem ? COL_S1 : COL_S0);
    const uint32_t t_o = tmem_base + ((uint32_t)(32 * quad) << 16) + (rem ? COL_O1 : COL_O0);
    const uint32_t p_row = smem_u32(smem_p) + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    const float sl2 = 0.125f * 1.4426950408889634f;                   // 1/sqrt(64) * log2(e)
    float inv_prev = 0.0f;

    auto epilogue = [&](int it, float inv) {
      mbar_wait(o_full, (uint32_t)it & 1u);
      tcgen05_fence_after();
      uint32_t o[HD];
#pragma unroll
      for (int c = 0; c < HD / 16; ++c) tmem_ld16(t_o + 16 * c, &o[16 * c]);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
      if (active) {
        const int problem = blockIdx.x + it * gridDim.x;
        const int b = problem / heads, h = problem - b * heads;
        uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)b * T + row) * E + h * HD);
#pragma unroll
        for (int c = 0; c < HD / 8; ++c) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[8 * c + 0]) * inv, __uint_as_float(o[8 * c + 1]) * inv);
          v.y = pack_bf16x2(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv);
          v.z = pack_bf16x2(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv);
          v.w = pack_bf16x2(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv);
          dst[c] = v;
        }
      }
    };

    for (int it = 0; it < n_local; ++it) {
      mbar_wait(s_full, (uint32_t)it & 1u);
      tcgen05_fence_after();
      uint32_t s[T];
#pragma unroll
      for (int c = 0; c < T / 16; ++c) tmem_ld16(t_s + 16 * c, &s[16 * c]);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);                            // the tensor core may overwrite S with the next problem

      float mx[4];                                                    // four independent chains: one warp per sub-partition
#pragma unroll                                                        // has only its own ILP to cover instruction latency
      for (int k = 0; k < 4; ++k) mx[k] = __uint_as_float(s[k]);
#pragma unroll
      for (int j = 4; j < T; ++j) mx[j & 3] = fmaxf(mx[j & 3], __uint_as_float(s[j]));
      const float off = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * sl2;
      float l0 = 0.0f, l1 = 0.0f, l2 = 0.0f, l3 = 0.0f;
#pragma unroll
      for (int j = 0; j < T; j += 4) {
        const float e0 = ex2_approx(fmaf(__uint_as_float(s[j]), sl2, -off));
        const float e1 = ex2_approx(fmaf(__uint_as_float(s[j + 1]), sl2, -off));
        const float e2 = ex2_approx(fmaf(__uint_as_float(s[j + 2]), sl2, -off));
        const float e3 = ex2_approx(fmaf(__uint_as_float(s[j + 3]), sl2, -off));
        l0 += e0; l1 += e1; l2 += e2; l3 += e3;
        s[j >> 1] = pack_bf16x2(e0, e1);                              // P in place: 72 packed words
        s[(j >> 1) + 1] = pack_bf16x2(e2, e3);
      }
      mbar_wait(p_free, ((uint32_t)it & 1u) ^ 1u);                    // P V of the previous problem has consumed P
      if (active) {
#pragma unroll
        for (int c = 0; c < T / 8; ++c) {                             // 18 chunks of 8 keys (16 B)
          const uint32_t addr = p_row + (uint32_t)(c >> 3) * TILE_BYTES + ((((uint32_t)c & 7u) ^ sw) << 4);
          sts128(addr, make_uint4(s[4 * c], s[4 * c + 1], s[4 * c + 2], s[4 * c + 3]));
        }
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy stores -> visible to the MMA's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      const float inv = 1.0f / ((l0 + l1) + (l2 + l3));
      if (it > 0) epilogue(it - 1, inv_prev);
      inv_prev = inv;
    }
    if (n_local > 0) epilogue(n_local - 1, inv_prev);
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 6) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace

extern "C" int msd_attention_tc5_bf16(const void* d_qkv, void* d_out, int B, int T_, int heads, void* stream) {
  MSD_REQUIRE(d_qkv && d_out && B > 0 && heads > 0, "attention: bad arguments");
  MSD_REQUIRE(msd_aligned16(d_qkv) && msd_aligned16(d_out), "attention: pointers must be 16-byte aligned");
  MSD_REQUIRE(T_ == T, "attention: only T=144 tokens (192x192 images, patch 16) is built; got T=%d", T_);
  MSD_REQUIRE((long long)B * T < (1ll << 31) / (3 * heads * HD), "attention: batch too large");
  CUtensorMap tm;
  const int E = heads * HD;
  int rc = make_map(&tm, d_qkv, (uint64_t)3 * E, (uint64_t)B * T, (uint64_t)3 * E, (uint32_t)T);
  if (rc) return rc;
  static PerDeviceOnce attr_set;
  if (attr_set.first()) {
    MSD_CHECK_CUDA(cudaFuncSetAttribute(attention_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_A));
  }
  const int problems = B * heads;
  const int grid = problems < num_sms() ? problems : num_sms();
  MSD_CHECK_CUDA(msd_launch_pdl(attention_tc5_kernel, dim3((unsigned)grid), dim3(kThreadsA), SMEM_A, (cudaStream_t)stream, tm, reinterpret_cast<__nv_bfloat16*>(d_out), heads, problems));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}
