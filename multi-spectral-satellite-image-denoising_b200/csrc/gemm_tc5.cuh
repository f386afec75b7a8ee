// tcgen05 / TMEM / TMA GEMM + implicit 3x3 convolution for sm_100a (the product path of every dense contraction
// in SatMAERestoration.forward; shapes in SURVEY Appendix A).
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0   : TMA producer  -- cp.async.bulk.tensor.2d of the A box [128 rows x 64 k] (row = m0 + tap shift) and
//              the W box [BN rows x 64 k] into a kStages-deep 128B-swizzled shared-memory ring (mbarrier full/empty)
//   warp 1   : MMA issuer    -- one thread issues 4 x tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) per
//              stage into one of two TMEM accumulator buffers, tcgen05.commit releases the stage / publishes the tile
//   warp 2   : TMEM allocator
//   warps 4-11: epilogue     -- two groups of four warps; group g drains TMEM accumulator buffer g (even / odd tiles), so
//              two tiles' epilogues are in flight and every SM sub-partition has two epilogue warps to interleave.
//              Per column slab: tcgen05.ld 32x32b.x16 (thread = one accumulator row) -> TMEM is
//              released as soon as the last slab is in registers -> fp32 slab staged in a per-warp padded smem tile
//              -> re-read with lanes running along the columns, so that bias / residual / skip loads and the
//              fused GELU / pixel-shuffle / halo-zero stores are coalesced row segments instead of 32 scattered rows;
//              all of it overlaps the next tile's MMAs (two TMEM accumulator buffers)
// Accumulators are fp32 in TMEM (2 x BN columns), operands bf16.
#pragma once
#include "gemm.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace gemm {
namespace tc5 {

constexpr int kEpiWarps = 16;    // 2 accumulator buffers x 4 TMEM lane quadrants x 2 column halves
constexpr int kThreads = 128 + 32 * kEpiWarps;   // 4 control warps (TMA producer, 2 MMA issuers, TMEM allocator) + epilogue
constexpr uint64_t kSpinLimit = 4000000000ull;             // ~2 s of SM clocks: trap instead of hanging the GPU

// MODE 0: one A box [128 rows x 64 k] + one W box per pipeline stage (linear layers, ConvTranspose, wide convs).
// MODE 1: 3x3 conv, "dx-merged": one A box of 136 rows per (dy, k-block) serves the three taps dx = -1, 0, +1 as row-
//         shifted windows of the same swizzled buffer (the UMMA descriptor start may sit anywhere inside a 1024-byte
//         swizzle atom as long as the base-offset field stays 0 -- verified by tools/desc_probe.py), so the activation
//         tile crosses L2->SM 3x instead of 9x; the stage also carries the three taps' W boxes.
// MODE 2: MODE 1 + the whole packed weight matrix resident in shared memory (N == BN, one k-block per tap: the
//         48-channel stage, 60 KB), so a stage is the A box only.
// MODE 3: MODE 2 with FOUR TMEM accumulator buffers and four epilogue groups (16 warps), nine taps only (a skip connection is
//         added by the epilogue, E_CONV_SKIPADD).  An experiment on the hypothesis that the 48-channel convs are bound by the
//         latency of one epilogue warp-tile (~540 dependent instructions at IPC 0.2): they are not (0.562 vs 0.540 ms), so the
//         mode is opt-in (MSD_RES4=1) and documents the negative result.
constexpr int MODE_PLAIN = 0, MODE_MERGE3 = 1, MODE_MERGE3_RESIDENT = 2, MODE_MERGE3_RESIDENT4 = 3;
__host__ __device__ constexpr bool mode_resident(int mode) { return mode == MODE_MERGE3_RESIDENT || mode == MODE_MERGE3_RESIDENT4; }
constexpr int MERGE_ROWS = 136;   // 128 + 2 shifted rows, rounded up to a whole 8-row swizzle atom

template <int BN, int MODE = MODE_PLAIN>
struct Cfg {
  static constexpr uint32_t W_BOX_BYTES = BN * BLOCK_K * 2;
  static constexpr uint32_t A_BOX_BYTES = (MODE == MODE_PLAIN ? BLOCK_M : MERGE_ROWS) * BLOCK_K * 2;
  static constexpr uint32_t A_STAGE_BYTES = (A_BOX_BYTES + 1023u) & ~1023u;
  static constexpr uint32_t B_STAGE_BYTES = MODE == MODE_PLAIN ? W_BOX_BYTES : MODE == MODE_MERGE3 ? 3 * W_BOX_BYTES : 0;
  static constexpr uint32_t STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr uint32_t RESIDENT_BYTES = MODE == MODE_MERGE3_RESIDENT ? 10 * W_BOX_BYTES : MODE == MODE_MERGE3_RESIDENT4 ? 9 * W_BOX_BYTES : 0;
  static constexpr int NBUF = MODE == MODE_MERGE3_RESIDENT4 ? 4 : 2;       // TMEM accumulator buffers = epilogue groups
  // epilogue staging: slabs of SC accumulator columns x 32 rows per epilogue warp (fp32 or bf16), row pitch padded by 16 B
  // (an odd number of 16-byte units: row-per-lane writes are conflict-free)
  // narrow tiles (one slab) run with a single column half = 8 epilogue warps: more warps would only take issue slots
  // from the producer / MMA issuer warps, which are the critical path there
  static constexpr int SC = (BN % 32 == 0) ? 32 : (BN % 48 == 0) ? 48 : 16;
  static constexpr int NSLAB = BN / SC;
  static constexpr int SLAB_SPLIT = (NSLAB + 1) / 2;         // column half 0 owns slabs [0, SLAB_SPLIT), half 1 the rest
  // plain mode may stage fp32 rows (fp32-output epilogues); the merged conv modes only ever stage bf16
  static constexpr uint32_t STG_PITCH = (MODE == MODE_PLAIN ? SC * 4 : SC * 2) + 16;
  static constexpr uint32_t STG_WARP_BYTES = 32 * STG_PITCH + SLAB_SPLIT * SC * 4;   // staged slab + this warp's bias slice
  static constexpr int EPI_WARPS = (NSLAB >= 2 || NBUF == 4) ? kEpiWarps : kEpiWarps / 2;
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;
  static constexpr uint32_t STG_BYTES = EPI_WARPS * STG_WARP_BYTES;
  static constexpr int kStagesRaw = (int)((227u * 1024u - 1024u - 256u - STG_BYTES - RESIDENT_BYTES) / STAGE_BYTES);
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr uint32_t TMEM_COLS = (NBUF * BN <= 32) ? 32 : (NBUF * BN <= 64) ? 64 : (NBUF * BN <= 128) ? 128 : (NBUF * BN <= 256) ? 256 : 512;
  static_assert(NBUF * BN <= 512, "accumulator buffers exceed the 512 TMEM columns");
  static constexpr uint32_t SMEM_BYTES = kStages * STAGE_BYTES + RESIDENT_BYTES + STG_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(BN % SC == 0 && SC % 16 == 0, "slab must tile BN in 16-column TMEM loads");
  static_assert(kStages >= 3, "pipeline too shallow");
  static_assert(SMEM_BYTES <= 227u * 1024u, "shared memory budget");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128 must be a multiple of 16 in [16,256]");
  static_assert(W_BOX_BYTES % 1024 == 0, "stage bases must stay 1024-byte aligned for SWIZZLE_128B");
  static_assert(MODE != MODE_MERGE3_RESIDENT || BN != 48 || (uint32_t)EPI_WARPS * 2u * BN * 4u <= W_BOX_BYTES, "column-sum slots must fit the tenth-tap slot");
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if ((uint64_t)(clock64() - t0) > kSpinLimit) {
      printf("libmsd_b200 gemm_tc5: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// one lane of a converged warp; keeps the surrounding loop warp-uniform so that descriptors live in uniform registers
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 16 consecutive fp32 accumulator columns of this thread's TMEM lane; completion via tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
// zero 16 consecutive fp32 accumulator columns of this thread's TMEM lane; completion via tmem_st_wait()
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// explicit shared-space accesses for the epilogue staging buffers (generic LD/ST on a pointer derived from the dynamic
// smem base go through the long-latency generic path and serialise against each other)
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

constexpr uint32_t kRowSkip = 0xFFFFFFFFu;   // row produces no output (beyond M, or a halo row of a ConvTranspose input)
constexpr uint32_t kRowHalo = 0x80000000u;   // EPI_CONV halo row: stored as zeros

// Per-row output coordinate, computed once per tile by the lane that owns the accumulator row:
//   EPI_ROWMAJOR: m              EPI_CONV: m | kRowHalo?        EPI_CONVT: output pixel of sub-position (0,0)
__device__ __forceinline__ uint32_t row_info(const Problem& p, int m) {
  if (m >= p.M) return kRowSkip;
  if (p.mode == EPI_ROWMAJOR) return (uint32_t)m;
  uint32_t img, rem, yy, xx;
  if (p.mode == EPI_CONV) {
    fd_divmod((uint32_t)m, p.plane, img, rem);
    fd_divmod(rem, p.pitch, yy, xx);
    const bool interior = yy >= 1 && yy <= (uint32_t)p.H && xx >= 1 && xx <= (uint32_t)p.W;
    return (uint32_t)m | (interior ? 0u : kRowHalo);
  }
  fd_divmod((uint32_t)m, p.in_plane, img, rem);
  fd_divmod(rem, p.in_pitch, yy, xx);
  if (p.in_pad) {
    if (!(yy >= 1 && yy <= (uint32_t)p.Hin && xx >= 1 && xx <= (uint32_t)p.Win)) return kRowSkip;
    yy -= 1; xx -= 1;
  }
  return img * (uint32_t)p.out_plane + (2 * yy + 1) * (uint32_t)p.out_pitch + (2 * xx + 1);
}

// E_CONV_HEAD: NCHW offset of channel 0 of the output pixel of accumulator row m (a padded pixel of [B, H+2, W+2]), or kRowSkip
__device__ __forceinline__ uint32_t row_info_head(const Problem& p, int m) {
  if (m >= p.M) return kRowSkip;
  uint32_t img, rem, yy, xx;
  fd_divmod((uint32_t)m, p.plane, img, rem);
  fd_divmod(rem, p.pitch, yy, xx);
  if (!(yy >= 1 && yy <= (uint32_t)p.H && xx >= 1 && xx <= (uint32_t)p.W)) return kRowSkip;
  return (img * (uint32_t)p.head_cout * (uint32_t)p.H + (yy - 1)) * (uint32_t)p.W + (xx - 1);
}
// Fused 1x1 head = a second, tiny tcgen05 contraction per tile:  D2[128 x 16] = A2[128 x 48] . W2^T, fp32 in 16 spare TMEM columns.
//   A2 : the tile's GELU'd activation, bf16, written by the epilogue lanes (one row each) as the K-major SWIZZLE_128B operand into
//        ring stage 6 + buffer (the nine-tap head conv cycles through stages 0..5 only)
//   W2 : the fp32 1x1 weights split into hi + lo bf16 (two MMAs per K step: ~16 mantissa bits of every weight), built once per CTA
//        in the free tenth-tap slot of the resident weight area, followed by the mbarriers of the hand-off and the 16 biases
// The issuer of a tile waits for its A2 (a2_full), issues the six M128 x N16 x K16 MMAs and commits d2_full; the epilogue lanes
// then read 16 columns and store NCHW fp32.  Measured alternatives for the same contraction in the epilogue warps -- FMA pipe with
// broadcast ld.shared weights (0.78 ms for the fused kernel vs ~0.5 for the conv alone), packed FFMA2 (0.81), mma.sync tiles (0.78:
// the legacy HMMA queues behind the in-flight tcgen05 MMAs), weights as kernel parameters (1.12: one LDCU per weight) -- all made
// the epilogue warps the bottleneck of the kernel.
constexpr uint32_t HEAD_W2_LO = 2048, HEAD_BARS = 4096, HEAD_BIAS = 4160, HEAD_SLOT_BYTES = 4224;
__device__ __forceinline__ void build_head_operand(const Problem& p, uint32_t slot_s, int tid, int nthreads) {
  // W2 as two [16 rows x 64 k] K-major SWIZZLE_128B tiles (rows >= head_cout and k >= N are zero)
  for (int i = tid; i < 16 * 32; i += nthreads) {
    const int n = i >> 5, kp = i & 31, k = 2 * kp;
    float w0 = 0.0f, w1 = 0.0f;
    if (n < p.head_cout && k < p.N) { w0 = __ldg(p.head_w + (size_t)n * p.N + k); w1 = __ldg(p.head_w + (size_t)n * p.N + k + 1); }
    const __nv_bfloat16 h0 = __float2bfloat16_rn(w0), h1 = __float2bfloat16_rn(w1);
    const uint32_t hi = pack_bf16x2(w0, w1), lo = pack_bf16x2(w0 - __bfloat162float(h0), w1 - __bfloat162float(h1));
    const uint32_t off = (uint32_t)(n >> 3) * 1024u + (uint32_t)(n & 7) * 128u + ((((uint32_t)k >> 3) ^ (uint32_t)(n & 7)) << 4) + ((uint32_t)k & 7u) * 2u;
    asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(slot_s + off), "r"(hi) : "memory");
    asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(slot_s + HEAD_W2_LO + off), "r"(lo) : "memory");
  }
  for (int i = tid; i < 16; i += nthreads) {
    const float b = (i < p.head_cout && p.head_b != nullptr) ? __ldg(p.head_b + i) : 0.0f;
    asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(slot_s + HEAD_BIAS + 4u * i), "f"(b) : "memory");
  }
}
struct HeadCtx {
  uint32_t a2_s;        // shared address of this accumulator buffer's A2 tile
  uint32_t bias_s;      // 16 fp32 biases
  uint32_t taddr_d2;    // TMEM address of D2 for this warp's lane quadrant
  uint64_t* a2_full;    // epilogue -> issuer: the A2 tile is written
  uint64_t* d2_full;    // issuer -> epilogue: D2 is complete
  uint32_t parity;      // phase parity of THIS tile on both barriers
  uint32_t row0;        // row of this warp's first lane inside the 128-row A2 tile (32 x TMEM lane quadrant)
  // software pipeline of the epilogue lanes: D2 of this buffer's previous tile is read while this tile is handled
  bool has_prev;
  uint32_t prev_rowpix; // output offset of this lane's pixel in the previous tile (kRowSkip: none)
};
// read D2 of the pending tile of this accumulator buffer and store the lane's pixel (NCHW fp32, + bias)
__device__ __forceinline__ void head_store(const Problem& p, const HeadCtx& hc, uint32_t pix) {
  uint32_t d2[16];
  tmem_ld16(hc.taddr_d2, d2);
  tmem_ld_wait();
  tcgen05_fence_before();
  if (pix != kRowSkip) {
    const uint32_t hw = (uint32_t)p.H * (uint32_t)p.W;
    float* outp = p.head_out + pix;
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      if (n < p.head_cout) {
        float bsc;
        asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(bsc) : "r"(hc.bias_s + 4u * n));
        outp[(uint32_t)n * hw] = __uint_as_float(d2[n]) + bsc;         // lanes = consecutive pixels: 128-byte rows per channel plane
      }
    }
  }
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (unused for one swizzle atom along K),
//   [32,46) stride byte offset >> 4 = 8 rows * 128 B, [46,48) version = 1 (sm_100), [61,64) layout = 2 (128B swizzle)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6)=1, a=bf16 [7,10)=1, b=bf16 [10,13)=1,
// a/b K-major (bits 15,16 = 0), n>>3 at [17,23), m>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// Up to four K=16 MMAs over one 64-wide k-block.  a_lo / b_lo = (shared address >> 4) of the operand tiles; the K-major
// SWIZZLE_128B descriptor's high word is a constant, and advancing 16 bf16 (32 bytes) inside the 128-byte swizzle atom
// is +2 in the address field.  One straight-line sequence per slice count keeps every operand in uniform registers.
__device__ __forceinline__ void mma_kblock(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi_extra, uint32_t idesc,
                                           int nk, uint32_t acc) {
  constexpr uint32_t HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
  const uint64_t da = ((uint64_t)(HI | a_hi_extra) << 32) | a_lo, db = ((uint64_t)HI << 32) | b_lo;
  if (nk == 4) {
    tcgen05_mma_bf16(tmem_d, da, db, idesc, acc);
    tcgen05_mma_bf16(tmem_d, da + 2, db + 2, idesc, 1u);
    tcgen05_mma_bf16(tmem_d, da + 4, db + 4, idesc, 1u);
    tcgen05_mma_bf16(tmem_d, da + 6, db + 6, idesc, 1u);
  } else if (nk == 3) {
    tcgen05_mma_bf16(tmem_d, da, db, idesc, acc);
    tcgen05_mma_bf16(tmem_d, da + 2, db + 2, idesc, 1u);
    tcgen05_mma_bf16(tmem_d, da + 4, db + 4, idesc, 1u);
  } else if (nk == 2) {
    tcgen05_mma_bf16(tmem_d, da, db, idesc, acc);
    tcgen05_mma_bf16(tmem_d, da + 2, db + 2, idesc, 1u);
  } else {
    tcgen05_mma_bf16(tmem_d, da, db, idesc, acc);
  }
}

// ------------------------------------------------------------------------------------------------ epilogue
// Fused epilogues, one instruction stream each (selected per launch from the Problem, see classify()):
enum Epi : int {
  E_RM_F32 = 0,     // out fp32 [M,N] = acc + bias                                  (patch embedding)
  E_RM_BF16,        // out bf16 = acc + bias                                         (qkv)
  E_RM_BF16_GELU,   // out bf16 = GELU(acc + bias)                                   (fc1)
  E_RM_F32_RESID,   // out fp32 = resid + acc + bias, in place                       (attention proj, fc2)
  E_CONV,           // haloed NHWC bf16 = GELU(acc + bias), halo rows zeroed         (every 3x3 conv; a ResidualBlock's skip is a 10th tap)
  E_CONVT,          // pixel-shuffle scatter of GELU(acc + bias)                     (ConvTranspose2d k2 s2)
  E_CONV_RAW,       // haloed NHWC bf16 = acc + bias, halo rows zeroed               (training: pre-BatchNorm output, dgrad)
  E_CONVT_RAW,      // pixel-shuffle scatter of acc + bias                            (training: pre-BatchNorm output)
  E_CONV_HEAD,      // NCHW fp32 = head_w . GELU(acc + bias) + head_b                 (head conv3x3 + 1x1, one kernel)
  E_CONV_SKIPADD,   // haloed NHWC bf16 = GELU(acc + bias + skip), skip read in the epilogue (ResidualBlock's second conv, 9 taps)
  E_RM_BF16_LN,         // out bf16 = rstd (acc - mean ln_s) + bias                  (qkv behind a folded LayerNorm, see Problem)
  E_RM_BF16_GELU_LN,    // out bf16 = GELU(rstd (acc - mean ln_s) + bias)            (fc1 behind a folded LayerNorm)
  E_RM_F32_RESID_STATS, // E_RM_F32_RESID + bf16 copy of the row + partial LayerNorm sums (proj / fc2 in front of a folded LayerNorm)
  E_RM_BF16_GELU_PRE,   // out_bf16 = bf16(acc + bias), out = GELU(that)                (training fc1: pre-activation + activation)
  E_RM_BF16_DGELU,      // out bf16 = bf16(acc + bias) * gelu'(gelu_pre)                (training fc2 data gradient)
  E_GENERIC         // any other combination of the Problem's flags (slower, runtime-branched)
};

__host__ inline int classify(const Problem& p) {
  const bool idx32 = (unsigned long long)p.M * (unsigned long long)p.ldo < (1ull << 31);
  if (!idx32) return E_GENERIC;
  if (p.mode == EPI_ROWMAJOR) {
    if (p.ln_stats != nullptr) return (!p.out_f32 && p.resid == nullptr) ? (p.act ? E_RM_BF16_GELU_LN : E_RM_BF16_LN) : E_GENERIC;
    if (p.stats_out != nullptr) return (p.out_f32 && !p.act && p.resid != nullptr && p.out_bf16 != nullptr) ? E_RM_F32_RESID_STATS : E_GENERIC;
    if (p.gelu_pre != nullptr) return (!p.out_f32 && !p.act && p.resid == nullptr && p.out_bf16 == nullptr) ? E_RM_BF16_DGELU : E_GENERIC;
    if (p.out_bf16 != nullptr) return (!p.out_f32 && p.act && p.resid == nullptr) ? E_RM_BF16_GELU_PRE : E_GENERIC;
    if (p.out_f32 && !p.act && p.resid == nullptr) return E_RM_F32;
    if (p.out_f32 && !p.act && p.resid != nullptr) return E_RM_F32_RESID;
    if (!p.out_f32 && p.resid == nullptr) return p.act ? E_RM_BF16_GELU : E_RM_BF16;
    return E_GENERIC;
  }
  if (p.mode == EPI_CONV && p.head_out != nullptr) return E_CONV_HEAD;
  if (p.mode == EPI_CONV && !p.out_f32 && p.skip && p.skip_tap < 0 && p.act) return E_CONV_SKIPADD;
  if (p.mode == EPI_CONV && !p.out_f32 && !(p.skip && p.skip_tap < 0)) return p.act ? E_CONV : E_CONV_RAW;   // raw + epilogue skip: generic path
  if (p.mode == EPI_CONVT && !p.out_f32 &&
      (unsigned long long)(p.M / (int)p.in_plane.d + 1) * p.out_plane * p.ldo < (1ull << 31)) return p.act ? E_CONVT : E_CONVT_RAW;
  return E_GENERIC;
}

// One epilogue warp = one TMEM lane quadrant (32 accumulator rows) x one column half of the tile.  Per slab of SC columns:
//   phase A (lane = accumulator row): tcgen05.ld -> + bias (broadcast ld.shared) -> [GELU, bf16 rounding, halo-row zeroing]
//            -> padded shared-memory row; the TMEM buffer is handed back as soon as the warp's last slab is in registers
//   phase B (lanes run along the columns): 16-byte units are read back and stored as coalesced row segments; the fp32
//            residual variant loads its operand there (batched, all loads in flight before the first add)
// arrive on the barrier at the same shared-memory offset in the EVEN CTA of a cta_group::2 pair (cute::Sm100MmaPeerBitMask)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(smem_u32(bar) & 0xFEFFFFFFu) : "memory");
}

// folded LayerNorm, consumer side: {rstd, -mean * rstd} of accumulator row m from the partial sums the producing GEMM left
__device__ __forceinline__ void ln_row_coeffs(const Problem& p, int m, float& a, float& c) {
  float S = 0.0f, Q = 0.0f;
  if (m < p.M) {
    // ln_parts is even (two column halves per tile) and <= 16: all loads are issued before the first add (one L2 latency)
    const float4* sp = reinterpret_cast<const float4*>(p.ln_stats + (size_t)m * p.ln_parts);
    const int n4 = p.ln_parts >> 1;
    float4 t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = i < n4 ? __ldg(sp + i) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { S += t[i].x + t[i].z; Q += t[i].y + t[i].w; }
  }
  const float mean = S * p.ln_inv_k;
  const float var = fmaxf(fmaf(Q, p.ln_inv_k, -mean * mean), 0.0f);
  a = rsqrtf(var + p.ln_eps);
  c = -mean * a;
}
// L2 prefetch of the rows a warp's epilogue is going to read in phase B (fp32 residual rows / bf16 skip rows of its 32 x ncols
// block): issued before the warp waits for its accumulator, so that the loads behind the MMAs find their lines in L2
__device__ __forceinline__ void prefetch_rows_l2(const void* base, size_t row_bytes_pitch, int row0, int rows_valid, uint32_t col_byte0,
                                                 uint32_t bytes_per_row, int lane) {
  const uint32_t lines = (bytes_per_row + 127u) >> 7;                   // 128-byte lines per row
  const uint32_t total = 32u * lines;
  for (uint32_t i = lane; i < total; i += 32) {
    const uint32_t rr = i / lines, l = i - rr * lines;
    if ((int)rr < rows_valid) {
      const char* a = reinterpret_cast<const char*>(base) + (size_t)(row0 + rr) * row_bytes_pitch + col_byte0 + l * 128u;
      asm volatile("prefetch.global.L2 [%0];\n" ::"l"(a));
    }
  }
}
// a warp's private copy of `ncols` (<= 32 * KMAX) consecutive floats of a per-column vector (the ln_s slice of its slabs)
template <int KMAX>
__device__ __forceinline__ void stage_column_slice(const float* __restrict__ src, uint32_t dst_s, int ncols, int lane) {
  float b[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) { const int i = k * 32 + lane; b[k] = i < ncols ? __ldg(src + i) : 0.0f; }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int i = k * 32 + lane;
    if (i < ncols) asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(dst_s + 4u * i), "f"(b[k]) : "memory");
  }
  __syncwarp();
}

// Column sums in a conv epilogue (Problem::bn_sums): every epilogue warp keeps a private fp32 slot [2][N] in shared memory (zeroed
// before its first tile); after the last tile the slots are added up in warp order and each CTA issues one fp64 atomic per entry.
// Named barrier 2 belongs to the `nwarps` epilogue warps.
__device__ __forceinline__ void bn_sums_zero(uint32_t slot_s, int N, int lane) {
  for (int i = lane; i < 2 * N; i += 32) asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(slot_s + 4u * (uint32_t)i), "f"(0.0f) : "memory");
  __syncwarp();
}
__device__ __forceinline__ void bn_sums_flush(uint32_t slots_s, double* __restrict__ sums, int N, int tid, int nwarps) {
  __syncwarp();
  asm volatile("bar.sync 2, %0;\n" ::"r"(nwarps * 32) : "memory");
  for (int i = tid; i < 2 * N; i += nwarps * 32) {
    float acc = 0.0f;
    for (int w = 0; w < nwarps; ++w) {
      float v;
      asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(slots_s + 4u * (uint32_t)(w * 2 * N + i)));
      acc += v;
    }
    atomicAdd(sums + i, (double)acc);
  }
}

// EPI_CONVT with Problem::zero_out_halo: zero the 1-pixel halo of the haloed NHWC output [B, 2Hin+2, 2Win+2, ldo] (16-byte units),
// grid-strided over the epilogue threads of all CTAs; `tid` / `nthreads` number the epilogue threads of this CTA
__device__ __forceinline__ void zero_convt_out_halo(const Problem& p, int tid, int nthreads) {
  const int Ho = 2 * p.Hin, Wo = 2 * p.Win, PW = Wo + 2, PH = Ho + 2;
  const int B = p.M / (int)p.in_plane.d;
  const int per_img = 2 * PW + 2 * Ho, vec = p.ldo / 8;
  const long long total = (long long)B * per_img * vec;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
  for (long long i = (long long)blockIdx.x * nthreads + tid; i < total; i += (long long)gridDim.x * nthreads) {
    const int v = (int)(i % vec);
    const long long tt = i / vec;
    const int k = (int)(tt % per_img), b = (int)(tt / per_img);
    int yy, xx;
    if (k < PW) { yy = 0; xx = k; }
    else if (k < 2 * PW) { yy = Ho + 1; xx = k - PW; }
    else { const int r = k - 2 * PW; yy = 1 + (r >> 1); xx = (r & 1) ? Wo + 1 : 0; }
    *reinterpret_cast<uint4*>(out + (((size_t)b * PH + yy) * PW + xx) * p.ldo + v * 8) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// ZERO: the accumulator columns are zeroed behind the read, before the buffer is handed back (conv_roll.cu: every MMA accumulates)
template <int BN, int EPI, bool PAIR = false, bool ZERO = false>
__device__ __forceinline__ void epilogue_tile(const Problem& p, uint32_t stg_s, uint32_t bias_s, uint32_t taddr,
                                              uint64_t* tmem_empty_bar, int mrow0, int n0, int lane, int slab_begin, int slab_end,
                                              HeadCtx* hc = nullptr, float ln_a = 0.0f, float ln_c = 0.0f, uint32_t sums_s = 0u) {
  using C = Cfg<BN>;
  constexpr bool kHead = EPI == E_CONV_HEAD;
  constexpr bool kLn = EPI == E_RM_BF16_LN || EPI == E_RM_BF16_GELU_LN;          // consumer of a folded LayerNorm
  constexpr bool kStats = EPI == E_RM_F32_RESID_STATS && C::SC == 32;             // producer in front of one (8 lanes per staged row)
  static_assert(!kHead || (C::NSLAB == 1 && C::SC == 48), "the fused head epilogue is built for one 48-column slab");
  constexpr bool kSkipAdd = EPI == E_CONV_SKIPADD;
  constexpr bool kConv = EPI == E_CONV || EPI == E_CONV_RAW || kSkipAdd;
  constexpr bool kConvT = EPI == E_CONVT || EPI == E_CONVT_RAW;
  constexpr bool kF32 = EPI == E_RM_F32 || EPI == E_RM_F32_RESID || EPI == E_RM_F32_RESID_STATS;
  constexpr bool kResid = EPI == E_RM_F32_RESID || EPI == E_RM_F32_RESID_STATS;
  constexpr bool kGelu = EPI == E_RM_BF16_GELU || EPI == E_RM_BF16_GELU_LN || EPI == E_CONV || EPI == E_CONVT || kHead;
  constexpr bool kPre = EPI == E_RM_BF16_GELU_PRE;          // two stores per slab: the bf16 pre-activation, then its GELU
  constexpr bool kDGelu = EPI == E_RM_BF16_DGELU;
  constexpr int UPR = kF32 ? C::SC / 4 : C::SC / 8;        // 16-byte units per staged row (a power of two)
  constexpr uint32_t PITCH = (kF32 ? C::SC * 4 : C::SC * 2) + 16;
  const uint32_t rowpix = kHead ? row_info_head(p, mrow0 + lane) : (EPI == E_GENERIC || kConvT || kConv) ? row_info(p, mrow0 + lane) : 0u;
  const bool zero_row = kConv && !kSkipAdd && (rowpix & kRowHalo) != 0;      // (the skip-add variant zeroes halo rows in phase B)
  const int rows_valid = p.M - mrow0;   // rows rr < rows_valid exist (row-major / conv)
  const uint32_t ldo = (uint32_t)p.ldo;
  if (slab_begin >= slab_end) {          // a one-slab tile leaves the second column half idle: just release the buffer
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) { if (PAIR) mbar_arrive_leader(tmem_empty_bar); else mbar_arrive(tmem_empty_bar); }
    return;
  }
  // partial LayerNorm sums of this warp's rows over the slabs it owns: lane (g = lane >> 3, lane & 7) holds, for j = 0..7, the
  // four columns `lane & 7` of row 4 j + g of every slab
  float ps[kStats ? 8 : 1], pq[kStats ? 8 : 1];
  if constexpr (kStats) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { ps[j] = 0.0f; pq[j] = 0.0f; }
  }
#pragma unroll 1
  for (int slab = slab_begin; slab < slab_end; ++slab) {
    uint32_t r[C::SC];
#pragma unroll
    for (int c = 0; c < C::SC / 16; ++c) tmem_ld16(taddr + slab * C::SC + c * 16, &r[c * 16]);
    tmem_ld_wait();
    if constexpr (ZERO) {
#pragma unroll
      for (int c = 0; c < C::SC / 16; ++c) tmem_st16_zero(taddr + slab * C::SC + c * 16);
      tmem_st_wait();
    }
    if (slab == slab_end - 1) {
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) { if (PAIR) mbar_arrive_leader(tmem_empty_bar); else mbar_arrive(tmem_empty_bar); }
    }
    const int nslab = slab * C::SC;
    if constexpr (EPI == E_GENERIC) {
      // any other flag combination: straight from registers, one accumulator row per lane (slow, rarely used)
#pragma unroll
      for (int c = 0; c < C::SC / 16; ++c) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[c * 16 + i]);
        epilogue_store<16>(p, mrow0 + lane, n0 + nslab + c * 16, v);
      }
    } else {
      // ---- phase A
      const uint32_t bsl = bias_s + (uint32_t)(slab - slab_begin) * (C::SC * 4);
      if constexpr (kLn) {
        // folded LayerNorm: out = rstd * (acc - mean * s[n]) + b'[n] = ln_a * acc + (ln_c * s[n] + b'[n]); this lane's row owns
        // ln_a = rstd, ln_c = -mean * rstd; the s slice sits behind the (bf16-pitched) staging rows of this warp
        const uint32_t lsl = stg_s + 32u * PITCH + (uint32_t)(slab - slab_begin) * (C::SC * 4);
#pragma unroll
        for (int c = 0; c < C::SC / 4; ++c) {
          const uint4 b = lds128(bsl + 16 * c), sv = lds128(lsl + 16 * c);
          r[4 * c] = __float_as_uint(fmaf(ln_a, __uint_as_float(r[4 * c]), fmaf(ln_c, __uint_as_float(sv.x), __uint_as_float(b.x))));
          r[4 * c + 1] = __float_as_uint(fmaf(ln_a, __uint_as_float(r[4 * c + 1]), fmaf(ln_c, __uint_as_float(sv.y), __uint_as_float(b.y))));
          r[4 * c + 2] = __float_as_uint(fmaf(ln_a, __uint_as_float(r[4 * c + 2]), fmaf(ln_c, __uint_as_float(sv.z), __uint_as_float(b.z))));
          r[4 * c + 3] = __float_as_uint(fmaf(ln_a, __uint_as_float(r[4 * c + 3]), fmaf(ln_c, __uint_as_float(sv.w), __uint_as_float(b.w))));
        }
      } else {
#pragma unroll
        for (int c = 0; c < C::SC / 4; ++c) {
          const uint4 b = lds128(bsl + 16 * c);               // warp-uniform address: broadcast
          r[4 * c] = __float_as_uint(__uint_as_float(r[4 * c]) + __uint_as_float(b.x));
          r[4 * c + 1] = __float_as_uint(__uint_as_float(r[4 * c + 1]) + __uint_as_float(b.y));
          r[4 * c + 2] = __float_as_uint(__uint_as_float(r[4 * c + 2]) + __uint_as_float(b.z));
          r[4 * c + 3] = __float_as_uint(__uint_as_float(r[4 * c + 3]) + __uint_as_float(b.w));
        }
      }
      if constexpr (kHead) {
        // fused 1x1 head: GELU, round to bf16 (registers); once the head MMAs of this buffer's PREVIOUS tile have finished with A2
        // and D2 (d2_full, normally long complete: they run right behind this tile's conv MMAs), write this lane's row of the A2
        // operand (16-byte chunk c of row r at chunk position c ^ (r & 7) of the row's 128 bytes: the SWIZZLE_128B pattern,
        // conflict-free for row-per-lane stores), store the previous tile's pixel from D2, then hand A2 to the issuer.  The lanes
        // never wait for the D2 of the tile they just produced.
        const uint32_t rr = hc->row0 + (uint32_t)lane;                         // row inside the 128-row tile
        const uint32_t row_s = hc->a2_s + (rr >> 3) * 1024u + (rr & 7u) * 128u;
        uint4 q[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const float4 v0 = make_float4(__uint_as_float(r[8 * c]), __uint_as_float(r[8 * c + 1]), __uint_as_float(r[8 * c + 2]), __uint_as_float(r[8 * c + 3]));
          const float4 v1 = make_float4(__uint_as_float(r[8 * c + 4]), __uint_as_float(r[8 * c + 5]), __uint_as_float(r[8 * c + 6]), __uint_as_float(r[8 * c + 7]));
          const uint2 o0 = gelu4_to_bf16(v0), o1 = gelu4_to_bf16(v1);
          q[c] = make_uint4(o0.x, o0.y, o1.x, o1.y);
        }
        if (hc->has_prev) {
          mbar_wait(hc->d2_full, hc->parity ^ 1u);
          tcgen05_fence_after();
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) sts128(row_s + ((((uint32_t)c) ^ (rr & 7u)) << 4), q[c]);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");       // generic-proxy stores -> visible to the MMA's async proxy
        if (hc->has_prev) head_store(p, *hc, hc->prev_rowpix);                // before the arrive: head(t) may then overwrite D2
        __syncwarp();
        if (lane == 0) mbar_arrive(hc->a2_full);
        hc->has_prev = true;
        hc->prev_rowpix = rowpix;
        continue;
      }
      if constexpr (kF32) {
#pragma unroll
        for (int c = 0; c < UPR; ++c) sts128(stg_s + lane * PITCH + c * 16, make_uint4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]));
      } else {
#pragma unroll
        for (int c = 0; c < UPR; ++c) {
          const float4 v0 = make_float4(__uint_as_float(r[8 * c]), __uint_as_float(r[8 * c + 1]), __uint_as_float(r[8 * c + 2]), __uint_as_float(r[8 * c + 3]));
          const float4 v1 = make_float4(__uint_as_float(r[8 * c + 4]), __uint_as_float(r[8 * c + 5]), __uint_as_float(r[8 * c + 6]), __uint_as_float(r[8 * c + 7]));
          uint2 o0, o1;
          if constexpr (kSkipAdd) {
            // pre-activation staged as fp16 (the precision GELU is evaluated in anyway); skip + GELU follow in phase B
            const __half2 h0 = __floats2half2_rn(v0.x, v0.y), h1 = __floats2half2_rn(v0.z, v0.w), h2 = __floats2half2_rn(v1.x, v1.y), h3 = __floats2half2_rn(v1.z, v1.w);
            o0 = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
            o1 = make_uint2(*reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
          } else {
            o0 = kGelu ? gelu4_to_bf16(v0) : make_uint2(pack_bf16x2(v0.x, v0.y), pack_bf16x2(v0.z, v0.w));
            o1 = kGelu ? gelu4_to_bf16(v1) : make_uint2(pack_bf16x2(v1.x, v1.y), pack_bf16x2(v1.z, v1.w));
          }
          if (zero_row) { o0 = make_uint2(0u, 0u); o1 = o0; }
          sts128(stg_s + lane * PITCH + c * 16, make_uint4(o0.x, o0.y, o1.x, o1.y));
        }
      }
      __syncwarp();
      if constexpr (EPI == E_CONV_RAW) {
        // column sums of the tensor being stored (Problem::bn_sums): the staged bf16 slab is read back transposed -- lane (g, cp) walks
        // rows g, g + G, ... of column pair cp -- and added to this warp's private fp32 slot [2][N] (no atomics: one fixed summation
        // order per warp, so a training step is reproducible)
        if (sums_s != 0u) {
          constexpr int CP = C::SC / 2, G = 32 / CP;                 // column pairs per slab, row groups per warp (48 -> 24 x 1, 32 -> 16 x 2)
          const int cp = lane % CP, g = lane / CP;
          const int rows_here = rows_valid < 32 ? rows_valid : 32;
          float s0 = 0.0f, s1 = 0.0f, q0 = 0.0f, q1 = 0.0f;
          if (g < G) {
            for (int rr = g; rr < rows_here; rr += G) {
              uint32_t w;
              asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(w) : "r"(stg_s + (uint32_t)rr * PITCH + 4u * (uint32_t)cp));
              const float f0 = __uint_as_float(w << 16), f1 = __uint_as_float(w & 0xFFFF0000u);
              s0 += f0; s1 += f1;
              q0 = fmaf(f0, f0, q0); q1 = fmaf(f1, f1, q1);
            }
          }
#pragma unroll
          for (int sh = CP; sh < 32 && G > 1; sh <<= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, sh); s1 += __shfl_xor_sync(0xffffffffu, s1, sh);
            q0 += __shfl_xor_sync(0xffffffffu, q0, sh); q1 += __shfl_xor_sync(0xffffffffu, q1, sh);
          }
          if (lane < CP) {
            const uint32_t a = sums_s + 4u * (uint32_t)(n0 + nslab + 2 * cp), aq = a + 4u * (uint32_t)p.N;
            float2 t, tq;
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];\n" : "=f"(t.x), "=f"(t.y) : "r"(a));
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];\n" : "=f"(tq.x), "=f"(tq.y) : "r"(aq));
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};\n" ::"r"(a), "f"(t.x + s0), "f"(t.y + s1) : "memory");
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};\n" ::"r"(aq), "f"(tq.x + q0), "f"(tq.y + q1) : "memory");
          }
        }
      }
      // ---- phase B
      constexpr int EPU = kF32 ? 4 : 8;                     // elements per 16-byte unit
      uint4 v[UPR];
      uint32_t off[UPR];
      bool live[UPR];
#pragma unroll
      for (int j = 0; j < UPR; ++j) {
        const int u = j * 32 + lane;
        const int rr = u / UPR, cu = u % UPR;
        v[j] = lds128(stg_s + rr * PITCH + cu * 16);
        const int n = n0 + nslab + cu * EPU;
        if constexpr (kConvT) {
          const uint32_t pix = __shfl_sync(0xffffffffu, rowpix, rr);
          uint32_t qq, o;
          fd_divmod((uint32_t)n, p.cout_div, qq, o);
          live[j] = pix != kRowSkip;
          off[j] = (pix + (qq >> 1) * (uint32_t)p.out_pitch + (qq & 1)) * ldo + o;
        } else {
          live[j] = rr < rows_valid;
          off[j] = (uint32_t)(mrow0 + rr) * ldo + (uint32_t)n;
        }
      }
      if constexpr (kSkipAdd) {
        // ResidualBlock (blocks.py:85): GELU(x + block(x)).  The skip tensor is read here, coalesced, instead of riding through the
        // tensor cores as a tenth identity tap (one A box and 1/9 of the MMAs less per tile); halo rows are stored as zeros.
        uint4 sk[UPR];
        bool halo[UPR];
#pragma unroll
        for (int j = 0; j < UPR; ++j) {
          const int rr = (j * 32 + lane) / UPR;
          halo[j] = (__shfl_sync(0xffffffffu, rowpix, rr) & kRowHalo) != 0;
          sk[j] = (live[j] && !halo[j]) ? __ldg(reinterpret_cast<const uint4*>(p.skip + off[j])) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int j = 0; j < UPR; ++j) {
          uint32_t w[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
          const uint32_t k4[4] = {sk[j].x, sk[j].y, sk[j].z, sk[j].w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 sf = unpack_bf16x2(k4[t]);
            const __half2 z = __hadd2(*reinterpret_cast<const __half2*>(&w[t]), __floats2half2_rn(sf.x, sf.y));
            const float2 g = __half22float2(gelu_h2(z));
            w[t] = halo[j] ? 0u : pack_bf16x2(g.x, g.y);
          }
          v[j] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      if constexpr (kResid) {
        float4 e[UPR];
#pragma unroll
        for (int j = 0; j < UPR; ++j) e[j] = live[j] ? *reinterpret_cast<const float4*>(p.resid + off[j]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < UPR; ++j) {
          v[j].x = __float_as_uint(__uint_as_float(v[j].x) + e[j].x); v[j].y = __float_as_uint(__uint_as_float(v[j].y) + e[j].y);
          v[j].z = __float_as_uint(__uint_as_float(v[j].z) + e[j].z); v[j].w = __float_as_uint(__uint_as_float(v[j].w) + e[j].w);
        }
      }
      if constexpr (kDGelu) {
        // fc2 data gradient times the GELU derivative at fc1's pre-activation (read here, coalesced; its lines were prefetched to L2)
        uint4 z[UPR];
#pragma unroll
        for (int j = 0; j < UPR; ++j) z[j] = live[j] ? __ldg(reinterpret_cast<const uint4*>(p.gelu_pre + off[j])) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int j = 0; j < UPR; ++j) {
          uint32_t w[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
          const uint32_t z4[4] = {z[j].x, z[j].y, z[j].z, z[j].w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 g = unpack_bf16x2(w[t]), zz = unpack_bf16x2(z4[t]);
            const float2 d = gelu_grad2(zz.x, zz.y);
            w[t] = pack_bf16x2(g.x * d.x, g.y * d.y);
          }
          v[j] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      if constexpr (kStats) {
        // the new residual row also leaves as bf16 (A operand of the GEMM behind the folded LayerNorm) and feeds the row sums
#pragma unroll
        for (int j = 0; j < UPR; ++j) {
          const float a = __uint_as_float(v[j].x), b = __uint_as_float(v[j].y), c = __uint_as_float(v[j].z), d = __uint_as_float(v[j].w);
          if (live[j]) {
            ps[j] += (a + b) + (c + d);
            pq[j] = fmaf(a, a, fmaf(b, b, fmaf(c, c, fmaf(d, d, pq[j]))));
            *reinterpret_cast<uint2*>(p.out_bf16 + off[j]) = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
          }
        }
      }
#pragma unroll
      for (int j = 0; j < UPR; ++j) {
        if (live[j]) {
          if constexpr (kF32) *reinterpret_cast<uint4*>(reinterpret_cast<float*>(p.out) + off[j]) = v[j];
          else if constexpr (kPre) *reinterpret_cast<uint4*>(p.out_bf16 + off[j]) = v[j];
          else *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off[j]) = v[j];
        }
      }
      __syncwarp();
      if constexpr (kPre) {
        // second pass over the same slab: GELU of the bf16-rounded pre-activation (what msd_gelu_fwd_bf16 computes from the stored tensor)
#pragma unroll
        for (int c = 0; c < UPR; ++c) {
          const float2 a0 = unpack_bf16x2(pack_bf16x2(__uint_as_float(r[8 * c]), __uint_as_float(r[8 * c + 1])));
          const float2 a1 = unpack_bf16x2(pack_bf16x2(__uint_as_float(r[8 * c + 2]), __uint_as_float(r[8 * c + 3])));
          const float2 a2 = unpack_bf16x2(pack_bf16x2(__uint_as_float(r[8 * c + 4]), __uint_as_float(r[8 * c + 5])));
          const float2 a3 = unpack_bf16x2(pack_bf16x2(__uint_as_float(r[8 * c + 6]), __uint_as_float(r[8 * c + 7])));
          const uint2 o0 = gelu4_to_bf16(make_float4(a0.x, a0.y, a1.x, a1.y)), o1 = gelu4_to_bf16(make_float4(a2.x, a2.y, a3.x, a3.y));
          sts128(stg_s + lane * PITCH + c * 16, make_uint4(o0.x, o0.y, o1.x, o1.y));
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < UPR; ++j) {
          const int u = j * 32 + lane;
          const uint4 h = lds128(stg_s + (u / UPR) * PITCH + (u % UPR) * 16);
          if (live[j]) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off[j]) = h;
        }
        __syncwarp();
      }
    }
  }
  if constexpr (kStats) {
    // reduce-scatter over the 8 lanes that share a row (xor 4, 2, 1): 14 shuffles for the 2 x 8 partial sums; afterwards lane l
    // holds the totals of row 4 (l & 7) + (l >> 3) over this warp's columns -> slot 2 (n0 / BN) + column half of that row
    const bool b2 = (lane & 4) != 0, b1 = (lane & 2) != 0, b0 = (lane & 1) != 0;
    float t[4], tq[4], u[2], uq[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      t[i] = (b2 ? ps[i + 4] : ps[i]) + __shfl_xor_sync(0xffffffffu, b2 ? ps[i] : ps[i + 4], 4);
      tq[i] = (b2 ? pq[i + 4] : pq[i]) + __shfl_xor_sync(0xffffffffu, b2 ? pq[i] : pq[i + 4], 4);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      u[i] = (b1 ? t[i + 2] : t[i]) + __shfl_xor_sync(0xffffffffu, b1 ? t[i] : t[i + 2], 2);
      uq[i] = (b1 ? tq[i + 2] : tq[i]) + __shfl_xor_sync(0xffffffffu, b1 ? tq[i] : tq[i + 2], 2);
    }
    const float w = (b0 ? u[1] : u[0]) + __shfl_xor_sync(0xffffffffu, b0 ? u[0] : u[1], 1);
    const float wq = (b0 ? uq[1] : uq[0]) + __shfl_xor_sync(0xffffffffu, b0 ? uq[0] : uq[1], 1);
    const int rr = (lane & 7) * 4 + (lane >> 3);
    if (rr < rows_valid) p.stats_out[(size_t)(mrow0 + rr) * p.stats_parts + (2 * (n0 / BN) + (slab_begin ? 1 : 0))] = make_float2(w, wq);
  }
}

// One whole tile of the resident-weight 3x3 conv mode with one k-block per tap (<= 64 channels), issued as straight-line code:
// STEPS = 3 (dy) or 4 (+ identity skip tap) pipeline stages, NK valid 16-wide K slices, windows dx = -1, 0, +1 at +0 / +128 /
// +256 bytes of the stage's A box, weights at tap * W_BOX in the resident matrix.  a_base / r_lo are (address >> 4).
template <int BN, int STEPS, int NK, uint32_t A_STAGE_BYTES_>
__device__ __forceinline__ void issue_tile_resident(uint32_t tmem_d, uint32_t a_base, uint32_t r_lo, uint32_t idesc, uint64_t* full,
                                                    uint64_t* empty, uint64_t* tmem_full_bar, uint32_t parity, int skip_tap) {
  constexpr uint32_t HI = (1024u >> 4) | (1u << 14) | (2u << 29);
  constexpr uint32_t WB = (BN * BLOCK_K * 2) >> 4;
  const uint64_t da0 = ((uint64_t)HI << 32) | a_base, db0 = ((uint64_t)HI << 32) | r_lo;
#pragma unroll
  for (int st = 0; st < STEPS; ++st) {
    mbar_wait(&full[st], parity);
    tcgen05_fence_after();
    if (elect_one()) {
      const uint64_t da = da0 + (uint64_t)(st * (A_STAGE_BYTES_ >> 4));
      if (st < 3) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int k = 0; k < NK; ++k)
            tcgen05_mma_bf16(tmem_d, da + (uint64_t)(8 * j + 2 * k), db0 + (uint64_t)((3 * st + j) * WB + 2 * k), idesc, (st | j | k) != 0 ? 1u : 0u);
      } else {                                           // identity skip tap through the middle window
        const uint64_t db = db0 + (uint64_t)((uint32_t)skip_tap * WB);
#pragma unroll
        for (int k = 0; k < NK; ++k) tcgen05_mma_bf16(tmem_d, da + (uint64_t)(8 + 2 * k), db + (uint64_t)(2 * k), idesc, 1u);
      }
      tcgen05_commit(&empty[st]);
      if (st == STEPS - 1) tcgen05_commit(tmem_full_bar);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ kernel
template <int BN, int MODE>
__global__ void __launch_bounds__((Cfg<BN, MODE>::THREADS), 1)
gemm_tc5_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmS, const Problem p, int tiles_m, int tiles_n, int epi) {
  using C = Cfg<BN, MODE>;
  constexpr uint32_t A_STAGE_BYTES = C::A_STAGE_BYTES;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* smem_a = smem;
  unsigned char* smem_b = smem + C::kStages * A_STAGE_BYTES;                 // per-stage W boxes (MODE 0 / 1)
  unsigned char* smem_res = smem + C::kStages * C::STAGE_BYTES;              // resident weights (MODE 2)
  unsigned char* smem_stg = smem_res + C::RESIDENT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + C::STG_BYTES);
  uint64_t* full = bars;                     // [2][kStages]: one set per MMA issuer (tile parity), see the issuer loop
  uint64_t* empty = bars + 2 * C::kStages;
  uint64_t* tmem_full = bars + 3 * C::kStages;
  uint64_t* tmem_empty = tmem_full + C::NBUF;
  uint64_t* res_full = tmem_empty + C::NBUF;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&full[C::kStages + i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < C::NBUF; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], C::EPI_WARPS / C::NBUF); }
    mbar_init(res_full, 1);
    if constexpr (MODE == MODE_MERGE3_RESIDENT && BN == 48) {
      static_assert(C::kStages >= 8 && 2 * BN + 32 <= (int)C::TMEM_COLS && HEAD_SLOT_BYTES <= C::W_BOX_BYTES, "fused head: needs ring stages 6/7, 32 spare TMEM columns and the tenth-tap slot");
      if (epi == E_CONV_HEAD) {          // hand-off barriers of the fused 1x1 head (in the free tenth-tap slot, see HeadCtx)
        uint64_t* hb = reinterpret_cast<uint64_t*>(smem_res + 9 * C::W_BOX_BYTES + HEAD_BARS);
        mbar_init(&hb[0], 4); mbar_init(&hb[1], 4);      // a2_full[buf]: the four epilogue warps of the buffer's group
        mbar_init(&hb[2], 1); mbar_init(&hb[3], 1);      // d2_full[buf]: tcgen05.commit
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_ptr)), "r"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  msd_pdl_enter();      // everything above ran under the previous kernel's tail; no global memory has been touched yet
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);   // same word for every lane; shuffle makes it provably uniform
  const int num_tiles = tiles_m * tiles_n;

  // merged modes walk (dy, k-block) steps, then the k-blocks of the identity skip tap
  const int kbpt = p.kb_per_tap;
  const int n_dy = 3, has_skip = p.skip_tap >= 0 ? 1 : 0;
  // 16-wide K slices of the last k-block of a tap that hold valid channels (K = 48 or 96 does not fill it)
  const int nk_last = min(BLOCK_K / 16, (p.kvalid - (kbpt - 1) * BLOCK_K + 15) / 16);
  // Ring length.  The resident-weight mode uses exactly TWO tiles' worth of stages (6 or 8 <= kStages): tile `it` then always
  // occupies stages (it & 1) * steps .. + steps, i.e. each MMA issuer (even / odd tiles) sees the SAME stages every tile and its
  // whole per-tile instruction stream has loop-invariant descriptors (issue_tile_resident below).
  const int ring = mode_resident(MODE) ? min(C::kStages, 2 * (n_dy + has_skip) * kbpt) : C::kStages;

  // Producer and MMA warps run their loops with all 32 lanes converged and let one elected lane issue the
  // asynchronous instructions.  Both loops are the serial critical path of a CTA (one warp each, nothing to hide
  // instruction latency behind), so they are written as plain nested counters: no division, no per-MMA descriptor
  // rebuild -- descriptors are a 32-bit base + small adds, the high word is a constant.
  if (warp == 0) {
    // ===== TMA producer =====
    int stage = 0; uint32_t phase = 0;
    if constexpr (mode_resident(MODE)) {
      // the whole packed weight matrix (<= 10 taps x one k-block) stays in shared memory for the life of the CTA
      if (elect_one()) {
        mbar_arrive_expect_tx(res_full, (uint32_t)p.num_kb * C::W_BOX_BYTES);
        for (int kb = 0; kb < p.num_kb; ++kb) tma_load_2d(&tmB, res_full, smem_res + kb * C::W_BOX_BYTES, kb * BLOCK_K, 0);
      }
      __syncwarp();
    }
    const int tap0 = p.tap_off[0], tap3 = p.tap_off[3], tap6 = p.tap_off[6];
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int tm = tile / tiles_n;
      const int m0 = tm * BLOCK_M, n0 = (tile - tm * tiles_n) * BN;
      uint64_t* full = bars + (it & 1u) * C::kStages;      // the issuer that owns this tile listens on its own barrier set
      if constexpr (MODE == MODE_PLAIN) {
        int tap = 0, kc = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full[stage], C::A_BOX_BYTES + C::W_BOX_BYTES);
            // the skip connection of a ResidualBlock rides along as one more tap: A = skip tensor, W = identity
            tma_load_2d(tap == p.skip_tap ? &tmS : &tmA, &full[stage], smem_a + stage * A_STAGE_BYTES, kc * BLOCK_K, m0 + p.tap_off[tap] - p.dbg_shift);
            tma_load_2d(&tmB, &full[stage], smem_b + stage * C::B_STAGE_BYTES, kb * BLOCK_K, n0);
          }
          __syncwarp();
          if (++kc == kbpt) { kc = 0; ++tap; }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      } else {
        // rows m0 + off(dy, dx=-1) .. +135: windows at +0 / +1 / +2 rows are the dx = -1 / 0 / +1 taps;
        // the skip tensor is loaded one row early too and read through the middle window
        for (int dy = 0; dy < n_dy + has_skip; ++dy) {
          const bool is_skip = dy == n_dy;
          const int row = is_skip ? m0 - 1 : m0 + (dy == 0 ? tap0 : dy == 1 ? tap3 : tap6);
          const int wcol0 = (is_skip ? p.skip_tap : 3 * dy) * kbpt * BLOCK_K;
          for (int kc = 0; kc < kbpt; ++kc) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one()) {
              if constexpr (MODE == MODE_MERGE3) {
                mbar_arrive_expect_tx(&full[stage], C::A_BOX_BYTES + (is_skip ? 1u : 3u) * C::W_BOX_BYTES);
                unsigned char* sb = smem_b + stage * C::B_STAGE_BYTES;
                tma_load_2d(&tmB, &full[stage], sb, wcol0 + kc * BLOCK_K, n0);
                if (!is_skip) {
                  tma_load_2d(&tmB, &full[stage], sb + C::W_BOX_BYTES, wcol0 + (kbpt + kc) * BLOCK_K, n0);
                  tma_load_2d(&tmB, &full[stage], sb + 2 * C::W_BOX_BYTES, wcol0 + (2 * kbpt + kc) * BLOCK_K, n0);
                }
              } else {
                mbar_arrive_expect_tx(&full[stage], C::A_BOX_BYTES);
              }
              tma_load_2d(is_skip ? &tmS : &tmA, &full[stage], smem_a + stage * A_STAGE_BYTES, kc * BLOCK_K, row);
            }
            __syncwarp();
            if (++stage == ring) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===== MMA issuers: warp 1 takes this CTA's even tiles (accumulator buffer 0), warp 3 the odd ones (buffer 1).
    // One warp's issue loop costs several hundred cycles of dependent scalar work per tile, which is more than the
    // tensor time of a narrow tile (N = 48: 27 MMAs x 24 cycles); two issuers overlap that latency.  Both walk the same
    // shared-memory ring: a warp skips the other's tile by advancing its ring position by one tile's worth of stages.
    constexpr uint32_t idesc = make_idesc(BLOCK_M, BN);
    const uint32_t par = warp == 1 ? 0u : 1u;
    const uint32_t a_lo0 = (smem_u32(smem_a) & 0x3FFFFu) >> 4, b_lo0 = (smem_u32(smem_b) & 0x3FFFFu) >> 4;
    const uint32_t r_lo0 = (smem_u32(smem_res) & 0x3FFFFu) >> 4;
    const int steps_per_tile = MODE == MODE_PLAIN ? p.num_kb : (n_dy + has_skip) * kbpt;
    int gstep = 0;
    uint32_t it = 0;
    // Each issuer has its own `full` barrier per stage (the producer signals the set of the tile's owner), and keeps one
    // parity bit per stage: a parity wait is only unambiguous for a waiter that has seen every earlier phase itself.
    uint64_t* full = bars + par * C::kStages;
    uint32_t pmask = 0;
    if constexpr (mode_resident(MODE)) { mbar_wait(res_full, 0); tcgen05_fence_after(); }
    // fused 1x1 head (E_CONV_HEAD): D2 = A2 . (W2hi + W2lo)^T into 16 spare TMEM columns once the tile's epilogue lanes wrote A2
    bool head_pending = false;
    uint32_t head_parity = 0;
    auto issue_head = [&](uint32_t hpar) {
      if constexpr (MODE == MODE_MERGE3_RESIDENT && BN == 48) {
        uint64_t* hb = reinterpret_cast<uint64_t*>(smem_res + 9 * C::W_BOX_BYTES + HEAD_BARS);
        mbar_wait(&hb[par], hpar);
        tcgen05_fence_after();
        if (elect_one()) {
          constexpr uint32_t HI = (1024u >> 4) | (1u << 14) | (2u << 29);
          constexpr uint32_t idesc2 = make_idesc(BLOCK_M, 16);
          const uint32_t a2_lo = a_lo0 + (6u + par) * (A_STAGE_BYTES >> 4);
          const uint32_t w2_lo = r_lo0 + 9u * (C::W_BOX_BYTES >> 4);
          const uint32_t d2 = tmem_base + 2 * BN + par * 16;
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) {
            const uint64_t da = ((uint64_t)HI << 32) | (a2_lo + 2u * kk);
            tcgen05_mma_bf16(d2, da, ((uint64_t)HI << 32) | (w2_lo + 2u * kk), idesc2, kk ? 1u : 0u);
            tcgen05_mma_bf16(d2, da, ((uint64_t)HI << 32) | (w2_lo + (HEAD_W2_LO >> 4) + 2u * kk), idesc2, 1u);
          }
          tcgen05_commit(&hb[2 + par]);
        }
        __syncwarp();
      }
    };
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it, gstep += steps_per_tile) {
      if ((it & 1u) != par) continue;
      int stage = gstep % ring;
      // accumulator buffer of tile `it`: it mod NBUF (this issuer's n-th tile alternates between its two buffers when there are four)
      const uint32_t buf = it & (uint32_t)(C::NBUF - 1), uses = it / (uint32_t)C::NBUF;
      const uint32_t tmem_d = tmem_base + buf * BN;
      mbar_wait(&tmem_empty[buf], (uses & 1u) ^ 1u);
      tcgen05_fence_after();
      uint32_t acc = 0;                                   // first MMA of the tile overwrites the accumulator
      if constexpr (MODE == MODE_PLAIN) {
        int kc = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full[stage], (pmask >> stage) & 1u);
          pmask ^= 1u << stage;
          tcgen05_fence_after();
          const int nk = (kc == kbpt - 1) ? nk_last : BLOCK_K / 16;
          if (elect_one()) {
            const uint32_t a_lo = a_lo0 + stage * (A_STAGE_BYTES >> 4) + 8u * p.dbg_shift;
            const uint32_t b_lo = b_lo0 + stage * (C::B_STAGE_BYTES >> 4);
            mma_kblock(tmem_d, a_lo, b_lo, (uint32_t)(p.dbg_bo & 7) << 17, idesc, nk, acc);
            tcgen05_commit(&empty[stage]);               // frees the smem stage once these MMAs retire
            if (kb == p.num_kb - 1) tcgen05_commit(&tmem_full[buf]);  // accumulator complete -> epilogue
          }
          __syncwarp();
          acc = 1;
          if (++kc == kbpt) kc = 0;
          if (++stage == C::kStages) stage = 0;
        }
      } else {
        if constexpr (mode_resident(MODE)) {
          if (kbpt == 1 && nk_last == 3 && ring == 2 * (n_dy + has_skip)) {
            // 48-channel stage: straight-line issue, every descriptor = a per-issuer base + a compile-time offset
            const uint32_t a_base = a_lo0 + par * (uint32_t)(n_dy + has_skip) * (A_STAGE_BYTES >> 4);
            uint64_t* fullp = full + par * (n_dy + has_skip);
            uint64_t* emptyp = empty + par * (n_dy + has_skip);
            const uint32_t parity = (it >> 1) & 1u;
            if (has_skip) issue_tile_resident<BN, 4, 3, A_STAGE_BYTES>(tmem_d, a_base, r_lo0, idesc, fullp, emptyp, &tmem_full[buf], parity, p.skip_tap);
            else issue_tile_resident<BN, 3, 3, A_STAGE_BYTES>(tmem_d, a_base, r_lo0, idesc, fullp, emptyp, &tmem_full[buf], parity, 0);
            if constexpr (BN == 48) {
              if (epi == E_CONV_HEAD) {
                // fused 1x1 head of this issuer's PREVIOUS tile: its epilogue lanes wrote A2 while this tile's MMAs were being
                // queued, so the conv MMAs never wait on an epilogue (issuing head(t) before tile t+2 put phase A of the epilogue,
                // ~2,300 cycles, on the tensor core's critical path: 0.79 ms instead of 0.5)
                if (head_pending) issue_head(head_parity);
                head_pending = true; head_parity = parity;
              }
            }
            pmask ^= ((1u << (n_dy + has_skip)) - 1u) << (par * (n_dy + has_skip));
            continue;
          }
        }
        for (int dy = 0; dy < n_dy + has_skip; ++dy) {
          const bool is_skip = dy == n_dy;
          for (int kc = 0; kc < kbpt; ++kc) {
            mbar_wait(&full[stage], (pmask >> stage) & 1u);
            pmask ^= 1u << stage;
            tcgen05_fence_after();
            const int nk = (kc == kbpt - 1) ? nk_last : BLOCK_K / 16;
            const bool last = (dy == n_dy + has_skip - 1) && (kc == kbpt - 1);
            if (elect_one()) {
              const uint32_t a_lo = a_lo0 + stage * (A_STAGE_BYTES >> 4);
              if (is_skip) {                                                 // identity tap through the middle window
                const uint32_t b_lo = mode_resident(MODE) ? r_lo0 + (p.skip_tap * kbpt + kc) * (C::W_BOX_BYTES >> 4)
                                                                     : b_lo0 + stage * (C::B_STAGE_BYTES >> 4);
                mma_kblock(tmem_d, a_lo + 8u, b_lo, 0u, idesc, nk, acc);
              } else {
                const uint32_t b_lo = mode_resident(MODE) ? r_lo0 + (3 * dy * kbpt + kc) * (C::W_BOX_BYTES >> 4)
                                                                     : b_lo0 + stage * (C::B_STAGE_BYTES >> 4);
                const uint32_t b_step = (mode_resident(MODE) ? kbpt : 1) * (C::W_BOX_BYTES >> 4);
                mma_kblock(tmem_d, a_lo, b_lo, 0u, idesc, nk, acc);                       // dx = -1
                mma_kblock(tmem_d, a_lo + 8u, b_lo + b_step, 0u, idesc, nk, 1u);           // dx =  0 (one row = 128 B down)
                mma_kblock(tmem_d, a_lo + 16u, b_lo + 2u * b_step, 0u, idesc, nk, 1u);     // dx = +1
              }
              tcgen05_commit(&empty[stage]);
              if (last) tcgen05_commit(&tmem_full[buf]);
            }
            __syncwarp();
            acc = 1;
            if (++stage == ring) stage = 0;
          }
        }
      }
    }
    if (head_pending) issue_head(head_parity);           // the last tile of this issuer
  } else if (warp >= 4) {
    // ===== epilogue: warp w may only read TMEM lanes 32*(w%4).. (hardware rule).  e = w-4: bit 2 selects the accumulator
    // buffer (even / odd tiles of this CTA), bit 3 the column half of the tile =====
    const int e = warp - 4;
    if (p.mode == EPI_CONVT && p.zero_out_halo) zero_convt_out_halo(p, (int)threadIdx.x - 128, C::EPI_WARPS * 32);
    // two buffers: bit 2 of e = accumulator buffer, bit 3 = column half; four buffers: bits 2-3 = accumulator buffer, one column range
    const int q = warp & 3, grp = (e >> 2) & (C::NBUF - 1), half = C::NBUF == 4 ? 0 : e >> 3;
    const uint32_t stg_s = smem_u32(smem_stg + e * C::STG_WARP_BYTES);
    const uint32_t bias_s = stg_s + 32 * C::STG_PITCH;
    const int slab_begin = half ? C::SLAB_SPLIT : 0, slab_end = half ? C::NSLAB : C::SLAB_SPLIT;
    const int ncols = (slab_end - slab_begin) * C::SC;
    if constexpr (MODE == MODE_MERGE3_RESIDENT && BN == 48) {
      if (epi == E_CONV_HEAD) {
        // the head conv has nine taps: the tenth (skip) slot of the resident weight area holds the 1x1 head's W2 operand
        build_head_operand(p, smem_u32(smem_res + 9 * C::W_BOX_BYTES), (int)threadIdx.x - 128, C::EPI_WARPS * 32);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("bar.sync 1, %0;\n" ::"n"(C::EPI_WARPS * 32) : "memory");
      }
    }
    // column sums of the stored tensor (training BatchNorm statistics): one slot per epilogue warp in the free tenth-tap slot
    uint32_t sums_s = 0u;
    if constexpr (MODE == MODE_MERGE3_RESIDENT && BN == 48) {
      if (p.bn_sums != nullptr && epi == E_CONV_RAW) {
        sums_s = smem_u32(smem_res + 9 * C::W_BOX_BYTES) + (uint32_t)e * (uint32_t)(2 * BN * 4);
        bn_sums_zero(sums_s, BN, lane);
      }
    }
    int it = 0, bias_n0 = -1;
    HeadCtx hc;                       // fused 1x1 head: this warp's pending tile (E_CONV_HEAD only)
    hc.has_prev = false; hc.prev_rowpix = kRowSkip; hc.parity = 0; hc.d2_full = nullptr;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      if ((it & (C::NBUF - 1)) != grp) continue;
      const uint32_t as = it & (C::NBUF - 1), aphase = (it / C::NBUF) & 1;
      const int tm = tile / tiles_n;
      const int m0 = tm * BLOCK_M, n0 = (tile - tm * tiles_n) * BN;
      // this warp's bias slice -> smem (private to the warp), fetched before waiting for the accumulator and only when
      // the tile column changed (never again for N == BN, i.e. every conv layer)
      if (n0 != bias_n0) {
        bias_n0 = n0;
        float b[(C::SLAB_SPLIT * C::SC + 31) / 32];
#pragma unroll
        for (int k = 0; k < (C::SLAB_SPLIT * C::SC + 31) / 32; ++k) {
          const int i = k * 32 + lane;
          b[k] = (p.bias != nullptr && i < ncols) ? __ldg(p.bias + n0 + slab_begin * C::SC + i) : 0.0f;
        }
        __syncwarp();                                     // the previous tile's readers are done
#pragma unroll
        for (int k = 0; k < (C::SLAB_SPLIT * C::SC + 31) / 32; ++k) {
          const int i = k * 32 + lane;
          if (i < ncols) asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(bias_s + 4u * i), "f"(b[k]) : "memory");
        }
        __syncwarp();
        if constexpr (MODE == MODE_PLAIN) {
          if (p.ln_s != nullptr)      // folded LayerNorm: the column sums of the weights, behind the bf16-pitched staging rows
            stage_column_slice<(C::SLAB_SPLIT * C::SC + 31) / 32>(p.ln_s + n0 + slab_begin * C::SC, stg_s + 32u * (C::SC * 2 + 16), ncols, lane);
        }
      }
      float ln_a = 0.0f, ln_c = 0.0f;
      if constexpr (MODE == MODE_PLAIN) {
        if (epi == E_RM_BF16_LN || epi == E_RM_BF16_GELU_LN) ln_row_coeffs(p, m0 + q * 32 + lane, ln_a, ln_c);
        if (epi == E_RM_F32_RESID || epi == E_RM_F32_RESID_STATS)
          prefetch_rows_l2(p.resid, (size_t)p.ldo * 4, m0 + q * 32, p.M - (m0 + q * 32), (uint32_t)(n0 + slab_begin * C::SC) * 4u, (uint32_t)ncols * 4u, lane);
      }
      if (epi == E_CONV_SKIPADD)
        prefetch_rows_l2(p.skip, (size_t)p.ldo * 2, m0 + q * 32, p.M - (m0 + q * 32), (uint32_t)(n0 + slab_begin * C::SC) * 2u, (uint32_t)ncols * 2u, lane);
      if constexpr (MODE == MODE_PLAIN) {
        if (epi == E_RM_BF16_DGELU)
          prefetch_rows_l2(p.gelu_pre, (size_t)p.ldo * 2, m0 + q * 32, p.M - (m0 + q * 32), (uint32_t)(n0 + slab_begin * C::SC) * 2u, (uint32_t)ncols * 2u, lane);
      }
      mbar_wait(&tmem_full[as], aphase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
      if constexpr (MODE == MODE_MERGE3_RESIDENT && BN == 48) {
        if (epi == E_CONV_HEAD) {
          uint64_t* hb = reinterpret_cast<uint64_t*>(smem_res + 9 * C::W_BOX_BYTES + HEAD_BARS);
          hc.a2_s = smem_u32(smem_a + (6 + as) * A_STAGE_BYTES);
          hc.bias_s = smem_u32(smem_res + 9 * C::W_BOX_BYTES + HEAD_BIAS);
          hc.taddr_d2 = tmem_base + ((uint32_t)(q * 32) << 16) + 2 * BN + as * 16;
          hc.a2_full = &hb[as]; hc.d2_full = &hb[2 + as]; hc.parity = aphase; hc.row0 = (uint32_t)(q * 32);
          epilogue_tile<BN, E_CONV_HEAD>(p, stg_s, bias_s, taddr, &tmem_empty[as], m0 + q * 32, n0, lane, slab_begin, slab_end, &hc);
          continue;
        }
      }
      if constexpr (MODE == MODE_MERGE3_RESIDENT4) {
        // 16 epilogue warps cap the kernel at 96 registers: only the three conv epilogues this mode is launched with are compiled in
        if (epi == E_CONV) epilogue_tile<BN, E_CONV>(p, stg_s, bias_s, taddr, &tmem_empty[as], m0 + q * 32, n0, lane, slab_begin, slab_end);
        else if (epi == E_CONV_SKIPADD) epilogue_tile<BN, E_CONV_SKIPADD>(p, stg_s, bias_s, taddr, &tmem_empty[as], m0 + q * 32, n0, lane, slab_begin, slab_end);
        else epilogue_tile<BN, E_CONV_RAW>(p, stg_s, bias_s, taddr, &tmem_empty[as], m0 + q * 32, n0, lane, slab_begin, slab_end);
        continue;
      }
      // one specialised instruction stream per fused epilogue (the switch is per tile, the loops inside are branch-free)
#define MSD_EPI_CASE(E) case E: epilogue_tile<BN, E>(p, stg_s, bias_s, taddr, &tmem_empty[as], m0 + q * 32, n0, lane, slab_begin, slab_end); break;
      switch (epi) {
        MSD_EPI_CASE(E_RM_F32)
        MSD_EPI_CASE(E_RM_BF16)
        MSD_EPI_CASE(E_RM_BF16_GELU)
        MSD_EPI_CASE(E_RM_F32_RESID)
        MSD_EPI_CASE(E_CONV)
        MSD_EPI_CASE(E_CONVT)
        case E_CONV_RAW: epilogue_tile<BN, E_CONV_RAW>(p, stg_s, bias_s, taddr, &tmem_empty[as], m0 + q * 32, n0, lane, slab_begin, slab_end, nullptr, 0.0f, 0.0f, sums_s); break;
        MSD_EPI_CASE(E_CONVT_RAW)
        MSD_EPI_CASE(E_CONV_SKIPADD)
        MSD_EPI_CASE(E_RM_F32_RESID_STATS)
        MSD_EPI_CASE(E_RM_BF16_GELU_PRE)
        MSD_EPI_CASE(E_RM_BF16_DGELU)
        case E_RM_BF16_LN: epilogue_tile<BN, E_RM_BF16_LN>(p, stg_s, bias_s, taddr, &tmem_empty[as], m0 + q * 32, n0, lane, slab_begin, slab_end, nullptr, ln_a, ln_c); break;
        case E_RM_BF16_GELU_LN: epilogue_tile<BN, E_RM_BF16_GELU_LN>(p, stg_s, bias_s, taddr, &tmem_empty[as], m0 + q * 32, n0, lane, slab_begin, slab_end, nullptr, ln_a, ln_c); break;
        default: epilogue_tile<BN, E_GENERIC>(p, stg_s, bias_s, taddr, &tmem_empty[as], m0 + q * 32, n0, lane, slab_begin, slab_end); break;
      }
#undef MSD_EPI_CASE
    }
    if constexpr (MODE == MODE_MERGE3_RESIDENT && BN == 48) {
      if (p.bn_sums != nullptr && epi == E_CONV_RAW)
        bn_sums_flush(smem_u32(smem_res + 9 * C::W_BOX_BYTES), p.bn_sums, BN, (int)threadIdx.x - 128, C::EPI_WARPS);
      if (epi == E_CONV_HEAD && hc.has_prev) {           // the last tile of this warp's accumulator buffer
        mbar_wait(hc.d2_full, hc.parity);
        tcgen05_fence_after();
        head_store(p, hc, hc.prev_rowpix);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}


// host: launch one BN instantiation (instantiated per BN in gemm_tc5_bn*.cu)
int make_map(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t rows, uint64_t ld, uint32_t box_rows);
int make_map_planes_f32(CUtensorMap* tm, const void* base, uint64_t W, uint64_t H, uint64_t planes, uint32_t box_w, uint32_t box_h);
int num_sms();

template <int BN, int MODE>
int launch_mode(const Problem& p, cudaStream_t s) {
  using C = Cfg<BN, MODE>;
  const uint32_t a_box_rows = MODE == MODE_PLAIN ? BLOCK_M : MERGE_ROWS;
  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, p.A, (uint64_t)p.kvalid, (uint64_t)p.a_rows, (uint64_t)p.ldA, a_box_rows);
  if (rc) return rc;
  rc = make_map(&tmB, p.Wt, (uint64_t)p.num_kb * BLOCK_K, (uint64_t)p.N, (uint64_t)(p.w_ld_kb ? p.w_ld_kb : p.num_kb) * BLOCK_K, BN);
  if (rc) return rc;
  CUtensorMap tmS = tmA;
  if (p.skip_tap >= 0) {
    rc = make_map(&tmS, p.skip, (uint64_t)p.kvalid, (uint64_t)p.a_rows, (uint64_t)p.ldA, a_box_rows);
    if (rc) return rc;
  }
  static PerDeviceOnce attr_set;
  if (attr_set.first()) {
    MSD_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc5_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
  }
  const int tiles_m = (p.M + BLOCK_M - 1) / BLOCK_M, tiles_n = p.N / BN;
  const int tiles = tiles_m * tiles_n;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  MSD_CHECK_CUDA(msd_launch_pdl(gemm_tc5_kernel<BN, MODE>, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, s, tmA, tmB, tmS, p, tiles_m, tiles_n, classify(p)));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

// is this Problem a 3x3 conv whose taps are laid out as (dy, dx) row shifts of one pitch?  (what MODE 1/2 assume)
__host__ inline bool is_conv3x3(const Problem& p) {
  if (p.mode != EPI_CONV || p.num_kb < 9 * p.kb_per_tap) return false;
  for (int dy = 0; dy < 3; ++dy)
    if (p.tap_off[3 * dy + 1] != p.tap_off[3 * dy] + 1 || p.tap_off[3 * dy + 2] != p.tap_off[3 * dy] + 2) return false;
  return p.tap_off[4] == 0 && p.dbg_shift == 0;
}

// can the 1x1 head ride in this conv's epilogue?  (48-channel 3x3 conv without skip in the resident-weight mode)
__host__ inline bool can_fuse_head(const Problem& p, int head_cout) {
  static const bool off = getenv("MSD_NO_MERGE3") != nullptr || getenv("MSD_NO_HEAD_FUSION") != nullptr;
  return !off && is_conv3x3(p) && p.N == 48 && p.kb_per_tap == 1 && p.num_kb == 9 && p.skip_tap < 0 && p.act && head_cout >= 1 &&
         head_cout <= 16 && (unsigned long long)(p.M / (int)p.plane.d + 1) * head_cout * p.H * p.W < (1ull << 31);
}

template <int BN>
int launch_bn(const Problem& p, cudaStream_t s) {
  if constexpr (BN == 48) {
    if (p.head_out != nullptr) {
      MSD_REQUIRE(can_fuse_head(p, p.head_cout) && p.head_w != nullptr, "gemm_tc5: this conv cannot carry the fused 1x1 head");
      return launch_mode<BN, MODE_MERGE3_RESIDENT>(p, s);
    }
  } else {
    MSD_REQUIRE(p.head_out == nullptr, "gemm_tc5: the fused 1x1 head needs the 48-channel conv kernel");
  }
  if constexpr (BN == 48 || BN == 96) {
    static const bool no_merge = getenv("MSD_NO_MERGE3") != nullptr;     // A/B switch for the profiles
    if (!no_merge && is_conv3x3(p)) {
      if constexpr (BN == 48) {
        // measured at B = 256: 0.562 ms vs 0.540 ms with two buffers / 8 epilogue warps (conv 192^2 x 48) and 0.643 vs 0.653 for the
        // skip variant -- the epilogue is NOT what bounds this conv (L2 -> SM delivery of the three dy boxes + the shared-memory
        // reads of N = 48 MMAs are); kept as an opt-in A/B mode (MSD_RES4=1), off by default
        static const bool res4 = getenv("MSD_RES4") != nullptr;
        const int e = classify(p);
        if (res4 && p.N == BN && p.kb_per_tap == 1 && p.num_kb == 9 && (e == E_CONV || e == E_CONV_RAW || e == E_CONV_SKIPADD))
          return launch_mode<BN, MODE_MERGE3_RESIDENT4>(p, s);
        if (p.N == BN && p.kb_per_tap == 1 && p.num_kb <= 10) return launch_mode<BN, MODE_MERGE3_RESIDENT>(p, s);
      }
      return launch_mode<BN, MODE_MERGE3>(p, s);
    }
  }
  return launch_mode<BN, MODE_PLAIN>(p, s);
}

}  // namespace tc5
}  // namespace gemm
