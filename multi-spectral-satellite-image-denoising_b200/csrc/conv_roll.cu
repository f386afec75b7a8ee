// Rolling-row 3x3 convolution for the 48-channel decoder stage (blocks.py:66-85, decoder stage 4: 192 x 192 x 48) on tcgen05.
//
// Why a second formulation: in the tap-major kernels of gemm_tc5.cuh every tap is its own M128 x N48 x K16 MMA, and each of them
// fetches its 4 KB A window + 1.5 KB of weights from shared memory for 24 cycles of math -- ncu puts that layer at 70 % of the
// tensor cores' shared-memory operand bandwidth with the tensor pipe 39 % busy.  N = 48 amortises an A fetch over too
// little math.  Here the three VERTICAL taps of a column dx are stacked along N instead:
//
//     E[y_in][x, (dy, n)] = sum_k A[y_in, x + dx, k] . W[(dy, dx)][n, k]           one M128 x N144 x K16 MMA per (dx, k slice)
//     out[y][x, n]        = sum_dy E[y + dy][x, (dy, n)]
//
// The sum over dy needs no data movement at all: a tile is 128 pixels of ONE image row, so the three 48-column groups of an input
// row's MMA belong to the SAME TMEM lanes of three different output rows -- and a TMEM column offset is free.  Output rows live in
// a ring of R accumulator slots (48 fp32 columns each); input row s accumulates into the slots of output rows s-2, s-1, s (one
// N = 144 MMA when the three slots are contiguous, two when the ring wraps between them), and output row s-2 is complete when
// input row s has been issued.  Per 128 output pixels: 11-12 MMAs instead of 27, A fetched from shared memory 1.2x instead of 9x,
// and every input row crosses L2 -> shared memory ONCE instead of three times (one 136-row TMA box; the dx taps are row-shifted
// windows of it, as in MODE_MERGE3 of gemm_tc5.cuh).
//
// Geometry: an image row of W = 128 a + 64 b pixels is cut into `a` full strips (128 lanes = 128 consecutive x) and, for b = 1,
// one PAIR strip: lanes 0-63 = the 64 remaining columns of row y, lanes 64-127 = the same columns of row y + H/2 (the two halves
// of the image march down together; the dx windows of a pair tile are three separately loaded tiles of two 64-row boxes).  A
// strip is a sequence of "steps" (output rows); all steps of the launch form one list that is cut into gridDim.x equal ranges, one
// per CTA; a range is walked segment by segment (a segment = consecutive output rows of one strip; it costs two extra input rows).
//
// Every MMA ACCUMULATES: a slot is zeroed by the epilogue warp that drained it (tcgen05.st behind the tcgen05.ld, before the slot
// is handed back), so the issuer's per-row instruction stream has no first-touch special case -- nine MMAs with one D address
// (eighteen when the ring wraps inside the row's three slots) and a few commits.  That stream must stay short: the issuer shares
// its scheduler and instruction cache with epilogue warps, and a first version with per-MMA descriptor arithmetic and a
// three-way split of each row's first MMA spent ~2,000 cycles per row in the issuer alone (0.63 ms per layer, slower than tap-major).
//
// Roles: warp 0 TMA producer, warp 1 (and 5) MMA issuer, warp 2 TMEM allocator, warp 3 issuer of the fused head's 1x1, then three
// or four epilogue groups of four warps (RollCfg below); group g drains output rows u = g (mod GROUPS) with the fused epilogues
// of gemm_tc5.cuh (bias + GELU [+ skip], bf16, coalesced).
// Only interior pixels are computed; the kernel zeroes the halo of its output buffer itself (epilogue threads, before their first tile).
#include "gemm_tc5.cuh"

namespace gemm {
namespace roll {

using namespace tc5;

constexpr int CH = 48;                       // channels in = out
constexpr int NK = CH / 16;                  // K = 16 slices per tap
constexpr int R = 8;                         // accumulator ring slots (output rows in flight): 384 of the 512 TMEM columns
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t UNIT_BYTES = MERGE_ROWS * BLOCK_K * 2;       // one A ring unit: a 136-row box (full strip) or two 64-row boxes
constexpr uint32_t FULL_TX = MERGE_ROWS * BLOCK_K * 2, PAIR_TX = 2 * 64 * BLOCK_K * 2;
constexpr uint32_t W_BOX = CH * BLOCK_K * 2;                    // 6144 B: one tap's weights, K-major SWIZZLE_128B
constexpr uint32_t RES_BYTES = 9 * W_BOX;
constexpr uint32_t STG_WARP_BYTES = 32 * (CH * 2 + 16) + CH * 4;   // bf16 staging rows + the warp's bias copy
constexpr uint32_t BAR_BYTES = 512;                               // <= 41 mbarriers + the TMEM base word
constexpr uint32_t A2_BYTES = BLOCK_M * BLOCK_K * 2;
// Two layouts of the 640 threads, chosen per epilogue:
//   three epilogue groups + TWO MMA issuers (warps 1 and 5) -- plain and fused-head epilogues.  One issuer's per-row stream (barrier
//     waits, descriptor arithmetic, 9-18 MMAs, commits) costs ~1,600 cycles of which ~500 are not spent waiting for the tensor
//     core's queue; the second issuer takes every other input row and hides them.  fp32 accumulation is not associative, so
//     the two streams must reach the tensor core in ROW ORDER: an issuer queues its row's MMAs only when the other has handed
//     over the turn behind its own row's last MMA.  That token only orders the MMAs if both warps sit on the SAME SM
//     sub-partition (warp % 4 equal: they share one dispatch path to the tensor core) -- with warps 1 and 3 results differed
//     in the last bit from run to run; with warps 1 and 5 they are bit-identical run to run and across batch compositions
//     (tools/roll_bench.py checks both, tests compare bit for bit).  0.487 -> 0.449 ms per layer, fused head 0.496 -> 0.454.
//   four epilogue groups + one issuer -- the skip-add epilogue, whose longer per-row epilogue needs the fourth group more than
//     the second issuer (0.570 vs 0.597 ms).
// Fused decoder head (E_CONV_HEAD, decoder.py:76-81; the scheme of gemm_tc5.cuh): the epilogue lanes write the GELU'd row as the
// bf16 A2 operand of a second contraction D2[128 x 16] = A2[128 x 48] . W2^T (hi + lo bf16 split of the fp32 1x1 weights), issued
// by a warp of its own (warp 3) into 16 spare TMEM columns per epilogue group, and store NCHW fp32 from D2 while they handle
// their next row.  No staging rows are needed then: the area holds one A2 tile per group, W2 + biases, one bias copy.
template <int EPI>
struct RollCfg {
  static constexpr bool kHead = EPI == E_CONV_HEAD;
  static constexpr int GROUPS = EPI == E_CONV_SKIPADD ? 4 : 3;
  static constexpr bool TWO = GROUPS == 3;                       // two MMA issuers (warps 1 and 5)
  static constexpr int EPI_WARPS = 4 * GROUPS, EPI_BASE = TWO ? 8 : 4, THREADS = 32 * (EPI_BASE + EPI_WARPS);
  static constexpr int UNITS = TWO ? 7 : 6;
  // (training convs, E_CONV_RAW: + one [2][48] fp32 slot per epilogue warp for the BatchNorm column sums, Problem::bn_sums)
  static constexpr uint32_t SUMS_BYTES = EPI == E_CONV_RAW ? EPI_WARPS * 2 * CH * 4 : 0;
  static constexpr uint32_t STG_BYTES = kHead ? GROUPS * A2_BYTES + 4352 /* HEAD_SLOT_BYTES rounded */ + CH * 4 : EPI_WARPS * STG_WARP_BYTES + SUMS_BYTES;
  static constexpr uint32_t SMEM_BYTES = UNITS * UNIT_BYTES + RES_BYTES + STG_BYTES + 1024 + BAR_BYTES;
  static_assert(THREADS == 640, "640 threads keep 96 registers per thread for the epilogue warps");
  static_assert(SMEM_BYTES <= 227u * 1024u, "shared memory budget");
  static_assert(HEAD_SLOT_BYTES <= 4352 && R * CH + GROUPS * 16 <= (int)TMEM_COLS && 2 * UNITS + 2 * R + 1 + 2 * GROUPS + 2 <= 63, "fused head / barrier budget");
};
static_assert(UNIT_BYTES % 1024 == 0 && W_BOX % 1024 == 0, "operand tiles must stay 1024-byte aligned for SWIZZLE_128B");
static_assert(R * CH <= (int)TMEM_COLS, "accumulator ring");

struct Geo {
  int B, H, W;
  int nfull, has_pair;      // full 128-pixel strips per row, one 64-pixel pair strip
  int steps_img;            // nfull * H + has_pair * H / 2
  long long total;          // B * steps_img
};

struct Seg {
  int img, strip;           // strip == nfull: the pair strip
  int o0, n;                // output rows o0 .. o0 + n - 1 of the strip (0-based; pair strips: of each half)
};

// next segment of the range [t, t1)
__device__ __forceinline__ bool next_seg(const Geo& g, long long& t, long long t1, Seg& s) {
  if (t >= t1) return false;
  const int img = (int)(t / g.steps_img);
  const int r = (int)(t - (long long)img * g.steps_img);
  int strip, o, ulen;
  if (r < g.nfull * g.H) { strip = r / g.H; o = r - strip * g.H; ulen = g.H; }
  else { strip = g.nfull; o = r - g.nfull * g.H; ulen = g.H >> 1; }
  const long long left = t1 - t;
  const int n = (long long)(ulen - o) < left ? ulen - o : (int)left;
  s.img = img; s.strip = strip; s.o0 = o; s.n = n;
  t += n;
  return true;
}

template <int EPI>
__global__ void __launch_bounds__(RollCfg<EPI>::THREADS, 1)
conv_roll_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA64,
                 const __grid_constant__ CUtensorMap tmB, const Problem p, const Geo g) {
  using K = RollCfg<EPI>;
  constexpr int GROUPS = K::GROUPS, EPI_WARPS = K::EPI_WARPS, EPI_BASE = K::EPI_BASE, UNITS = K::UNITS;
  constexpr bool TWO = K::TWO;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* smem_a = smem;
  unsigned char* smem_res = smem + UNITS * UNIT_BYTES;
  unsigned char* smem_stg = smem_res + RES_BYTES;
  constexpr bool kHead = K::kHead;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stg + K::STG_BYTES);
  uint64_t* full = bars;                    // [UNITS]  producer -> issuer
  uint64_t* empty = full + UNITS;           // [UNITS]  issuer (tcgen05.commit) -> producer
  uint64_t* slot_full = empty + UNITS;      // [R]      issuer (tcgen05.commit) -> epilogue group
  uint64_t* slot_empty = slot_full + R;     // [R]      epilogue group (4 warps) -> issuer
  uint64_t* res_full = slot_empty + R;
  uint64_t* a2_full = res_full + 1;         // [GROUPS] epilogue group -> issuer: the A2 tile of the fused head is written
  uint64_t* d2_full = a2_full + GROUPS;     // [GROUPS] issuer (tcgen05.commit) -> epilogue group: D2 is complete
  uint64_t* turn = d2_full + GROUPS;        // [2] issuer -> the other issuer: my row's MMAs are issued
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(turn + 2);
  unsigned char* smem_a2 = smem_stg;                                   // fused head: [GROUPS] A2 tiles, then W2 hi/lo + biases, then the conv bias
  unsigned char* smem_hslot = smem_stg + GROUPS * A2_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmA64)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < UNITS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < R; ++i) { mbar_init(&slot_full[i], 3); mbar_init(&slot_empty[i], 4); }     // three input rows feed an output row
    mbar_init(&turn[0], 1); mbar_init(&turn[1], 1);
    mbar_init(res_full, 1);
    for (int i = 0; i < GROUPS; ++i) { mbar_init(&a2_full[i], 4); mbar_init(&d2_full[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  msd_pdl_enter();      // the prologue above ran under the previous kernel's tail; no global memory touched yet
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  // this CTA's range of steps
  const long long t_begin = g.total * (long long)blockIdx.x / (long long)gridDim.x;
  const long long t_end = g.total * (long long)(blockIdx.x + 1) / (long long)gridDim.x;
  const int PW = g.W + 2, PH = g.H + 2;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      // weights resident for the life of the CTA; slot dx * 3 + j holds tap (ky = 2 - j, kx = dx): the three slots of a dx form one
      // 144-row K-major B operand whose 48-row groups j = 0, 1, 2 feed output rows s-2, s-1, s of input row s
      mbar_arrive_expect_tx(res_full, RES_BYTES);
      for (int dx = 0; dx < 3; ++dx)
        for (int j = 0; j < 3; ++j) tma_load_2d(&tmB, res_full, smem_res + (dx * 3 + j) * W_BOX, ((2 - j) * 3 + dx) * BLOCK_K, 0);
    }
    __syncwarp();
    int unit = 0; uint32_t phase = 0;
    long long t = t_begin;
    Seg sg;
    while (next_seg(g, t, t_end, sg)) {
      const bool pair = sg.strip == g.nfull;
      const int x0 = sg.strip * 128;                              // padded x of the first box column (= x - 1 of the first output pixel)
      const int row_base = (sg.img * PH + sg.o0) * PW + x0;       // flattened padded pixel of input row s = 0
      const int row_half = (g.H >> 1) * PW;
      for (int s = 0; s < sg.n + 2; ++s) {
        const int row = row_base + s * PW;
        if (!pair) {
          mbar_wait(&empty[unit], phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full[unit], FULL_TX);
            tma_load_2d(&tmA, &full[unit], smem_a + unit * UNIT_BYTES, 0, row);
          }
          __syncwarp();
          if (++unit == UNITS) { unit = 0; phase ^= 1; }
        } else {
          for (int dx = 0; dx < 3; ++dx) {
            mbar_wait(&empty[unit], phase ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&full[unit], PAIR_TX);
              tma_load_2d(&tmA64, &full[unit], smem_a + unit * UNIT_BYTES, 0, row + dx);
              tma_load_2d(&tmA64, &full[unit], smem_a + unit * UNIT_BYTES + 64 * BLOCK_K * 2, 0, row + row_half + dx);
            }
            __syncwarp();
            if (++unit == UNITS) { unit = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 || (TWO && warp == 5)) {
    // ===== MMA issuer(s): warp 1 alone, or warps 1 and 5 on alternate input rows with the turn token (RollCfg). =====
    constexpr uint32_t HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
    constexpr uint32_t IDESC0 = make_idesc(BLOCK_M, 0);              // + (N >> 3) << 17
    constexpr uint32_t WB = W_BOX >> 4, UB = UNIT_BYTES >> 4;
    const uint32_t a_lo0 = (smem_u32(smem_a) & 0x3FFFFu) >> 4, r_lo0 = (smem_u32(smem_res) & 0x3FFFFu) >> 4;
    mbar_wait(res_full, 0);
    tcgen05_fence_after();
    int unit = 0; uint32_t phase = 0;
    uint32_t ubase = 0;                      // output rows this CTA has started before the current segment
    long long t = t_begin;
    Seg sg;
    const uint32_t me = warp == 1 ? 0u : 1u;
    uint32_t rowctr = 0, turn_par = me ^ 1u;
    while (next_seg(g, t, t_end, sg)) {
      const bool pair = sg.strip == g.nfull;
      const int n = sg.n;
      for (int s = 0; s < n + 2; ++s) {
        // valid groups j of this input row (output row o = s - 2 + j inside [0, n)); interior rows: 0 .. 2
        const int jlo = s >= 2 ? 0 : 2 - s, jhi = s < n ? 2 : n + 1 - s;
        const uint32_t u0 = ubase + (uint32_t)s;              // u0 - 2 + j = ring index of group j's output row
        if (TWO && ((rowctr++) & 1u) != me) {                 // the other issuer's row: step over its ring units
          unit += pair ? 3 : 1;
          if (unit >= UNITS) { unit -= UNITS; phase ^= 1; }
          continue;
        }
        if (jhi == 2) {                                       // a slot is touched for the first time: its previous use must be drained and zeroed
          mbar_wait(&slot_empty[u0 % R], (u0 / R) & 1u);
          tcgen05_fence_after();
        }
        // the groups as one or (when the ring wraps between them) two runs of contiguous slots
        const uint32_t slot_lo = (u0 + (uint32_t)jlo + (uint32_t)(R - 2)) % R;
        const int ng = jhi - jlo + 1;
        const int n_a = (int)slot_lo + ng <= R ? ng : R - (int)slot_lo, n_b = ng - n_a;
        const uint32_t d_a = tmem_base + slot_lo * CH, id_a = IDESC0 | ((uint32_t)(n_a * CH >> 3) << 17), id_b = IDESC0 | ((uint32_t)(n_b * CH >> 3) << 17);
        const uint32_t b_a = r_lo0 + (uint32_t)jlo * WB, b_b = b_a + (uint32_t)n_a * WB;
        if (!pair) {
          mbar_wait(&full[unit], phase);
          if (TWO) { mbar_wait(&turn[me], turn_par); turn_par ^= 1u; }
          tcgen05_fence_after();
          if (elect_one()) {
            const uint32_t a_lo = a_lo0 + (uint32_t)unit * UB;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
              for (int k = 0; k < NK; ++k) {
                const uint64_t da = ((uint64_t)HI << 32) | (a_lo + 8u * dx + 2u * k);        // window dx: one row = 128 B down
                tcgen05_mma_bf16(d_a, da, ((uint64_t)HI << 32) | (b_a + (uint32_t)(dx * 3) * WB + 2u * k), id_a, 1u);
                if (n_b > 0) tcgen05_mma_bf16(tmem_base, da, ((uint64_t)HI << 32) | (b_b + (uint32_t)(dx * 3) * WB + 2u * k), id_b, 1u);
              }
            if (TWO) mbar_arrive(&turn[me ^ 1u]);
            tcgen05_commit(&empty[unit]);
            for (int j = jlo; j <= jhi; ++j) tcgen05_commit(&slot_full[(u0 + (uint32_t)j + (uint32_t)(R - 2)) % R]);
          }
          __syncwarp();
          if (++unit == UNITS) { unit = 0; phase ^= 1; }
        } else {
          for (int dx = 0; dx < 3; ++dx) {
            mbar_wait(&full[unit], phase);
            if (TWO && dx == 0) { mbar_wait(&turn[me], turn_par); turn_par ^= 1u; }
            tcgen05_fence_after();
            if (elect_one()) {
              const uint32_t a_lo = a_lo0 + (uint32_t)unit * UB;
#pragma unroll
              for (int k = 0; k < NK; ++k) {
                const uint64_t da = ((uint64_t)HI << 32) | (a_lo + 2u * k);
                tcgen05_mma_bf16(d_a, da, ((uint64_t)HI << 32) | (b_a + (uint32_t)(dx * 3) * WB + 2u * k), id_a, 1u);
                if (n_b > 0) tcgen05_mma_bf16(tmem_base, da, ((uint64_t)HI << 32) | (b_b + (uint32_t)(dx * 3) * WB + 2u * k), id_b, 1u);
              }
              if (TWO && dx == 2) mbar_arrive(&turn[me ^ 1u]);
              tcgen05_commit(&empty[unit]);
              if (dx == 2)
                for (int j = jlo; j <= jhi; ++j) tcgen05_commit(&slot_full[(u0 + (uint32_t)j + (uint32_t)(R - 2)) % R]);
            }
            __syncwarp();
            if (++unit == UNITS) { unit = 0; phase ^= 1; }
          }
        }
      }
      ubase += (uint32_t)n;
    }
  } else if (warp == 3) {
    // ===== fused head: its own issuer warp.  D2 of an epilogue group is a separate accumulator written by this thread only, so
    // the 1x1 contractions need no ordering against the conv MMAs, and the conv issuer never waits for an epilogue. =====
    if constexpr (kHead) {
      constexpr uint32_t HI = (1024u >> 4) | (1u << 14) | (2u << 29);
      asm volatile("bar.sync 1, %0;\n" ::"n"(EPI_WARPS * 32 + 32) : "memory");     // W2 operand built by the epilogue warps
      tcgen05_fence_after();
      const uint32_t nrows = (uint32_t)(t_end - t_begin);
      for (uint32_t v = 0; v < nrows; ++v) {
        const uint32_t grp = v % GROUPS;
        mbar_wait(&a2_full[grp], (v / GROUPS) & 1u);
        tcgen05_fence_after();
        if (elect_one()) {
          constexpr uint32_t idesc2 = make_idesc(BLOCK_M, 16);
          const uint32_t a2_lo = ((smem_u32(smem_a2) & 0x3FFFFu) >> 4) + grp * (A2_BYTES >> 4);
          const uint32_t w2_lo = (smem_u32(smem_hslot) & 0x3FFFFu) >> 4;
          const uint32_t d2 = tmem_base + R * CH + grp * 16;
#pragma unroll
          for (int kk = 0; kk < NK; ++kk) {
            const uint64_t da = ((uint64_t)HI << 32) | (a2_lo + 2u * kk);
            tcgen05_mma_bf16(d2, da, ((uint64_t)HI << 32) | (w2_lo + 2u * kk), idesc2, kk ? 1u : 0u);
            tcgen05_mma_bf16(d2, da, ((uint64_t)HI << 32) | (w2_lo + (HEAD_W2_LO >> 4) + 2u * kk), idesc2, 1u);
          }
          tcgen05_commit(&d2_full[grp]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= EPI_BASE) {
    // ===== epilogue: warp w reads TMEM lanes 32 (w % 4) ..; group (w - EPI_BASE) / 4 owns output rows u = group (mod GROUPS) =====
    const int e = warp - EPI_BASE, q = warp & 3, grp = e >> 2;
    const uint32_t stg_s = kHead ? 0u : smem_u32(smem_stg + e * STG_WARP_BYTES);
    const uint32_t bias_s = kHead ? smem_u32(smem_hslot + 4352) : stg_s + 32 * (CH * 2 + 16);    // fused head: one copy for all warps
    if (!kHead || e == 0) {
      float b[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) { const int i = k * 32 + lane; b[k] = (p.bias != nullptr && i < CH) ? __ldg(p.bias + i) : 0.0f; }
#pragma unroll
      for (int k = 0; k < 2; ++k) { const int i = k * 32 + lane; if (i < CH) asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(bias_s + 4u * i), "f"(b[k]) : "memory"); }
      __syncwarp();
    }
    // the accumulator ring starts zeroed: this warp's 32 lanes of the slots its group drains (phase 0 of slot_empty)
    for (int slot = grp; slot < R; slot += GROUPS) {
      const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)slot * CH;
#pragma unroll
      for (int c = 0; c < CH / 16; ++c) tmem_st16_zero(ta + c * 16);
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&slot_empty[slot]);
    }
    uint32_t sums_s = 0u;
    if constexpr (EPI == E_CONV_RAW) {
      if (p.bn_sums != nullptr) {
        sums_s = smem_u32(smem_stg + EPI_WARPS * STG_WARP_BYTES) + (uint32_t)e * (uint32_t)(2 * CH * 4);
        bn_sums_zero(sums_s, CH, lane);
      }
    }
    HeadCtx hc;
    hc.has_prev = false; hc.prev_rowpix = kRowSkip; hc.parity = 0; hc.d2_full = nullptr;
    if constexpr (kHead) {
      build_head_operand(p, smem_u32(smem_hslot), (int)threadIdx.x - 32 * EPI_BASE, EPI_WARPS * 32);
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      asm volatile("bar.sync 1, %0;\n" ::"n"(EPI_WARPS * 32 + 32) : "memory");     // + the issuer warp; also publishes the shared bias copy
      hc.a2_s = smem_u32(smem_a2 + grp * A2_BYTES);
      hc.bias_s = smem_u32(smem_hslot) + HEAD_BIAS;
      hc.taddr_d2 = tmem_base + ((uint32_t)(q * 32) << 16) + R * CH + grp * 16;
      hc.a2_full = &a2_full[grp]; hc.d2_full = &d2_full[grp]; hc.row0 = (uint32_t)(q * 32);
    }
    // zero the halo of the output buffer (only interior pixels are stored below): 2 (W + 2) + 2 H pixels per image, 16-byte units
    if constexpr (!kHead) {
      __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
      const int per_img = 2 * PW + 2 * g.H, vec = CH / 8;
      const long long total = (long long)g.B * per_img * vec;
      for (long long i = (long long)blockIdx.x * (EPI_WARPS * 32) + (threadIdx.x - 32 * EPI_BASE); i < total; i += (long long)gridDim.x * (EPI_WARPS * 32)) {
        const int v = (int)(i % vec);
        const long long tt = i / vec;
        const int k = (int)(tt % per_img), b = (int)(tt / per_img);
        int yy, xx;
        if (k < PW) { yy = 0; xx = k; }
        else if (k < 2 * PW) { yy = g.H + 1; xx = k - PW; }
        else { const int r = k - 2 * PW; yy = 1 + (r >> 1); xx = (r & 1) ? g.W + 1 : 0; }
        *reinterpret_cast<uint4*>(out + (((size_t)b * PH + yy) * PW + xx) * CH + v * 8) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    uint32_t ubase = 0;
    long long t = t_begin;
    Seg sg;
    while (next_seg(g, t, t_end, sg)) {
      const bool pair = sg.strip == g.nfull;
      // flattened padded pixel of this warp's first accumulator row, output row o = 0 of the segment
      int m_base = (sg.img * PH + sg.o0 + 1) * PW + sg.strip * 128 + 1;
      if (!pair) m_base += 32 * q;
      else m_base += (q < 2 ? 32 * q : (g.H >> 1) * PW + 32 * (q - 2));
      // this group's rows of the segment: u = ubase + o with u = grp (mod GROUPS)
      for (int o = (int)(((uint32_t)grp + (uint32_t)GROUPS - ubase % (uint32_t)GROUPS) % (uint32_t)GROUPS); o < sg.n; o += GROUPS) {
        const uint32_t u = ubase + (uint32_t)o;
        const uint32_t slot = u % R;
        const int mrow0 = m_base + o * PW;
        if constexpr (EPI == E_CONV_SKIPADD)
          prefetch_rows_l2(p.skip, (size_t)p.ldo * 2, mrow0, 32, 0u, (uint32_t)CH * 2u, lane);
        mbar_wait(&slot_full[slot], (u / R) & 1u);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + slot * CH;
        if constexpr (kHead) {
          hc.parity = (u / GROUPS) & 1u;
          epilogue_tile<CH, EPI, false, true>(p, stg_s, bias_s, taddr, &slot_empty[slot], mrow0, 0, lane, 0, 1, &hc);
        } else if constexpr (EPI == E_CONV_RAW) {
          epilogue_tile<CH, EPI, false, true>(p, stg_s, bias_s, taddr, &slot_empty[slot], mrow0, 0, lane, 0, 1, nullptr, 0.0f, 0.0f, sums_s);
        } else {
          epilogue_tile<CH, EPI, false, true>(p, stg_s, bias_s, taddr, &slot_empty[slot], mrow0, 0, lane, 0, 1);
        }
      }
      ubase += (uint32_t)sg.n;
    }
    if constexpr (EPI == E_CONV_RAW) {
      if (p.bn_sums != nullptr) bn_sums_flush(smem_u32(smem_stg + EPI_WARPS * STG_WARP_BYTES), p.bn_sums, CH, (int)threadIdx.x - 32 * EPI_BASE, EPI_WARPS);
    }
    if constexpr (kHead) {
      if (hc.has_prev) {                                   // the last row of this warp's group
        mbar_wait(hc.d2_full, hc.parity);
        tcgen05_fence_after();
        head_store(p, hc, hc.prev_rowpix);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------- host
bool geometry_ok(int H, int W, int C) {
  // (read on every call, not cached: the parity test switches between this kernel and the tap-major one inside one process)
  const char* off = getenv("MSD_NO_ROLL");
  if ((off != nullptr && off[0] == '1') || C != CH || H < 1) return false;
  const int rem = W % 128;
  return W >= 64 && (rem == 0 || (rem == 64 && H % 2 == 0));
}

bool eligible(const Problem& p) {
  if (p.mode != EPI_CONV || p.N != CH || p.kvalid != CH || p.ldA != CH || p.ldo != CH || p.kb_per_tap != 1 || p.num_kb != 9 || p.skip_tap >= 0 ||
      p.out_f32 || p.dbg_shift != 0 || p.dbg_bo != 0 || !is_conv3x3(p)) return false;
  if (!geometry_ok(p.H, p.W, p.N)) return false;
  const int e = classify(p);
  if (e == E_CONV_HEAD) return p.act && p.skip == nullptr && p.head_w != nullptr && p.head_cout >= 1 && p.head_cout <= 16 &&
                               (unsigned long long)(p.M / (int)p.plane.d + 1) * p.head_cout * p.H * p.W < (1ull << 31);
  if (p.bn_sums != nullptr && e != E_CONV_RAW) return false;
  return e == E_CONV || e == E_CONV_SKIPADD || e == E_CONV_RAW;     // (raw: the training step's pre-BatchNorm convs and data gradients)
}

int launch(const Problem& p, cudaStream_t s) {
  Geo g;
  g.H = p.H; g.W = p.W; g.B = p.M / (int)p.plane.d;
  MSD_REQUIRE(g.B >= 1 && (long long)g.B * (int)p.plane.d == p.M && (int)p.pitch.d == p.W + 2, "conv_roll: inconsistent geometry");
  g.nfull = p.W / 128; g.has_pair = (p.W % 128) ? 1 : 0;
  g.steps_img = g.nfull * g.H + g.has_pair * (g.H / 2);
  g.total = (long long)g.B * g.steps_img;
  CUtensorMap tmA, tmA64, tmB;
  int rc = make_map(&tmA, p.A, (uint64_t)p.kvalid, (uint64_t)p.a_rows, (uint64_t)p.ldA, MERGE_ROWS);
  if (rc) return rc;
  rc = make_map(&tmA64, p.A, (uint64_t)p.kvalid, (uint64_t)p.a_rows, (uint64_t)p.ldA, 64);
  if (rc) return rc;
  rc = make_map(&tmB, p.Wt, (uint64_t)9 * BLOCK_K, (uint64_t)p.N, (uint64_t)(p.w_ld_kb ? p.w_ld_kb : p.num_kb) * BLOCK_K, CH);
  if (rc) return rc;
  const int epi = classify(p);
  static PerDeviceOnce attr_set;
  if (attr_set.first()) {
    MSD_CHECK_CUDA(cudaFuncSetAttribute(conv_roll_kernel<E_CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RollCfg<E_CONV>::SMEM_BYTES));
    MSD_CHECK_CUDA(cudaFuncSetAttribute(conv_roll_kernel<E_CONV_SKIPADD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RollCfg<E_CONV_SKIPADD>::SMEM_BYTES));
    MSD_CHECK_CUDA(cudaFuncSetAttribute(conv_roll_kernel<E_CONV_HEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RollCfg<E_CONV_HEAD>::SMEM_BYTES));
    MSD_CHECK_CUDA(cudaFuncSetAttribute(conv_roll_kernel<E_CONV_RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RollCfg<E_CONV_RAW>::SMEM_BYTES));
  }
  const long long sms = num_sms();
  const int grid = (int)(g.total < sms ? g.total : sms);
  if (epi == E_CONV_HEAD) MSD_CHECK_CUDA(msd_launch_pdl(conv_roll_kernel<E_CONV_HEAD>, dim3(grid), dim3(RollCfg<E_CONV_HEAD>::THREADS), RollCfg<E_CONV_HEAD>::SMEM_BYTES, s, tmA, tmA64, tmB, p, g));
  else if (epi == E_CONV_RAW) MSD_CHECK_CUDA(msd_launch_pdl(conv_roll_kernel<E_CONV_RAW>, dim3(grid), dim3(RollCfg<E_CONV_RAW>::THREADS), RollCfg<E_CONV_RAW>::SMEM_BYTES, s, tmA, tmA64, tmB, p, g));
  else if (epi == E_CONV_SKIPADD) MSD_CHECK_CUDA(msd_launch_pdl(conv_roll_kernel<E_CONV_SKIPADD>, dim3(grid), dim3(RollCfg<E_CONV_SKIPADD>::THREADS), RollCfg<E_CONV_SKIPADD>::SMEM_BYTES, s, tmA, tmA64, tmB, p, g));
  else MSD_CHECK_CUDA(msd_launch_pdl(conv_roll_kernel<E_CONV>, dim3(grid), dim3(RollCfg<E_CONV>::THREADS), RollCfg<E_CONV>::SMEM_BYTES, s, tmA, tmA64, tmB, p, g));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

}  // namespace roll

bool roll_geometry_ok(int H, int W, int C) { return roll::geometry_ok(H, W, C); }

}  // namespace gemm
