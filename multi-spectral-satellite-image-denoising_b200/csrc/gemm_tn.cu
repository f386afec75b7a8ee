// Weight-gradient GEMMs of the training step on tcgen05: contractions over the ROW index of two activation matrices,
//
//     D[i, j] += sum_p  A[p (+ tap shift), i] * B[p, j]             (bf16 x bf16 -> fp32, split over p across CTAs)
//
// Both operands are "MN-major" for the tensor core (the contracted index p is the slow one in memory), which UMMA reads
// directly from 128B-swizzled shared memory: a TMA box {64 columns, KB rows} IS the canonical MN-major SWIZZLE_128B atom
// stack (8 rows x 128 B per atom, SBO = 1024 B between 8-row groups, LBO = distance between 64-column atoms), so no
// transposed copies of the activations are ever written (reference: autograd of nn.Linear / nn.Conv2d / ConvTranspose2d,
// src/models/encoder.py:154-176, blocks.py:66-72, 113-118 -- torch computes the same sums with cuDNN/cuBLAS wgrad).
//
//   MODE_LIN : dW[out, in] = dY^T X.  A = dY [P, out], B = X [P, in]; tile 128 (two 64-column atoms of A) x 64*NB.
//   MODE_CONV: 3x3 conv.  dWp[tap][ci][co] = sum_p x[p + off_tap, ci] dc[p, co] over the halo-padded pixel list.  One
//              CTA owns (64 input channels, 64 output channels, a range of p) and ALL nine taps: per 64-row block it loads
//              three 72-row boxes of x (dy = -1, 0, +1; the dx taps are 128-byte shifted windows of the same box) and one box
//              of dc, and issues five M=128 MMAs per 16-row K step whose two 64-row halves are two consecutive taps (LBO = the
//              byte distance between the two taps' windows) into five TMEM accumulators.
// Partial sums are added to the fp32 output with red.global.add (the caller zeroes it first).
#include <cstdlib>
#include "gemm_tc5.cuh"

namespace gemm {
namespace tn {

using namespace tc5;

constexpr int KB = 64;                 // contraction rows per pipeline stage
constexpr int ATOM_BYTES = KB * 128;   // one 64-column atom: 64 rows x 128 B
constexpr int XBOX_ROWS = 72;          // conv: 64 + 2 shifted rows, rounded up to whole 8-row swizzle atoms
constexpr int XBOX_BYTES = XBOX_ROWS * 128;
constexpr int kThreadsTn = 192;        // TMA producer, MMA issuer, 4 epilogue warps (warp % 4 = TMEM lane quadrant)

struct TnProblem {
  const __nv_bfloat16 *A, *B;
  int a_cols, ldA, b_cols, ldB;   // valid columns / pitch (elements)
  long long rows;                 // rows of both matrices (contraction length)
  int pitch;                      // conv: W + 2
  float* out;
  int ldo;
  int out_rows, out_cols;         // lin: valid output extent (= a_cols, b_cols)
  int cinpad;                     // conv: rows per tap in the packed output [9][cinpad][ldo]
  int tiles_m, tiles_n, splitk, kb_per_split, num_kb;
};

// MN-major SWIZZLE_128B descriptor: LBO = byte distance between the two 64-wide atoms of the M (or N) extent,
// SBO = 1024 B between 8-row groups along K
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_mn(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;\n" ::"l"(addr), "f"(v) : "memory");
}
// four consecutive fp32 adds in one 16-byte reduction (sm_90+): a quarter of the L2 atomic transactions
__device__ __forceinline__ void red_add_v4_f32(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// MODE 0 = linear, 1 = conv3x3; NB = 64-column atoms of B loaded per tile; NCOLS = the MMA's N (<= 64 NB; the conv runs N = 96 over
// two atoms -- every decoder width but 48 is a multiple of 96 -- or N = 48)
template <int MODE, int NB, int NCOLS = 64 * NB>
struct TnCfg {
  static constexpr uint32_t A_BYTES = MODE == 0 ? 2 * ATOM_BYTES : 3 * XBOX_BYTES;
  static constexpr uint32_t B_BYTES = NB * ATOM_BYTES;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int kStagesRaw = (int)((227u * 1024u - 2048u) / STAGE_BYTES);
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr uint32_t SMEM_BYTES = kStages * STAGE_BYTES + 1024 + 256;
  static constexpr int NACC = MODE == 0 ? 1 : 5;
  static constexpr int BN = NCOLS;
  static_assert(NCOLS % 16 == 0 && NCOLS <= 64 * NB && NCOLS > 64 * (NB - 1), "N must end inside the last loaded atom");
  static constexpr uint32_t TMEM_COLS = (NACC * BN <= 64) ? 64 : (NACC * BN <= 128) ? 128 : (NACC * BN <= 256) ? 256 : 512;
  static_assert(NACC * BN <= 512, "accumulators exceed TMEM");
};

template <int MODE, int NB, int NCOLS = 64 * NB>
__global__ void __launch_bounds__(kThreadsTn, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TnProblem p) {
  msd_pdl_enter();
  using C = TnCfg<MODE, NB, NCOLS>;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::kStages;
  uint64_t* acc_full = bars + 2 * C::kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    for (int i = 0; i < C::kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_ptr)), "r"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  // work item: (tile_m, tile_n, k split)
  int w = blockIdx.x;
  const int ks = w % p.splitk; w /= p.splitk;
  const int tn_ = w % p.tiles_n;
  const int tm_ = w / p.tiles_n;
  const int kb0 = ks * p.kb_per_split;
  const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
  const int nkb = kb1 - kb0;                       // may be <= 0 for the last split: nothing to add
  const int a_col0 = tm_ * (MODE == 0 ? 128 : 64), b_col0 = tn_ * C::BN;

  if (warp == 0) {
    // ===== TMA producer =====
    int stage = 0; uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&empty[stage], phase ^ 1);
      if (elect_one()) {
        unsigned char* sa = smem + stage * C::STAGE_BYTES;
        unsigned char* sb = sa + C::A_BYTES;
        mbar_arrive_expect_tx(&full[stage], C::STAGE_BYTES);
        const int row = kb * KB;
        if constexpr (MODE == 0) {
          tma_load_2d(&tmA, &full[stage], sa, a_col0, row);
          tma_load_2d(&tmA, &full[stage], sa + ATOM_BYTES, a_col0 + 64, row);
        } else {
#pragma unroll
          for (int d = 0; d < 3; ++d) tma_load_2d(&tmA, &full[stage], sa + d * XBOX_BYTES, a_col0, row + (d - 1) * p.pitch - 1);
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) tma_load_2d(&tmB, &full[stage], sb + j * ATOM_BYTES, b_col0 + 64 * j, row);
      }
      __syncwarp();
      if (++stage == C::kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc_mn(128, C::BN);
    int stage = 0; uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&full[stage], phase);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES), sb = sa + C::A_BYTES;
#pragma unroll
        for (int k = 0; k < KB / 16; ++k) {
          const uint64_t db = make_desc_mn(sb + k * 2048, ATOM_BYTES);
          const uint32_t accum = (kb > kb0 || k > 0) ? 1u : 0u;
          if constexpr (MODE == 0) {
            tcgen05_mma_bf16(tmem_base, make_desc_mn(sa + k * 2048, ATOM_BYTES), db, idesc, accum);
          } else {
#pragma unroll
            for (int g = 0; g < 5; ++g) {
              // rows 0..63 of the M extent = tap 2g, rows 64..127 = tap 2g+1 (tap 9 does not exist: the next window of the
              // same box is read instead and its rows are dropped by the epilogue)
              const int t0 = 2 * g, t1 = 2 * g + 1;
              const uint32_t a0 = sa + (t0 / 3) * XBOX_BYTES + (t0 % 3) * 128 + k * 2048;
              const uint32_t a1 = g < 4 ? sa + (t1 / 3) * XBOX_BYTES + (t1 % 3) * 128 + k * 2048 : a0 + 128;
              tcgen05_mma_bf16(tmem_base + g * C::BN, make_desc_mn(a0, a1 - a0), db, idesc, accum);
            }
          }
        }
        tcgen05_commit(&empty[stage]);
        if (kb == kb1 - 1) tcgen05_commit(acc_full);
      }
      __syncwarp();
      if (++stage == C::kStages) { stage = 0; phase ^= 1; }
    }
  } else if (nkb > 0) {
    // ===== epilogue: warp w reads TMEM lanes 32*(w%4).., adds its rows to the fp32 output =====
    const int q = warp & 3;
    mbar_wait(acc_full, 0);
    tcgen05_fence_after();
    const int r = q * 32 + lane;
    const bool vec_ok = (p.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15u) == 0;
#pragma unroll 1
    for (int g = 0; g < C::NACC; ++g) {
      long long orow;
      bool row_ok;
      if constexpr (MODE == 0) {
        orow = a_col0 + r;
        row_ok = orow < p.out_rows;
      } else {
        const int tap = 2 * g + (r >> 6), ci = a_col0 + (r & 63);
        row_ok = tap < 9 && ci < p.a_cols;
        orow = (long long)tap * p.cinpad + ci;
      }
#pragma unroll 1
      for (int c = 0; c < C::BN / 16; ++c) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + g * C::BN + c * 16, v);
        tmem_ld_wait();
        if (row_ok) {
          float* dst = p.out + orow * p.ldo + b_col0 + c * 16;
          if (vec_ok && b_col0 + c * 16 + 16 <= p.out_cols) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              red_add_v4_f32(dst + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (b_col0 + c * 16 + j < p.out_cols) red_add_f32(dst + j, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

template <int MODE, int NB, int NCOLS = 64 * NB>
int launch_tn(TnProblem& p, cudaStream_t s) {
  using C = TnCfg<MODE, NB, NCOLS>;
  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, p.A, (uint64_t)p.a_cols, (uint64_t)p.rows, (uint64_t)p.ldA, MODE == 0 ? KB : XBOX_ROWS);
  if (rc) return rc;
  rc = make_map(&tmB, p.B, (uint64_t)p.b_cols, (uint64_t)p.rows, (uint64_t)p.ldB, KB);
  if (rc) return rc;
  static PerDeviceOnce attr_set;
  if (attr_set.first()) {
    MSD_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<MODE, NB, NCOLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
  }
  p.num_kb = (int)((p.rows + KB - 1) / KB);
  const int tiles = p.tiles_m * p.tiles_n;
  // Split the contraction across CTAs (fp32 atomics merge the partial sums).  Every CTA pays a fixed prologue (TMA latency)
  // and epilogue (its whole tile goes out through atomics), so the split is chosen to fill WHOLE waves of the machine with
  // as few, as long CTAs as possible: the smallest number of waves whose CTAs still contract >= 48 k-blocks each wins,
  // else one wave.
  const int sms = num_sms();
  int splitk = 1;
  for (int waves = 1; waves <= 4; ++waves) {
    int sk = (waves * sms) / tiles;
    if (sk < 1) sk = 1;
    if (sk > p.num_kb) sk = p.num_kb;
    splitk = sk;
    if (tiles * sk >= (waves * sms * 9) / 10 || p.num_kb / sk < 48) break;   // this wave count is (nearly) full, or CTAs get short
  }
  p.kb_per_split = (p.num_kb + splitk - 1) / splitk;
  p.splitk = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
  MSD_CHECK_CUDA(msd_launch_pdl((gemm_tn_kernel<MODE, NB, NCOLS>), dim3(tiles * p.splitk), dim3(kThreadsTn), C::SMEM_BYTES, s, tmA, tmB, p));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

}  // namespace tn
}  // namespace gemm

// dW[out_f, in_f] (fp32, row pitch ldo, ACCUMULATED into) += dY[rows, out_f]^T . X[rows, in_f]   (bf16 operands)
extern "C" int msd_wgrad_linear_bf16(const void* d_dy, int ld_dy, const void* d_x, int ld_x, float* d_dw, int ldo, long long rows,
                                     int out_f, int in_f, void* stream) {
  MSD_REQUIRE(d_dy && d_x && d_dw && rows > 0 && out_f > 0 && in_f > 0, "wgrad_linear: bad arguments");
  MSD_REQUIRE(ld_dy % 8 == 0 && ld_x % 8 == 0 && msd_aligned16(d_dy) && msd_aligned16(d_x), "wgrad_linear: operands must be 16-byte aligned / pitched");
  gemm::tn::TnProblem p{};
  p.A = reinterpret_cast<const __nv_bfloat16*>(d_dy); p.a_cols = out_f; p.ldA = ld_dy;
  p.B = reinterpret_cast<const __nv_bfloat16*>(d_x); p.b_cols = in_f; p.ldB = ld_x;
  p.rows = rows; p.out = d_dw; p.ldo = ldo; p.out_rows = out_f; p.out_cols = in_f;
  p.tiles_m = (out_f + 127) / 128;
  const int atoms_n = (in_f + 63) / 64;
  if (atoms_n >= 4) { p.tiles_n = (atoms_n + 3) / 4; return gemm::tn::launch_tn<0, 4>(p, (cudaStream_t)stream); }
  if (atoms_n == 3) { p.tiles_n = 1; return gemm::tn::launch_tn<0, 3>(p, (cudaStream_t)stream); }
  if (atoms_n == 2) { p.tiles_n = 1; return gemm::tn::launch_tn<0, 2>(p, (cudaStream_t)stream); }
  p.tiles_n = 1;
  return gemm::tn::launch_tn<0, 1>(p, (cudaStream_t)stream);
}

// 3x3 conv weight gradient over halo-padded NHWC bf16 tensors x [B,H+2,W+2,Cin], dc [B,H+2,W+2,Cout] (halo rows of dc
// must be zero): d_dw fp32 [9][cinpad][ldo] (tap-major, ACCUMULATED into) += sum_p x[p + off_tap, ci] dc[p, co]
extern "C" int msd_wgrad_conv3x3_bf16(const void* d_x, const void* d_dc, float* d_dw, int B, int H, int W, int Cin, int Cout,
                                      int cinpad, int ldo, void* stream) {
  MSD_REQUIRE(d_x && d_dc && d_dw && B > 0 && H > 0 && W > 0, "wgrad_conv3x3: bad arguments");
  MSD_REQUIRE(Cin % 8 == 0 && Cout % 8 == 0 && cinpad >= Cin && ldo >= Cout, "wgrad_conv3x3: bad channel counts");
  gemm::tn::TnProblem p{};
  p.A = reinterpret_cast<const __nv_bfloat16*>(d_x); p.a_cols = Cin; p.ldA = Cin;
  p.B = reinterpret_cast<const __nv_bfloat16*>(d_dc); p.b_cols = Cout; p.ldB = Cout;
  p.rows = (long long)B * (H + 2) * (W + 2);
  MSD_REQUIRE(p.rows < (1ll << 31), "wgrad_conv3x3: too many pixels");
  p.pitch = W + 2;
  p.out = d_dw; p.ldo = ldo; p.out_cols = Cout; p.cinpad = cinpad;
  p.tiles_m = (Cin + 63) / 64;
  // output-channel tile: the five M = 128 MMAs of a K step fetch 5 x 4 KB of x windows from shared memory whatever N is, so a wider N
  // amortises them: N = 96 over two 64-column atoms of dc (5 x 96 accumulator columns fill TMEM) when Cout is a multiple of 96 -- the
  // decoder's 96 / 192 / 384 -- measured at micro-batch 64: 205 -> 135 us at 96 channels, 133 -> 111 at 192, 139 -> 115 at 384; N = 48 for
  // the 48-channel stage (203 -> 194 us); 64 otherwise
  static const bool n64 = getenv("MSD_WGRAD_N64") != nullptr;          // A/B switch: the 64-wide tiles
  if (!n64 && Cout % 96 == 0) { p.tiles_n = Cout / 96; return gemm::tn::launch_tn<1, 2, 96>(p, (cudaStream_t)stream); }
  if (!n64 && Cout == 48) { p.tiles_n = 1; return gemm::tn::launch_tn<1, 1, 48>(p, (cudaStream_t)stream); }
  p.tiles_n = (Cout + 63) / 64;
  return gemm::tn::launch_tn<1, 1>(p, (cudaStream_t)stream);
}
