// tcgen05 / TMEM / TMA GEMM + implicit 3x3 convolution for sm_100a (the product path of every dense contraction
// in SatMAERestoration.forward; shapes in SURVEY Appendix A).
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0   : TMA producer  -- cp.async.bulk.tensor.2d of the A box [128 rows x 64 k] (row = m0 + tap shift) and
//              the W box [BN rows x 64 k] into a kStages-deep 128B-swizzled shared-memory ring (mbarrier full/empty)
//   warp 1   : MMA issuer    -- one thread issues 4 x tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) per
//              stage into one of two TMEM accumulator buffers, tcgen05.commit releases the stage / publishes the tile
//   warp 2   : TMEM allocator
//   warps 4-7: epilogue      -- tcgen05.ld 32x32b.x16 (thread = one output row, 16 columns), fused bias / GELU /
//              residual / skip / pixel-shuffle / halo-zero store (gemm.cuh), overlapping the next tile's MMAs
// Accumulators are fp32 in TMEM (2 x BN columns), operands bf16.
#include "gemm.cuh"
#include <cuda.h>

namespace gemm {
namespace {

constexpr int kThreads = 256;
constexpr uint32_t A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr uint64_t kSpinLimit = 4000000000ull;             // ~2 s of SM clocks: trap instead of hanging the GPU

template <int BN>
struct Cfg {
  static constexpr uint32_t B_STAGE_BYTES = BN * BLOCK_K * 2;
  static constexpr uint32_t STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int kStagesRaw = (int)((200u * 1024u) / STAGE_BYTES);
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr uint32_t SMEM_BYTES = kStages * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128 must be a multiple of 16 in [16,256]");
  static_assert(B_STAGE_BYTES % 1024 == 0, "stage bases must stay 1024-byte aligned for SWIZZLE_128B");
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
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

// ------------------------------------------------------------------------------------------------ kernel
template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc5_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Problem p,
                int tiles_m, int tiles_n) {
  using C = Cfg<BN>;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* smem_a = smem;
  unsigned char* smem_b = smem + C::kStages * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::kStages;
  uint64_t* tmem_full = bars + 2 * C::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_ptr)), "r"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int num_tiles = tiles_m * tiles_n;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * BLOCK_M, n0 = (tile % tiles_n) * BN;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        const int tap = kb / p.kb_per_tap, kc = kb - tap * p.kb_per_tap;
        mbar_arrive_expect_tx(&full[stage], C::STAGE_BYTES);
        tma_load_2d(&tmA, &full[stage], smem_a + stage * A_STAGE_BYTES, kc * BLOCK_K, m0 + p.tap_off[tap]);
        tma_load_2d(&tmB, &full[stage], smem_b + stage * C::B_STAGE_BYTES, kb * BLOCK_K, n0);
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc(BLOCK_M, BN);
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t as = it & 1, aphase = (it >> 1) & 1;
      mbar_wait(&tmem_empty[as], aphase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tcgen05_fence_after();
        const uint64_t da = make_smem_desc(smem_u32(smem_a + stage * A_STAGE_BYTES));
        const uint64_t db = make_smem_desc(smem_u32(smem_b + stage * C::B_STAGE_BYTES));
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k) {
          // advance 16 bf16 = 32 bytes inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
          tcgen05_mma_bf16(tmem_d, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        tcgen05_commit(&empty[stage]);               // frees the smem stage once these MMAs retire
        if (kb == p.num_kb - 1) tcgen05_commit(&tmem_full[as]);  // accumulator complete -> epilogue
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: warp (4+q) owns TMEM lanes 32q..32q+31 = output rows 32q..32q+31 of the tile =====
    const int q = warp - 4;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t as = it & 1, aphase = (it >> 1) & 1;
      const int m0 = (tile / tiles_n) * BLOCK_M, n0 = (tile % tiles_n) * BN;
      mbar_wait(&tmem_full[as], aphase);
      tcgen05_fence_after();
      const int m = m0 + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 16; ++c) {
        float v[16];
        tmem_ld16(taddr + c * 16, v);
        epilogue_store<16>(p, m, n0 + c * 16, v);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || ptr == nullptr) {
    msd_set_error("cuTensorMapEncodeTiled not available from the driver");
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

// 2-D bf16 tensor map: inner extent `inner` elements (valid K), `rows` rows of pitch `ld` elements, box 64 x box_rows
int make_map(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t rows, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return MSD_ERR_CUDA;
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    msd_set_error("cuTensorMapEncodeTiled failed (%d): inner=%llu rows=%llu ld=%llu box_rows=%u base=%p", (int)r,
                  (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)ld, box_rows, base);
    return MSD_ERR_CUDA;
  }
  return MSD_OK;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN>
int launch_bn(const Problem& p, cudaStream_t s) {
  using C = Cfg<BN>;
  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, p.A, (uint64_t)p.kvalid, (uint64_t)p.a_rows, (uint64_t)p.ldA, BLOCK_M);
  if (rc) return rc;
  rc = make_map(&tmB, p.Wt, (uint64_t)p.num_kb * BLOCK_K, (uint64_t)p.N, (uint64_t)p.num_kb * BLOCK_K, BN);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    MSD_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc5_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
    attr_set = true;
  }
  const int tiles_m = (p.M + BLOCK_M - 1) / BLOCK_M, tiles_n = p.N / BN;
  const int tiles = tiles_m * tiles_n;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_tc5_kernel<BN><<<grid, kThreads, C::SMEM_BYTES, s>>>(tmA, tmB, p, tiles_m, tiles_n);
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

}  // namespace

int pick_bn_tc5(int N) {
  // largest tile that divides N; 2*BN fp32 accumulator columns must fit the 512 TMEM columns
  const int cand[] = {256, 192, 128, 96, 64, 48, 32, 16};
  for (int c : cand)
    if (N % c == 0) return c;
  return 0;
}

int launch_tc5(const Problem& p, cudaStream_t s) {
  MSD_REQUIRE(p.N % 16 == 0, "gemm_tc5: N=%d must be a multiple of 16", p.N);
  MSD_REQUIRE(p.ldA % 8 == 0 && msd_aligned16(p.A) && msd_aligned16(p.Wt), "gemm_tc5: operands must be 16-byte aligned / pitched");
  switch (pick_bn_tc5(p.N)) {
    case 256: return launch_bn<256>(p, s);
    case 192: return launch_bn<192>(p, s);
    case 128: return launch_bn<128>(p, s);
    case 96: return launch_bn<96>(p, s);
    case 64: return launch_bn<64>(p, s);
    case 48: return launch_bn<48>(p, s);
    case 32: return launch_bn<32>(p, s);
    case 16: return launch_bn<16>(p, s);
  }
  msd_set_error("gemm_tc5: no tile for N=%d", p.N);
  return MSD_ERR_INVALID;
}

}  // namespace gemm
