// Host side of the tcgen05 GEMM: tensor-map encoding through the driver entry point (no libcuda link dependency),
// tile-size selection and dispatch to the per-BN instantiations (gemm_tc5_bn*.cu, compiled in parallel).
#include "gemm_tc5.cuh"

namespace gemm {
namespace tc5 {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
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

// 3-D fp32 tensor map over [planes, H, W] (W contiguous), box {box_w, box_h, 1}, no swizzle: halo tiles of an image plane;
// out-of-image coordinates (negative or >= extent) read as zero, which is exactly the zero padding of the SSIM window
int make_map_planes_f32(CUtensorMap* tm, const void* base, uint64_t W, uint64_t H, uint64_t planes, uint32_t box_w, uint32_t box_h) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return MSD_ERR_CUDA;
  cuuint64_t dims[3] = {W, H, planes};
  cuuint64_t strides[2] = {W * 4, W * H * 4};
  cuuint32_t box[3] = {box_w, box_h, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    msd_set_error("cuTensorMapEncodeTiled (fp32 planes) failed (%d): W=%llu H=%llu planes=%llu base=%p", (int)r,
                  (unsigned long long)W, (unsigned long long)H, (unsigned long long)planes, base);
    return MSD_ERR_CUDA;
  }
  return MSD_OK;
}

int num_sms() {
  static int cache[64] = {};          // per device: one process may drive several GPUs
  int dev = 0;
  cudaGetDevice(&dev);
  int& n = cache[dev & 63];
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN> int launch_pair(const Problem&, cudaStream_t);   // gemm_tc5_pair.cu (cta_group::2)
extern template int launch_pair<256>(const Problem&, cudaStream_t);
extern template int launch_pair<192>(const Problem&, cudaStream_t);
extern template int launch_pair<96>(const Problem&, cudaStream_t);
extern template int launch_pair<48>(const Problem&, cudaStream_t);
extern template int launch_bn<256>(const Problem&, cudaStream_t);
extern template int launch_bn<192>(const Problem&, cudaStream_t);
extern template int launch_bn<128>(const Problem&, cudaStream_t);
extern template int launch_bn<96>(const Problem&, cudaStream_t);
extern template int launch_bn<64>(const Problem&, cudaStream_t);
extern template int launch_bn<48>(const Problem&, cudaStream_t);
extern template int launch_bn<32>(const Problem&, cudaStream_t);
extern template int launch_bn<16>(const Problem&, cudaStream_t);

}  // namespace tc5

namespace roll {   // conv_roll.cu: rolling-row formulation of the 48-channel 3x3 conv
bool eligible(const Problem& p);
int launch(const Problem& p, cudaStream_t s);
}  // namespace roll

int pick_bn_tc5(int N, int M, bool wave_aware, int min_bn = 16) {
  // Tile width: a divisor of N (2*BN fp32 accumulator columns must fit the 512 TMEM columns).  Wider tiles amortise the A
  // operand better, but the persistent grid runs whole waves of 148 CTAs: for the small GEMMs of the training step
  // (M = 9,216 rows at micro-batch 64) a narrower tile that fills its last wave beats a wider one that leaves it half empty.
  const int cand[] = {256, 192, 128, 96, 64, 48, 32, 16};
  const double speed[] = {1.00, 0.97, 0.90, 0.84, 0.74, 0.62, 0.45, 0.25};   // relative per-tile efficiency (tools/gemm_bench.py)
  const int sms = tc5::num_sms();
  const long long tiles_m = (M + BLOCK_M - 1) / BLOCK_M;
  int best = 0;
  double best_score = -1.0;
  for (int i = 0; i < 8; ++i) {
    const int c = cand[i];
    if (N % c != 0 || c < min_bn) continue;
    if (!wave_aware) return c;                                 // largest divisor
    const long long tiles = tiles_m * (N / c);
    const long long waves = (tiles + sms - 1) / sms;
    const double score = speed[i] * (double)tiles / (double)(waves * sms);
    if (score > best_score + 1e-9) { best_score = score; best = c; }
  }
  return best;
}

bool tc5_can_fuse_head(const Problem& p, int head_cout) { return tc5::can_fuse_head(p, head_cout); }

// A GEMM in front of a folded LayerNorm leaves 2 * N / BN partial row sums per row (one per column tile and column half); its
// tile is at least 96 wide (32-column slabs, two halves), so at most 16 partials for N = 768.  Mirrors launch_tc5's tile choice.
constexpr int kStatsMinBN = 96;
int tc5_stats_parts(int N, int M) {
  const int bn = pick_bn_tc5(N, M, true, kStatsMinBN);
  return bn > 0 ? 2 * (N / bn) : 0;
}

// Column sums in the conv epilogue (Problem::bn_sums) exist in the two kernels that run the large decoder stages: the 48-channel
// resident-weight conv (single CTAs) and the 96-channel dx-merged pair kernel; both keep their per-warp slots where the skip tap's
// weights would sit.  Mirrors the routing of launch_tc5 / launch_bn below.
bool tc5_supports_bn_sums(const Problem& p) {
  static const bool off = getenv("MSD_NO_MERGE3") != nullptr || getenv("MSD_RES4") != nullptr || getenv("MSD_NO_PAIR") != nullptr ||
                          getenv("MSD_PAIR48") != nullptr || getenv("MSD_NO_EPI_SUMS") != nullptr;
  if (off || tc5::classify(p) != tc5::E_CONV_RAW || p.skip_tap >= 0 || p.head_out != nullptr || !tc5::is_conv3x3(p) || p.dbg_bo != 0) return false;
  if (p.N == 48) return p.kb_per_tap == 1 && p.num_kb == 9;
  if (p.N == 96) return p.M >= 2048 && p.kb_per_tap == 2 && p.num_kb == 18;
  return false;
}

int launch_tc5(const Problem& p, cudaStream_t s) {
  MSD_REQUIRE(p.N % 16 == 0, "gemm_tc5: N=%d must be a multiple of 16", p.N);
  MSD_REQUIRE(p.bn_sums == nullptr || tc5_supports_bn_sums(p), "gemm_tc5: this problem has no column-sum epilogue (ask tc5_supports_bn_sums first)");
  MSD_REQUIRE(p.ldA % 8 == 0 && msd_aligned16(p.A) && msd_aligned16(p.Wt), "gemm_tc5: operands must be 16-byte aligned / pitched");
  if (roll::eligible(p)) return roll::launch(p, s);
  // multi-tap (3x3 conv) problems keep the largest tile: their summation order (tap-major vs dx-merged) depends on the tile
  // width through the kernel mode, and results must not depend on the batch size (sharding invariance is tested bit-exactly)
  const bool ln_consumer = p.ln_stats != nullptr, ln_producer = p.stats_out != nullptr;
  if (ln_consumer)
    MSD_REQUIRE(p.mode == EPI_ROWMAJOR && !p.out_f32 && p.resid == nullptr && p.ln_s != nullptr && p.bias != nullptr && p.ln_parts > 0 &&
                p.ln_parts <= 64 && msd_aligned16(p.ln_s) && !ln_producer, "gemm_tc5: a folded LayerNorm needs a bf16 row-major output, ln_s, bias and its partial sums");
  if (ln_producer)
    MSD_REQUIRE(p.mode == EPI_ROWMAJOR && p.out_f32 && !p.act && p.resid != nullptr && p.out_bf16 != nullptr && msd_aligned16(p.out_bf16) &&
                p.stats_parts == tc5_stats_parts(p.N, p.M) && p.stats_parts > 0, "gemm_tc5: LayerNorm partial sums need the fp32 residual epilogue, a bf16 copy and stats_parts = tc5_stats_parts(N, M)");
  if (p.gelu_pre != nullptr)
    MSD_REQUIRE(tc5::classify(p) == tc5::E_RM_BF16_DGELU && msd_aligned16(p.gelu_pre) && p.ldo % 8 == 0, "gemm_tc5: gelu_pre needs a plain bf16 row-major output");
  if (p.out_bf16 != nullptr && !ln_producer)
    MSD_REQUIRE(tc5::classify(p) == tc5::E_RM_BF16_GELU_PRE && msd_aligned16(p.out_bf16) && p.ldo % 8 == 0, "gemm_tc5: a pre-activation copy needs a bf16 row-major GELU output");
  // (narrower tiles for the short-K residual GEMM, proj, were measured and are not faster: 0.066 ms at BN = 256, 0.064 at 192,
  //  0.073 at 128, 0.078 at 96, 0.092 at 64 for M = 36,864)
  const int bn = pick_bn_tc5(p.N, p.M, p.num_kb == p.kb_per_tap, ln_producer ? kStatsMinBN : 16);
  // wide tiles of large problems run on CTA pairs (cta_group::2): half the W traffic per CTA, deeper ring
  static const bool no_pair = getenv("MSD_NO_PAIR") != nullptr;
  const bool pair_ok = !no_pair && p.M >= 2048 && p.dbg_shift == 0 && p.dbg_bo == 0 &&
                       !(p.mode == EPI_CONV && p.skip != nullptr && p.skip_tap < 0 && !p.act);   // raw + epilogue skip: generic single-CTA path
  if (pair_ok && (bn == 256 || bn == 192))
    return bn == 256 ? tc5::launch_pair<256>(p, s) : tc5::launch_pair<192>(p, s);
  // 3x3 convs with 96 / 48 channels: dx-merged windows + the weight matrix resident in shared memory, split over the pair
  static const bool no_merge = getenv("MSD_NO_MERGE3") != nullptr;
  // (the 48-channel stage measures slightly FASTER on single CTAs -- 0.555 vs 0.586 ms at B = 256 -- and stays there;
  //  MSD_PAIR48 switches it for A/B runs)
  static const bool pair48 = getenv("MSD_PAIR48") != nullptr;
  if (pair_ok && !no_merge && p.head_out == nullptr && (bn == 96 || (bn == 48 && pair48)) && p.N == bn && tc5::is_conv3x3(p) && p.num_kb <= 20)
    return bn == 96 ? tc5::launch_pair<96>(p, s) : tc5::launch_pair<48>(p, s);
  switch (bn) {
    case 256: return tc5::launch_bn<256>(p, s);
    case 192: return tc5::launch_bn<192>(p, s);
    case 128: return tc5::launch_bn<128>(p, s);
    case 96: return tc5::launch_bn<96>(p, s);
    case 64: return tc5::launch_bn<64>(p, s);
    case 48: return tc5::launch_bn<48>(p, s);
    case 32: return tc5::launch_bn<32>(p, s);
    case 16: return tc5::launch_bn<16>(p, s);
  }
  msd_set_error("gemm_tc5: no tile for N=%d", p.N);
  return MSD_ERR_INVALID;
}

}  // namespace gemm
