// Fused evaluation metrics (reference: src/training/metrics.py).
//   MetricsTracker.update (:247-256) loops over the batch and calls calculate_psnr/ssim/sam/rmse per image,
//   each ending in .item() (4*B device->host syncs).  Here: two streaming kernels accumulate raw per-image
//   sums in fp64, a third turns them into {psnr, ssim, sam, rmse} per image and the tracker's running sums,
//   all on the device; the host reads 5 doubles once per batch.
//
//   metrics_fused_kernel (W % 4 == 0, the product path): one CTA per (image, 32x64 tile) walks the C bands.  The halo tiles
//                    of pred / target arrive by TMA (3-D tensor map, out-of-image = zero = the window's zero padding) into a
//                    two-deep shared-memory ring, band c+1 in flight while band c is blurred; the SSIM map, (p-t)^2, |p-t| of
//                    the band and the per-pixel spectral dot / norms of SAM come from the same tile, so pred and target
//                    cross HBM exactly once (algorithmic bytes: 2 tensors * 4 B per element) and a CTA does one fp64
//                    reduction per 13 bands.
//   ssim_mse_kernel + sam_kernel: the same sums for any W, one CTA per (plane, tile) / one thread per pixel (two passes).
#include "gemm_tc5.cuh"
#include "ssim_tile.cuh"

namespace {

using namespace ssim;

struct SsimSmem {
  RawTile<2> raw;
  HBlur<4> hb;
  double red[3][kThreads / 32];
};

__global__ void __launch_bounds__(kThreads, 2)
ssim_mse_kernel(const float* __restrict__ pred, const float* __restrict__ target, int C, int H, int W, int tiles_x,
                int tiles_y, SsimConsts consts, double* __restrict__ acc /*[B][4]*/) {
  msd_pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SsimSmem& sm = *reinterpret_cast<SsimSmem*>(smem_raw);
  const int tile = blockIdx.x % (tiles_x * tiles_y);
  const int plane = blockIdx.x / (tiles_x * tiles_y);
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const int y0 = ty * TH, x0 = tx * TW;
  const size_t poff = (size_t)plane * H * W;

  load_raw_plane(sm.raw.v[0], pred + poff, y0, x0, H, W);
  load_raw_plane(sm.raw.v[1], target + poff, y0, x0, H, W);
  __syncthreads();
  // maps: p, t, p*p + t*t, p*t.  The reference blurs p*p and t*t separately (metrics.py:105-117) but its ssim_map only uses
  // sigma1_sq + sigma2_sq (:118-121), and the blur is linear: one map of the sum saves a fifth of the stencil work.
  hpass<2, 4>(sm.raw, sm.hb, [](int m, const float (&px)[2]) -> float {
    return m == 0 ? px[0] : m == 1 ? px[1] : m == 2 ? fmaf(px[0], px[0], px[1] * px[1]) : px[0] * px[1];
  });
  __syncthreads();
  float res[4][8];
  vpass<4>(sm.hb, res);

  const int x = threadIdx.x & 63, h = threadIdx.x >> 6;
  float s_ssim = 0.0f, s_sq = 0.0f, s_abs = 0.0f;
  if (x0 + x < W) {
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      int yy = y0 + 8 * h + o;
      if (yy < H) {
        s_ssim += ssim_value_sum(res[0][o], res[1][o], res[2][o], res[3][o], consts);
        float d = sm.raw.v[0][8 * h + o + R][x + R] - sm.raw.v[1][8 * h + o + R][x + R];
        s_sq = fmaf(d, d, s_sq);
        s_abs += fabsf(d);
      }
    }
  }
  double d_ssim = warp_sum_d((double)s_ssim), d_sq = warp_sum_d((double)s_sq), d_abs = warp_sum_d((double)s_abs);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sm.red[0][warp] = d_sq; sm.red[1][warp] = d_ssim; sm.red[2][warp] = d_abs; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += sm.red[threadIdx.x][w];
    // scratch slots: 0 = sum (p-t)^2, 1 = sum ssim_map, 2 = sum SAM degrees (sam_kernel), 3 = sum |p-t|
    atomicAdd(acc + (size_t)(plane / C) * 4 + (threadIdx.x == 2 ? 3 : threadIdx.x), t);
  }
}

// ---- fused product path -----------------------------------------------------------------------------------------
struct FusedSmem {
  RawTile<2> raw[2];          // ring of two bands: [buf].v[0] = pred halo tile, .v[1] = target halo tile  (TMA destinations)
  HBlur2<2> hb;               // horizontally blurred maps as packed pairs: (p, t) and (p*p + t*t, p*t)
  double red[4][kThreads / 32];
  uint64_t bar[2];
};
constexpr uint32_t kTileBytes = IH * PW * 4;     // one TMA box: 42 rows x 80 floats
// A TMA box must start on a 16-byte boundary of the row (an inner coordinate of x0 - 5 floats raises an illegal-instruction
// fault): the box starts at x0 - 8 and the stencil reads it three columns further in.
constexpr int kXoff = 3;

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
      ::"r"(gemm::tc5::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(gemm::tc5::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__global__ void __launch_bounds__(kThreads, 2)
metrics_fused_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmT, int C, int H, int W,
                     int tiles_x, int tiles_y, SsimConsts consts, float eps, double* __restrict__ acc /*[B][4]*/) {
  msd_pdl_enter();
  using gemm::tc5::mbar_init; using gemm::tc5::mbar_wait; using gemm::tc5::mbar_arrive_expect_tx;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FusedSmem& sm = *reinterpret_cast<FusedSmem*>(smem_raw);
  const int tile = blockIdx.x % (tiles_x * tiles_y);
  const int b = blockIdx.x / (tiles_x * tiles_y);
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const int y0 = ty * TH, x0 = tx * TW;
  if (threadIdx.x == 0) {
    mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  // warp 0 issues the loads, one elected lane, the warp converged around it (as the GEMM producers do)
#define MSD_ISSUE_BAND(c_)                                                                                            \
  do {                                                                                                                 \
    if (threadIdx.x < 32) {                                                                                            \
      if (gemm::tc5::elect_one()) {                                                                                    \
        const int buf_ = (c_) & 1;                                                                                     \
        mbar_arrive_expect_tx(&sm.bar[buf_], 2 * kTileBytes);                                                          \
        tma_load_3d(&tmP, &sm.bar[buf_], &sm.raw[buf_].v[0][0][0], x0 - R - kXoff, y0 - R, b * C + (c_));                     \
        tma_load_3d(&tmT, &sm.bar[buf_], &sm.raw[buf_].v[1][0][0], x0 - R - kXoff, y0 - R, b * C + (c_));                     \
      }                                                                                                                \
      __syncwarp();                                                                                                    \
    }                                                                                                                  \
  } while (0)
  MSD_ISSUE_BAND(0);

  const int x = threadIdx.x & 63, h = threadIdx.x >> 6;
  const bool col_ok = x0 + x < W;
  float dot[8], pp[8], tt[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) dot[o] = pp[o] = tt[o] = 0.0f;
  double d_ssim = 0.0, d_sq = 0.0, d_abs = 0.0;

  for (int c = 0; c < C; ++c) {
    const int buf = c & 1;
    // ring slot (c+1)&1 was last read for band c-1; every thread passed the barrier that closes that iteration
    if (c + 1 < C) MSD_ISSUE_BAND(c + 1);
    mbar_wait(&sm.bar[buf], (uint32_t)(c >> 1) & 1u);
    const RawTile<2>& raw = sm.raw[buf];
    // maps: p, t, p*p + t*t, p*t.  The reference blurs p*p and t*t separately (metrics.py:105-117) but its ssim_map only uses
    // sigma1_sq + sigma2_sq (:118-121), and the blur is linear: one map of the sum saves a fifth of the stencil work.
    // Two maps per 64-bit register, one FFMA2 per tap (ssim_tile.cuh): same per-lane arithmetic, half the FMA issue slots.
    // Pair 0 is (p + t, p - t), not (p, t): both lanes are then freshly computed values the register allocator can place side
    // by side (a pair of two LOADED values costs two MOVs per use); the blur is linear, so blur(p +- t) = mu1 +- mu2.
    hpass2<2, 2, kXoff>(raw, sm.hb, [](int mp, const float (&px)[2]) -> f2 {
      return mp == 0 ? pack2(px[0] + px[1], px[0] - px[1]) : pack2(fmaf(px[0], px[0], px[1] * px[1]), px[0] * px[1]);
    });
    __syncthreads();
    f2 res2[2][8];
    vpass2<2>(sm.hb, res2);
    float s_ssim = 0.0f, s_sq = 0.0f, s_abs = 0.0f;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const float a = raw.v[0][8 * h + o + R][x + R + kXoff], bb = raw.v[1][8 * h + o + R][x + R + kXoff];
      dot[o] = fmaf(a, bb, dot[o]);            // SAM: spectral dot product and norms, band by band (metrics.py:155-160)
      pp[o] = fmaf(a, a, pp[o]);
      tt[o] = fmaf(bb, bb, tt[o]);
      // branch-free: pixels outside the image (ragged last tile) are computed and masked out
      const float keep = (col_ok && y0 + 8 * h + o < H) ? 1.0f : 0.0f;
      float ms, md, esum, e12;
      unpack2(res2[0][o], ms, md);             // mu1 + mu2, mu1 - mu2
      unpack2(res2[1][o], esum, e12);
      s_ssim = fmaf(keep, ssim_value_sumdiff(ms, md, esum, e12, consts), s_ssim);
      const float d = (a - bb) * keep;
      s_sq = fmaf(d, d, s_sq);
      s_abs += fabsf(d);
    }
    d_ssim += (double)s_ssim; d_sq += (double)s_sq; d_abs += (double)s_abs;
    __syncthreads();                           // hb and ring slot `buf` are free again
  }
#undef MSD_ISSUE_BAND
  // SAM angle of this thread's pixels (metrics.py:161-170)
  float sum_deg = 0.0f;
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    if (col_ok && y0 + 8 * h + o < H) {
      const float pn = sqrtf(pp[o]) + eps, tn = sqrtf(tt[o]) + eps;
      float cosang = dot[o] / (pn * tn);
      cosang = fminf(fmaxf(cosang, -1.0f), 1.0f);
      sum_deg += acosf(cosang) * 57.29577951308232f;  // torch.rad2deg
    }
  }
  const double r_sq = warp_sum_d(d_sq), r_ssim = warp_sum_d(d_ssim), r_sam = warp_sum_d((double)sum_deg), r_abs = warp_sum_d(d_abs);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sm.red[0][warp] = r_sq; sm.red[1][warp] = r_ssim; sm.red[2][warp] = r_sam; sm.red[3][warp] = r_abs; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += sm.red[threadIdx.x][w];
    atomicAdd(acc + (size_t)b * 4 + threadIdx.x, t);   // slots: sum (p-t)^2, sum ssim_map, sum SAM degrees, sum |p-t|
  }
}

// SAM: angle between the C-band spectra of each pixel (metrics.py:150-170)
__global__ void __launch_bounds__(256)
sam_kernel(const float* __restrict__ pred, const float* __restrict__ target, int C, int HW, int chunks_per_image,
           float eps, double* __restrict__ acc) {
  msd_pdl_enter();
  const int b = blockIdx.x / chunks_per_image;
  const int chunk = blockIdx.x - b * chunks_per_image;
  const float* p = pred + (size_t)b * C * HW;
  const float* t = target + (size_t)b * C * HW;
  float sum_deg = 0.0f;
  for (int pix = chunk * 1024 + threadIdx.x; pix < min(HW, (chunk + 1) * 1024); pix += 256) {
    float dot = 0.0f, pp = 0.0f, tt = 0.0f;
    for (int c = 0; c < C; ++c) {
      float a = __ldg(p + (size_t)c * HW + pix), bb = __ldg(t + (size_t)c * HW + pix);
      dot = fmaf(a, bb, dot);
      pp = fmaf(a, a, pp);
      tt = fmaf(bb, bb, tt);
    }
    float pn = sqrtf(pp) + eps, tn = sqrtf(tt) + eps;
    float cosang = dot / (pn * tn);
    cosang = fminf(fmaxf(cosang, -1.0f), 1.0f);
    sum_deg += acosf(cosang) * 57.29577951308232f;  // torch.rad2deg
  }
  __shared__ double red[8];
  double d = warp_sum_d((double)sum_deg);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tsum = 0.0;
    for (int w = 0; w < 8; ++w) tsum += red[w];
    atomicAdd(acc + (size_t)b * 4 + 2, tsum);
  }
}

// per-image metrics + the tracker's running sums (metrics.py:39-46, 121, 170, 189-191, 252-258)
__global__ void metrics_finalize_kernel(const double* __restrict__ acc, int B, int C, int HW, double data_range,
                                        double* __restrict__ per_image, double* __restrict__ sums) {
  msd_pdl_enter();
  __shared__ double sh[4][32];
  double loc[4] = {0, 0, 0, 0};
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const double n = (double)C * HW;
    // F.mse_loss(...).item() is an fp32 value; the dB / sqrt are done in fp64 on it
    const double mse = (double)(float)(acc[b * 4 + 0] / n);
    const double psnr = (mse == 0.0) ? (double)INFINITY : 10.0 * log10(data_range * data_range / mse);
    const double ssimv = (double)(float)(acc[b * 4 + 1] / n);
    const double sam = (double)(float)(acc[b * 4 + 2] / (double)HW);
    const double rmse = sqrt(mse);
    per_image[b * 4 + 0] = psnr; per_image[b * 4 + 1] = ssimv; per_image[b * 4 + 2] = sam; per_image[b * 4 + 3] = rmse;
    loc[0] += psnr; loc[1] += ssimv; loc[2] += sam; loc[3] += rmse;
  }
  for (int k = 0; k < 4; ++k) {
    double v = warp_sum_d(loc[k]);
    if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[threadIdx.x][w];
    sums[threadIdx.x] = t;
  }
  if (threadIdx.x == 0) sums[4] = (double)B;
}

}  // namespace

extern "C" int msd_metrics_f32(const float* d_pred, const float* d_target, int B, int C, int H, int W, float data_range,
                               double* d_per_image, double* d_sums, double* d_scratch, void* stream) {
  MSD_REQUIRE(d_pred && d_target && d_per_image && d_sums && d_scratch, "metrics: null pointer");
  MSD_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "metrics: bad shape [%d,%d,%d,%d]", B, C, H, W);
  MSD_REQUIRE((int64_t)B * C * H * W < (1ll << 31), "metrics: tensor too large");
  cudaStream_t s = (cudaStream_t)stream;
  MSD_CHECK_CUDA(cudaMemsetAsync(d_scratch, 0, sizeof(double) * 4 * (size_t)B, s));
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  SsimConsts consts;
  consts.C1 = (float)((0.01 * (double)data_range) * (0.01 * (double)data_range));
  consts.C2 = (float)((0.03 * (double)data_range) * (0.03 * (double)data_range));
  const int HW = H * W;
  CUtensorMap tmP, tmT;
  // TMA needs 16-byte aligned rows; anything else (and a driver that refuses the map) takes the two-pass kernels below
  const bool fused = W % 4 == 0 && msd_aligned16(d_pred) && msd_aligned16(d_target) &&
                     gemm::tc5::make_map_planes_f32(&tmP, d_pred, (uint64_t)W, (uint64_t)H, (uint64_t)B * C, PW, IH) == MSD_OK &&
                     gemm::tc5::make_map_planes_f32(&tmT, d_target, (uint64_t)W, (uint64_t)H, (uint64_t)B * C, PW, IH) == MSD_OK;
  if (fused) {
    // product path: TMA-staged, one pass over pred and target
    static PerDeviceOnce fused_attr_set;
    if (fused_attr_set.first()) {
      MSD_CHECK_CUDA(cudaFuncSetAttribute(metrics_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem)));
    }
    MSD_CHECK_CUDA(msd_launch_pdl((metrics_fused_kernel), dim3((unsigned)B * tiles_x * tiles_y), dim3(kThreads), sizeof(FusedSmem), s, tmP, tmT, C, H, W, tiles_x, tiles_y, consts,
                                                                                             1e-8f, d_scratch));
    MSD_LAUNCH_CHECK();
  } else {
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
      MSD_CHECK_CUDA(cudaFuncSetAttribute(ssim_mse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SsimSmem)));
    }
    const unsigned blocks = (unsigned)B * C * tiles_x * tiles_y;
    MSD_CHECK_CUDA(msd_launch_pdl((ssim_mse_kernel), dim3(blocks), dim3(kThreads), sizeof(SsimSmem), s, d_pred, d_target, C, H, W, tiles_x, tiles_y, consts, d_scratch));
    MSD_LAUNCH_CHECK();
    const int chunks = (HW + 1023) / 1024;
    MSD_CHECK_CUDA(msd_launch_pdl((sam_kernel), dim3((unsigned)B * chunks), dim3(256), 0, s, d_pred, d_target, C, HW, chunks, 1e-8f, d_scratch));
    MSD_LAUNCH_CHECK();
  }
  MSD_CHECK_CUDA(msd_launch_pdl((metrics_finalize_kernel), dim3(1), dim3(256), 0, s, d_scratch, B, C, HW, (double)data_range, d_per_image, d_sums));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}
