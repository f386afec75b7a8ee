// Training loss forward + analytic backward (reference: src/training/losses.py).
//   CombinedLoss.forward (:55-77): total = w_mse * mean((p-t)^2) + w_ssim * (1 - mean(ssim_map))
//   SSIMLoss._ssim (:123-163): five zero-padded Gaussian blurs (p, t, p*p, t*t, p*t), C1/C2 from data_range.
// The reference gets d(total)/d(pred) from autograd through 5 depthwise convs; here it is closed-form:
//   S = A1*A2/(B1*B2),  A1 = 2 mu1 mu2 + C1, A2 = 2 (e12 - mu1 mu2) + C2, B1 = mu1^2 + mu2^2 + C1,
//   B2 = (e11 - mu1^2) + (e22 - mu2^2) + C2
//   dS/de12 = 2 A1/(B1 B2);  dS/de11 = -S/B2;  dS/dmu1 = 2 mu2 (A2 - A1)/(B1 B2) - 2 mu1 S (1/B1 - 1/B2)
//   d mean(S)/dp = (1/N) [ blur(dS/dmu1) + 2 p blur(dS/de11) + t blur(dS/de12) ]   (the window is symmetric and
//   the padding is zero, so the blur is self-adjoint).
// Pass 1 (fwd): blur 5 maps, reduce sum((p-t)^2) and sum(S), write the 3 derivative maps.
// Pass 2 (bwd): blur the 3 derivative maps and combine with p, t into grad_pred.
#include "gemm_tc5.cuh"
#include "ssim_tile.cuh"

namespace {

using namespace ssim;

// Halo tiles arrive by TMA when the rows are 16-byte aligned (W % 4 == 0): one 3-D box {80, 42, 1} per map over [planes, H, W],
// out-of-image coordinates read as zero = the window's zero padding.  A box must start on a 16-byte boundary of the row, so it
// starts at x0 - 8 and the stencil reads it kXoff = 3 columns further in (see metrics.cu).  Otherwise: per-thread guarded loads.
constexpr uint32_t kTileBytes = IH * PW * 4;
constexpr int kXoff = 3;

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
      ::"r"(gemm::tc5::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(gemm::tc5::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

struct FwdSmem {
  RawTile<2> raw;
  HBlur2<2> hb;               // packed pairs (p + t, p - t) and (p*p + t*t, p*t): one FFMA2 per tap (ssim_tile.cuh)
  double red[2][kThreads / 32];
  uint64_t bar;
};
struct BwdSmem {
  RawTile<3> raw;
  HBlur2<1> hb2;              // (dmu1 + de11, dmu1 - de11), packed
  HBlur<1> hb;                // de12
  float pt[2][TH][TW];        // the tile's own pred / target pixels (TMA path): no dependent global loads at the tail of the CTA
  uint64_t bar;
};

template <bool kWriteDeriv, bool kTma>
__global__ void __launch_bounds__(kThreads, 2)
loss_fwd_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmT, const float* __restrict__ pred,
                const float* __restrict__ target, int H, int W, int tiles_x, int tiles_y,
                SsimConsts consts, size_t numel, float* __restrict__ deriv /*[3][numel]*/, double* __restrict__ acc /*[2]*/) {
  msd_pdl_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw);
  constexpr int XO = kTma ? kXoff : 0;
  const int tile = blockIdx.x % (tiles_x * tiles_y);
  const int plane = blockIdx.x / (tiles_x * tiles_y);
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const int y0 = ty * TH, x0 = tx * TW;
  const size_t poff = (size_t)plane * H * W;
  if constexpr (kTma) {
    if (threadIdx.x == 0) {
      gemm::tc5::mbar_init(&sm.bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
      gemm::tc5::mbar_arrive_expect_tx(&sm.bar, 2 * kTileBytes);
      tma_load_3d(&tmP, &sm.bar, &sm.raw.v[0][0][0], x0 - R - kXoff, y0 - R, plane);
      tma_load_3d(&tmT, &sm.bar, &sm.raw.v[1][0][0], x0 - R - kXoff, y0 - R, plane);
    }
    __syncthreads();                       // the barrier's initialisation is visible to every waiter
    gemm::tc5::mbar_wait(&sm.bar, 0);
  } else {
    load_raw_plane(sm.raw.v[0], pred + poff, y0, x0, H, W);
    load_raw_plane(sm.raw.v[1], target + poff, y0, x0, H, W);
    __syncthreads();
  }
  // Four maps, not the reference's five: S and its derivatives only use e11 + e22 (B2 below; dS/de11 = dS/de22), and the blur is
  // linear.  Two maps per 64-bit register, one FFMA2 per tap; pair 0 is (p + t, p - t) so that both lanes are computed values
  // (blur(p +- t) = mu1 +- mu2), see metrics.cu.
  hpass2<2, 2, XO>(sm.raw, sm.hb, [](int mp, const float (&px)[2]) -> f2 {
    return mp == 0 ? pack2(px[0] + px[1], px[0] - px[1]) : pack2(fmaf(px[0], px[0], px[1] * px[1]), px[0] * px[1]);
  });
  __syncthreads();
  f2 res2[2][8];
  vpass2<2>(sm.hb, res2);
  const int x = threadIdx.x & 63, h = threadIdx.x >> 6;
  float s_ssim = 0.0f, s_sq = 0.0f;
  if (x0 + x < W) {
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const int yy = y0 + 8 * h + o;
      if (yy >= H) continue;
      float ms, md, esum, e12;
      unpack2(res2[0][o], ms, md);
      unpack2(res2[1][o], esum, e12);
      const float mu1 = 0.5f * (ms + md), mu2 = 0.5f * (ms - md);
      const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu1_mu2 = mu1 * mu2;
      const float A1 = 2.0f * mu1_mu2 + consts.C1;
      const float A2 = 2.0f * (e12 - mu1_mu2) + consts.C2;
      const float B1 = mu1_sq + mu2_sq + consts.C1;
      const float B2 = (esum - mu1_sq - mu2_sq) + consts.C2;
      const float S = (A1 * A2) / (B1 * B2);
      s_ssim += S;
      const float d = sm.raw.v[0][8 * h + o + R][x + R + XO] - sm.raw.v[1][8 * h + o + R][x + R + XO];
      s_sq = fmaf(d, d, s_sq);
      if (kWriteDeriv) {
        const float inv = 1.0f / (B1 * B2);
        const float dmu1 = 2.0f * mu2 * (A2 - A1) * inv - 2.0f * mu1 * S * (1.0f / B1 - 1.0f / B2);
        const float de11 = -S / B2;
        const float de12 = 2.0f * A1 * inv;
        const size_t e = poff + (size_t)yy * W + (x0 + x);
        deriv[e] = dmu1;
        deriv[numel + e] = de11;
        deriv[2 * numel + e] = de12;
      }
    }
  }
  double d_ssim = warp_sum_d((double)s_ssim), d_sq = warp_sum_d((double)s_sq);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sm.red[0][warp] = d_sq; sm.red[1][warp] = d_ssim; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += sm.red[threadIdx.x][w];
    atomicAdd(acc + threadIdx.x, t);
  }
}

template <bool kTma>
__global__ void __launch_bounds__(kThreads, 2)
loss_bwd_kernel(const __grid_constant__ CUtensorMap tmD /* deriv as [3 * planes, H, W] */, const __grid_constant__ CUtensorMap tmPc,
                const __grid_constant__ CUtensorMap tmTc /* pred / target, box {64, 32, 1} */, int planes, const float* __restrict__ pred,
                const float* __restrict__ target, const float* __restrict__ deriv, int H,
                int W, int tiles_x, int tiles_y, size_t numel, float k_ssim /* -w_ssim*scale/N */,
                float k_mse /* 2*w_mse*scale/N */, float* __restrict__ grad) {
  msd_pdl_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
  constexpr int XO = kTma ? kXoff : 0;
  const int tile = blockIdx.x % (tiles_x * tiles_y);
  const int plane = blockIdx.x / (tiles_x * tiles_y);
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const int y0 = ty * TH, x0 = tx * TW;
  const size_t poff = (size_t)plane * H * W;
  if constexpr (kTma) {
    if (threadIdx.x == 0) {
      gemm::tc5::mbar_init(&sm.bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
      gemm::tc5::mbar_arrive_expect_tx(&sm.bar, 3 * kTileBytes + 2 * TH * TW * 4);
#pragma unroll
      for (int i = 0; i < 3; ++i) tma_load_3d(&tmD, &sm.bar, &sm.raw.v[i][0][0], x0 - R - kXoff, y0 - R, i * planes + plane);
      tma_load_3d(&tmPc, &sm.bar, &sm.pt[0][0][0], x0, y0, plane);
      tma_load_3d(&tmTc, &sm.bar, &sm.pt[1][0][0], x0, y0, plane);
    }
    __syncthreads();
    gemm::tc5::mbar_wait(&sm.bar, 0);
  } else {
#pragma unroll
    for (int i = 0; i < 3; ++i) load_raw_plane(sm.raw.v[i], deriv + (size_t)i * numel + poff, y0, x0, H, W);
    __syncthreads();
  }
  // blur(dmu1) and blur(de11) ride together as the packed pair (sum, difference); blur(de12) takes the scalar stencil
  hpass2<3, 1, XO>(sm.raw, sm.hb2, [](int, const float (&px)[3]) -> f2 { return pack2(px[0] + px[1], px[0] - px[1]); });
  hpass<3, 1, XO>(sm.raw, sm.hb, [](int, const float (&px)[3]) -> float { return px[2]; });
  __syncthreads();
  f2 res2[1][8];
  vpass2<1>(sm.hb2, res2);
  float res12[1][8];
  vpass<1>(sm.hb, res12);
  const int x = threadIdx.x & 63, h = threadIdx.x >> 6;
  if (x0 + x < W) {
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const int yy = y0 + 8 * h + o;
      if (yy >= H) continue;
      const size_t e = poff + (size_t)yy * W + (x0 + x);
      const float p = kTma ? sm.pt[0][8 * h + o][x] : __ldg(pred + e), t = kTma ? sm.pt[1][8 * h + o][x] : __ldg(target + e);
      float bs, bd;
      unpack2(res2[0][o], bs, bd);
      const float b_mu1 = 0.5f * (bs + bd), b_e11 = 0.5f * (bs - bd);
      const float gs = b_mu1 + 2.0f * p * b_e11 + t * res12[0][o];
      grad[e] = k_ssim * gs + k_mse * (p - t);
    }
  }
}

__global__ void loss_finalize_kernel(const double* __restrict__ acc, double n, float w_mse, float w_ssim, float* out) {
  msd_pdl_enter();
  const float mse = (float)(acc[0] / n);
  const float ssim_loss = 1.0f - (float)(acc[1] / n);
  out[0] = w_mse * mse + w_ssim * ssim_loss;
  out[1] = mse;
  out[2] = ssim_loss;
}

}  // namespace

extern "C" size_t msd_loss_workspace_bytes(int B, int C, int H, int W) {
  size_t numel = (size_t)B * C * H * W;
  return 256 + 3 * numel * sizeof(float);
}

extern "C" int msd_loss_fwd_bwd_f32(const float* d_pred, const float* d_target, int B, int C, int H, int W, float w_mse,
                                    float w_ssim, float data_range, float grad_scale, float* d_out, float* d_grad_pred,
                                    void* d_ws, size_t ws_bytes, void* stream) {
  MSD_REQUIRE(d_pred && d_target && d_out && d_ws, "loss: null pointer");
  MSD_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "loss: bad shape [%d,%d,%d,%d]", B, C, H, W);
  const size_t numel = (size_t)B * C * H * W;
  MSD_REQUIRE(numel < (1ull << 31), "loss: tensor too large");
  const size_t need = d_grad_pred ? msd_loss_workspace_bytes(B, C, H, W) : 256;
  if (ws_bytes < need) {
    msd_set_error("loss: workspace %zu < %zu bytes", ws_bytes, need);
    return MSD_ERR_WORKSPACE;
  }
  MSD_REQUIRE((reinterpret_cast<uintptr_t>(d_ws) & 255u) == 0, "loss: workspace must be 256-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  double* acc = reinterpret_cast<double*>(d_ws);
  float* deriv = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(d_ws) + 256);
  MSD_CHECK_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(double), s));
  SsimConsts consts;
  consts.C1 = (float)((0.01 * (double)data_range) * (0.01 * (double)data_range));
  consts.C2 = (float)((0.03 * (double)data_range) * (0.03 * (double)data_range));
  static PerDeviceOnce attr_set;
  if (attr_set.first()) {
    MSD_CHECK_CUDA(cudaFuncSetAttribute(loss_fwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwdSmem)));
    MSD_CHECK_CUDA(cudaFuncSetAttribute(loss_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwdSmem)));
    MSD_CHECK_CUDA(cudaFuncSetAttribute(loss_fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwdSmem)));
    MSD_CHECK_CUDA(cudaFuncSetAttribute(loss_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwdSmem)));
    MSD_CHECK_CUDA(cudaFuncSetAttribute(loss_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem)));
    MSD_CHECK_CUDA(cudaFuncSetAttribute(loss_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem)));
  }
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  const unsigned blocks = (unsigned)B * C * tiles_x * tiles_y;
  const int planes = B * C;
  CUtensorMap tmP, tmT, tmD, tmPc, tmTc;
  // TMA needs 16-byte aligned rows; anything else (or a driver that refuses a map) takes the per-thread loads
  bool tma = W % 4 == 0 && msd_aligned16(d_pred) && msd_aligned16(d_target) &&
             gemm::tc5::make_map_planes_f32(&tmP, d_pred, (uint64_t)W, (uint64_t)H, (uint64_t)planes, PW, IH) == MSD_OK &&
             gemm::tc5::make_map_planes_f32(&tmT, d_target, (uint64_t)W, (uint64_t)H, (uint64_t)planes, PW, IH) == MSD_OK;
  if (tma && d_grad_pred)
    tma = gemm::tc5::make_map_planes_f32(&tmD, deriv, (uint64_t)W, (uint64_t)H, 3ull * planes, PW, IH) == MSD_OK &&
          gemm::tc5::make_map_planes_f32(&tmPc, d_pred, (uint64_t)W, (uint64_t)H, (uint64_t)planes, TW, TH) == MSD_OK &&
          gemm::tc5::make_map_planes_f32(&tmTc, d_target, (uint64_t)W, (uint64_t)H, (uint64_t)planes, TW, TH) == MSD_OK;
  if (!tma) { memset(&tmP, 0, sizeof(tmP)); tmT = tmP; tmD = tmP; }
  if (!tma || !d_grad_pred) { tmD = tmP; tmPc = tmP; tmTc = tmP; }
#define MSD_LOSS_FWD(WD, TM) MSD_CHECK_CUDA(msd_launch_pdl((loss_fwd_kernel<WD, TM>), dim3(blocks), dim3(kThreads), sizeof(FwdSmem), s, tmP, tmT, d_pred, d_target, H, W, tiles_x, tiles_y, consts, numel, d_grad_pred ? deriv : nullptr, acc))
  if (d_grad_pred) { if (tma) MSD_LOSS_FWD(true, true); else MSD_LOSS_FWD(true, false); }
  else { if (tma) MSD_LOSS_FWD(false, true); else MSD_LOSS_FWD(false, false); }
#undef MSD_LOSS_FWD
  MSD_LAUNCH_CHECK();
  MSD_CHECK_CUDA(msd_launch_pdl((loss_finalize_kernel), dim3(1), dim3(1), 0, s, acc, (double)numel, w_mse, w_ssim, d_out));
  MSD_LAUNCH_CHECK();
  if (d_grad_pred) {
    const float k_ssim = (float)(-(double)w_ssim * (double)grad_scale / (double)numel);
    const float k_mse = (float)(2.0 * (double)w_mse * (double)grad_scale / (double)numel);
    if (tma) MSD_CHECK_CUDA(msd_launch_pdl((loss_bwd_kernel<true>), dim3(blocks), dim3(kThreads), sizeof(BwdSmem), s, tmD, tmPc, tmTc, planes, d_pred, d_target, deriv, H, W, tiles_x, tiles_y, numel, k_ssim, k_mse, d_grad_pred));
    else MSD_CHECK_CUDA(msd_launch_pdl((loss_bwd_kernel<false>), dim3(blocks), dim3(kThreads), sizeof(BwdSmem), s, tmD, tmPc, tmTc, planes, d_pred, d_target, deriv, H, W, tiles_x, tiles_y, numel, k_ssim, k_mse, d_grad_pred));
    MSD_LAUNCH_CHECK();
  }
  return MSD_OK;
}
