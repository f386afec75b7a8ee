// Bandwidth-bound kernels of the training step (everything around the tensor-core GEMMs): training-mode BatchNorm
// (batch statistics, reference blocks.py:66-72 / 113-118 nn.BatchNorm2d in .train()), its backward, LayerNorm backward,
// GELU backward, bias gradients (column sums), the pixel un-shuffle of the ConvTranspose backward, the head 1x1 backward
// and the positional-embedding gradient.  Each makes one pass over its operands with 128-bit accesses; per-channel sums
// are accumulated in registers, reduced per block in shared memory and merged with fp64 / fp32 atomics.
//
// Layout conventions (see model.cu): decoder activations are NHWC bf16 [B, H+2, W+2, C] with a zero 1-pixel halo,
// "rows" = padded pixels; encoder activations are [tokens, features].
#include <cstdlib>
#include "common.cuh"

namespace {

constexpr int kBlock = 256;

__device__ __forceinline__ bool interior_row(uint32_t p, const FastDiv& plane, const FastDiv& pitch, int H, int W) {
  uint32_t img, rem, yy, xx;
  fd_divmod(p, plane, img, rem);
  fd_divmod(rem, pitch, yy, xx);
  return yy >= 1 && yy <= (uint32_t)H && xx >= 1 && xx <= (uint32_t)W;
}

// ---------------------------------------------------------------------------------------------------
// Per-column sums over the rows of a bf16 [rows, C] matrix:  NS = 1: sum a;  NS = 2: sum a, sum a*a;
// with a second matrix b and per-channel (mean, rstd): sum a, sum a * (b - mean) * rstd     (BatchNorm backward)
// Thread t owns the 8-channel vector t % VPR of rows t / VPR, t / VPR + RPI, ...; block partials -> fp64 atomics.
// ---------------------------------------------------------------------------------------------------
template <int MODE>   // 0: sum a | 1: sum a, sum a^2 | 2: sum a, sum a*xhat(b)
__global__ void __launch_bounds__(kBlock) colsum_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                        const float* __restrict__ stats /* [4][C]: .., .., mean, rstd */,
                                                        double* __restrict__ sums /* [2][C] */, long long rows, int C, int ld) {
  msd_pdl_enter();
  extern __shared__ float sh[];   // [2][kBlock][8]
  const int vpr = C >> 3;
  const int rpi = kBlock / vpr;
  const int v = threadIdx.x % vpr, r0 = threadIdx.x / vpr;
  float s1[8], s2[8], mean[8], rstd[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; mean[i] = 0.f; rstd[i] = 1.f; }
  if (MODE == 2 && r0 < rpi) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { mean[i] = stats[2 * C + v * 8 + i]; rstd[i] = stats[3 * C + v * 8 + i]; }
  }
  if (r0 < rpi) {
    const long long stride = (long long)gridDim.x * rpi;
    if (MODE == 2) {
      for (long long row = (long long)blockIdx.x * rpi + r0; row < rows; row += stride) {
        float fa[8], fb[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(a + row * ld + v * 8)), fa);
        unpack8(__ldg(reinterpret_cast<const uint4*>(b + row * ld + v * 8)), fb);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s1[i] += fa[i]; s2[i] = fmaf(fa[i], (fb[i] - mean[i]) * rstd[i], s2[i]); }
      }
    } else {
      // four rows in flight per thread (one 16-byte load each): with a single load the pass was latency-bound at 3.4 TB/s
      for (long long row = (long long)blockIdx.x * rpi + r0; row < rows; row += 4 * stride) {
        uint4 q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long long r = row + u * stride;
          q[u] = r < rows ? __ldg(reinterpret_cast<const uint4*>(a + r * ld + v * 8)) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float fa[8];
          unpack8(q[u], fa);
#pragma unroll
          for (int i = 0; i < 8; ++i) { s1[i] += fa[i]; if (MODE == 1) s2[i] = fmaf(fa[i], fa[i], s2[i]); }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { sh[threadIdx.x * 8 + i] = s1[i]; sh[(kBlock + threadIdx.x) * 8 + i] = s2[i]; }
  __syncthreads();
  // thread (v, i) of the first C threads folds the rpi row-groups
  for (int c = threadIdx.x; c < C; c += kBlock) {
    const int vv = c >> 3, ii = c & 7;
    double t1 = 0.0, t2 = 0.0;
    for (int g = 0; g < rpi; ++g) {
      t1 += sh[(g * vpr + vv) * 8 + ii];
      t2 += sh[(kBlock + g * vpr + vv) * 8 + ii];
    }
    atomicAdd(&sums[c], t1);
    if (MODE != 0) atomicAdd(&sums[C + c], t2);
  }
}

// BatchNorm finalize (training): batch mean / biased variance -> scale, shift, mean, rstd; running statistics are
// updated in place with momentum (unbiased variance), as nn.BatchNorm2d does.
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ stats, int C) {
  msd_pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = sums[c] / count;
  double var = sums[C + c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * rstd;
  stats[c] = sc;
  stats[C + c] = beta[c] - (float)mean * sc;
  stats[2 * C + c] = (float)mean;
  stats[3 * C + c] = rstd;
  if (running_mean != nullptr) {
    running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)(var * count / (count > 1.0 ? count - 1.0 : 1.0));
  }
}

// y = GELU(scale * c + shift [+ skip]) on interior pixels, 0 on the halo.  Thread t owns the 8-channel vector t % VPR of rows
// t / VPR, + RPI, ... (no per-element division; scale / shift live in registers)
__global__ void __launch_bounds__(kBlock) bn_act_fwd_kernel(const __nv_bfloat16* __restrict__ c, const __nv_bfloat16* __restrict__ skip,
                                                            const float* __restrict__ stats, __nv_bfloat16* __restrict__ y,
                                                            long long rows, int C, FastDiv plane, FastDiv pitch, int H, int W) {
  msd_pdl_enter();
  const int vpr = C >> 3;
  const int rpi = kBlock / vpr;
  const int v = threadIdx.x % vpr, r0 = threadIdx.x / vpr;
  if (r0 >= rpi) return;
  float sc[8], sf[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] = stats[v * 8 + k]; sf[k] = stats[C + v * 8 + k]; }
  for (long long row = (long long)blockIdx.x * rpi + r0; row < rows; row += (long long)gridDim.x * rpi) {
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (interior_row((uint32_t)row, plane, pitch, H, W)) {
      float f[8], s[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(c + row * C + v * 8)), f);
      if (skip != nullptr) unpack8(__ldg(reinterpret_cast<const uint4*>(skip + row * C + v * 8)), s);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        f[k] = fmaf(f[k], sc[k], sf[k]);
        if (skip != nullptr) f[k] += s[k];
      }
      // the same half2 tanh-form GELU the eval-path GEMM epilogues apply (common.cuh), so train and eval activations agree
      const uint2 a = gelu4_to_bf16(make_float4(f[0], f[1], f[2], f[3])), b = gelu4_to_bf16(make_float4(f[4], f[5], f[6], f[7]));
      o = make_uint4(a.x, a.y, b.x, b.y);
    }
    *reinterpret_cast<uint4*>(y + row * C + v * 8) = o;
  }
}

// BatchNorm + GELU backward, pass 1: dz = dy * gelu'(scale c + shift [+ skip]), written as bf16, and the per-channel sums
// sum dz, sum dz * c are accumulated (sum dz * xhat = rstd (sum dz c - mean sum dz) is formed by the finalize kernel, so the loop
// needs only scale / shift per channel: three blocks per SM).
// The pass was issue-bound (ncu: 30.6 thread-instructions per element, issue-active 67 %, 52 % of the HBM rate), so the loop is
// written for instruction count: no interior test -- dy has zero halo rows (every producer in the library writes them), so dz and both
// sums come out zero there by themselves; 32-bit element offsets shared by the four tensors; bf16 pairs unpacked with one shift and
// one mask; the +-8 clamp of gelu' done on the packed half2 value.
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
// gelu'(z) for a packed pair, tanh form (see gelu_grad2 in common.cuh), |z| clamped to 8 in half2
__device__ __forceinline__ float2 gelu_grad_pair(float z0, float z1) {
  const __half2 lim = __float2half2_rn(8.0f);
  const __half2 z = __hmax2(__hmin2(__floats2half2_rn(z0, z1), lim), __hneg2(lim));
  const __half2 c0 = __float2half2_rn(0.7978845608f), c1 = __float2half2_rn(0.0356774081f), c3 = __float2half2_rn(0.1070322243f);
  const __half2 half = __float2half2_rn(0.5f), one = __float2half2_rn(1.0f);
  const __half2 z2 = __hmul2(z, z);
  const __half2 u = __hmul2(z, __hfma2(z2, c1, c0));
  uint32_t ui = *reinterpret_cast<const uint32_t*>(&u), ti;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(ti) : "r"(ui));
  const __half2 t = *reinterpret_cast<const __half2*>(&ti);
  const __half2 sech2 = __hfma2(__hneg2(t), t, one);
  const __half2 inner = __hmul2(__hmul2(half, z), __hfma2(z2, c3, c0));
  return __half22float2(__hfma2(sech2, inner, __hfma2(half, t, half)));
}
template <bool kSkip>
__global__ void __launch_bounds__(kBlock, 3) bn_act_bwd1_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ c,
                                                                const __nv_bfloat16* __restrict__ skip, const float* __restrict__ stats,
                                                                __nv_bfloat16* __restrict__ dz, double* __restrict__ sums, uint32_t total,
                                                                int C) {
  msd_pdl_enter();
  extern __shared__ float sh[];
  const int vpr = C >> 3;
  const int rpi = kBlock / vpr;
  const int v = threadIdx.x % vpr, r0 = threadIdx.x / vpr;
  float s1[8], s2[8], sc[8], sf[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; sc[i] = sf[i] = 0.f; }
  if (r0 < rpi) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = stats[v * 8 + i]; sf[i] = stats[C + v * 8 + i]; }
    // two rows per iteration, every load issued before the first use; `total` = rows * C elements (< 2^31)
    const uint32_t stride = gridDim.x * (uint32_t)(rpi * C);
    for (uint32_t off = (blockIdx.x * (uint32_t)rpi + (uint32_t)r0) * (uint32_t)C + (uint32_t)v * 8u; off < total; off += 2u * stride) {
      uint4 rg[2], rc[2], rs[2];
      const bool two = off + stride < total;
      rg[0] = __ldg(reinterpret_cast<const uint4*>(dy + off));
      rc[0] = __ldg(reinterpret_cast<const uint4*>(c + off));
      if (kSkip) rs[0] = __ldg(reinterpret_cast<const uint4*>(skip + off));
      if (two) {
        rg[1] = __ldg(reinterpret_cast<const uint4*>(dy + off + stride));
        rc[1] = __ldg(reinterpret_cast<const uint4*>(c + off + stride));
        if (kSkip) rs[1] = __ldg(reinterpret_cast<const uint4*>(skip + off + stride));
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !two) break;
        const uint32_t g4[4] = {rg[u].x, rg[u].y, rg[u].z, rg[u].w}, c4[4] = {rc[u].x, rc[u].y, rc[u].z, rc[u].w};
        const uint32_t k4[4] = {kSkip ? rs[u].x : 0u, kSkip ? rs[u].y : 0u, kSkip ? rs[u].z : 0u, kSkip ? rs[u].w : 0u};
        uint32_t o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float f0 = bf16_lo(c4[t]), f1 = bf16_hi(c4[t]);
          float z0 = fmaf(f0, sc[2 * t], sf[2 * t]), z1 = fmaf(f1, sc[2 * t + 1], sf[2 * t + 1]);
          if (kSkip) { z0 += bf16_lo(k4[t]); z1 += bf16_hi(k4[t]); }
          const float2 gg = gelu_grad_pair(z0, z1);
          const float d0 = bf16_lo(g4[t]) * gg.x, d1 = bf16_hi(g4[t]) * gg.y;
          s1[2 * t] += d0; s1[2 * t + 1] += d1;
          s2[2 * t] = fmaf(d0, f0, s2[2 * t]);
          s2[2 * t + 1] = fmaf(d1, f1, s2[2 * t + 1]);
          o[t] = pack_bf16x2(d0, d1);
        }
        *reinterpret_cast<uint4*>(dz + off + (u ? stride : 0u)) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { sh[threadIdx.x * 8 + i] = s1[i]; sh[(kBlock + threadIdx.x) * 8 + i] = s2[i]; }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += kBlock) {
    const int vv = ch >> 3, ii = ch & 7;
    double t1 = 0.0, t2 = 0.0;
    for (int g = 0; g < rpi; ++g) {
      t1 += sh[(g * vpr + vv) * 8 + ii];
      t2 += sh[(kBlock + g * vpr + vv) * 8 + ii];
    }
    atomicAdd(&sums[ch], t1);
    atomicAdd(&sums[C + ch], t2);
  }
}

// BatchNorm backward finalize: sums = {sum dz, sum dz * c} -> sum dz * xhat = rstd (sum dz c - mean sum dz) (fp64);
// dgamma = sum dz*xhat, dbeta = sum dz, coefficients k1 = sum dz / N, k2 = sum dz*xhat / N
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ stats, double count,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef, int C) {
  msd_pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double sdz = sums[c];
  const double sxh = (double)stats[3 * C + c] * (sums[C + c] - (double)stats[2 * C + c] * sdz);
  if (dgamma != nullptr) dgamma[c] = (float)sxh;
  if (dbeta != nullptr) dbeta[c] = (float)sdz;
  coef[c] = (float)(sdz / count);
  coef[C + c] = (float)(sxh / count);
}

// pass 2: dc = gamma*rstd * (dz - k1 - xhat * k2) on interior pixels, 0 on the halo (same thread layout as the forward)
__global__ void __launch_bounds__(kBlock) bn_act_bwd2_kernel(const __nv_bfloat16* __restrict__ dz, const __nv_bfloat16* __restrict__ c,
                                                             const float* __restrict__ stats, const float* __restrict__ coef,
                                                             __nv_bfloat16* __restrict__ dc, long long rows, int C, FastDiv plane,
                                                             FastDiv pitch, int H, int W) {
  msd_pdl_enter();
  const int vpr = C >> 3;
  const int rpi = kBlock / vpr;
  const int v = threadIdx.x % vpr, r0 = threadIdx.x / vpr;
  if (r0 >= rpi) return;
  // dc = a * dz + b * c + d  with  a = scale, b = -scale*rstd*k2, d = scale * (mean*rstd*k2 - k1)
  float ca[8], cb[8], cd[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = v * 8 + k;
    const float sc = stats[ch], mean = stats[2 * C + ch], rstd = stats[3 * C + ch], k1 = coef[ch], k2 = coef[C + ch];
    ca[k] = sc; cb[k] = -sc * rstd * k2; cd[k] = sc * (mean * rstd * k2 - k1);
  }
  for (long long row = (long long)blockIdx.x * rpi + r0; row < rows; row += (long long)gridDim.x * rpi) {
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (interior_row((uint32_t)row, plane, pitch, H, W)) {
      float g[8], f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(dz + row * C + v * 8)), g);
      unpack8(__ldg(reinterpret_cast<const uint4*>(c + row * C + v * 8)), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] = fmaf(ca[k], g[k], fmaf(cb[k], f[k], cd[k]));
      o = pack8(g);
    }
    *reinterpret_cast<uint4*>(dc + row * C + v * 8) = o;
  }
}

// fp64 [n] -> fp32 [n] (bias gradients out of the column-sum accumulators)
__global__ void f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, int n) {
  msd_pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

// dh_pre = dh * gelu'(h_pre), bf16 elementwise (in place allowed)
__global__ void __launch_bounds__(kBlock) gelu_bwd_kernel(const __nv_bfloat16* __restrict__ dh, const __nv_bfloat16* __restrict__ hpre,
                                                          __nv_bfloat16* __restrict__ out, long long nvec) {
  msd_pdl_enter();
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < nvec; i += (long long)gridDim.x * kBlock) {
    float g[8], z[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dh) + i), g);
    unpack8(__ldg(reinterpret_cast<const uint4*>(hpre) + i), z);
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      const float2 gg = gelu_grad2(z[k], z[k + 1]);
      g[k] *= gg.x; g[k + 1] *= gg.y;
    }
    reinterpret_cast<uint4*>(out)[i] = pack8(g);
  }
}

// The same over a [rows, N] matrix, and colsum[N] += column sums of the (bf16-rounded) result: the bias gradient of the Linear that
// produced h_pre (fc1) comes out of the pass that forms its output gradient.  A block owns a stripe of 256 columns (32 vectors) and
// every eighth row of its row group (one warp per row: 512 contiguous bytes); the eight warps are folded through shared memory
// and each thread issues ONE atomic -- same-address atomics from different CTAs serialise in L2 (~20 ns each), so their number
// per column (= gridDim.x), not their total, is what must stay small.
__global__ void __launch_bounds__(kBlock) gelu_bwd_colsum_kernel(const __nv_bfloat16* __restrict__ dh, const __nv_bfloat16* __restrict__ hpre,
                                                                 __nv_bfloat16* __restrict__ out, float* __restrict__ colsum, int rows, int vpr) {
  msd_pdl_enter();
  __shared__ float red[kBlock / 32][256];
  const int cv = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int v = blockIdx.y * 32 + cv;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int row = blockIdx.x * (kBlock / 32) + rl; row < rows; row += gridDim.x * (kBlock / 32)) {
    const size_t i = (size_t)row * vpr + v;
    float g[8], z[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dh) + i), g);
    unpack8(__ldg(reinterpret_cast<const uint4*>(hpre) + i), z);
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      const float2 gg = gelu_grad2(z[k], z[k + 1]);
      g[k] *= gg.x; g[k + 1] *= gg.y;
    }
    const uint4 o = pack8(g);
    reinterpret_cast<uint4*>(out)[i] = o;
    unpack8(o, g);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += g[k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[rl][cv * 8 + k] = acc[k];
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kBlock / 32; ++w) t += red[w][threadIdx.x];
  atomicAdd(colsum + blockIdx.y * 256 + threadIdx.x, t);
}

// h = GELU(h_pre), bf16 elementwise (the training forward keeps both: h_pre for the backward, h as the fc2 operand)
__global__ void __launch_bounds__(kBlock) gelu_fwd_kernel(const __nv_bfloat16* __restrict__ hpre, __nv_bfloat16* __restrict__ out, long long nvec) {
  msd_pdl_enter();
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < nvec; i += (long long)gridDim.x * kBlock) {
    float z[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(hpre) + i), z);
    const uint2 a = gelu4_to_bf16(make_float4(z[0], z[1], z[2], z[3])), b = gelu4_to_bf16(make_float4(z[4], z[5], z[6], z[7]));
    reinterpret_cast<uint4*>(out)[i] = make_uint4(a.x, a.y, b.x, b.y);
  }
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm backward over rows of width D = 768:
//   dxh = dy * gamma;  dx = rstd * (dxh - mean(dxh) - xhat * mean(dxh * xhat));  dgamma += dy * xhat;  dbeta += dy
// dy is bf16 (GRAD_F32 = 0) or fp32.  g_out fp32 (+= when `accumulate`), optional bf16 copy of the result.
// COLSUM: the column sums of that bf16 copy are accumulated into colsum[D] as well -- the copy is the dY operand of the Linear
// whose output this gradient belongs to (fc2 / proj / patch embedding), so its column sums ARE that layer's bias gradient.
//
// TWO warps share a row (384 columns each, 12 per lane): the per-column accumulators (dgamma, dbeta, column sums) are what fills
// the register file, and with a whole row per warp (72 accumulators + 48 row values per lane, 184 registers) one 8-warp block
// fit an SM -- ncu: 12.6 % of the warp slots, issue-active 14 %, 43 us for 113 MB.  Half rows need 36 + 24 and run 16 warps per
// SM; the two halves exchange their three partial row sums through shared memory (a 64-thread named barrier per exchange).
// One persistent block per SM: 148 atomics per column and array at the end.
// ---------------------------------------------------------------------------------------------------
constexpr int kLnThreads = 512, kLnD = 768, kLnHalf = 384, kLnNV = 3;
// sum of `v` over the two warps of a row pair; slot = this pair's [2] floats for this exchange, bar = the pair's named barrier
__device__ __forceinline__ float pair_sum(float v, float* slot, int h, int bar, int lane) {
  v = warp_sum(v);
  if (lane == 0) slot[h] = v;
  asm volatile("bar.sync %0, 64;\n" ::"r"(bar) : "memory");
  return v + slot[h ^ 1];
}
template <int GRAD_F32, int COLSUM>
__global__ void __launch_bounds__(kLnThreads, 1) ln_bwd_kernel(const void* __restrict__ dy_, const float* __restrict__ x,
                                                               const float* __restrict__ gamma, float eps, float* __restrict__ g_out,
                                                               int accumulate, __nv_bfloat16* __restrict__ g_bf16,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                               float* __restrict__ colsum, int M) {
  msd_pdl_enter();
  constexpr int NV = kLnNV, D = kLnD, PAIRS = kLnThreads / 64;
  __shared__ float s_red[kLnThreads / 32][kLnHalf];
  __shared__ float s_x[PAIRS][4][2];                 // per pair: {sum x, sum (x-mean)^2, m1, m2} x {half 0, half 1}
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp >> 1, h = warp & 1, bar = 1 + pair;
  const int col0 = h * kLnHalf + lane * 4;           // this lane's columns: col0 + 128 v + {0..3}
  float gam[NV][4], dg[NV][4], db[NV][4], cs[COLSUM ? NV : 1][4];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const float4 t = *reinterpret_cast<const float4*>(gamma + col0 + v * 128);
    gam[v][0] = t.x; gam[v][1] = t.y; gam[v][2] = t.z; gam[v][3] = t.w;
#pragma unroll
    for (int k = 0; k < 4; ++k) { dg[v][k] = db[v][k] = 0.f; if (COLSUM) cs[v][k] = 0.f; }
  }
  for (int row = blockIdx.x * PAIRS + pair; row < M; row += gridDim.x * PAIRS) {
    float xv[NV][4], dyv[NV][4], go[NV][4];
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const size_t at = (size_t)row * D + col0 + v * 128;
      const float4 t = *reinterpret_cast<const float4*>(x + at);
      xv[v][0] = t.x; xv[v][1] = t.y; xv[v][2] = t.z; xv[v][3] = t.w;
      s += (t.x + t.y) + (t.z + t.w);
      if (GRAD_F32) {
        const float4 u = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + at);
        dyv[v][0] = u.x; dyv[v][1] = u.y; dyv[v][2] = u.z; dyv[v][3] = u.w;
      } else {
        const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy_) + at);
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
        dyv[v][0] = a.x; dyv[v][1] = a.y; dyv[v][2] = b.x; dyv[v][3] = b.y;
      }
      if (accumulate) {                               // issued with the other loads of the row, used at the very end
        const float4 t2 = *reinterpret_cast<const float4*>(g_out + at);
        go[v][0] = t2.x; go[v][1] = t2.y; go[v][2] = t2.z; go[v][3] = t2.w;
      } else {
        go[v][0] = go[v][1] = go[v][2] = go[v][3] = 0.f;
      }
    }
    const float mean = pair_sum(s, s_x[pair][0], h, bar, lane) * (1.0f / D);
    float var = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int k = 0; k < 4; ++k) { xv[v][k] -= mean; var = fmaf(xv[v][k], xv[v][k], var); }
    const float rstd = rsqrtf(pair_sum(var, s_x[pair][1], h, bar, lane) * (1.0f / D) + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        xv[v][k] *= rstd;                       // xhat
        const float dxh = dyv[v][k] * gam[v][k];
        m1 += dxh;
        m2 = fmaf(dxh, xv[v][k], m2);
        dg[v][k] = fmaf(dyv[v][k], xv[v][k], dg[v][k]);
        db[v][k] += dyv[v][k];
        dyv[v][k] = dxh;
      }
    // both partial sums cross in one exchange
    m1 = warp_sum(m1); m2 = warp_sum(m2);
    if (lane == 0) { s_x[pair][2][h] = m1; s_x[pair][3][h] = m2; }
    asm volatile("bar.sync %0, 64;\n" ::"r"(bar) : "memory");
    m1 = (m1 + s_x[pair][2][h ^ 1]) * (1.0f / D);
    m2 = (m2 + s_x[pair][3][h ^ 1]) * (1.0f / D);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const size_t at = (size_t)row * D + col0 + v * 128;
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = fmaf(rstd, dyv[v][k] - m1 - xv[v][k] * m2, go[v][k]);
      *reinterpret_cast<float4*>(g_out + at) = make_float4(o[0], o[1], o[2], o[3]);
      const uint2 ob = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
      if (g_bf16 != nullptr) *reinterpret_cast<uint2*>(g_bf16 + at) = ob;
      if (COLSUM) {
        const float2 a = unpack_bf16x2(ob.x), b = unpack_bf16x2(ob.y);
        cs[v][0] += a.x; cs[v][1] += a.y; cs[v][2] += b.x; cs[v][3] += b.y;
      }
    }
  }
  // fold the 8 warps of each half (one shared buffer, used for dgamma, dbeta, then the column sums), then one atomic per column
  // per block; a frozen LayerNorm (dgamma == NULL) has no parameter gradients
#pragma unroll 1
  for (int pass = (dgamma == nullptr ? 2 : 0); pass < (COLSUM ? 3 : 2); ++pass) {
    __syncthreads();
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int k = 0; k < 4; ++k) s_red[warp][lane * 4 + v * 128 + k] = pass == 0 ? dg[v][k] : pass == 1 ? db[v][k] : cs[COLSUM ? v : 0][k];
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += kLnThreads) {
      const int hh = c >= kLnHalf ? 1 : 0, cc = c - hh * kLnHalf;
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < PAIRS; ++w) a += s_red[2 * w + hh][cc];
      atomicAdd((pass == 0 ? dgamma : pass == 1 ? dbeta : colsum) + c, a);
    }
  }
}

// ConvTranspose2d k2 s2 backward gather: dys[m, q*C + o] = dy[outpix(m, q), o] for interior input pixels m, 0 for halo
// rows (input layout [B, Hin(+2), Win(+2)] pixel list, output haloed [B, 2Hin+2, 2Win+2, C])
__global__ void __launch_bounds__(kBlock) unshuffle_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dys,
                                                           long long rows, int C, FastDiv in_plane, FastDiv in_pitch, int in_pad,
                                                           int Hin, int Win) {
  msd_pdl_enter();
  const int vpr = C >> 3;
  const long long total = rows * 4 * vpr;
  const int out_pitch = 2 * Win + 2;
  const long long out_plane = (long long)(2 * Hin + 2) * out_pitch;
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < total; i += (long long)gridDim.x * kBlock) {
    const long long m = i / (4 * vpr);
    const int rem = (int)(i - m * 4 * vpr);
    const int q = rem / vpr, v = rem - q * vpr;
    uint32_t img, r2, yy, xx;
    fd_divmod((uint32_t)m, in_plane, img, r2);
    fd_divmod(r2, in_pitch, yy, xx);
    bool ok = true;
    if (in_pad) { ok = yy >= 1 && yy <= (uint32_t)Hin && xx >= 1 && xx <= (uint32_t)Win; yy -= 1; xx -= 1; }
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (ok) {
      const long long op = (long long)img * out_plane + (long long)(2 * yy + (q >> 1) + 1) * out_pitch + (2 * xx + (q & 1) + 1);
      o = __ldg(reinterpret_cast<const uint4*>(dy + op * C + v * 8));
    }
    *reinterpret_cast<uint4*>(dys + m * (4 * C) + q * C + v * 8) = o;
  }
}

// head 1x1 backward: dy fp32 NCHW [B,Cout,H,W] ->  dh bf16 haloed NHWC [B,H+2,W+2,Cin] = sum_o dy[o] w[o,c] (halo 0)
// and dyt bf16 [rows, 16] = dy transposed to the pixel list (columns >= Cout and halo rows zero), the operand of the
// weight / bias gradient contractions.  One thread per padded pixel.
template <int CIN>
__global__ void __launch_bounds__(kBlock) head1x1_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                             __nv_bfloat16* __restrict__ dh, __nv_bfloat16* __restrict__ dyt, int B,
                                                             int H, int W, int Cin, int Cout) {
  msd_pdl_enter();
  __shared__ __align__(16) float sw[16 * CIN];
  for (int i = threadIdx.x; i < 16 * CIN; i += kBlock) sw[i] = (i / CIN < Cout && i % CIN < Cin) ? w[(i / CIN) * Cin + i % CIN] : 0.f;
  __syncthreads();
  const long long rows = (long long)B * (H + 2) * (W + 2);
  for (long long p = (long long)blockIdx.x * kBlock + threadIdx.x; p < rows; p += (long long)gridDim.x * kBlock) {
    const int b = (int)(p / ((H + 2) * (W + 2)));
    const int rem = (int)(p - (long long)b * (H + 2) * (W + 2));
    const int yy = rem / (W + 2), xx = rem - yy * (W + 2);
    const bool ok = yy >= 1 && yy <= H && xx >= 1 && xx <= W;
    float g[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) g[o] = (ok && o < Cout) ? __ldg(dy + (((size_t)b * Cout + o) * H + (yy - 1)) * W + (xx - 1)) : 0.f;
    uint4* t = reinterpret_cast<uint4*>(dyt + p * 16);
    t[0] = make_uint4(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]), pack_bf16x2(g[4], g[5]), pack_bf16x2(g[6], g[7]));
    t[1] = make_uint4(pack_bf16x2(g[8], g[9]), pack_bf16x2(g[10], g[11]), pack_bf16x2(g[12], g[13]), pack_bf16x2(g[14], g[15]));
#pragma unroll
    for (int c8 = 0; c8 < CIN / 8; ++c8) {
      if (c8 * 8 < Cin) {
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
        for (int o = 0; o < 16; ++o) {
          const float4 w0 = *reinterpret_cast<const float4*>(&sw[o * CIN + c8 * 8]);
          const float4 w1 = *reinterpret_cast<const float4*>(&sw[o * CIN + c8 * 8 + 4]);
          acc[0] = fmaf(g[o], w0.x, acc[0]); acc[1] = fmaf(g[o], w0.y, acc[1]); acc[2] = fmaf(g[o], w0.z, acc[2]); acc[3] = fmaf(g[o], w0.w, acc[3]);
          acc[4] = fmaf(g[o], w1.x, acc[4]); acc[5] = fmaf(g[o], w1.y, acc[5]); acc[6] = fmaf(g[o], w1.z, acc[6]); acc[7] = fmaf(g[o], w1.w, acc[7]);
        }
        *reinterpret_cast<uint4*>(dh + p * Cin + c8 * 8) = pack8(acc);
      }
    }
  }
}

// out[i] = sum_b in[b * n + i]   (positional-embedding gradient: sum over the batch)
__global__ void __launch_bounds__(kBlock) batch_sum_kernel(const float* __restrict__ in, float* __restrict__ out, int B, long long n) {
  msd_pdl_enter();
  for (long long i = ((long long)blockIdx.x * kBlock + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * kBlock * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < B; ++b) {
      const float4 t = *reinterpret_cast<const float4*>(in + (size_t)b * n + i);
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    *reinterpret_cast<float4*>(out + i) = acc;
  }
}

unsigned grid_for(long long work_items, int per_sm = 8) {
  const long long blocks = (work_items + kBlock - 1) / kBlock;
  const long long cap = 148ll * per_sm;
  return (unsigned)(blocks < 1 ? 1 : blocks < cap ? blocks : cap);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------- C ABI
extern "C" int msd_colsum_bf16(const void* d_a, long long rows, int C, double* d_sums, void* stream) {
  MSD_REQUIRE(d_a && d_sums && rows > 0 && C % 8 == 0, "colsum: bad arguments (C=%d)", C);
  MSD_CHECK_CUDA(cudaMemsetAsync(d_sums, 0, sizeof(double) * C, (cudaStream_t)stream));
  for (int c0 = 0; c0 < C; c0 += 2048) {      // 256 threads x 8 channels per pass over the rows
    const int cc = C - c0 < 2048 ? C - c0 : 2048;
    const int rpi = kBlock / (cc / 8);
    MSD_CHECK_CUDA(msd_launch_pdl((colsum_kernel<0>), dim3(grid_for((rows + rpi - 1) / rpi * kBlock, 4)), dim3(kBlock), 2 * kBlock * 8 * sizeof(float), (cudaStream_t)stream, 
        reinterpret_cast<const __nv_bfloat16*>(d_a) + c0, nullptr, nullptr, d_sums + c0, rows, cc, C));
    MSD_LAUNCH_CHECK();
  }
  return MSD_OK;
}

extern "C" int msd_f64_to_f32(const double* d_in, float* d_out, int n, void* stream) {
  MSD_REQUIRE(d_in && d_out && n > 0, "f64_to_f32: bad arguments");
  MSD_CHECK_CUDA(msd_launch_pdl((f64_to_f32_kernel), dim3((n + 255) / 256), dim3(256), 0, (cudaStream_t)stream, d_in, d_out, n));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

// training-mode BatchNorm statistics of a raw conv output c [B,H+2,W+2,C] (halo rows zero), in two steps so that the first one can
// ride in the epilogue of the conv that writes c (msd_conv3x3_stats_nhwc_bf16):
//   msd_bn_sums_bf16     d_sums fp64 [2][C] = {sum c, sum c^2} over all rows
//   msd_bn_finalize      d_stats fp32 [4][C] = {gamma*rstd, beta - mean*gamma*rstd, mean, rstd}; running statistics updated in
//                        place (momentum, unbiased variance)
extern "C" int msd_bn_sums_bf16(const void* d_c, long long rows, int C, double* d_sums, void* stream) {
  MSD_REQUIRE(d_c && d_sums && rows > 0, "bn_sums: null pointer");
  MSD_REQUIRE(C % 8 == 0 && C / 8 <= kBlock && C <= 1024, "bn_sums: unsupported channel count %d", C);
  cudaStream_t s = (cudaStream_t)stream;
  MSD_CHECK_CUDA(cudaMemsetAsync(d_sums, 0, sizeof(double) * 2 * C, s));
  const int rpi = kBlock / (C / 8);
  MSD_CHECK_CUDA(msd_launch_pdl((colsum_kernel<1>), dim3(grid_for((rows + rpi - 1) / rpi * kBlock, 4)), dim3(kBlock), 2 * kBlock * 8 * sizeof(float), s, 
      reinterpret_cast<const __nv_bfloat16*>(d_c), nullptr, nullptr, d_sums, rows, C, C));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

extern "C" int msd_bn_finalize(const double* d_sums, double count, int C, const float* d_gamma, const float* d_beta, float eps, float momentum,
                               float* d_running_mean, float* d_running_var, float* d_stats, void* stream) {
  MSD_REQUIRE(d_sums && d_gamma && d_beta && d_stats && count > 0 && C > 0, "bn_finalize: bad arguments");
  MSD_CHECK_CUDA(msd_launch_pdl((bn_finalize_kernel), dim3((C + 127) / 128), dim3(128), 0, (cudaStream_t)stream, d_sums, count, d_gamma, d_beta, eps, momentum, d_running_mean,
                                                                         d_running_var, d_stats, C));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

extern "C" int msd_bn_stats_bf16(const void* d_c, int B, int H, int W, int C, const float* d_gamma, const float* d_beta, float eps,
                                 float momentum, float* d_running_mean, float* d_running_var, float* d_stats, double* d_sums,
                                 void* stream) {
  MSD_REQUIRE(d_c && d_gamma && d_beta && d_stats && d_sums, "bn_stats: null pointer");
  int rc = msd_bn_sums_bf16(d_c, (long long)B * (H + 2) * (W + 2), C, d_sums, stream);
  if (rc) return rc;
  return msd_bn_finalize(d_sums, (double)B * H * W, C, d_gamma, d_beta, eps, momentum, d_running_mean, d_running_var, d_stats, stream);
}

extern "C" int msd_bn_act_fwd_bf16(const void* d_c, const void* d_skip, const float* d_stats, void* d_y, int B, int H, int W, int C,
                                   void* stream) {
  MSD_REQUIRE(d_c && d_stats && d_y && C % 8 == 0 && C / 8 <= kBlock, "bn_act_fwd: bad arguments");
  const long long rows = (long long)B * (H + 2) * (W + 2);
  MSD_REQUIRE(rows < (1ll << 31), "bn_act_fwd: too many pixels");
  MSD_CHECK_CUDA(msd_launch_pdl((bn_act_fwd_kernel), dim3(grid_for((rows + kBlock / (C / 8) - 1) / (kBlock / (C / 8)) * kBlock, 16)), dim3(kBlock), 0, (cudaStream_t)stream, 
      reinterpret_cast<const __nv_bfloat16*>(d_c), reinterpret_cast<const __nv_bfloat16*>(d_skip), d_stats,
      reinterpret_cast<__nv_bfloat16*>(d_y), rows, C, make_fastdiv((uint32_t)((H + 2) * (W + 2))), make_fastdiv((uint32_t)(W + 2)), H, W));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

// BatchNorm(train) + GELU backward: d_dy grad of y = GELU(BN(c) [+ skip]); outputs d_dz (grad of the pre-activation: also
// the gradient of the skip input), d_dc (grad of the raw conv output; may alias d_dz), d_dgamma / d_dbeta fp32 [C].
extern "C" int msd_bn_act_bwd_bf16(const void* d_dy, const void* d_c, const void* d_skip, const float* d_stats, void* d_dz, void* d_dc,
                                   float* d_dgamma, float* d_dbeta, float* d_coef, double* d_sums, int B, int H, int W, int C,
                                   void* stream) {
  MSD_REQUIRE(d_dy && d_c && d_stats && d_dz && d_dc && d_coef && d_sums, "bn_act_bwd: null pointer");
  MSD_REQUIRE(C % 8 == 0 && C / 8 <= kBlock, "bn_act_bwd: unsupported channel count %d", C);
  cudaStream_t s = (cudaStream_t)stream;
  const long long rows = (long long)B * (H + 2) * (W + 2);
  MSD_REQUIRE(rows < (1ll << 31), "bn_act_bwd: too many pixels");
  const FastDiv plane = make_fastdiv((uint32_t)((H + 2) * (W + 2))), pitch = make_fastdiv((uint32_t)(W + 2));
  MSD_CHECK_CUDA(cudaMemsetAsync(d_sums, 0, sizeof(double) * 2 * C, s));
  const int rpi = kBlock / (C / 8);
  MSD_REQUIRE(rows * C < (1ll << 31), "bn_act_bwd: tensor too large for 32-bit element offsets");
  const unsigned grid1 = grid_for((rows + rpi - 1) / rpi * kBlock, 12);
  const size_t sh1 = 2 * kBlock * 8 * sizeof(float);
  if (d_skip != nullptr)
    MSD_CHECK_CUDA(msd_launch_pdl((bn_act_bwd1_kernel<true>), dim3(grid1), dim3(kBlock), sh1, s, reinterpret_cast<const __nv_bfloat16*>(d_dy), reinterpret_cast<const __nv_bfloat16*>(d_c),
                                                        reinterpret_cast<const __nv_bfloat16*>(d_skip), d_stats,
                                                        reinterpret_cast<__nv_bfloat16*>(d_dz), d_sums, (uint32_t)(rows * C), C));
  else
    MSD_CHECK_CUDA(msd_launch_pdl((bn_act_bwd1_kernel<false>), dim3(grid1), dim3(kBlock), sh1, s, reinterpret_cast<const __nv_bfloat16*>(d_dy), reinterpret_cast<const __nv_bfloat16*>(d_c),
                                                         nullptr, d_stats, reinterpret_cast<__nv_bfloat16*>(d_dz), d_sums, (uint32_t)(rows * C), C));
  MSD_LAUNCH_CHECK();
  MSD_CHECK_CUDA(msd_launch_pdl((bn_bwd_finalize_kernel), dim3((C + 127) / 128), dim3(128), 0, s, d_sums, d_stats, (double)B * H * W, d_dgamma, d_dbeta, d_coef, C));
  MSD_LAUNCH_CHECK();
  MSD_CHECK_CUDA(msd_launch_pdl((bn_act_bwd2_kernel), dim3(grid_for((rows + rpi - 1) / rpi * kBlock, 16)), dim3(kBlock), 0, s, reinterpret_cast<const __nv_bfloat16*>(d_dz),
                                                                 reinterpret_cast<const __nv_bfloat16*>(d_c), d_stats, d_coef,
                                                                 reinterpret_cast<__nv_bfloat16*>(d_dc), rows, C, plane, pitch, H, W));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

extern "C" int msd_gelu_fwd_bf16(const void* d_hpre, void* d_out, long long n, void* stream) {
  MSD_REQUIRE(d_hpre && d_out && n > 0 && n % 8 == 0, "gelu_fwd: bad arguments");
  MSD_CHECK_CUDA(msd_launch_pdl((gelu_fwd_kernel), dim3(grid_for(n / 8)), dim3(kBlock), 0, (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(d_hpre),
                                                                        reinterpret_cast<__nv_bfloat16*>(d_out), n / 8));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

extern "C" int msd_gelu_bwd_bf16(const void* d_dh, const void* d_hpre, void* d_out, long long n, void* stream) {
  MSD_REQUIRE(d_dh && d_hpre && d_out && n > 0 && n % 8 == 0, "gelu_bwd: bad arguments");
  MSD_CHECK_CUDA(msd_launch_pdl((gelu_bwd_kernel), dim3(grid_for(n / 8)), dim3(kBlock), 0, (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(d_dh),
                                                                        reinterpret_cast<const __nv_bfloat16*>(d_hpre),
                                                                        reinterpret_cast<__nv_bfloat16*>(d_out), n / 8));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

extern "C" int msd_gelu_bwd_colsum_bf16(const void* d_dh, const void* d_hpre, void* d_out, float* d_colsum, int rows, int N, void* stream) {
  MSD_REQUIRE(d_dh && d_hpre && d_out && d_colsum && rows > 0, "gelu_bwd_colsum: null pointer");
  MSD_REQUIRE(N % 256 == 0 && N > 0, "gelu_bwd_colsum: N=%d must be a multiple of 256", N);
  const int stripes = N / 256;
  int row_groups = (148 * 8 + stripes - 1) / stripes;                 // ~8 resident blocks per SM
  const int max_groups = (rows + kBlock / 32 - 1) / (kBlock / 32);
  if (row_groups > max_groups) row_groups = max_groups;
  MSD_CHECK_CUDA(msd_launch_pdl((gelu_bwd_colsum_kernel), dim3(dim3((unsigned)row_groups, (unsigned)stripes)), dim3(kBlock), 0, (cudaStream_t)stream, 
      reinterpret_cast<const __nv_bfloat16*>(d_dh), reinterpret_cast<const __nv_bfloat16*>(d_hpre), reinterpret_cast<__nv_bfloat16*>(d_out),
      d_colsum, rows, N / 8));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

// LayerNorm backward over rows of width D (768 in the reference model): see ln_bwd_kernel.  d_dgamma/d_dbeta are
// ACCUMULATED into (zero them first); NULL skips the parameter gradients (frozen blocks).
extern "C" int msd_layernorm_bwd_colsum(const void* d_dy, int dy_is_f32, const float* d_x, const float* d_gamma, float eps, float* d_g,
                                        int accumulate, void* d_g_bf16, float* d_dgamma, float* d_dbeta, float* d_colsum, int M, int D,
                                        void* stream) {
  MSD_REQUIRE(d_dy && d_x && d_gamma && d_g && M > 0, "layernorm_bwd: null pointer");
  MSD_REQUIRE(D == 768, "layernorm_bwd: only D=768 is built (got %d)", D);
  MSD_REQUIRE((d_dgamma == nullptr) == (d_dbeta == nullptr), "layernorm_bwd: dgamma and dbeta go together");
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((M + 7) / 8 < 148 ? (M + 7) / 8 : 148);
  __nv_bfloat16* gb = reinterpret_cast<__nv_bfloat16*>(d_g_bf16);
#define MSD_LN_BWD(F32, CS) MSD_CHECK_CUDA(msd_launch_pdl((ln_bwd_kernel<F32, CS>), dim3(blocks), dim3(kLnThreads), 0, s, d_dy, d_x, d_gamma, eps, d_g, accumulate, gb, d_dgamma, d_dbeta, d_colsum, M))
  if (d_colsum != nullptr) { if (dy_is_f32) MSD_LN_BWD(1, 1); else MSD_LN_BWD(0, 1); }
  else { if (dy_is_f32) MSD_LN_BWD(1, 0); else MSD_LN_BWD(0, 0); }
#undef MSD_LN_BWD
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

extern "C" int msd_layernorm_bwd(const void* d_dy, int dy_is_f32, const float* d_x, const float* d_gamma, float eps, float* d_g,
                                 int accumulate, void* d_g_bf16, float* d_dgamma, float* d_dbeta, int M, int D, void* stream) {
  return msd_layernorm_bwd_colsum(d_dy, dy_is_f32, d_x, d_gamma, eps, d_g, accumulate, d_g_bf16, d_dgamma, d_dbeta, nullptr, M, D, stream);
}

extern "C" int msd_convt_unshuffle_bf16(const void* d_dy, void* d_dys, int B, int Hin, int Win, int in_padded, int C, void* stream) {
  MSD_REQUIRE(d_dy && d_dys && C % 8 == 0, "unshuffle: bad arguments");
  const int ph = in_padded ? Hin + 2 : Hin, pw = in_padded ? Win + 2 : Win;
  const long long rows = (long long)B * ph * pw;
  MSD_REQUIRE(rows < (1ll << 31), "unshuffle: too many pixels");
  MSD_CHECK_CUDA(msd_launch_pdl((unshuffle_kernel), dim3(grid_for(rows * 4 * (C / 8))), dim3(kBlock), 0, (cudaStream_t)stream, 
      reinterpret_cast<const __nv_bfloat16*>(d_dy), reinterpret_cast<__nv_bfloat16*>(d_dys), rows, C,
      make_fastdiv((uint32_t)(ph * pw)), make_fastdiv((uint32_t)pw), in_padded, Hin, Win));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

extern "C" int msd_head1x1_bwd(const float* d_dy, const float* d_w, void* d_dh, void* d_dyt, int B, int H, int W, int Cin, int Cout,
                               void* stream) {
  MSD_REQUIRE(d_dy && d_w && d_dh && d_dyt, "head1x1_bwd: null pointer");
  MSD_REQUIRE(Cin % 8 == 0 && Cin <= 64 && Cout <= 16 && Cout > 0, "head1x1_bwd: Cin=%d (<=64, %%8) Cout=%d (<=16)", Cin, Cout);
  const long long rows = (long long)B * (H + 2) * (W + 2);
  if (Cin <= 48)
    MSD_CHECK_CUDA(msd_launch_pdl((head1x1_bwd_kernel<48>), dim3(grid_for(rows)), dim3(kBlock), 0, (cudaStream_t)stream, d_dy, d_w, reinterpret_cast<__nv_bfloat16*>(d_dh),
                                                                               reinterpret_cast<__nv_bfloat16*>(d_dyt), B, H, W, Cin, Cout));
  else
    MSD_CHECK_CUDA(msd_launch_pdl((head1x1_bwd_kernel<64>), dim3(grid_for(rows)), dim3(kBlock), 0, (cudaStream_t)stream, d_dy, d_w, reinterpret_cast<__nv_bfloat16*>(d_dh),
                                                                               reinterpret_cast<__nv_bfloat16*>(d_dyt), B, H, W, Cin, Cout));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

extern "C" int msd_batch_sum_f32(const float* d_in, float* d_out, int B, long long n, void* stream) {
  MSD_REQUIRE(d_in && d_out && B > 0 && n > 0 && n % 4 == 0, "batch_sum: bad arguments");
  MSD_CHECK_CUDA(msd_launch_pdl((batch_sum_kernel), dim3(grid_for(n / 4)), dim3(kBlock), 0, (cudaStream_t)stream, d_in, d_out, B, n));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}
