// Separable 11-tap Gaussian blur of NMAPS maps over one 32x64 output tile, zero padding 5 -- the stencil shared
// by the SSIM metric (src/training/metrics.py:105-117) and the SSIM loss (src/training/losses.py:138-156).
//
// The reference builds the window as outer(g, g) in fp32 and runs 5 depthwise 11x11 convs (121 taps each);
// outer(g,g) is exactly separable, so the tile does a horizontal then a vertical 11-tap pass (22 taps).
//   smem: raw[NRAW][42][80] (zero-padded input halo tile) + hb[NMAPS][42][64] (horizontally blurred rows)
//   horizontal pass: one thread -> 8 adjacent outputs of one row (18 inputs held in registers)
//   vertical pass:   one thread -> 8 adjacent rows of one column (18 inputs per map from smem, conflict-free)
#pragma once
#include "common.cuh"

namespace ssim {

constexpr int R = 5;             // window radius (window_size 11)
constexpr int TH = 32, TW = 64;  // output tile
constexpr int IH = TH + 2 * R;   // 42 input rows
constexpr int IW = TW + 2 * R;   // 74 input cols
constexpr int PW = 80;           // padded raw row (16-byte aligned rows)
constexpr int kThreads = 256;

// fp32 taps of torch: g = exp(-(x^2)/(2*1.5^2)); g /= g.sum()  (metrics.py:76-79, losses.py:111-114)
__device__ __forceinline__ constexpr float tap(int k) {
  return k == 0 || k == 10 ? 0x1.0d957p-10f
       : k == 1 || k == 9  ? 0x1.f1fe02p-8f
       : k == 2 || k == 8  ? 0x1.26eb18p-5f
       : k == 3 || k == 7  ? 0x1.bff0fep-4f
       : k == 4 || k == 6  ? 0x1.b43c3ep-3f
                           : 0x1.10656p-2f;
}

template <int NRAW>
struct RawTile {
  float v[NRAW][IH][PW];
};
template <int NMAPS>
struct HBlur {
  float v[NMAPS][IH][TW];
};

// load the (TH+10) x (TW+10) halo tile of `plane` (row pitch W) into raw[i]; out-of-image -> 0 (zero padding)
// All loads of the thread are issued before the first shared-memory store (14 independent requests in flight).
__device__ __forceinline__ void load_raw_plane(float (*dst)[PW], const float* __restrict__ plane, int y0, int x0, int H,
                                               int W) {
  constexpr int PER = (IH * PW + kThreads - 1) / kThreads;
  float tmp[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int idx = threadIdx.x + i * kThreads;
    const int r = idx / PW, j = idx - r * PW;
    const int gy = y0 - R + r, gx = x0 - R + j;
    tmp[i] = (idx < IH * PW && j < IW && gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(plane + (size_t)gy * W + gx) : 0.0f;
  }
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int idx = threadIdx.x + i * kThreads;
    if (idx < IH * PW) dst[idx / PW][idx % PW] = tmp[i];
  }
}

// 11-tap dot products for 8 adjacent outputs from 18 inputs
__device__ __forceinline__ void fir8(const float (&in)[18], float (&out)[8]) {
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    float a = 0.0f;
#pragma unroll
    for (int k = 0; k < 11; ++k) a = fmaf(tap(k), in[o + k], a);
    out[o] = a;
  }
}

// Horizontal pass.  MapFn(m, raw values at one pixel) -> value of map m at that pixel.
// XOFF: tile column of the first halo pixel (0 when the tile starts at x0 - R; 3 when it starts at x0 - 8, the 16-byte
// aligned start a TMA box needs).
template <int NRAW, int NMAPS, int XOFF = 0, typename MapFn>
__device__ __forceinline__ void hpass(const RawTile<NRAW>& raw, HBlur<NMAPS>& hb, MapFn mapfn) {
  constexpr int NV = (18 + XOFF + 3) / 4;      // float4 loads covering the 18 inputs of 8 adjacent outputs
  static_assert(8 * (TW / 8 - 1) + 4 * NV <= PW, "padded raw row too short");
  for (int item = threadIdx.x; item < IH * (TW / 8); item += kThreads) {
    const int r = item >> 3, g = item & 7;
    float in[NRAW][4 * NV];
#pragma unroll
    for (int i = 0; i < NRAW; ++i) {
      const float4* src = reinterpret_cast<const float4*>(&raw.v[i][r][8 * g]);
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        float4 t = src[q];
        in[i][4 * q + 0] = t.x; in[i][4 * q + 1] = t.y; in[i][4 * q + 2] = t.z; in[i][4 * q + 3] = t.w;
      }
    }
#pragma unroll
    for (int m = 0; m < NMAPS; ++m) {
      float cur[18], out[8];
#pragma unroll
      for (int j = 0; j < 18; ++j) {
        float px[NRAW];
#pragma unroll
        for (int i = 0; i < NRAW; ++i) px[i] = in[i][XOFF + j];
        cur[j] = mapfn(m, px);
      }
      fir8(cur, out);
      float4* dst = reinterpret_cast<float4*>(&hb.v[m][r][8 * g]);
      dst[0] = make_float4(out[0], out[1], out[2], out[3]);
      dst[1] = make_float4(out[4], out[5], out[6], out[7]);
    }
  }
}

// Vertical pass for this thread's column x = tid & 63 and rows 8h..8h+7 (h = tid >> 6): res[m][o]
template <int NMAPS>
__device__ __forceinline__ void vpass(const HBlur<NMAPS>& hb, float (&res)[NMAPS][8]) {
  const int x = threadIdx.x & 63, h = threadIdx.x >> 6;
#pragma unroll
  for (int m = 0; m < NMAPS; ++m) {
    float in[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) in[i] = hb.v[m][8 * h + i][x];
    fir8(in, res[m]);
  }
}

// ssim_map value and (optionally) its partial derivatives wrt the blurred statistics
struct SsimConsts {
  float C1, C2;
};
__device__ __forceinline__ float ssim_value(float mu1, float mu2, float e11, float e22, float e12, const SsimConsts& c) {
  float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu1_mu2 = mu1 * mu2;
  float s1 = e11 - mu1_sq, s2 = e22 - mu2_sq, s12 = e12 - mu1_mu2;
  float num = (2.0f * mu1_mu2 + c.C1) * (2.0f * s12 + c.C2);
  float den = (mu1_sq + mu2_sq + c.C1) * (s1 + s2 + c.C2);
  return num / den;
}

// same map from the blurred statistics mu1, mu2, E[p^2 + t^2], E[pt]
__device__ __forceinline__ float ssim_value_sum(float mu1, float mu2, float esum, float e12, const SsimConsts& c) {
  const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu1_mu2 = mu1 * mu2;
  const float s12 = e12 - mu1_mu2;
  const float num = (2.0f * mu1_mu2 + c.C1) * (2.0f * s12 + c.C2);
  const float den = (mu1_sq + mu2_sq + c.C1) * ((esum - mu1_sq - mu2_sq) + c.C2);
  return num / den;
}

}  // namespace ssim
