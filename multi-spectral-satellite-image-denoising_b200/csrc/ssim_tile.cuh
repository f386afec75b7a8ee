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

// ------------------------------------------------------------------------------------------------------------------
// Packed variant: TWO maps ride in the two fp32 lanes of a 64-bit register and every tap is one FFMA2 (fma.rn.f32x2, two IEEE
// fp32 FMAs per issue slot on sm_100).  The stencil kernels are bound by instruction issue, not by the FMA pipe or HBM (ncu:
// 72-80 % issue-active), so halving the FMA instruction count is worth more than anything else here.  Per lane the arithmetic
// and its order are exactly those of hpass / vpass above: results are bit-identical.
//   hb2[NPAIRS][42][64] float2 (same bytes as HBlur<2 * NPAIRS>); a row's 16-byte chunks are rotated by (group >> 1) inside each
//   8-pixel group so that the horizontal pass's 128-bit stores (8 threads of a row, 64 B apart) spread over all banks; the
//   vertical pass reads 8-byte pairs with lanes along x, a permutation inside each 64-byte group -- conflict-free both ways.
// ------------------------------------------------------------------------------------------------------------------
struct f2 {
  unsigned long long v;
};
__device__ __forceinline__ f2 pack2(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f2 a, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return d;
}

template <int NPAIRS>
struct HBlur2 {
  float2 v[NPAIRS][IH][TW];
};

__device__ __forceinline__ void fir8x2(const f2 (&in)[18], f2 (&out)[8]) {
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    f2 a = pack2(0.0f, 0.0f);
#pragma unroll
    for (int k = 0; k < 11; ++k) a = fma2(pack2(tap(k), tap(k)), in[o + k], a);
    out[o] = a;
  }
}

// Horizontal pass, packed.  PairFn(mp, raw values at one pixel) -> f2 holding maps (2 mp, 2 mp + 1) at that pixel.
template <int NRAW, int NPAIRS, int XOFF = 0, typename PairFn>
__device__ __forceinline__ void hpass2(const RawTile<NRAW>& raw, HBlur2<NPAIRS>& hb, PairFn pairfn) {
  constexpr int NV = (18 + XOFF + 3) / 4;
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
    for (int mp = 0; mp < NPAIRS; ++mp) {
      f2 cur[18], out[8];
#pragma unroll
      for (int j = 0; j < 18; ++j) {
        float px[NRAW];
#pragma unroll
        for (int i = 0; i < NRAW; ++i) px[i] = in[i][XOFF + j];
        cur[j] = pairfn(mp, px);
      }
      fir8x2(cur, out);
      // pixels 8g .. 8g+7 of the row = four 16-byte chunks (two float2 each); chunk q is stored at position (q + (g >> 1)) & 3
      ulonglong2* dst = reinterpret_cast<ulonglong2*>(&hb.v[mp][r][8 * g]);
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[(q + (g >> 1)) & 3] = make_ulonglong2(out[2 * q].v, out[2 * q + 1].v);
    }
  }
}

// Vertical pass, packed: this thread's column x = tid & 63, rows 8h .. 8h+7; res[mp][o] = maps (2 mp, 2 mp + 1)
template <int NPAIRS>
__device__ __forceinline__ void vpass2(const HBlur2<NPAIRS>& hb, f2 (&res)[NPAIRS][8]) {
  const int x = threadIdx.x & 63, h = threadIdx.x >> 6;
  const int g = x >> 3, j = x & 7;
  const int xs = 8 * g + 2 * (((j >> 1) + (g >> 1)) & 3) + (j & 1);       // where hpass2 put pixel x of a row
#pragma unroll
  for (int mp = 0; mp < NPAIRS; ++mp) {
    f2 in[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) in[i].v = *reinterpret_cast<const unsigned long long*>(&hb.v[mp][8 * h + i][xs]);
    fir8x2(in, res[mp]);
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

// the same map from blur(p + t) = mu1 + mu2 and blur(p - t) = mu1 - mu2:  mu1^2 + mu2^2 = (s^2 + d^2) / 2,  mu1 mu2 = (s^2 - d^2) / 4
__device__ __forceinline__ float ssim_value_sumdiff(float ms, float md, float esum, float e12, const SsimConsts& c) {
  const float s2 = ms * ms, d2 = md * md;
  const float musq = 0.5f * (s2 + d2), mu12 = 0.25f * (s2 - d2);
  const float num = (2.0f * mu12 + c.C1) * (2.0f * (e12 - mu12) + c.C2);
  const float den = (musq + c.C1) * ((esum - musq) + c.C2);
  return __fdividef(num, den);      // den >= C1 * C2 > 0 up to rounding; 2 ulp, no slow-path branch (the map is averaged to 1e-4)
}

}  // namespace ssim
