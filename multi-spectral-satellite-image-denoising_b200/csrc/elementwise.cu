// Bandwidth-bound glue kernels of the restoration forward pass: patchify (im2col for the k16/s16 patch
// embedding, src/models/encoder.py:48-50), LayerNorm (encoder.py:39,51,136 and the blocks' norm1/norm2),
// and the head's 1x1 convolution + NHWC->NCHW store (src/models/decoder.py:80).
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------
// patchify: x [B,C,S,S] fp32 -> A [B*G*G, C*P*P] bf16, column (c,ky,kx), row token n = G*py + px
// one thread: 8 consecutive kx of one (token, c, ky) -> two float4 loads, one 16-byte store
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ a, int B,
                                                       int C, int S, int P, int G) {
  msd_pdl_enter();
  const int K = C * P * P;
  const int vec_per_row = K / 8;
  const size_t total = (size_t)B * G * G * vec_per_row;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const size_t row = i / vec_per_row;
    const int col = (int)(i - row * vec_per_row) * 8;
    const int c = col / (P * P), rem = col - c * P * P, ky = rem / P, kx = rem - ky * P;
    const int b = (int)(row / (G * G)), n = (int)(row - (size_t)b * G * G), py = n / G, px = n - py * G;
    const float* src = x + (((size_t)b * C + c) * S + (py * P + ky)) * S + px * P + kx;
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(src));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
    uint4 o;
    o.x = pack_bf16x2(v0.x, v0.y); o.y = pack_bf16x2(v0.z, v0.w);
    o.z = pack_bf16x2(v1.x, v1.y); o.w = pack_bf16x2(v1.z, v1.w);
    *reinterpret_cast<uint4*>(a + row * K + col) = o;
  }
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm over the last dim, one warp per row, two-pass (mean, then centred variance) in registers.
//   kMode 0: out_bf16[m] = LN(x[m])                      (block norm1/norm2, encoder.norm)
//   kMode 1: x[m] = LN(x[m]) + pos[m % period]  (fp32)   (patch_embed.norm then + pos_embed, encoder.py:51,309)
//   kMode 2: kMode 1 + out_bf16[m] = bf16(x[m]) and the row's {sum x, sum x^2} in stats[m][0] (slots 1..parts-1 zeroed): the
//            first block's norm1 is folded into its qkv GEMM, which reads exactly what a stats-producing GEMM would have left
// ---------------------------------------------------------------------------------------------------
template <int kChunks, int kMode>
__global__ void __launch_bounds__(256) layernorm_kernel(float* __restrict__ x, const float* __restrict__ g,
                                                        const float* __restrict__ bta, float eps,
                                                        __nv_bfloat16* __restrict__ out, const float* __restrict__ pos,
                                                        int period, int M, float2* __restrict__ stats = nullptr, int parts = 0) {
  msd_pdl_enter();
  constexpr int D = kChunks * 128;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  float* xr = x + (size_t)row * D;
  float4 v[kChunks];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < kChunks; ++i) {
    v[i] = *reinterpret_cast<const float4*>(xr + i * 128 + lane * 4);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < kChunks; ++i) {
    float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + eps);
  float so = 0.0f, qo = 0.0f;
#pragma unroll
  for (int i = 0; i < kChunks; ++i) {
    const int col = i * 128 + lane * 4;
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g + col));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bta + col));
    float4 o;
    o.x = (v[i].x - mean) * rstd * gg.x + bb.x;
    o.y = (v[i].y - mean) * rstd * gg.y + bb.y;
    o.z = (v[i].z - mean) * rstd * gg.z + bb.z;
    o.w = (v[i].w - mean) * rstd * gg.w + bb.w;
    if (kMode == 0) {
      uint2 pk;
      pk.x = pack_bf16x2(o.x, o.y);
      pk.y = pack_bf16x2(o.z, o.w);
      *reinterpret_cast<uint2*>(out + (size_t)row * D + col) = pk;
    } else {
      const float4 pp = __ldg(reinterpret_cast<const float4*>(pos + (size_t)(row % period) * D + col));
      o.x += pp.x; o.y += pp.y; o.z += pp.z; o.w += pp.w;
      *reinterpret_cast<float4*>(xr + col) = o;
      if (kMode == 2) {
        uint2 pk;
        pk.x = pack_bf16x2(o.x, o.y);
        pk.y = pack_bf16x2(o.z, o.w);
        *reinterpret_cast<uint2*>(out + (size_t)row * D + col) = pk;
        so += (o.x + o.y) + (o.z + o.w);
        qo = fmaf(o.x, o.x, fmaf(o.y, o.y, fmaf(o.z, o.z, fmaf(o.w, o.w, qo))));
      }
    }
  }
  if (kMode == 2) {
    so = warp_sum(so);
    qo = warp_sum(qo);
    for (int i = lane; i < parts; i += 32) stats[(size_t)row * parts + i] = i == 0 ? make_float2(so, qo) : make_float2(0.0f, 0.0f);
  }
}

template <int kMode>
int launch_layernorm(float* x, const float* g, const float* b, float eps, __nv_bfloat16* out, const float* pos,
                     int period, int M, int D, cudaStream_t s, float2* stats = nullptr, int parts = 0) {
  MSD_REQUIRE(D % 128 == 0 && D >= 128 && D <= 1024, "layernorm: D=%d must be a multiple of 128 in [128,1024]", D);
  const unsigned blocks = (unsigned)((M + 7) / 8);
  switch (D / 128) {
#define MSD_LN_CASE(n) \
    case n: MSD_CHECK_CUDA(msd_launch_pdl((layernorm_kernel<n, kMode>), dim3(blocks), dim3(256), 0, s, x, g, b, eps, out, pos, period, M, stats, parts)); break;
    MSD_LN_CASE(1) MSD_LN_CASE(2) MSD_LN_CASE(3) MSD_LN_CASE(4) MSD_LN_CASE(5) MSD_LN_CASE(6) MSD_LN_CASE(7) MSD_LN_CASE(8)
#undef MSD_LN_CASE
  }
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

// ---------------------------------------------------------------------------------------------------
// head 1x1 conv: haloed NHWC bf16 [B,H+2,W+2,Cin] -> NCHW fp32 [B,Cout,H,W]; one thread per pixel,
// weights [Cout,Cin] fp32 broadcast from smem.  Cin <= 64, Cout <= 16.
// ---------------------------------------------------------------------------------------------------
template <int CIN, int COUT>   // compile-time upper bounds of Cin (multiple of 8) / Cout: <48,13> is the reference model
__global__ void __launch_bounds__(256) head1x1_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ y, int B, int H,
                                                      int W, int Cin, int Cout) {
  msd_pdl_enter();
  // Two pixels per thread: every broadcast 128-bit weight load feeds 8 FMAs.  With one pixel per thread the kernel was bound
  // by the shared-memory return path (an LDS.128 occupies it for 4 cycles, the same as the 4 FMAs it fed), not by HBM or FMA.
  constexpr int PPT = 2;
  __shared__ __align__(16) float sw[COUT * CIN];
  __shared__ float sb[16];
  for (int i = threadIdx.x; i < COUT * CIN; i += 256) {
    const int o = i / CIN, c = i - o * CIN;
    sw[i] = (o < Cout && c < Cin) ? w[o * Cin + c] : 0.0f;
  }
  if (threadIdx.x < 16) sb[threadIdx.x] = threadIdx.x < Cout ? bias[threadIdx.x] : 0.0f;
  __syncthreads();
  const uint32_t npix = (uint32_t)B * H * W, hw = (uint32_t)H * W, step = gridDim.x * (256 * PPT);
  // software pipeline: the raw bf16 rows of the NEXT pair of pixels are requested before the current pair is computed
  // (a thread holds ~1000 instructions of work per pair; only 8 warps fit per SM, too few to hide HBM latency otherwise)
  uint4 raw[PPT][CIN / 8], nxt[PPT][CIN / 8];
  auto fetch = [&](uint32_t p0, uint4 (&dst)[PPT][CIN / 8]) {
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      const uint32_t p = p0 + q * 256u;
      const uint32_t pp = p < npix ? p : 0u;
      const uint32_t b = pp / hw, rem = pp - b * hw;
      const uint32_t yy = rem / (uint32_t)W, xx = rem - yy * (uint32_t)W;
      const __nv_bfloat16* src = x + (((size_t)b * (H + 2) + (yy + 1)) * (W + 2) + (xx + 1)) * Cin;
#pragma unroll
      for (int c8 = 0; c8 < CIN / 8; ++c8)
        dst[q][c8] = (p < npix && c8 * 8 < Cin) ? __ldg(reinterpret_cast<const uint4*>(src + c8 * 8)) : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  uint32_t p0 = blockIdx.x * (256u * PPT) + threadIdx.x;
  if (p0 < npix) fetch(p0, raw);
  for (; p0 < npix; p0 += step) {
    if (p0 + step < npix) fetch(p0 + step, nxt);
    float in[PPT][CIN];
#pragma unroll
    for (int q = 0; q < PPT; ++q)
#pragma unroll
      for (int c8 = 0; c8 < CIN / 8; ++c8) {
        const uint4 u = raw[q][c8];
        float2 f;
        f = unpack_bf16x2(u.x); in[q][c8 * 8 + 0] = f.x; in[q][c8 * 8 + 1] = f.y;
        f = unpack_bf16x2(u.y); in[q][c8 * 8 + 2] = f.x; in[q][c8 * 8 + 3] = f.y;
        f = unpack_bf16x2(u.z); in[q][c8 * 8 + 4] = f.x; in[q][c8 * 8 + 5] = f.y;
        f = unpack_bf16x2(u.w); in[q][c8 * 8 + 6] = f.x; in[q][c8 * 8 + 7] = f.y;
      }
    float acc[PPT][COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o)
#pragma unroll
      for (int q = 0; q < PPT; ++q) acc[q][o] = sb[o];
#pragma unroll
    for (int c4 = 0; c4 < CIN / 4; ++c4) {
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        const float4 w4 = *reinterpret_cast<const float4*>(&sw[o * CIN + c4 * 4]);
#pragma unroll
        for (int q = 0; q < PPT; ++q)
          acc[q][o] = fmaf(in[q][c4 * 4 + 3], w4.w, fmaf(in[q][c4 * 4 + 2], w4.z, fmaf(in[q][c4 * 4 + 1], w4.y, fmaf(in[q][c4 * 4], w4.x, acc[q][o]))));
      }
    }
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      const uint32_t p = p0 + q * 256u;
      if (p < npix) {
        const uint32_t b = p / hw, rem = p - b * hw;
        float* dst = y + (size_t)b * Cout * hw + rem;
#pragma unroll
        for (int o = 0; o < COUT; ++o)
          if (o < Cout) dst[(size_t)o * hw] = acc[q][o];
      }
    }
#pragma unroll
    for (int q = 0; q < PPT; ++q)
#pragma unroll
      for (int c8 = 0; c8 < CIN / 8; ++c8) raw[q][c8] = nxt[q][c8];
  }
}

}  // namespace

extern "C" int msd_patchify_bf16(const float* d_x, void* d_a, int B, int C, int S, int P, void* stream) {
  MSD_REQUIRE(d_x && d_a, "patchify: null pointer");
  MSD_REQUIRE(B > 0 && C > 0 && S > 0 && P > 0 && S % P == 0 && P % 8 == 0, "patchify: bad shape S=%d P=%d", S, P);
  MSD_REQUIRE(msd_aligned16(d_x) && msd_aligned16(d_a), "patchify: pointers must be 16-byte aligned");
  const int G = S / P;
  const size_t total = (size_t)B * G * G * (C * P * P / 8);
  const unsigned blocks = (unsigned)((total + 255) / 256 < 148u * 16u ? (total + 255) / 256 : 148u * 16u);
  MSD_CHECK_CUDA(msd_launch_pdl((patchify_kernel), dim3(blocks), dim3(256), 0, (cudaStream_t)stream, d_x, reinterpret_cast<__nv_bfloat16*>(d_a), B, C, S, P, G));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

extern "C" int msd_layernorm_f32_bf16(const float* d_x, const float* d_g, const float* d_b, float eps, void* d_out_bf16,
                                      int M, int D, void* stream) {
  MSD_REQUIRE(d_x && d_g && d_b && d_out_bf16 && M > 0, "layernorm: bad arguments");
  return launch_layernorm<0>(const_cast<float*>(d_x), d_g, d_b, eps, reinterpret_cast<__nv_bfloat16*>(d_out_bf16), nullptr,
                             1, M, D, (cudaStream_t)stream);
}

extern "C" int msd_layernorm_pos_f32(float* d_x, const float* d_g, const float* d_b, float eps, const float* d_pos,
                                     int period, int M, int D, void* stream) {
  MSD_REQUIRE(d_x && d_g && d_b && d_pos && M > 0 && period > 0, "layernorm_pos: bad arguments");
  return launch_layernorm<1>(d_x, d_g, d_b, eps, nullptr, d_pos, period, M, D, (cudaStream_t)stream);
}

extern "C" int msd_layernorm_pos_stats_f32(float* d_x, const float* d_g, const float* d_b, float eps, const float* d_pos,
                                          int period, void* d_x_bf16, float* d_stats, int parts, int M, int D, void* stream) {
  MSD_REQUIRE(d_x && d_g && d_b && d_pos && d_x_bf16 && d_stats && M > 0 && period > 0 && parts > 0 && parts <= 64, "layernorm_pos_stats: bad arguments");
  return launch_layernorm<2>(d_x, d_g, d_b, eps, reinterpret_cast<__nv_bfloat16*>(d_x_bf16), d_pos, period, M, D, (cudaStream_t)stream,
                             reinterpret_cast<float2*>(d_stats), parts);
}

extern "C" int msd_head1x1_nchw_f32(const void* d_x, const float* d_w, const float* d_b, float* d_y, int B, int H, int W,
                                    int Cin, int Cout, void* stream) {
  MSD_REQUIRE(d_x && d_w && d_b && d_y, "head1x1: null pointer");
  MSD_REQUIRE(Cin % 8 == 0 && Cin <= 64 && Cout <= 16 && Cout > 0, "head1x1: Cin=%d (<=64, %%8) Cout=%d (<=16)", Cin, Cout);
  const size_t npix = (size_t)B * H * W;
  MSD_REQUIRE(npix < (1ull << 30), "head1x1: too many pixels");
  const unsigned blocks = (unsigned)((npix + 511) / 512 < 148u * 8u ? (npix + 511) / 512 : 148u * 8u);
  if (Cin == 48 && Cout == 13)
    MSD_CHECK_CUDA(msd_launch_pdl((head1x1_kernel<48, 13>), dim3(blocks), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(d_x), d_w, d_b, d_y, B, H, W, Cin, Cout));
  else
    MSD_CHECK_CUDA(msd_launch_pdl((head1x1_kernel<64, 16>), dim3(blocks), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const __nv_bfloat16*>(d_x), d_w, d_b, d_y, B, H, W, Cin, Cout));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}
