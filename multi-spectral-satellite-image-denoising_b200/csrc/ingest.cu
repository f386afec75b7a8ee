// The data formats either side of the restoration path (BASELINE config C5):
//   * tile ingest   : Sentinel-2 digital numbers (uint16 [B,C,h,w]) -> the normalised fp32 tile the model consumes, i.e.
//                     EuroSATMultiSpectral.__getitem__'s preprocessing (reference src/data/dataset.py:160-186):
//                     scale by 1/10000 (or 1/255, or 1, chosen per image from its maximum) -> bilinear resize to SxS
//                     (align_corners=False) -> clamp(0,1) -> (x - 0.5) / 0.25
//   * scene tiling  : a full scene [C,H,W] is cut into SxS tiles on a regular grid with overlap (the last row / column is
//                     anchored at the scene edge), the restored tiles are blended back with separable linear-ramp weights.
//                     The reference has no scene path: the blend below IS the definition, restated in
//                     oracle/scene.py.
// All kernels are pure streaming gathers: one pass, coalesced along x, no atomics (every output pixel gathers its <= 4 tiles).
#include "common.cuh"

namespace {

constexpr int kBlock = 256;

__global__ void __launch_bounds__(kBlock) image_max_u16_kernel(const uint16_t* __restrict__ dn, float* __restrict__ scale, long long per_image) {
  msd_pdl_enter();
  __shared__ unsigned smax[kBlock / 32];
  const uint16_t* src = dn + (long long)blockIdx.x * per_image;
  unsigned m = 0;
  for (long long i = threadIdx.x; i < per_image; i += kBlock) m = max(m, (unsigned)src[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kBlock / 32; ++w) m = max(m, smax[w]);
    // dataset.py:160-165: raw reflectance (max > 10) / 10000; 8-bit range (max > 1.5) / 255; already normalised otherwise
    scale[blockIdx.x] = m > 10u ? 10000.0f : (m > 1u ? 255.0f : 1.0f);
  }
}

// torch's area_pixel_compute_source_index for align_corners=False: src = scale * (dst + 0.5) - 0.5, clamped at 0
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int& i0, int& i1, float& lam) {
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  lam = s - (float)i0;
}

__global__ void __launch_bounds__(kBlock) ingest_kernel(const uint16_t* __restrict__ dn, const float* __restrict__ scale,
                                                        float* __restrict__ out, int planes_per_image, int h, int w, int S,
                                                        float sh, float sw, long long total) {
  msd_pdl_enter();
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < total; i += (long long)gridDim.x * kBlock) {
    const int x = (int)(i % S);
    const long long t = i / S;
    const int y = (int)(t % S);
    const long long plane = t / S;
    const float div = __ldg(scale + plane / planes_per_image);
    int y0, y1, x0, x1;
    float ly, lx;
    src_index(y, sh, h, y0, y1, ly);
    src_index(x, sw, w, x0, x1, lx);
    const uint16_t* p = dn + plane * (long long)h * w;
    const float v00 = __fdiv_rn((float)__ldg(p + y0 * w + x0), div), v01 = __fdiv_rn((float)__ldg(p + y0 * w + x1), div);
    const float v10 = __fdiv_rn((float)__ldg(p + y1 * w + x0), div), v11 = __fdiv_rn((float)__ldg(p + y1 * w + x1), div);
    // same association as ATen's upsample_bilinear2d: h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11)
    const float hx0 = 1.0f - lx, hy0 = 1.0f - ly;
    float v = hy0 * (hx0 * v00 + lx * v01) + ly * (hx0 * v10 + lx * v11);
    v = fminf(fmaxf(v, 0.0f), 1.0f);
    out[i] = __fdiv_rn(v - 0.5f, 0.25f);
  }
}

// The same arithmetic, one block per plane (the product path for S % 4 == 0 and source planes that fit shared memory): the
// h x w source plane is scaled ONCE into shared memory (the generic kernel above divides the four neighbours of every output
// pixel again: 4 IEEE divisions + 64-bit index arithmetic per element kept it at 0.98 ms per 256 tiles, 8 % of the HBM write
// rate), each thread owns four adjacent output columns (their source indices / weights are computed once, reused for every
// row) and stores 16 bytes.  Same expression tree per output as the generic kernel, so the results are identical.
__global__ void __launch_bounds__(kBlock) ingest_plane_kernel(const uint16_t* __restrict__ dn, const float* __restrict__ scale,
                                                              float* __restrict__ out, int planes_per_image, int h, int w, int S,
                                                              float sh, float sw) {
  msd_pdl_enter();
  extern __shared__ float tile[];                       // [h][w], already divided by the image's scale
  const int plane = blockIdx.x;
  const float div = __ldg(scale + plane / planes_per_image);
  const uint16_t* p = dn + (long long)plane * h * w;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthreads = blockDim.x * blockDim.y;
  for (int i = tid; i < h * w; i += nthreads) tile[i] = __fdiv_rn((float)__ldg(p + i), div);
  __syncthreads();
  int x0[4], x1[4];
  float lx[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) src_index(4 * threadIdx.x + k, sw, w, x0[k], x1[k], lx[k]);
  float* dst = out + (long long)plane * S * S;
  for (int y = threadIdx.y; y < S; y += blockDim.y) {
    int y0, y1;
    float ly;
    src_index(y, sh, h, y0, y1, ly);
    const float* r0 = tile + y0 * w;
    const float* r1 = tile + y1 * w;
    const float hy0 = 1.0f - ly;
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float hx0 = 1.0f - lx[k];
      float v = hy0 * (hx0 * r0[x0[k]] + lx[k] * r0[x1[k]]) + ly * (hx0 * r1[x0[k]] + lx[k] * r1[x1[k]]);
      v = fminf(fmaxf(v, 0.0f), 1.0f);
      o[k] = __fdiv_rn(v - 0.5f, 0.25f);
    }
    *reinterpret_cast<float4*>(dst + (long long)y * S + 4 * threadIdx.x) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// tiles[t, c, y, x] = scene[c, y0[t] + y, x0[t] + x]
__global__ void __launch_bounds__(kBlock) tile_extract_kernel(const float* __restrict__ scene, const int* __restrict__ ty0,
                                                              const int* __restrict__ tx0, float* __restrict__ tiles, int C, int H,
                                                              int W, int S, long long total4) {
  msd_pdl_enter();
  const int s4 = S / 4;
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < total4; i += (long long)gridDim.x * kBlock) {
    const int xv = (int)(i % s4);
    long long t = i / s4;
    const int y = (int)(t % S); t /= S;
    const int c = (int)(t % C);
    const int tile = (int)(t / C);
    const float* src = scene + ((long long)c * H + (__ldg(ty0 + tile) + y)) * W + __ldg(tx0 + tile) + xv * 4;
    float4 v;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) v = __ldg(reinterpret_cast<const float4*>(src));
    else v = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
    reinterpret_cast<float4*>(tiles)[i] = v;
  }
}

// separable linear ramp: 1 in the tile interior, falling to 1/ramp at the tile border over `ramp` pixels
__device__ __forceinline__ float ramp_w(int i, int S, int ramp) {
  const int d = min(i + 1, S - i);
  return (float)min(d, ramp) / (float)ramp;
}

// out[c, Y, X] = sum_{tiles covering (Y,X)} w * restored[tile, c, Y - y0, X - x0] / sum w, for scene rows [row0, row1).
// Tile (r, k) of the grid sits at (gy[r], gx[k]); tile_index[r * nx + k] = its slot in `restored` (-1: not available on this
// rank, which must not happen for the rows it owns).  A pixel gathers at most 3 x 3 tiles (2 x 2 away from the scene edge), in a fixed order (deterministic).
__global__ void __launch_bounds__(kBlock) tile_blend_kernel(const float* __restrict__ restored, const int* __restrict__ gy,
                                                            const int* __restrict__ gx, const int* __restrict__ tile_index, int ny, int nx,
                                                            float* __restrict__ out, int C, int H, int W, int S, int ramp, int row0,
                                                            int row1, long long total) {
  msd_pdl_enter();
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < total; i += (long long)gridDim.x * kBlock) {
    const int X = (int)(i % W);
    long long t = i / W;
    const int Y = row0 + (int)(t % (row1 - row0));
    const int c = (int)(t / (row1 - row0));
    // last grid row/column whose origin is <= Y / X (grids are sorted); the one before it may still cover the pixel
    int r = ny - 1;
    while (r > 0 && __ldg(gy + r) > Y) --r;
    int k = nx - 1;
    while (k > 0 && __ldg(gx + k) > X) --k;
    float acc = 0.f, wsum = 0.f;
#pragma unroll
    for (int dr = 2; dr >= 0; --dr) {   // the edge-anchored last tile can sit closer than `overlap` to its neighbour: 3 candidates
      const int rr = r - dr;
      if (rr < 0) continue;
      const int y = Y - __ldg(gy + rr);
      if (y < 0 || y >= S) continue;
      const float wy = ramp_w(y, S, ramp);
#pragma unroll
      for (int dk = 2; dk >= 0; --dk) {
        const int kk = k - dk;
        if (kk < 0) continue;
        const int x = X - __ldg(gx + kk);
        if (x < 0 || x >= S) continue;
        const int slot = __ldg(tile_index + rr * nx + kk);
        if (slot < 0) continue;
        const float wgt = wy * ramp_w(x, S, ramp);
        acc = fmaf(wgt, __ldg(restored + (((long long)slot * C + c) * S + y) * S + x), acc);
        wsum += wgt;
      }
    }
    out[((long long)c * H + Y) * W + X] = acc / wsum;
  }
}

// inverse of the ingest normalisation: restored fp32 in [-2, 2] -> Sentinel-2 digital numbers (reflectance x 10000, uint16)
__global__ void __launch_bounds__(kBlock) export_kernel(const float* __restrict__ x, uint16_t* __restrict__ dn, long long n4) {
  msd_pdl_enter();
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n4; i += (long long)gridDim.x * kBlock) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    auto q = [](float t) -> uint32_t {
      t = fminf(fmaxf(fmaf(t, 0.25f, 0.5f), 0.0f), 1.0f);       // (x - 0.5) / 0.25 undone, clamp as dataset.py:182 does
      return (uint32_t)__float2int_rn(t * 10000.0f);
    };
    reinterpret_cast<uint2*>(dn)[i] = make_uint2(q(v.x) | (q(v.y) << 16), q(v.z) | (q(v.w) << 16));
  }
}

unsigned grid_for(long long items) {
  const long long blocks = (items + kBlock - 1) / kBlock;
  return (unsigned)(blocks < 1 ? 1 : blocks < 148ll * 16 ? blocks : 148ll * 16);
}

}  // namespace

extern "C" int msd_ingest_u16_f32(const uint16_t* d_dn, float* d_out, float* d_scale_ws, int B, int C, int h, int w, int S, void* stream) {
  MSD_REQUIRE(d_dn && d_out && d_scale_ws && B > 0 && C > 0 && h > 0 && w > 0 && S > 0, "ingest: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  MSD_CHECK_CUDA(msd_launch_pdl((image_max_u16_kernel), dim3((unsigned)B), dim3(kBlock), 0, s, d_dn, d_scale_ws, (long long)C * h * w));
  MSD_LAUNCH_CHECK();
  const long long total = (long long)B * C * S * S;
  const size_t plane_smem = (size_t)h * w * sizeof(float);
  if (S % 4 == 0 && S / 4 <= kBlock && plane_smem <= 48 * 1024 && msd_aligned16(d_out) && (long long)B * C < (1ll << 31)) {
    const dim3 block((unsigned)(S / 4), (unsigned)(kBlock / (S / 4) > 0 ? kBlock / (S / 4) : 1));
    MSD_CHECK_CUDA(msd_launch_pdl((ingest_plane_kernel), dim3((unsigned)(B * C)), dim3(block), plane_smem, s, d_dn, d_scale_ws, d_out, C, h, w, S, (float)h / (float)S, (float)w / (float)S));
  } else {
    MSD_CHECK_CUDA(msd_launch_pdl((ingest_kernel), dim3(grid_for(total)), dim3(kBlock), 0, s, d_dn, d_scale_ws, d_out, C, h, w, S, (float)h / (float)S, (float)w / (float)S, total));
  }
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

extern "C" int msd_tile_extract_f32(const float* d_scene, const int* d_ty0, const int* d_tx0, float* d_tiles, int n_tiles, int C, int H,
                                    int W, int S, void* stream) {
  MSD_REQUIRE(d_scene && d_ty0 && d_tx0 && d_tiles && n_tiles > 0 && S % 4 == 0 && H >= S && W >= S, "tile_extract: bad arguments");
  const long long total4 = (long long)n_tiles * C * S * (S / 4);
  MSD_CHECK_CUDA(msd_launch_pdl((tile_extract_kernel), dim3(grid_for(total4)), dim3(kBlock), 0, (cudaStream_t)stream, d_scene, d_ty0, d_tx0, d_tiles, C, H, W, S, total4));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

extern "C" int msd_tile_blend_f32(const float* d_restored, const int* d_gy, const int* d_gx, const int* d_tile_index, int ny, int nx,
                                  float* d_out, int C, int H, int W, int S, int ramp, int row0, int row1, void* stream) {
  MSD_REQUIRE(d_restored && d_gy && d_gx && d_tile_index && d_out, "tile_blend: null pointer");
  MSD_REQUIRE(ny > 0 && nx > 0 && ramp >= 1 && 0 <= row0 && row0 < row1 && row1 <= H, "tile_blend: bad grid / rows");
  const long long total = (long long)C * (row1 - row0) * W;
  MSD_CHECK_CUDA(msd_launch_pdl((tile_blend_kernel), dim3(grid_for(total)), dim3(kBlock), 0, (cudaStream_t)stream, d_restored, d_gy, d_gx, d_tile_index, ny, nx, d_out, C, H, W, S,
                                                                         ramp, row0, row1, total));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

extern "C" int msd_export_u16(const float* d_x, uint16_t* d_dn, long long numel, void* stream) {
  MSD_REQUIRE(d_x && d_dn && numel > 0 && numel % 4 == 0, "export: bad arguments");
  MSD_REQUIRE(msd_aligned16(d_x) && (reinterpret_cast<uintptr_t>(d_dn) & 7u) == 0, "export: pointers must be 16- / 8-byte aligned");
  MSD_CHECK_CUDA(msd_launch_pdl((export_kernel), dim3(grid_for(numel / 4)), dim3(kBlock), 0, (cudaStream_t)stream, d_x, d_dn, numel / 4));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}
