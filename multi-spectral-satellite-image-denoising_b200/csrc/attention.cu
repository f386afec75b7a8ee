// Multi-head self-attention for the 144-token ViT-B blocks (reference: nn.MultiheadAttention in the fallback
// CustomBlock, src/models/encoder.py:158-169, or timm Attention -> SDPA, encoder.py:126-134):
//   out = softmax(q k^T / sqrt(64)) v per (image, head); no mask, no dropout.
// One CTA per (image, head).  The whole problem (3 x 144 x 64 bf16 = 54 KB) sits in shared memory, the 144x144
// score matrix never leaves registers: each warp owns 16 query rows, keeps S = q k^T (fp32) in registers, does the
// row softmax with quad shuffles, and feeds P (bf16) straight back as the A operand of P v.  Attention is 3% of a
// block's FLOPs; this kernel uses warp-level mma.sync tiles -- the tcgen05 path is reserved for the
// GEMM/conv kernels that carry 97% of the work.
#include "common.cuh"

namespace {

constexpr int HD = 64;       // head dim
constexpr int LDS = HD;      // dense 128-byte rows; the 16-byte chunk index is XOR-swizzled with (row & 7): conflict-free ldmatrix

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(dst)), "l"(src));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int kT>
__global__ void __launch_bounds__(kT * 2) attention_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                           __nv_bfloat16* __restrict__ out, int heads, int problems) {
  static_assert(kT % 16 == 0, "token count must be a multiple of 16");
  constexpr int NT = kT / 8;    // key tiles of 8
  constexpr int KS = kT / 16;   // key steps of 16 for P.V
  constexpr int kThreads = kT * 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // Persistent CTA with two q/k/v buffers: the next (image, head) problem streams in through cp.async while the current
  // one is being computed, so the load latency (the kernel used to sit in the barrier behind it) is off the critical path.
  typedef __nv_bfloat16 (*Tile)[LDS];
  const int E = heads * HD;
  auto issue_load = [&](int problem, int buf) {
    Tile dst = reinterpret_cast<Tile>(smem_raw) + buf * 3 * kT;
    const int bb = problem / heads, hh = problem - bb * heads;
    const __nv_bfloat16* base = qkv + (size_t)bb * kT * 3 * E + hh * HD;
    // q, k, v: 3 * kT rows of 128 B, 16-byte chunks
    for (int i = threadIdx.x; i < 3 * kT * 8; i += kThreads) {
      const int which = i / (kT * 8), rem = i - which * kT * 8, row = rem >> 3, ch = rem & 7;
      cp_async16(&dst[which * kT + row][(ch ^ (row & 7)) * 8], base + (size_t)row * 3 * E + which * E + ch * 8);
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  if ((int)blockIdx.x < problems) issue_load(blockIdx.x, 0);
  int it = 0;
  for (int problem = blockIdx.x; problem < problems; problem += gridDim.x, ++it) {
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();              // this problem's tile has landed AND every warp is done with the buffer about to be refilled
  if (problem + (int)gridDim.x < problems) issue_load(problem + gridDim.x, (it + 1) & 1);
  Tile Qs = reinterpret_cast<Tile>(smem_raw) + (it & 1) * 3 * kT;
  Tile Ks = Qs + kT;
  Tile Vs = Ks + kT;
  const int b = problem / heads, h = problem - b * heads;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = warp * 16;

  uint32_t qa[4][4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) ldmatrix_x4(qa[kk], &Qs[row0 + (lane & 15)][((kk * 2 + (lane >> 4)) ^ (lane & 7)) * 8]);

  float s[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0f;
    uint32_t kb0[4], kb1[4];
    ldmatrix_x4(kb0, &Ks[8 * j + (lane & 7)][((lane >> 3) ^ (lane & 7)) * 8]);
    ldmatrix_x4(kb1, &Ks[8 * j + (lane & 7)][((4 + (lane >> 3)) ^ (lane & 7)) * 8]);
    mma_bf16(s[j], qa[0], kb0[0], kb0[1]);
    mma_bf16(s[j], qa[1], kb0[2], kb0[3]);
    mma_bf16(s[j], qa[2], kb1[0], kb1[1]);
    mma_bf16(s[j], qa[3], kb1[2], kb1[3]);
  }
  // row softmax; rows r0 = lane/4 (regs 0,1) and r0+8 (regs 2,3); a row is spread over the 4 lanes of a quad
  const float sl2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    m0 = fmaxf(m0, fmaxf(s[j][0], s[j][1]));
    m1 = fmaxf(m1, fmaxf(s[j][2], s[j][3]));
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  const float o0 = m0 * sl2, o1 = m1 * sl2;
  float l0 = 0.0f, l1 = 0.0f;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    s[j][0] = exp2f(fmaf(s[j][0], sl2, -o0)); s[j][1] = exp2f(fmaf(s[j][1], sl2, -o0));
    s[j][2] = exp2f(fmaf(s[j][2], sl2, -o1)); s[j][3] = exp2f(fmaf(s[j][3], sl2, -o1));
    l0 += s[j][0] + s[j][1];
    l1 += s[j][2] + s[j][3];
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.0f;
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) {
    uint32_t pa[4];
    pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
    pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
    pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {  // two 8-wide dim tiles per ldmatrix.x4.trans
      uint32_t vb[4];
      ldmatrix_x4_trans(vb, &Vs[16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8][((2 * n2 + (lane >> 4)) ^ (lane & 7)) * 8]);
      mma_bf16(o[2 * n2], pa, vb[0], vb[1]);
      mma_bf16(o[2 * n2 + 1], pa, vb[2], vb[3]);
    }
  }
  const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
  const int r0 = row0 + (lane >> 2), c0 = (lane & 3) * 2;
  __nv_bfloat16* ob = out + (size_t)b * kT * E + h * HD;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    *reinterpret_cast<uint32_t*>(ob + (size_t)r0 * E + n * 8 + c0) = pack_bf16x2(o[n][0] * inv0, o[n][1] * inv0);
    *reinterpret_cast<uint32_t*>(ob + (size_t)(r0 + 8) * E + n * 8 + c0) = pack_bf16x2(o[n][2] * inv1, o[n][3] * inv1);
  }
  }   // problem loop
}

}  // namespace

extern "C" int msd_attention_bf16(const void* d_qkv, void* d_out, int B, int T, int heads, void* stream) {
  MSD_REQUIRE(d_qkv && d_out && B > 0 && heads > 0, "attention: bad arguments");
  MSD_REQUIRE(msd_aligned16(d_qkv) && msd_aligned16(d_out), "attention: pointers must be 16-byte aligned");
  MSD_REQUIRE(T == 144, "attention: only T=144 tokens (192x192 images, patch 16) is built; got T=%d", T);
  constexpr int kT = 144;
  const size_t smem = 2 * 3 * kT * LDS * sizeof(__nv_bfloat16);   // two q/k/v buffers (double-buffered persistent CTAs)
  static PerDeviceOnce attr_set;
  if (attr_set.first()) {
    MSD_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<kT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int problems = B * heads;
  const int grid = problems < 2 * 148 ? problems : 2 * 148;       // two resident CTAs per SM (2 x 108 KB of shared memory)
  attention_kernel<kT><<<(unsigned)grid, kT * 2, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(d_qkv), reinterpret_cast<__nv_bfloat16*>(d_out), heads, problems);
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}
