// Backward of the 144-token multi-head self-attention (forward: attention.cu; reference: autograd through
// nn.MultiheadAttention / timm Attention, src/models/encoder.py:126-134, 158-169):
//   P = softmax(q k^T / 8);  dV = P^T dO;  dP = dO V^T;  dS = P * (dP - rowsum(P * dP)) / 8;  dQ = dS K;  dK = dS^T Q
// One CTA per (image, head), nine warps of 16 query rows.  q, k, v, dO sit in shared memory; each warp recomputes its
// rows of S and P in registers (nothing but q, k, v was saved by the forward), forms dP and dS there, writes its dQ rows,
// and parks P and dS (bf16) in shared memory so that, after one barrier, each warp can contract over ALL queries for its
// 16 key rows of dV and dK (the transposed operands come straight out of ldmatrix.trans).  mma.sync m16n8k16 tiles.
#include "common.cuh"

namespace {

constexpr int HD = 64;
constexpr int LDS = HD + 8;   // padded q/k/v/dO row
constexpr int kT = 144;
constexpr int LDP = kT + 8;   // padded P / dS row
constexpr int NT = kT / 8, KS = kT / 16;
constexpr int kThreads = kT * 2;

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

__global__ void __launch_bounds__(kThreads, 1) attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                    const __nv_bfloat16* __restrict__ dout,
                                                                    __nv_bfloat16* __restrict__ dqkv, int heads) {
  msd_pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16(*Qs)[LDS] = reinterpret_cast<__nv_bfloat16(*)[LDS]>(smem_raw);
  __nv_bfloat16(*Ks)[LDS] = Qs + kT;
  __nv_bfloat16(*Vs)[LDS] = Ks + kT;
  __nv_bfloat16(*Os)[LDS] = Vs + kT;                                                   // dO
  __nv_bfloat16(*Ps)[LDP] = reinterpret_cast<__nv_bfloat16(*)[LDP]>(Os + kT);          // P   [query][key]
  __nv_bfloat16(*Ss)[LDP] = Ps + kT;                                                   // dS  [query][key]

  const int b = blockIdx.x / heads, h = blockIdx.x - b * heads;
  const int E = heads * HD;
  const __nv_bfloat16* base = qkv + (size_t)b * kT * 3 * E + h * HD;
  const __nv_bfloat16* dob = dout + (size_t)b * kT * E + h * HD;
  for (int i = threadIdx.x; i < 4 * kT * 8; i += kThreads) {
    const int which = i / (kT * 8), rem = i - which * kT * 8, row = rem >> 3, ch = rem & 7;
    const __nv_bfloat16* src = which < 3 ? base + (size_t)row * 3 * E + which * E + ch * 8 : dob + (size_t)row * E + ch * 8;
    cp_async16(&Qs[which * kT + row][ch * 8], src);
  }
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = warp * 16;
  const int r0 = row0 + (lane >> 2), c0 = (lane & 3) * 2;
  __nv_bfloat16* dq = dqkv + (size_t)b * kT * 3 * E + h * HD;

  {
    uint32_t qa[4][4], oa[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      ldmatrix_x4(qa[kk], &Qs[row0 + (lane & 15)][kk * 16 + (lane >> 4) * 8]);
      ldmatrix_x4(oa[kk], &Os[row0 + (lane & 15)][kk * 16 + (lane >> 4) * 8]);
    }
    // S = q k^T and dP = dO v^T for this warp's 16 query rows, all 144 keys
    float s[NT][4], dp[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.0f;
      dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.0f;
      uint32_t b0[4], b1[4];
      ldmatrix_x4(b0, &Ks[8 * j + (lane & 7)][(lane >> 3) * 8]);
      ldmatrix_x4(b1, &Ks[8 * j + (lane & 7)][32 + (lane >> 3) * 8]);
      mma_bf16(s[j], qa[0], b0[0], b0[1]);
      mma_bf16(s[j], qa[1], b0[2], b0[3]);
      mma_bf16(s[j], qa[2], b1[0], b1[1]);
      mma_bf16(s[j], qa[3], b1[2], b1[3]);
      ldmatrix_x4(b0, &Vs[8 * j + (lane & 7)][(lane >> 3) * 8]);
      ldmatrix_x4(b1, &Vs[8 * j + (lane & 7)][32 + (lane >> 3) * 8]);
      mma_bf16(dp[j], oa[0], b0[0], b0[1]);
      mma_bf16(dp[j], oa[1], b0[2], b0[3]);
      mma_bf16(dp[j], oa[2], b1[0], b1[1]);
      mma_bf16(dp[j], oa[3], b1[2], b1[3]);
    }
    // softmax rows (regs 0,1: row r0; regs 2,3: row r0+8), exactly as the forward
    const float sl2 = 0.125f * 1.4426950408889634f;
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
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    float d0 = 0.0f, d1 = 0.0f;   // rowsum(P * dP)
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      s[j][0] *= inv0; s[j][1] *= inv0; s[j][2] *= inv1; s[j][3] *= inv1;
      d0 += s[j][0] * dp[j][0] + s[j][1] * dp[j][1];
      d1 += s[j][2] * dp[j][2] + s[j][3] * dp[j][3];
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    // P and dS = P (dP - D) / 8 -> bf16 in shared memory [query][key]; dS stays in dp for the dQ product
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      dp[j][0] = s[j][0] * (dp[j][0] - d0) * 0.125f; dp[j][1] = s[j][1] * (dp[j][1] - d0) * 0.125f;
      dp[j][2] = s[j][2] * (dp[j][2] - d1) * 0.125f; dp[j][3] = s[j][3] * (dp[j][3] - d1) * 0.125f;
      *reinterpret_cast<uint32_t*>(&Ps[r0][8 * j + c0]) = pack_bf16x2(s[j][0], s[j][1]);
      *reinterpret_cast<uint32_t*>(&Ps[r0 + 8][8 * j + c0]) = pack_bf16x2(s[j][2], s[j][3]);
      *reinterpret_cast<uint32_t*>(&Ss[r0][8 * j + c0]) = pack_bf16x2(dp[j][0], dp[j][1]);
      *reinterpret_cast<uint32_t*>(&Ss[r0 + 8][8 * j + c0]) = pack_bf16x2(dp[j][2], dp[j][3]);
    }
    // dQ = dS K  (contraction over keys; K read as [key][dim] through ldmatrix.trans, like V in the forward's P V)
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.0f;
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(dp[2 * kk][0], dp[2 * kk][1]);
      pa[1] = pack_bf16x2(dp[2 * kk][2], dp[2 * kk][3]);
      pa[2] = pack_bf16x2(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
      pa[3] = pack_bf16x2(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
      for (int n2 = 0; n2 < 4; ++n2) {
        uint32_t vb[4];
        ldmatrix_x4_trans(vb, &Ks[16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8][16 * n2 + (lane >> 4) * 8]);
        mma_bf16(o[2 * n2], pa, vb[0], vb[1]);
        mma_bf16(o[2 * n2 + 1], pa, vb[2], vb[3]);
      }
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      *reinterpret_cast<uint32_t*>(dq + (size_t)r0 * 3 * E + n * 8 + c0) = pack_bf16x2(o[n][0], o[n][1]);
      *reinterpret_cast<uint32_t*>(dq + (size_t)(r0 + 8) * 3 * E + n * 8 + c0) = pack_bf16x2(o[n][2], o[n][3]);
    }
  }
  __syncthreads();
  // dV = P^T dO and dK = dS^T Q for this warp's 16 KEY rows: A = transposed [query][key] tiles (ldmatrix.trans),
  // B = dO / Q as [query][dim] (ldmatrix.trans), contraction over the 144 queries
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    __nv_bfloat16(*As)[LDP] = which == 0 ? Ps : Ss;
    __nv_bfloat16(*Bs)[LDS] = which == 0 ? Os : Qs;
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.0f;
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      uint32_t pa[4];
      // A[m = key, k = query] = As[query][key]: matrices (m 0-7,k 0-7), (m 8-15,k 0-7), (m 0-7,k 8-15), (m 8-15,k 8-15)
      ldmatrix_x4_trans(pa, &As[16 * kk + (lane & 7) + ((lane >> 4) & 1) * 8][row0 + ((lane >> 3) & 1) * 8]);
#pragma unroll
      for (int n2 = 0; n2 < 4; ++n2) {
        uint32_t vb[4];
        ldmatrix_x4_trans(vb, &Bs[16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8][16 * n2 + (lane >> 4) * 8]);
        mma_bf16(o[2 * n2], pa, vb[0], vb[1]);
        mma_bf16(o[2 * n2 + 1], pa, vb[2], vb[3]);
      }
    }
    __nv_bfloat16* dst = dq + (which == 0 ? 2 * E : E);      // dV -> third E columns, dK -> second
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      *reinterpret_cast<uint32_t*>(dst + (size_t)r0 * 3 * E + n * 8 + c0) = pack_bf16x2(o[n][0], o[n][1]);
      *reinterpret_cast<uint32_t*>(dst + (size_t)(r0 + 8) * 3 * E + n * 8 + c0) = pack_bf16x2(o[n][2], o[n][3]);
    }
  }
}

}  // namespace

// qkv bf16 [B*T, 3*heads*64], d_dout bf16 [B*T, heads*64] (grad of the attention output) -> d_dqkv bf16 [B*T, 3*heads*64]
extern "C" int msd_attention_bwd_bf16(const void* d_qkv, const void* d_dout, void* d_dqkv, int B, int T, int heads, void* stream) {
  MSD_REQUIRE(d_qkv && d_dout && d_dqkv && B > 0 && heads > 0, "attention_bwd: bad arguments");
  MSD_REQUIRE(msd_aligned16(d_qkv) && msd_aligned16(d_dout) && msd_aligned16(d_dqkv), "attention_bwd: pointers must be 16-byte aligned");
  MSD_REQUIRE(T == kT, "attention_bwd: only T=144 tokens is built; got T=%d", T);
  const size_t smem = (size_t)4 * kT * LDS * sizeof(__nv_bfloat16) + (size_t)2 * kT * LDP * sizeof(__nv_bfloat16);
  static PerDeviceOnce attr_set;
  if (attr_set.first()) {
    MSD_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  MSD_CHECK_CUDA(msd_launch_pdl((attention_bwd_kernel), dim3((unsigned)(B * heads)), dim3(kThreads), smem, (cudaStream_t)stream, 
      reinterpret_cast<const __nv_bfloat16*>(d_qkv), reinterpret_cast<const __nv_bfloat16*>(d_dout),
      reinterpret_cast<__nv_bfloat16*>(d_dqkv), heads));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}
