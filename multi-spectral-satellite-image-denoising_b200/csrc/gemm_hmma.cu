// Legacy tensor-core path: the same Problem (gemm.cuh) computed with warp-level mma.sync.m16n8k16 tiles fed by
// cp.async.  This is NOT the product path -- it is the "recompiled Ampere-style kernel" baseline the tcgen05
// kernel is measured against (msd_model_desc.gemm_backend = 1) and an independent implementation the parity
// tests cross-check the tcgen05 results with.
#include "gemm.cuh"

namespace gemm {
namespace {

constexpr int kThreads = 256;         // 8 warps, each 16 rows x BN
constexpr int LDSM = BLOCK_K + 8;     // padded smem row: 144 B, conflict-free ldmatrix

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16_zfill(void* dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(dst)), "l"(src), "r"(n));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int BN>
__global__ void __launch_bounds__(kThreads) gemm_hmma_kernel(const Problem p, int tiles_n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef __nv_bfloat16 (*TileA)[BLOCK_M][LDSM];
  typedef __nv_bfloat16 (*TileB)[BN][LDSM];
  TileA As = reinterpret_cast<TileA>(smem_raw);
  TileB Bs = reinterpret_cast<TileB>(smem_raw + 2 * BLOCK_M * LDSM * sizeof(__nv_bfloat16));

  const int m0 = (blockIdx.x / tiles_n) * BLOCK_M, n0 = (blockIdx.x % tiles_n) * BN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ldB = (p.w_ld_kb ? p.w_ld_kb : p.num_kb) * BLOCK_K;

  auto load_stage = [&](int kb, int buf) {
    const int tap = kb / p.kb_per_tap, kc = kb - tap * p.kb_per_tap;
    const int row_shift = p.tap_off[tap];
    const __nv_bfloat16* Abase = (tap == p.skip_tap) ? p.skip : p.A;   // skip connection rides along as an identity tap
    for (int i = threadIdx.x; i < BLOCK_M * 8; i += kThreads) {
      const int r = i >> 3, ch = i & 7;
      const long long grow = (long long)m0 + r + row_shift;
      const int kcol = kc * BLOCK_K + ch * 8;
      const bool valid = grow >= 0 && grow < p.a_rows && kcol < p.kvalid;
      const __nv_bfloat16* src = valid ? Abase + (size_t)grow * p.ldA + kcol : Abase;
      cp_async16_zfill(&As[buf][r][ch * 8], src, valid);
    }
    for (int i = threadIdx.x; i < BN * 8; i += kThreads) {
      const int r = i >> 3, ch = i & 7;
      const bool valid = (n0 + r) < p.N;
      const __nv_bfloat16* src = valid ? p.Wt + (size_t)(n0 + r) * ldB + kb * BLOCK_K + ch * 8 : p.Wt;
      cp_async16_zfill(&Bs[buf][r][ch * 8], src, valid);
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };

  float acc[BN / 8][4];
#pragma unroll
  for (int j = 0; j < BN / 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.0f;

  load_stage(0, 0);
  for (int kb = 0; kb < p.num_kb; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < p.num_kb) {
      load_stage(kb + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;\n" ::);
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::);
    }
    __syncthreads();
    uint32_t a[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) ldmatrix_x4(a[kk], &As[buf][warp * 16 + (lane & 15)][kk * 16 + (lane >> 4) * 8]);
#pragma unroll
    for (int j = 0; j < BN / 8; ++j) {
      uint32_t b0[4], b1[4];
      ldmatrix_x4(b0, &Bs[buf][8 * j + (lane & 7)][(lane >> 3) * 8]);
      ldmatrix_x4(b1, &Bs[buf][8 * j + (lane & 7)][32 + (lane >> 3) * 8]);
      mma_bf16(acc[j], a[0], b0[0], b0[1]);
      mma_bf16(acc[j], a[1], b0[2], b0[3]);
      mma_bf16(acc[j], a[2], b1[0], b1[1]);
      mma_bf16(acc[j], a[3], b1[2], b1[3]);
    }
    __syncthreads();
  }
  const int r0 = m0 + warp * 16 + (lane >> 2), c0 = (lane & 3) * 2;
#pragma unroll
  for (int j = 0; j < BN / 8; ++j) {
    float v0[2] = {acc[j][0], acc[j][1]};
    float v1[2] = {acc[j][2], acc[j][3]};
    epilogue_store<2>(p, r0, n0 + 8 * j + c0, v0);
    epilogue_store<2>(p, r0 + 8, n0 + 8 * j + c0, v1);
  }
}

template <int BN>
int launch_bn(const Problem& p, cudaStream_t s) {
  const size_t smem = 2 * (size_t)(BLOCK_M + BN) * LDSM * sizeof(__nv_bfloat16);
  static PerDeviceOnce attr_set;
  if (attr_set.first()) {
    MSD_CHECK_CUDA(cudaFuncSetAttribute(gemm_hmma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int tiles_m = (p.M + BLOCK_M - 1) / BLOCK_M, tiles_n = p.N / BN;
  gemm_hmma_kernel<BN><<<(unsigned)(tiles_m * tiles_n), kThreads, smem, s>>>(p, tiles_n);
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}

}  // namespace

int launch_hmma(const Problem& p, cudaStream_t s) {
  MSD_REQUIRE(p.N % 16 == 0, "gemm_hmma: N=%d must be a multiple of 16", p.N);
  MSD_REQUIRE(p.ldA % 8 == 0 && p.kvalid % 8 == 0, "gemm_hmma: ldA/kvalid must be multiples of 8");
  if (p.N % 128 == 0) return launch_bn<128>(p, s);
  if (p.N % 96 == 0) return launch_bn<96>(p, s);
  if (p.N % 64 == 0) return launch_bn<64>(p, s);
  if (p.N % 48 == 0) return launch_bn<48>(p, s);
  if (p.N % 32 == 0) return launch_bn<32>(p, s);
  return launch_bn<16>(p, s);
}

}  // namespace gemm
