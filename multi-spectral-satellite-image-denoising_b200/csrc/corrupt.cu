// Corruption simulator (reference: src/data/transforms.py:70-154, AddSensorNoise.forward/_add_striping).
//
// Parity contract: run on a CUDA tensor, the reference draws every random number from PyTorch's CUDA
// generator = Philox4x32-10 with (seed, subsequence = thread index, offset), float4 per call, laid out by
// ATen's distribution_elementwise_grid_stride_kernel (ATen/native/cuda/DistributionTemplates.h:65-90):
//   T = 256 * grid threads, element e is produced by thread (e % T) in its call number (e / (4T)),
//   component ((e / T) % 4).
// This file regenerates exactly that stream inside ONE fused pass (read clean once, write noisy once)
// instead of the reference's three full-size randn_like tensors + five temporaries, and applies the
// reference's arithmetic with every fp32 operation rounded separately (no FMA contraction), so that
// masks AND noisy pixels are bit-identical to the reference on a CUDA tensor.
//
// HBM traffic: 4 B read + 4 B write per element (+2 B if the bf16 patch operand is emitted).
#include "common.cuh"
#include <curand_kernel.h>

namespace {

constexpr int kThreads = 256;

// Philox4x32-10 (Salmon et al., Random123; the generator behind curandStatePhilox4_32_10_t): counter = 128 bits,
// key = 64 bits, 10 rounds of two 32x32->64 multiplies + xors, key bumped by the Weyl constants between rounds.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// The 4 words torch-thread `tid` obtains from its `round`-th curand4 call: curand_init(seed, subsequence = tid,
// offset) puts tid in counter words 2..3 and offset/4 in words 0..1; every call advances the low counter by one.
__device__ __forceinline__ uint4 torch_draw4(uint64_t seed, uint64_t offset, uint32_t tid, uint32_t round) {
  const uint64_t lo = (offset >> 2) + round;
  return philox4x32_10(make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), tid, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
// curand_normal4 = two Box-Muller pairs (curand_normal.h:70-87, cuRAND's own device routine, so the bits match)
__device__ __forceinline__ float4 torch_normal4(uint64_t seed, uint64_t offset, uint32_t tid, uint32_t round) {
  const uint4 x = torch_draw4(seed, offset, tid, round);
  const float2 a = _curand_box_muller(x.x, x.y), b = _curand_box_muller(x.z, x.w);
  return make_float4(a.x, a.y, b.x, b.y);
}
// curand_uniform4 = _curand_uniform on each word: x * 2^-32 + 2^-33 -> (0, 1]
__device__ __forceinline__ float4 torch_uniform4(uint64_t seed, uint64_t offset, uint32_t tid, uint32_t round) {
  const uint4 x = torch_draw4(seed, offset, tid, round);
  return make_float4(_curand_uniform(x.x), _curand_uniform(x.y), _curand_uniform(x.z), _curand_uniform(x.w));
}
// torch's normal_ transform is rand * std + mean with std = 1, mean = 0 (one fused multiply-add)
__device__ __forceinline__ float torch_normal_xf(float r) { return __fmaf_rn(r, 1.0f, 0.0f); }
// torch's uniform_ transform for [0,1): value = rand * 1 + 0, then 1.0 -> 0.0 (curand gives (0,1])
__device__ __forceinline__ float torch_uniform_xf(float r) {
  float v = __fmaf_rn(r, 1.0f, 0.0f);
  return v == 1.0f ? 0.0f : v;
}

template <bool kNormal>
__global__ void __launch_bounds__(kThreads) philox_fill_kernel(float* __restrict__ out, int64_t numel, uint64_t seed,
                                                               uint64_t offset, uint32_t T) {
  msd_pdl_enter();
  // work item = (round = blockIdx.y, tid); component ii goes to element round*4T + ii*T + tid  (T % 256 == 0)
  const uint32_t round = blockIdx.y;
  const uint32_t tid = blockIdx.x * kThreads + threadIdx.x;
  float4 r = kNormal ? torch_normal4(seed, offset, tid, round) : torch_uniform4(seed, offset, tid, round);
  const float v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int ii = 0; ii < 4; ++ii) {
    int64_t e = (int64_t)round * 4 * T + (int64_t)ii * T + tid;
    if (e < numel) out[e] = kNormal ? torch_normal_xf(v[ii]) : torch_uniform_xf(v[ii]);
  }
}

// dead-band mask: one torch.rand(C) call per sample (transforms.py:101-106); element c of sample b is the
// .x component of thread c's first draw at offset off0 + 4*b; mask = u < (float)p.
__global__ void dead_mask_kernel(uint8_t* __restrict__ mask, int B, int C, float p, uint64_t seed, uint64_t off0,
                                 uint32_t T) {
  msd_pdl_enter();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  int b = i / C, c = i - b * C;
  // torch.rand(C): grid = ceil(C/256) blocks -> T = 256*grid; element c -> thread c % T, round c / (4T), comp (c/T)%4
  uint32_t round = (uint32_t)c / (4u * T);
  uint32_t comp = ((uint32_t)c / T) & 3u;
  uint32_t tid = (uint32_t)c % T;
  float4 r = torch_uniform4(seed, off0 + 4ull * (uint64_t)b * ((uint64_t)(C - 1) / (4ull * T) + 1ull), tid, round);
  float u = comp == 0 ? r.x : comp == 1 ? r.y : comp == 2 ? r.z : r.w;
  u = torch_uniform_xf(u);
  mask[i] = (u < p) ? 1 : 0;
}

struct CorruptParams {
  const float* clean;
  float* noisy;
  const float* thermal_w;   // [C]
  const uint8_t* dead;      // [B*C] or null
  const float* stripe;      // [B*C*W] normals or null
  __nv_bfloat16* patch;     // optional bf16 patch operand
  uint32_t numel;
  uint32_t T;               // torch threads of the full-size draws
  uint32_t rounds;
  uint32_t hw, dplane, dpix; // H*W and T = dplane * hw + dpix: element e + T sits dplane planes and dpix pixels further
  FastDiv div_hw, div_c, div_w;
  int C, H, W;
  float sg, ss, st;
  uint64_t seed, off_g, off_s, off_t;
  int gridw;                // patches per row (patch size 16)
};

// kFast: every noise term on, dead-band mask present, no striping -- the configuration of every reference config
// (configs/*.yaml); the flags are compile-time there, so the per-element code carries no predicates.  kPatch: also emit the
// bf16 patch-embedding operand.  The generic instantiation <false, false> reads the flags at run time.
// Instruction budget (the kernel is issue-bound, not HBM-bound): 3 Philox4x32-10 blocks (~130) + 6 Box-Muller pairs with
// cuRAND's full-precision logf / sqrtf (~330, fixed by bit-exactness with torch) per 4 elements, + ~45 per element of index
// arithmetic, fp32 ops and stores.
template <bool kFast, bool kPatch>
__global__ void __launch_bounds__(kThreads) corrupt_kernel(const CorruptParams p) {
  msd_pdl_enter();
  const bool has_g = kFast || p.sg > 0.0f, has_s = kFast || p.ss > 0.0f, has_t = kFast || p.st > 0.0f;
  const bool has_dead = kFast || p.dead != nullptr, has_stripe = !kFast && p.stripe != nullptr;
  const bool has_patch = kPatch || (!kFast && p.patch != nullptr);
  // (the C thermal weights are read through the read-only path per element: staging them in shared memory put a block-wide
  //  barrier in front of a 4-element-per-thread kernel -- 17 % of its stall samples)
  const uint32_t round = blockIdx.y;                               // torch-thread call number
  const uint32_t tid = blockIdx.x * kThreads + threadIdx.x;       // torch-thread index (T % 256 == 0)
  const uint32_t e0 = round * 4u * p.T + tid;                      // < 2^31 + 4T: fits 32 bits

  // element ii = e0 + ii * T: (plane, pixel) of the first one by division, the others by adding T's (planes, pixels)
  uint32_t plane[4], pix[4];
  fd_divmod(e0, p.div_hw, plane[0], pix[0]);
#pragma unroll
  for (int ii = 1; ii < 4; ++ii) {
    uint32_t q = pix[ii - 1] + p.dpix, pl = plane[ii - 1] + p.dplane;
    if (q >= p.hw) { q -= p.hw; ++pl; }
    pix[ii] = q; plane[ii] = pl;
  }
  float x[4];
  bool ok[4], dead[4];
#pragma unroll
  for (int ii = 0; ii < 4; ++ii) {
    const uint32_t e = e0 + (uint32_t)ii * p.T;
    ok[ii] = e < p.numel;
    x[ii] = ok[ii] ? __ldg(p.clean + e) : 0.0f;
    dead[ii] = has_dead && ok[ii] && __ldg(p.dead + plane[ii]) != 0;
  }
  float n1[4] = {0, 0, 0, 0}, n2[4] = {0, 0, 0, 0}, n3[4] = {0, 0, 0, 0};
  if (has_g) {
    float4 r = torch_normal4(p.seed, p.off_g, tid, round);
    n1[0] = r.x; n1[1] = r.y; n1[2] = r.z; n1[3] = r.w;
  }
  if (has_s) {
    float4 r = torch_normal4(p.seed, p.off_s, tid, round);
    n2[0] = r.x; n2[1] = r.y; n2[2] = r.z; n2[3] = r.w;
  }
  if (has_t) {
    float4 r = torch_normal4(p.seed, p.off_t, tid, round);
    n3[0] = r.x; n3[1] = r.y; n3[2] = r.z; n3[3] = r.w;
  }
#pragma unroll
  for (int ii = 0; ii < 4; ++ii) {
    if (!ok[ii]) continue;
    const uint32_t e = e0 + (uint32_t)ii * p.T;
    uint32_t b, c;                  // plane = b*C + c
    fd_divmod(plane[ii], p.div_c, b, c);
    float v = x[ii];
    // 1. additive Gaussian: corrupted + randn*sigma            (transforms.py:90-92)
    if (has_g) v = __fadd_rn(v, __fmul_rn(torch_normal_xf(n1[ii]), p.sg));
    // 2. speckle: corrupted * (1.0 + randn*sigma)               (transforms.py:95-97)
    if (has_s) v = __fmul_rn(v, __fadd_rn(__fmul_rn(torch_normal_xf(n2[ii]), p.ss), 1.0f));
    // 3. dead bands zeroed BEFORE thermal noise                 (transforms.py:100-106)
    if (dead[ii]) v = 0.0f;
    // 4. thermal: corrupted + (randn*scale)*w_c                 (transforms.py:110-121)
    if (has_t) v = __fadd_rn(v, __fmul_rn(__fmul_rn(torch_normal_xf(n3[ii]), p.st), __ldg(p.thermal_w + c)));
    uint32_t yy, xx;
    fd_divmod(pix[ii], p.div_w, yy, xx);
    // 5. striping: x + randn(B,C,1,W)*0.02                      (transforms.py:136-154)
    if (has_stripe) v = __fadd_rn(v, __fmul_rn(__ldg(p.stripe + plane[ii] * (uint32_t)p.W + xx), 0.02f));
    // clamp(-3, 3), NaN-propagating like torch.clamp            (transforms.py:128)
    v = (v != v) ? v : fminf(fmaxf(v, -3.0f), 3.0f);
    p.noisy[e] = v;
    if (has_patch) {
      // A[(b*gh + py)*gw + px][c*P*P + ky*P + kx]  (encoder.py:48-50), patch size 16: offset = b * (C*H*W) [= b*gw*gw * C*256]
      // + (py*gw + px) * C*256 + c*256 + ky*16 + kx -- all of it < numel < 2^31, 32-bit arithmetic
      const uint32_t py = yy >> 4, ky = yy & 15u, px = xx >> 4, kx = xx & 15u;
      const uint32_t off = ((b * (uint32_t)p.gridw + py) * (uint32_t)p.gridw + px) * ((uint32_t)p.C << 8) + (c << 8) + (ky << 4) + kx;
      p.patch[off] = __float2bfloat16_rn(v);
    }
  }
}

}  // namespace

extern "C" uint64_t msd_philox_offset_increment(int64_t numel, uint32_t torch_grid) {
  if (numel <= 0 || torch_grid == 0) return 0;
  // ((numel-1)/(256*grid*4)+1)*4   (DistributionTemplates.h:60)
  return (uint64_t)(((numel - 1) / (256ll * torch_grid * 4ll) + 1) * 4ll);
}

template <bool kNormal>
static int philox_fill(float* d_out, int64_t numel, uint64_t seed, uint64_t offset, uint32_t torch_grid, void* stream) {
  MSD_REQUIRE(d_out != nullptr && numel > 0 && torch_grid > 0, "philox_fill: bad arguments");
  MSD_REQUIRE(numel < (1ll << 31), "philox_fill: numel must be < 2^31");
  const uint32_t T = 256u * torch_grid;
  const int64_t rounds = (numel - 1) / (4ll * T) + 1;
  MSD_REQUIRE(rounds <= 65535, "philox_fill: too many rounds (%lld)", (long long)rounds);
  MSD_CHECK_CUDA(msd_launch_pdl((philox_fill_kernel<kNormal>), dim3(dim3(T / kThreads, (unsigned)rounds)), dim3(kThreads), 0, (cudaStream_t)stream, d_out, numel, seed, offset, T));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}
extern "C" int msd_philox_normal_f32(float* d_out, int64_t numel, uint64_t seed, uint64_t offset, uint32_t torch_grid,
                                     void* stream) {
  return philox_fill<true>(d_out, numel, seed, offset, torch_grid, stream);
}
extern "C" int msd_philox_uniform_f32(float* d_out, int64_t numel, uint64_t seed, uint64_t offset,
                                      uint32_t torch_grid, void* stream) {
  return philox_fill<false>(d_out, numel, seed, offset, torch_grid, stream);
}

extern "C" int msd_corrupt_f32(const float* d_clean, float* d_noisy, const float* d_thermal_w, uint8_t* d_dead_mask,
                               float* d_stripe_ws, void* d_patch_bf16, int B, int C, int H, int W, float sg, float ss,
                               float p, float st, int stripe, uint64_t seed, uint64_t offset, uint32_t torch_grid_full,
                               uint32_t torch_grid_mask, uint32_t torch_grid_stripe, uint64_t* offset_consumed,
                               void* stream) {
  MSD_REQUIRE(d_clean && d_noisy, "corrupt: null tensor");
  MSD_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "corrupt: bad shape [%d,%d,%d,%d]", B, C, H, W);
  const int64_t numel = (int64_t)B * C * H * W;
  MSD_REQUIRE(numel < (1ll << 31), "corrupt: B*C*H*W must be < 2^31 (got %lld)", (long long)numel);
  MSD_REQUIRE(torch_grid_full > 0, "corrupt: torch_grid_full must be > 0");
  MSD_REQUIRE(st <= 0.0f || d_thermal_w != nullptr, "corrupt: thermal weights required when thermal_scale > 0");
  MSD_REQUIRE(p <= 0.0f || (d_dead_mask != nullptr && torch_grid_mask > 0), "corrupt: dead mask buffer required when p > 0");
  MSD_REQUIRE(!stripe || (d_stripe_ws != nullptr && torch_grid_stripe > 0), "corrupt: stripe scratch required");
  if (d_patch_bf16) MSD_REQUIRE(H == W && H % 16 == 0, "corrupt: patch operand needs square images with S %% 16 == 0");
  cudaStream_t s = (cudaStream_t)stream;

  // offsets in the reference's draw order
  uint64_t off = offset;
  const uint64_t inc_full = msd_philox_offset_increment(numel, torch_grid_full);
  uint64_t off_g = 0, off_s = 0, off_m = 0, off_t = 0, off_stripe = 0;
  if (sg > 0.0f) { off_g = off; off += inc_full; }
  if (ss > 0.0f) { off_s = off; off += inc_full; }
  if (p > 0.0f) { off_m = off; off += (uint64_t)B * msd_philox_offset_increment(C, torch_grid_mask); }
  if (st > 0.0f) { off_t = off; off += inc_full; }
  if (stripe) { off_stripe = off; off += msd_philox_offset_increment((int64_t)B * C * W, torch_grid_stripe); }
  if (offset_consumed) *offset_consumed = off - offset;

  if (p > 0.0f) {
    int n = B * C;
    MSD_CHECK_CUDA(msd_launch_pdl((dead_mask_kernel), dim3((n + 127) / 128), dim3(128), 0, s, d_dead_mask, B, C, p, seed, off_m, 256u * torch_grid_mask));
    MSD_LAUNCH_CHECK();
  }
  if (stripe) {
    int rc = msd_philox_normal_f32(d_stripe_ws, (int64_t)B * C * W, seed, off_stripe, torch_grid_stripe, stream);
    if (rc) return rc;
  }
  CorruptParams cp;
  cp.clean = d_clean; cp.noisy = d_noisy; cp.thermal_w = d_thermal_w;
  cp.dead = p > 0.0f ? d_dead_mask : nullptr;
  cp.stripe = stripe ? d_stripe_ws : nullptr;
  cp.patch = reinterpret_cast<__nv_bfloat16*>(d_patch_bf16);
  cp.numel = (uint32_t)numel;
  cp.T = 256u * torch_grid_full;
  cp.rounds = (uint32_t)((numel - 1) / (4ll * cp.T) + 1);
  MSD_REQUIRE((uint64_t)cp.rounds * 4ull * cp.T < (1ull << 32), "corrupt: element index range exceeds 32 bits");
  cp.hw = (uint32_t)(H * W); cp.dplane = cp.T / cp.hw; cp.dpix = cp.T % cp.hw;
  cp.div_hw = make_fastdiv((uint32_t)(H * W));
  cp.div_c = make_fastdiv((uint32_t)C);
  cp.div_w = make_fastdiv((uint32_t)W);
  cp.C = C; cp.H = H; cp.W = W;
  cp.sg = sg; cp.ss = ss; cp.st = st;
  cp.seed = seed; cp.off_g = off_g; cp.off_s = off_s; cp.off_t = off_t;
  cp.gridw = W / 16;
  MSD_REQUIRE(cp.rounds <= 65535, "corrupt: too many rounds");
  const dim3 grid(cp.T / kThreads, cp.rounds);
  const bool fast = sg > 0.0f && ss > 0.0f && st > 0.0f && p > 0.0f && !stripe;
  if (fast && cp.patch != nullptr) MSD_CHECK_CUDA(msd_launch_pdl((corrupt_kernel<true, true>), dim3(grid), dim3(kThreads), 0, s, cp));
  else if (fast) MSD_CHECK_CUDA(msd_launch_pdl((corrupt_kernel<true, false>), dim3(grid), dim3(kThreads), 0, s, cp));
  else MSD_CHECK_CUDA(msd_launch_pdl((corrupt_kernel<false, false>), dim3(grid), dim3(kThreads), 0, s, cp));
  MSD_LAUNCH_CHECK();
  return MSD_OK;
}
