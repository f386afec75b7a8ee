// Shared host/device helpers for libmsd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/msd.h"

// ---------------------------------------------------------------------------------------------------
// error plumbing: nothing throws across the C ABI; text is kept per thread
// ---------------------------------------------------------------------------------------------------
void msd_set_error(const char* fmt, ...);

#define MSD_CHECK_CUDA(expr)                                                                          \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) {                                                                          \
      msd_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));            \
      return MSD_ERR_CUDA;                                                                            \
    }                                                                                                 \
  } while (0)

#define MSD_REQUIRE(cond, ...)                                                                        \
  do {                                                                                                \
    if (!(cond)) {                                                                                    \
      msd_set_error(__VA_ARGS__);                                                                     \
      return MSD_ERR_INVALID;                                                                         \
    }                                                                                                 \
  } while (0)

#define MSD_LAUNCH_CHECK() MSD_CHECK_CUDA(cudaGetLastError())

static inline bool msd_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// cudaFuncSetAttribute belongs to a function ON A DEVICE: a process that drives several GPUs has to set it once per device,
// not once per process.  `static PerDeviceOnce guard; if (guard.first()) cudaFuncSetAttribute(...)`.
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

// ---------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL) for the chain of persistent tensor-core kernels of a forward / backward pass: a kernel
// launched through msd_launch_pdl may become resident on an SM as soon as the previous kernel's CTA there has exited, run its
// prologue (mbarrier init, TMEM allocation, tensor-map prefetch) under the previous kernel's tail, and only then wait for the
// previous grid to complete with its memory visible (msd_pdl_enter(), called by every thread before the kernel's first access to
// global memory -- reads AND writes, so no buffer the predecessor still uses is touched early).  msd_pdl_enter() first releases
// this kernel's own dependents.  MSD_NO_PDL=1 launches everything fully serialised (A/B runs).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void msd_pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
}
bool msd_pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t msd_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = msd_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------------
// exact unsigned division by a runtime constant (n < 2^31): q = (umulhi(n, m) + n) >> s
// ---------------------------------------------------------------------------------------------------
struct FastDiv {
  uint32_t d, m, s;
};
static inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;
  f.s = s;
  f.m = (uint32_t)((((1ull << 32) * ((1ull << s) - d)) / d) + 1);
  return f;
}
__device__ __forceinline__ uint32_t fd_div(uint32_t n, const FastDiv& f) {
  return (__umulhi(n, f.m) + n) >> f.s;
}
__device__ __forceinline__ void fd_divmod(uint32_t n, const FastDiv& f, uint32_t& q, uint32_t& r) {
  q = fd_div(n, f);
  r = n - q * f.d;
}

// ---------------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// GELU(erf), the nn.GELU() default the reference uses (blocks.py:70, encoder.py:164):  0.5 x (1 + erf(x/sqrt2)).
// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the bf16 ulp the result is rounded to):
// 5 FMAs + one MUFU.RCP + one MUFU.EX2 instead of erff()'s branchy ~40 instructions -- this sits in GEMM epilogues.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = x * 0.70710678118654752f;
  const float az = fabsf(z);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, az, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = __expf(-az * az);
  const float erf_abs = fmaf(-poly * t, e, 1.0f);
  const float hx = 0.5f * x;
  return fmaf(hx, copysignf(erf_abs, z), hx);
}

// GELU of four values that are about to be rounded to bf16, evaluated two-at-a-time in half2:
//   0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))  with MUFU.TANH.F16x2  (7 instructions per pair instead of ~30).
// |tanh-form - erf-form| <= 3e-4 and the fp16 evaluation error (~1e-3 relative) are both below the bf16 rounding
// (4e-3 relative) applied right after; used only by the bf16-output GEMM epilogues.
__device__ __forceinline__ __half2 gelu_h2(__half2 x) {
  const __half2 c0 = __float2half2_rn(0.7978845608f), c1 = __float2half2_rn(0.0356774081f), half = __float2half2_rn(0.5f);
  const __half2 u = __hmul2(x, __hfma2(__hmul2(x, x), c1, c0));
  uint32_t ui = *reinterpret_cast<const uint32_t*>(&u), ti;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(ti) : "r"(ui));
  const __half2 t = *reinterpret_cast<const __half2*>(&ti);
  const __half2 hx = __hmul2(x, half);
  return __hfma2(hx, t, hx);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint2 gelu4_to_bf16(float4 v) {
  const float2 a = __half22float2(gelu_h2(__floats2half2_rn(v.x, v.y)));
  const float2 b = __half22float2(gelu_h2(__floats2half2_rn(v.z, v.w)));
  return make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(b.x, b.y));
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
// exact-erf GELU and its derivative (nn.GELU() default): gelu'(z) = Phi(z) + z phi(z)
// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the bf16 rounding of every consumer; common.cuh gelu_erf):
// Phi(z) and phi(z) share one exp(-z^2/2)
__device__ __forceinline__ float gelu_fwd(float z) { return gelu_erf(z); }
__device__ __forceinline__ float gelu_grad(float z) {
  const float az = fabsf(z) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, az, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = __expf(-az * az);                       // = exp(-z^2 / 2)
  const float erf_abs = fmaf(-poly * t, e, 1.0f);
  return fmaf(0.5f, copysignf(erf_abs, z), 0.5f) + z * 0.3989422804014327f * e;
}
// dy * gelu'(z) for two elements at a time in half2 (tanh form, MUFU.TANH.F16x2):
//   u = z (c0 + c1 z^2), t = tanh(u), gelu'(z) = 0.5 (1 + t) + 0.5 z (1 - t^2) (c0 + 3 c1 z^2)
// |tanh form - erf form| <= 1e-3 on the derivative and fp16 evaluation (1e-3 relative) are both below the bf16 rounding of
// the product that follows; z is clamped to +-8 (the derivative is 0 / 1 beyond, and fp16 z^2 must not overflow).
__device__ __forceinline__ float2 gelu_grad2(float z0, float z1) {
  const __half2 z = __floats2half2_rn(fminf(fmaxf(z0, -8.f), 8.f), fminf(fmaxf(z1, -8.f), 8.f));
  const __half2 c0 = __float2half2_rn(0.7978845608f), c1 = __float2half2_rn(0.0356774081f), c3 = __float2half2_rn(0.1070322243f);
  const __half2 half = __float2half2_rn(0.5f), one = __float2half2_rn(1.0f);
  const __half2 z2 = __hmul2(z, z);
  const __half2 u = __hmul2(z, __hfma2(z2, c1, c0));
  uint32_t ui = *reinterpret_cast<const uint32_t*>(&u), ti;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(ti) : "r"(ui));
  const __half2 t = *reinterpret_cast<const __half2*>(&ti);
  const __half2 sech2 = __hfma2(__hneg2(t), t, one);                          // 1 - t^2
  const __half2 inner = __hmul2(__hmul2(half, z), __hfma2(z2, c3, c0));       // 0.5 z (c0 + 3 c1 z^2)
  const __half2 g = __hfma2(sech2, inner, __hfma2(half, t, half));
  return __half22float2(g);
}

