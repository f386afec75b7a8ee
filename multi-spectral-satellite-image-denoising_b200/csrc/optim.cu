// Optimizer step of the training loop (reference scripts/train.py:144-160 builds torch.optim.AdamW over two LR
// groups, src/training/trainer.py:188-213 clips the global gradient norm to 1.0 and steps).  Two multi-tensor kernels:
//   1. sum of squared gradients over every tensor -> one fp64 on the device (block partials in registers, one atomic per block)
//   2. AdamW with the clip coefficient min(1, max_norm / (norm + 1e-6)) applied to the gradient on the fly -- no host round
//      trip for the norm, no separate pass that rescales the gradients.
// Pure HBM streaming: 16 B read + 12 B written per parameter (p, g, m, v in; p, m, v out), 128-bit accesses.
// A launch takes up to 48 tensors through its parameter block; every block walks all of them with a grid stride.
#include "common.cuh"

namespace {

constexpr int kMaxT = 48;
constexpr int kBlock = 256;

struct TensorTable {
  float* p[kMaxT];
  const float* g[kMaxT];
  float* m[kMaxT];
  float* v[kMaxT];
  long long n[kMaxT];
  float lr[kMaxT], wd[kMaxT];
  float bc1[kMaxT], isbc2[kMaxT];       // per-tensor bias corrections: 1 - beta1^step, 1 / sqrt(1 - beta2^step)
  int count;
};

__global__ void __launch_bounds__(kBlock) sqnorm_kernel(const TensorTable t, double* __restrict__ out) {
  msd_pdl_enter();
  float acc = 0.f;
  const long long tid = (long long)blockIdx.x * kBlock + threadIdx.x, stride = (long long)gridDim.x * kBlock;
  for (int k = 0; k < t.count; ++k) {
    const float* g = t.g[k];
    const long long n4 = t.n[k] >> 2;
    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
      for (long long i = tid; i < n4; i += stride) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(g) + i);
        acc = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, acc))));
      }
      for (long long i = (n4 << 2) + tid; i < t.n[k]; i += stride) acc = fmaf(g[i], g[i], acc);
    } else {
      for (long long i = tid; i < t.n[k]; i += stride) acc = fmaf(g[i], g[i], acc);
    }
  }
  double d = warp_sum_d((double)acc);
  __shared__ double red[kBlock / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kBlock / 32; ++w) s += red[w];
    atomicAdd(out, s);
  }
}

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, float clip, float decay, float b1, float b2,
                                          float step_size, float inv_sqrt_bc2, float eps) {
  g *= clip;
  p *= decay;
  m = fmaf(b1, m, (1.0f - b1) * g);
  v = fmaf(b2, v, (1.0f - b2) * g * g);
  const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
  p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(kBlock) adamw_kernel(const TensorTable t, float b1, float b2, float eps, float max_norm,
                                                       const double* __restrict__ sqnorm) {
  msd_pdl_enter();
  float clip = 1.0f;
  if (max_norm > 0.0f && sqnorm != nullptr) {
    const double sq = *sqnorm;
    if (!isfinite(sq)) return;                                       // inf / nan gradients: skip the step (see msd.h)
    const float c = max_norm / ((float)sqrt(sq) + 1e-6f);            // torch.nn.utils.clip_grad_norm_
    clip = c < 1.0f ? c : 1.0f;
  }
  const long long tid = (long long)blockIdx.x * kBlock + threadIdx.x, stride = (long long)gridDim.x * kBlock;
  for (int k = 0; k < t.count; ++k) {
    float* p = t.p[k]; const float* g = t.g[k]; float* m = t.m[k]; float* v = t.v[k];
    const float lr = t.lr[k], decay = 1.0f - lr * t.wd[k], step_size = lr / t.bc1[k], inv_sqrt_bc2 = t.isbc2[k];
    const long long n = t.n[k], n4 = n >> 2;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
    if (vec) {
      for (long long i = tid; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
        adamw_one(pp.x, gg.x, mm.x, vv.x, clip, decay, b1, b2, step_size, inv_sqrt_bc2, eps);
        adamw_one(pp.y, gg.y, mm.y, vv.y, clip, decay, b1, b2, step_size, inv_sqrt_bc2, eps);
        adamw_one(pp.z, gg.z, mm.z, vv.z, clip, decay, b1, b2, step_size, inv_sqrt_bc2, eps);
        adamw_one(pp.w, gg.w, mm.w, vv.w, clip, decay, b1, b2, step_size, inv_sqrt_bc2, eps);
        reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
      }
    }
    for (long long i = (vec ? (n4 << 2) : 0) + tid; i < n; i += stride)
      adamw_one(p[i], g[i], m[i], v[i], clip, decay, b1, b2, step_size, inv_sqrt_bc2, eps);
  }
}

int fill(TensorTable& t, const msd_adamw_tensor* src, int n, float beta1 = 0.0f, float beta2 = 0.0f, int step = 1) {
  t.count = n;
  for (int k = 0; k < n; ++k) {
    const int st = src[k].step > 0 ? src[k].step : step;
    t.bc1[k] = 1.0f - powf(beta1, (float)st);
    t.isbc2[k] = 1.0f / sqrtf(1.0f - powf(beta2, (float)st));
    MSD_REQUIRE(src[k].param && src[k].grad && src[k].exp_avg && src[k].exp_avg_sq && src[k].numel > 0, "adamw: null tensor %d", k);
    t.p[k] = src[k].param; t.g[k] = src[k].grad; t.m[k] = src[k].exp_avg; t.v[k] = src[k].exp_avg_sq;
    t.n[k] = src[k].numel; t.lr[k] = src[k].lr; t.wd[k] = src[k].weight_decay;
  }
  return MSD_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Weight repack of the training step: after every optimizer step the fp32 master weights have to reappear as the bf16 /
// transposed / tap-flipped operands the forward and the data-gradient GEMMs read.  One table-driven gather kernel (one thread
// per DESTINATION element: coalesced bf16 stores, source reads served by L2 -- every weight matrix fits it) replaces the ~100
// small cast / permute / flip / pad launches a step spent on this.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxOps = 24;
struct RepackTable {
  const float* src[kMaxOps];
  void* dst[kMaxOps];
  long long n[kMaxOps];          // destination elements
  int kind[kMaxOps], d0[kMaxOps], d1[kMaxOps], kpad[kMaxOps], flag[kMaxOps];
  int count;
};

__device__ __forceinline__ float repack_value(const float* __restrict__ src, int kind, int d0, int d1, int kpad, int flag, long long i) {
  switch (kind) {
    case MSD_REPACK_CAST: return src[i];
    case MSD_REPACK_TRANSPOSE: {            // src [d0 = N, d1 = K] -> dst [K, N]
      const int k = (int)(i / d0), n = (int)(i - (long long)k * d0);
      return src[(size_t)n * d1 + k];
    }
    case MSD_REPACK_CONV3: {                // src [d0 = Cout, d1 = Cin, 3, 3] -> dst [Cout, 9 * kpad] tap-major (ky, kx, ci)
      const int row = 9 * kpad, co = (int)(i / row), r = (int)(i - (long long)co * row), tap = r / kpad, ci = r - tap * kpad;
      return ci < d1 ? src[((size_t)co * d1 + ci) * 9 + tap] : 0.0f;
    }
    case MSD_REPACK_CONV3_DGRAD: {          // dst [Cin, T * kpad(Cout)]: Wd[ci, tap, co] = W[co, ci, 8 - tap]; tap 9 (flag) = identity
      const int taps = flag ? 10 : 9, row = taps * kpad, ci = (int)(i / row), r = (int)(i - (long long)ci * row), tap = r / kpad, co = r - tap * kpad;
      if (tap == 9) return co == ci ? 1.0f : 0.0f;
      return co < d0 ? src[((size_t)co * d1 + ci) * 9 + (8 - tap)] : 0.0f;
    }
    case MSD_REPACK_CONVT: {                // src [d0 = Cin, d1 = Cout, 2, 2] -> dst [4 * Cout, kpad(Cin)] rows (di, dj, o)
      const int rowi = (int)(i / kpad), ci = (int)(i - (long long)rowi * kpad), q = rowi / d1, o = rowi - q * d1;
      return ci < d0 ? src[((size_t)ci * d1 + o) * 4 + q] : 0.0f;
    }
    case MSD_REPACK_CONVT_T: {              // dst [Cin, 4 * Cout] = the transpose of the un-padded matrix above
      const int cols = 4 * d1, ci = (int)(i / cols), c = (int)(i - (long long)ci * cols), q = c / d1, o = c - q * d1;
      return src[((size_t)ci * d1 + o) * 4 + q];
    }
    default: return src[i % d0];            // MSD_REPACK_REPEAT_F32: dst fp32 [flag * d0] = src repeated
  }
}

__global__ void __launch_bounds__(kBlock) repack_kernel(const RepackTable t) {
  msd_pdl_enter();
  const long long tid = (long long)blockIdx.x * kBlock + threadIdx.x, stride = (long long)gridDim.x * kBlock;
  for (int k = 0; k < t.count; ++k) {
    const float* src = t.src[k];
    const int kind = t.kind[k], d0 = t.d0[k], d1 = t.d1[k], kpad = t.kpad[k], flag = t.flag[k];
    if (kind == MSD_REPACK_REPEAT_F32) {
      float* dst = reinterpret_cast<float*>(t.dst[k]);
      for (long long i = tid; i < t.n[k]; i += stride) dst[i] = repack_value(src, kind, d0, d1, kpad, flag, i);
    } else {
      // two destination elements per thread: 4-byte stores (every packed row has an even length)
      __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(t.dst[k]);
      for (long long j = tid; 2 * j < t.n[k]; j += stride)
        dst[j] = __floats2bfloat162_rn(repack_value(src, kind, d0, d1, kpad, flag, 2 * j), repack_value(src, kind, d0, d1, kpad, flag, 2 * j + 1));
    }
  }
}

}  // namespace

extern "C" int msd_repack_weights(const msd_repack_op* ops, int n, void* stream) {
  MSD_REQUIRE(ops != nullptr && n > 0, "repack_weights: bad arguments");
  for (int k0 = 0; k0 < n; k0 += kMaxOps) {
    RepackTable t;
    t.count = n - k0 < kMaxOps ? n - k0 : kMaxOps;
    for (int k = 0; k < t.count; ++k) {
      const msd_repack_op& o = ops[k0 + k];
      MSD_REQUIRE(o.src && o.dst && o.kind >= MSD_REPACK_CAST && o.kind <= MSD_REPACK_REPEAT_F32 && o.d0 > 0, "repack_weights: bad op %d", k0 + k);
      long long cnt = 0;
      switch (o.kind) {
        case MSD_REPACK_CAST: cnt = (long long)o.d0 * (o.d1 > 0 ? o.d1 : 1); break;
        case MSD_REPACK_TRANSPOSE: cnt = (long long)o.d0 * o.d1; break;
        case MSD_REPACK_CONV3: cnt = (long long)o.d0 * 9 * o.kpad; break;
        case MSD_REPACK_CONV3_DGRAD: cnt = (long long)o.d1 * (o.flag ? 10 : 9) * o.kpad; break;
        case MSD_REPACK_CONVT: cnt = 4ll * o.d1 * o.kpad; break;
        case MSD_REPACK_CONVT_T: cnt = 4ll * o.d1 * o.d0; break;
        default: cnt = (long long)o.d0 * (o.flag > 0 ? o.flag : 1); break;
      }
      MSD_REQUIRE(o.kind == MSD_REPACK_REPEAT_F32 || cnt % 2 == 0, "repack_weights: op %d has an odd element count", k0 + k);
      MSD_REQUIRE(o.kind < MSD_REPACK_CONV3 || o.kind > MSD_REPACK_CONVT || o.kpad > 0, "repack_weights: op %d needs kpad", k0 + k);
      t.src[k] = o.src; t.dst[k] = o.dst; t.n[k] = cnt; t.kind[k] = o.kind; t.d0[k] = o.d0; t.d1[k] = o.d1; t.kpad[k] = o.kpad;
      t.flag[k] = o.kind == MSD_REPACK_CAST ? 0 : o.flag;
      if (o.kind == MSD_REPACK_CAST) { t.d0[k] = 1; }
    }
    MSD_CHECK_CUDA(msd_launch_pdl((repack_kernel), dim3(148 * 8), dim3(kBlock), 0, (cudaStream_t)stream, t));
    MSD_LAUNCH_CHECK();
  }
  return MSD_OK;
}

extern "C" int msd_grad_sqnorm_f32(const msd_adamw_tensor* tensors, int n, double* d_sqnorm, void* stream) {
  MSD_REQUIRE(tensors && d_sqnorm && n > 0, "grad_sqnorm: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  MSD_CHECK_CUDA(cudaMemsetAsync(d_sqnorm, 0, sizeof(double), s));
  for (int k0 = 0; k0 < n; k0 += kMaxT) {
    TensorTable t;
    int rc = fill(t, tensors + k0, n - k0 < kMaxT ? n - k0 : kMaxT);
    if (rc) return rc;
    MSD_CHECK_CUDA(msd_launch_pdl((sqnorm_kernel), dim3(148 * 4), dim3(kBlock), 0, s, t, d_sqnorm));
    MSD_LAUNCH_CHECK();
  }
  return MSD_OK;
}

extern "C" int msd_adamw_step_f32(const msd_adamw_tensor* tensors, int n, float beta1, float beta2, float eps, int step, float max_grad_norm,
                                  const double* d_sqnorm, void* stream) {
  MSD_REQUIRE(tensors && n > 0 && step >= 1, "adamw_step: bad arguments");
  MSD_REQUIRE(max_grad_norm <= 0.0f || d_sqnorm != nullptr, "adamw_step: clipping needs the squared gradient norm");
  for (int k0 = 0; k0 < n; k0 += kMaxT) {
    TensorTable t;
    int rc = fill(t, tensors + k0, n - k0 < kMaxT ? n - k0 : kMaxT, beta1, beta2, step);
    if (rc) return rc;
    MSD_CHECK_CUDA(msd_launch_pdl((adamw_kernel), dim3(148 * 8), dim3(kBlock), 0, (cudaStream_t)stream, t, beta1, beta2, eps, max_grad_norm, d_sqnorm));
    MSD_LAUNCH_CHECK();
  }
  return MSD_OK;
}
