// Problem description + fused epilogues shared by the tcgen05 kernel (gemm_tc5.cu) and the legacy mma.sync
// kernel (gemm_hmma.cu).  Every dense contraction of the restoration forward pass is ONE shape of this:
//
//     D[m, n] = sum_{tap} sum_{k} A[m + tap_off[tap], k] * Wt[n, tap*Kpad + k]        (bf16 x bf16 -> fp32)
//
//   linear layers (patch-embed, qkv, proj, fc1, fc2)        : 1 tap, tap_off = {0}
//   ConvTranspose2d k2 s2 (blocks.py:113-118)               : 1 tap, N = 4*Cout, pixel-shuffle store
//   Conv2d 3x3 pad 1 (blocks.py:66-72, decoder.py:77)       : 9 taps over the halo-padded NHWC pixel list:
//        activations are stored [B, H+2, W+2, C] with a zero 1-pixel halo, flattened to rows p; the tap
//        (dy,dx) is then a constant row shift dy*(W+2)+dx, so the A tile of every tap is a plain 2-D box and
//        TMA's out-of-bounds zero fill covers the ends of the buffer.  Halo rows compute garbage and are
//        stored as zeros, which keeps the halo invariant for the next layer.
#pragma once
#include "common.cuh"

namespace gemm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // bf16 elements = one 128-byte swizzle atom

enum EpiMode : int { EPI_ROWMAJOR = 0, EPI_CONV = 1, EPI_CONVT = 2 };

struct Problem {
  // operands
  const __nv_bfloat16* A;   // [a_rows, ldA]
  const __nv_bfloat16* Wt;  // [N, num_kb*64]
  int M, N;                 // output rows / cols (N % 16 == 0)
  int a_rows, ldA, kvalid;  // A extent; kvalid = valid K per tap (<= kb_per_tap*64), the rest reads as zero
  int num_kb, kb_per_tap;
  int w_ld_kb;              // row pitch of Wt in 64-wide k-blocks (0: num_kb) -- a packed matrix may hold more taps than are contracted
  int tap_off[10];
  int dbg_shift, dbg_bo;    // experiment knobs (tools/desc_probe.py): unaligned A window start / descriptor base offset
  int skip_tap;             // >= 0: tap whose A operand is `skip` (same shape as A) and whose weights are the identity
  // epilogue
  int mode, act, out_f32;
  const float* bias;        // [N] or null
  void* out;
  int ldo;                  // row pitch of out (elements)
  const float* resid;       // EPI_ROWMAJOR: fp32 residual [M, ldo] (may alias out) or null
  const __nv_bfloat16* skip;// EPI_CONV: skip tensor (same layout as out) or null
  // EPI_CONV: rows are padded pixels p of [B, H+2, W+2]
  FastDiv plane, pitch;
  int H, W;
  // EPI_CONVT: rows are input pixels ([B,Hin,Win] or padded), cols n = (di*2+dj)*Cout + o
  FastDiv in_plane, in_pitch, cout_div;
  int in_pad, Hin, Win;
  int out_plane, out_pitch; // (2Hin+2)*(2Win+2), 2Win+2
  int zero_out_halo;        // EPI_CONVT, tcgen05 kernels: the epilogue threads zero the 1-pixel halo of the output buffer before their
                            // first tile (the pixel-shuffle scatter only writes the interior) -- no separate zero-halo launch
  // EPI_CONV with the decoder's 1x1 head fused behind it (decoder.py:76-81; tcgen05 resident-weight conv mode only):
  // the GELU'd tile never goes to HBM -- each epilogue lane contracts its accumulator row (rounded to bf16, as the two-kernel
  // path stores it) with the fp32 1x1 weights and stores NCHW fp32 straight into the model output.  `out` is unused then.
  const float* head_w;      // [head_cout, N] fp32, or null
  const float* head_b;      // [head_cout]
  float* head_out;          // [B, head_cout, H, W]
  int head_cout;            // <= 16
  // LayerNorm folded into the GEMMs around it (encoder.py:39,136,154-176; tcgen05 path, EPI_ROWMAJOR only).
  //   producer (proj / fc2, fp32 residual epilogue): additionally writes the bf16 copy of the new residual row (the next GEMM's A
  //     operand) and, per (row, column tile, column half), the partial sums {sum x, sum x^2} of the columns it owns:
  //     stats_out[m * stats_parts + slot], slot = 2 * (n0 / BN) + half; stats_parts = 2 * N / BN (see tc5_stats_parts())
  //   consumer (qkv / fc1): A = bf16(x) un-normalised, Wt = bf16(gamma (.) W); the epilogue applies
  //     out = rstd * (acc - mean * ln_s[n]) + bias[n]  with  ln_s[n] = sum_k Wt[n,k],  bias[n] = b[n] + sum_k beta[k] W[n,k]
  //     and {mean, rstd} of row m from the ln_parts partial sums in ln_stats (fp32; var = E[x^2] - mean^2, + ln_eps)
  __nv_bfloat16* out_bf16;  // producer: [M, ldo] bf16 copy of out, or null
  float2* stats_out;        // producer: [M, stats_parts] partial {sum, sum sq}, or null
  int stats_parts;
  const float2* ln_stats;   // consumer: [M, ln_parts], or null
  int ln_parts;
  const float* ln_s;        // consumer: [N] column sums of the gamma-folded weights
  float ln_eps, ln_inv_k;   // consumer: eps, 1 / (LayerNorm width)
  // Training-mode BatchNorm statistics out of the conv that produces the tensor (EPI_CONV, raw bf16 output, tcgen05 kernels that
  // support it, see tc5_supports_bn_sums()): bn_sums[n] += sum over stored rows of c[., n], bn_sums[N + n] += sum of c^2 -- the
  // bf16 values the epilogue stores (halo rows are stored as zeros and add nothing).  fp32 per-CTA partials in shared memory,
  // one fp64 atomic per channel and CTA at the end.  The caller zeroes bn_sums.
  double* bn_sums;
  // Training-step MLP (EPI_ROWMAJOR, bf16 output, tcgen05 backend):
  //   act = 1 with out_bf16 set (and no LayerNorm fields): the pre-activation acc + bias, rounded to bf16, goes to out_bf16 and
  //     GELU of that rounded value to out -- fc1 leaves both tensors the backward needs, no separate GELU pass;
  //   gelu_pre set (act = 0): out = bf16(bf16(acc + bias) * gelu'(gelu_pre[m, n])) -- the fc2 data gradient arrives already
  //     multiplied by the GELU derivative (same roundings as the GEMM followed by msd_gelu_bwd_bf16).
  const __nv_bfloat16* gelu_pre;
};

__device__ __forceinline__ float apply_act(float v, int act) { return act ? gelu_erf(v) : v; }

// store NV (2, 8 or 16) consecutive output columns n..n+NV-1 of output row m; v = raw accumulators
template <int NV>
__device__ __forceinline__ void epilogue_store(const Problem& p, int m, int n, float (&v)[NV]) {
  if (m >= p.M || n >= p.N) return;
  if (p.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] += __ldg(p.bias + n + i);
  }
  if (p.mode == EPI_ROWMAJOR) {
    const size_t off = (size_t)m * p.ldo + n;
    if (p.resid != nullptr) {
#pragma unroll
      for (int i = 0; i < NV; i += 2) {
        const float2 r = *reinterpret_cast<const float2*>(p.resid + off + i);
        v[i] += r.x; v[i + 1] += r.y;
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = apply_act(v[i], p.act);
    if (p.out_f32) {
      float* o = reinterpret_cast<float*>(p.out) + off;
      if constexpr (NV % 4 == 0) {
#pragma unroll
        for (int i = 0; i < NV; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < NV; i += 2) *reinterpret_cast<float2*>(o + i) = make_float2(v[i], v[i + 1]);
      }
    } else {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
      if constexpr (NV % 8 == 0) {
#pragma unroll
        for (int i = 0; i < NV; i += 8)
          *reinterpret_cast<uint4*>(o + i) = make_uint4(pack_bf16x2(v[i], v[i + 1]), pack_bf16x2(v[i + 2], v[i + 3]),
                                                        pack_bf16x2(v[i + 4], v[i + 5]), pack_bf16x2(v[i + 6], v[i + 7]));
      } else {
#pragma unroll
        for (int i = 0; i < NV; i += 2) *reinterpret_cast<uint32_t*>(o + i) = pack_bf16x2(v[i], v[i + 1]);
      }
    }
    return;
  }
  size_t off;
  bool interior;
  if (p.mode == EPI_CONV) {
    uint32_t img, rem, yy, xx;
    fd_divmod((uint32_t)m, p.plane, img, rem);
    fd_divmod(rem, p.pitch, yy, xx);
    interior = yy >= 1 && yy <= (uint32_t)p.H && xx >= 1 && xx <= (uint32_t)p.W;
    off = (size_t)m * p.ldo + n;
    if (interior && p.skip != nullptr && p.skip_tap < 0) {
#pragma unroll
      for (int i = 0; i < NV; i += 2) {
        const float2 s = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p.skip + off + i));
        v[i] += s.x; v[i + 1] += s.y;
      }
    }
  } else {  // EPI_CONVT
    uint32_t img, rem, yy, xx;
    if (p.in_pad) {
      fd_divmod((uint32_t)m, p.in_plane, img, rem);
      fd_divmod(rem, p.in_pitch, yy, xx);
      interior = yy >= 1 && yy <= (uint32_t)p.Hin && xx >= 1 && xx <= (uint32_t)p.Win;
      yy -= 1; xx -= 1;
    } else {
      fd_divmod((uint32_t)m, p.in_plane, img, rem);
      fd_divmod(rem, p.in_pitch, yy, xx);
      interior = true;
    }
    if (!interior) return;  // nothing to scatter from a halo row
    uint32_t q, o;
    fd_divmod((uint32_t)n, p.cout_div, q, o);
    const uint32_t opix = img * (uint32_t)p.out_plane + (2 * yy + (q >> 1) + 1) * (uint32_t)p.out_pitch + (2 * xx + (q & 1) + 1);
    off = (size_t)opix * p.ldo + o;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = interior ? apply_act(v[i], p.act) : 0.0f;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
  if constexpr (NV % 8 == 0) {
#pragma unroll
    for (int i = 0; i < NV; i += 8)
      *reinterpret_cast<uint4*>(o + i) = make_uint4(pack_bf16x2(v[i], v[i + 1]), pack_bf16x2(v[i + 2], v[i + 3]),
                                                    pack_bf16x2(v[i + 4], v[i + 5]), pack_bf16x2(v[i + 6], v[i + 7]));
  } else {
#pragma unroll
    for (int i = 0; i < NV; i += 2) *reinterpret_cast<uint32_t*>(o + i) = pack_bf16x2(v[i], v[i + 1]);
  }
}

// host launchers (one per backend)
int launch_tc5(const Problem& p, cudaStream_t s);
int launch_hmma(const Problem& p, cudaStream_t s);
bool tc5_can_fuse_head(const Problem& p, int head_cout);   // may this 3x3 conv carry the 1x1 head in its epilogue?
int tc5_stats_parts(int N, int M);                         // partial LayerNorm sums per row a stats-producing GEMM of this shape writes
bool roll_geometry_ok(int H, int W, int C);                // does the rolling-row conv kernel (conv_roll.cu) take a C-channel H x W conv?
bool tc5_supports_bn_sums(const Problem& p);               // will launch_tc5 route this conv to a kernel that accumulates bn_sums?

inline int launch(const Problem& p, int backend, cudaStream_t s) {
  if (backend == 1) {
    MSD_REQUIRE(p.ln_stats == nullptr && p.stats_out == nullptr && p.bn_sums == nullptr && p.gelu_pre == nullptr && (p.out_bf16 == nullptr || p.stats_out != nullptr),
                "gemm: the folded LayerNorm / column-sum / training-MLP epilogues exist on the tcgen05 backend only");
    return launch_hmma(p, s);
  }
  return launch_tc5(p, s);
}

}  // namespace gemm
