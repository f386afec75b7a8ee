/*
 * msd.h -- C ABI of libmsd_b200.so, the sm_100a kernel library behind the restoration hot path
 * (corrupt -> SatMAE ViT-B + CNN-decoder restore -> MSE+SSIM loss / PSNR.SSIM.SAM metrics).
 *
 * The reference (ajammoussi/Multi-Spectral-Satellite-Image-Denoising) is pure Python/PyTorch and has no FFI;
 * the "interface each entry point replaces" is therefore a reference Python function or nn.Module.forward,
 * cited per entry as <file>:<lines> relative to the reference tree.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - every pointer named d_* is a DEVICE pointer owned by the caller (PyTorch allocates everything);
 *    the library owns only the opaque msd_model (tensor maps, tap tables), never frees caller memory;
 *  - every call is asynchronous and ordered on `stream` (a cudaStream_t passed as void*); no call
 *    synchronises the device;
 *  - return value: 0 on success, a negative msd_status otherwise; text via msd_last_error() (thread-local);
 *  - no C++ types or exceptions cross this boundary.
 */
#ifndef MSD_B200_H_
#define MSD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum msd_status {
  MSD_OK = 0,
  MSD_ERR_INVALID = -1,  /* bad shape / null pointer / misaligned pointer / unsupported option */
  MSD_ERR_CUDA = -2,     /* a CUDA runtime or driver call failed; see msd_last_error() */
  MSD_ERR_ARCH = -3,     /* device is not compute capability 10.x (library is sm_100a only) */
  MSD_ERR_WORKSPACE = -4 /* caller-provided workspace too small */
} msd_status;

/* Library ABI version (bumped on any signature change). */
int msd_version(void);
/* Last error text of the calling thread; valid until the next failing call on that thread. */
const char* msd_last_error(void);
/* 0 if the current device is sm_100 (B200); MSD_ERR_ARCH otherwise. */
int msd_check_device(void);

/* ---------------------------------------------------------------------------------------------------
 * Counter-based RNG, bit-compatible with the Philox4x32-10 stream PyTorch's CUDA generator produces
 * (what `torch.randn_like` / `torch.rand` draw when the reference's AddSensorNoise runs on a CUDA
 * tensor: src/data/transforms.py:90-121).  `torch_grid` is the grid PyTorch would launch for `numel`
 * (min(SMs * maxThreadsPerSM/256, ceil(numel/256))); the offset increment PyTorch applies afterwards is
 * returned by msd_philox_offset_increment().
 * ------------------------------------------------------------------------------------------------- */
uint64_t msd_philox_offset_increment(int64_t numel, uint32_t torch_grid);
int msd_philox_normal_f32(float* d_out, int64_t numel, uint64_t seed, uint64_t offset, uint32_t torch_grid,
                          void* stream);
int msd_philox_uniform_f32(float* d_out, int64_t numel, uint64_t seed, uint64_t offset, uint32_t torch_grid,
                           void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Corruption simulator.  Replaces AddSensorNoise.forward (src/data/transforms.py:70-134) and
 * _add_striping (:136-154) for a batch [B,C,H,W] fp32 NCHW.
 *   order of draws (identical to the reference): gaussian randn_like (if sg>0), speckle randn_like (if ss>0),
 *   B x rand(C) (if p>0, one call per sample), thermal randn_like (if st>0), stripe randn(B,C,1,W) (if stripe).
 *   d_thermal_w : [C] fp32 device, torch.linspace(1,2,C) as the reference builds it (:112-114)
 *   d_dead_mask : [B*C] uint8 device, OUT (1 = band zeroed); required when p>0
 *   d_stripe_ws : [B*C*W] fp32 device scratch, required when stripe != 0
 *   d_patch_bf16: optional OUT [B*(H/16)*(W/16), C*256] bf16, the patch-embed A operand of the noisy image
 *                 (src/models/encoder.py:48-50 im2col order (c,ky,kx)); NULL to skip
 *   offset_consumed: OUT, total Philox offset the reference would have consumed (host pointer)
 * ------------------------------------------------------------------------------------------------- */
int msd_corrupt_f32(const float* d_clean, float* d_noisy, const float* d_thermal_w, uint8_t* d_dead_mask,
                    float* d_stripe_ws, void* d_patch_bf16, int B, int C, int H, int W, float sg, float ss, float p,
                    float st, int stripe, uint64_t seed, uint64_t offset, uint32_t torch_grid_full,
                    uint32_t torch_grid_mask, uint32_t torch_grid_stripe, uint64_t* offset_consumed, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Fused evaluation metrics.  Replaces MetricsTracker.update's per-sample loop
 * (src/training/metrics.py:247-256) and calculate_psnr/ssim/sam/rmse (:19-46, 49-121, 124-172, 175-191).
 *   d_per_image: OUT [B,4] fp64 = {psnr_dB (inf if mse==0), ssim, sam_degrees, rmse} per image
 *   d_sums     : OUT [5]  fp64 = {sum psnr, sum ssim, sum sam, sum rmse, B}  (the tracker's running sums)
 *   d_scratch  : OUT [B,4] fp64 raw per-image sums {sum (p-t)^2, sum ssim_map, sum SAM degrees, sum |p-t|}
 *                (whole-batch calculate_* and calculate_mae (:194-209) are derived from these)
 * ------------------------------------------------------------------------------------------------- */
int msd_metrics_f32(const float* d_pred, const float* d_target, int B, int C, int H, int W, float data_range,
                    double* d_per_image, double* d_sums, double* d_scratch, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Training loss.  Replaces CombinedLoss.forward (src/training/losses.py:55-77) = w_mse*MSE + w_ssim*(1-SSIM)
 * with SSIMLoss._ssim (:123-163): 11x11 Gaussian sigma 1.5, zero padding 5, global mean.
 *   d_out      : OUT [3] fp32 = {total, mse, 1-ssim}
 *   d_grad_pred: optional OUT [B,C,H,W] fp32 = grad_scale * d(total)/d(pred); NULL for forward only
 *   d_ws       : scratch, msd_loss_workspace_bytes() bytes (3 derivative maps + 2 fp64 accumulators)
 * ------------------------------------------------------------------------------------------------- */
size_t msd_loss_workspace_bytes(int B, int C, int H, int W);
int msd_loss_fwd_bwd_f32(const float* d_pred, const float* d_target, int B, int C, int H, int W, float w_mse,
                         float w_ssim, float data_range, float grad_scale, float* d_out, float* d_grad_pred,
                         void* d_ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Restoration model forward.  Replaces SatMAERestoration.forward (src/models/satmae_restoration.py:81-110):
 * PatchEmbed (encoder.py:41-52) -> +pos_embed -> 12 pre-norm blocks (encoder.py:120-134 | 154-176) -> LayerNorm
 * -> 4 x {ConvTranspose2d k2s2 + BN + GELU, 2 x ResidualBlock} (blocks.py:74-85, 113-123) -> head
 * (decoder.py:76-81).  Eval-mode BatchNorm is folded into the packed weights by the host shim.
 * ------------------------------------------------------------------------------------------------- */
typedef struct msd_model msd_model;

typedef struct msd_model_desc {
  int in_channels;  /* 13 */
  int image_size;   /* 192 */
  int patch_size;   /* 16 */
  int embed_dim;    /* 768 */
  int depth;        /* 12 */
  int num_heads;    /* 12 */
  int mlp_hidden;   /* 3072 */
  int dec_channels[4]; /* 384,192,96,48 */
  int out_channels; /* 13 */
  float ln_eps_block; /* 1e-6 (timm Block) or 1e-5 (reference fallback CustomBlock) */
  float ln_eps_outer; /* 1e-5: patch_embed.norm and encoder.norm */
  int gemm_backend;   /* 0 = tcgen05/TMEM/TMA (product path), 1 = legacy mma.sync (A/B baseline) */
} msd_model_desc;

/* Packed weights, all DEVICE pointers that must outlive the model (the shim keeps the tensors alive).
 * bf16 matrices are row-major [N, K] (K contiguous).  Conv weights are [Cout, 9*Kpad] tap-major with
 * Kpad = Cin rounded up to 64 (10*Kpad, last tap = identity, for the second conv of a ResidualBlock); BN already folded (w*gamma/sqrt(var+eps), bias = beta + (b-mean)*scale). */
typedef struct msd_block_weights {
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  const void *w_qkv;  const float *b_qkv;   /* [2304,768]  */
  const void *w_proj; const float *b_proj;  /* [768,768]   */
  const void *w_fc1;  const float *b_fc1;   /* [3072,768]  */
  const void *w_fc2;  const float *b_fc2;   /* [768,3072]  */
  /* LayerNorm folded into the GEMM behind it (optional; all six set or all NULL; tcgen05 backend only):
   *   w_*_ln bf16 [N,768] = bf16(gamma[k] * W[n,k]),  s_* fp32 [N] = sum_k float(w_*_ln[n,k]),
   *   b_*_ln fp32 [N] = b[n] + sum_k beta[k] * W[n,k]   (gamma, beta = norm1 for qkv, norm2 for fc1).
   * When present the eval forward runs no per-block LayerNorm kernel: proj / fc2 emit the bf16 copy of the residual row and
   * its partial {sum, sum of squares}, qkv / fc1 apply rstd * (acc - mean * s) + b_ln in their epilogues (encoder.py:154-176). */
  const void *w_qkv_ln; const float *s_qkv; const float *b_qkv_ln;
  const void *w_fc1_ln; const float *s_fc1; const float *b_fc1_ln;
} msd_block_weights;

typedef struct msd_stage_weights {
  const void *w_up;  const float *b_up;     /* [4*Cout, Cin] rows ordered (di,dj,o); bias [4*Cout] */
  const void *w_conv[4]; const float *b_conv[4]; /* rb1.conv1, rb1.conv2, rb2.conv1, rb2.conv2 */
} msd_stage_weights;

typedef struct msd_model_weights {
  const void *w_pe; const float *b_pe;      /* [768, 3328] */
  const float *pe_ln_g, *pe_ln_b;
  const float *pos_embed;                   /* [144,768] fp32 */
  const msd_block_weights* blocks;          /* [depth] host array */
  const float *norm_g, *norm_b;
  msd_stage_weights stages[4];
  const void *w_head3; const float *b_head3; /* [48, 9*64] */
  const float *w_head1; const float *b_head1;/* [13,48] fp32, [13] */
} msd_model_weights;

int msd_model_create(const msd_model_desc* desc, msd_model** out);
int msd_model_set_weights(msd_model* m, const msd_model_weights* w);
size_t msd_model_workspace_bytes(const msd_model* m, int B);
/* d_x: [B,C,S,S] fp32 NCHW (or NULL if d_patch_bf16 given); d_patch_bf16: optional pre-patchified A operand
 * written by msd_corrupt_f32; d_y: [B,out_channels,S,S] fp32 NCHW; d_ws: workspace (256-byte aligned). */
int msd_model_forward(msd_model* m, const float* d_x, const void* d_patch_bf16, int B, float* d_y, void* d_ws,
                      size_t ws_bytes, void* stream);
/* number of kernels one forward launches (for bench.py's gpu_launches claim) */
int msd_model_launch_count(const msd_model* m);
/* Optional live profile: when enabled, msd_model_forward brackets every launch with CUDA events on `stream`.
 * After the stream is synchronised, msd_model_profile_read returns the number of launches of the LAST forward and
 * fills per-launch milliseconds, kind (0 = tensor-core GEMM/conv kernel, 1 = attention, 2 = LayerNorm, 3 = other)
 * and algorithmic FLOPs (2*M*N*K on the unpadded shapes of SURVEY Appendix A). */
int msd_model_profile_enable(msd_model* m, int enable);
int msd_model_profile_read(msd_model* m, float* ms_out, int* kind_out, double* flops_out, int capacity);
void msd_model_destroy(msd_model* m);

/* ---------------------------------------------------------------------------------------------------
 * Building blocks, exported so tests can pin each kernel against the oracle separately.
 * ------------------------------------------------------------------------------------------------- */
/* y[M,N] = act(A[M,K] . W[N,K]^T + bias); A,W bf16 row-major; out bf16 (out_f32=0) or fp32 (out_f32=1);
 * act: 0 none, 1 GELU(erf).  resid_f32 != NULL -> out fp32 = resid + A.W^T + bias (may alias out). */
int msd_linear_bf16(const void* d_a, const void* d_w, const float* d_bias, void* d_out, const float* d_resid,
                    int M, int N, int K, int act, int out_f32, int backend, void* stream);
/* The MLP of a training step (encoder.py:160-166 under autograd), GELU folded into the GEMMs either side of it (tcgen05 backend):
 *   msd_linear_gelu_pre_bf16   d_pre = bf16(A . W^T + bias) and d_out = GELU(d_pre): fc1 leaves the pre-activation the backward
 *                              needs and the activation fc2 reads from one epilogue (bit-identical to msd_linear_bf16 with act = 0
 *                              followed by msd_gelu_fwd_bf16)
 *   msd_linear_dgelu_bf16      d_out = bf16(bf16(A . W^T + bias) * gelu'(d_pre)): the fc2 data gradient, already multiplied by the
 *                              GELU derivative (bit-identical to msd_linear_bf16 followed by msd_gelu_bwd_bf16)               */
int msd_linear_gelu_pre_bf16(const void* d_a, const void* d_w, const float* d_bias, void* d_out, void* d_pre, int M, int N, int K,
                             void* stream);
int msd_linear_dgelu_bf16(const void* d_a, const void* d_w, const float* d_bias, const void* d_pre, void* d_out, int M, int N, int K,
                          void* stream);
/* 3x3 conv, pad 1, + bias + GELU, NHWC bf16 with a 1-pixel zero halo: x,y,skip are [B,(H+2),(W+2),C*] bf16;
 * w [Cout, 9*Kpad]; with a skip tensor (Cin == Cout) w is [Cout, 10*Kpad], the tenth tap being the identity, and the
 * result is GELU(conv + bias + skip) -- the skip connection is accumulated by the tensor cores like any other tap. */
int msd_conv3x3_nhwc_bf16(const void* d_x, const void* d_w, const float* d_bias, const void* d_skip, void* d_y,
                          int B, int H, int W, int Cin, int Cout, int backend, void* stream);
/* The decoder head in one kernel (decoder.py:76-81): y NCHW fp32 [B,Cout,H,W] = head_w . GELU(conv3x3(x) + bias) + head_b with
 * x haloed NHWC bf16 [B,H+2,W+2,48], w [48, 9*64] (BN folded), head_w fp32 [Cout,48], Cout <= 16; tcgen05 path only. */
int msd_conv3x3_head_nchw_f32(const void* d_x, const void* d_w, const float* d_bias, const float* d_head_w,
                              const float* d_head_b, float* d_y, int B, int H, int W, int Cin, int Cout, void* stream);
/* ConvTranspose2d k2 s2 + bias + GELU: x [B,Hin(+2p),Win(+2p),Cin] bf16 -> y [B,2Hin+2,2Win+2,Cout] haloed. */
int msd_convt2x2_nhwc_bf16(const void* d_x, const void* d_w, const float* d_bias, void* d_y, int B, int Hin, int Win,
                           int in_padded, int Cin, int Cout, int backend, void* stream);
/* out bf16 [M,D] = LayerNorm(x fp32 [M,D]); if pos != NULL first x = LN(x)+pos[m % period] is written back (fp32). */
int msd_layernorm_f32_bf16(const float* d_x, const float* d_g, const float* d_b, float eps, void* d_out_bf16, int M,
                           int D, void* stream);
int msd_layernorm_pos_f32(float* d_x, const float* d_g, const float* d_b, float eps, const float* d_pos, int period,
                          int M, int D, void* stream);
/* LayerNorm folded into the GEMMs around it (tcgen05 path; what msd_model_forward runs when the msd_block_weights carry *_ln):
 *   msd_ln_stats_parts        : partial sums per row a producing GEMM of shape [M,N] writes (depends on its tile width)
 *   msd_layernorm_pos_stats_f32: msd_layernorm_pos_f32 + d_x_bf16 [M,D] = bf16(x) + d_stats [M,parts,2] (slot 0 = {sum x, sum x^2})
 *   msd_linear_resid_stats_f32: x fp32 [M,N] += A.W^T + bias (in place) + d_x_bf16 = bf16(x) + d_stats [M,parts,2] partial sums
 *   msd_linear_lnfold_bf16    : out bf16 [M,N] = act(rstd_m * (xb.Wln^T - mean_m * s) + bias_ln), {mean, rstd} of row m from d_stats
 *                               (var = E[x^2] - mean^2 over K columns, + eps) */
int msd_ln_stats_parts(int N, int M);
int msd_layernorm_pos_stats_f32(float* d_x, const float* d_g, const float* d_b, float eps, const float* d_pos, int period,
                                void* d_x_bf16, float* d_stats, int parts, int M, int D, void* stream);
int msd_linear_resid_stats_f32(const void* d_a, const void* d_w, const float* d_bias, float* d_x, void* d_x_bf16, float* d_stats,
                               int parts, int M, int N, int K, void* stream);
int msd_linear_lnfold_bf16(const void* d_x_bf16, const void* d_w_ln, const float* d_bias_ln, const float* d_ln_s,
                           const float* d_stats, int parts, float eps, void* d_out, int M, int N, int K, int act, void* stream);
/* qkv bf16 [B*T, 3*heads*64] -> out bf16 [B*T, heads*64]; softmax(q k^T / 8) v per (image, head); T <= 256. */
int msd_attention_bf16(const void* d_qkv, void* d_out, int B, int T, int heads, void* stream);
/* the same contraction on tcgen05 / TMEM / TMA (product path; msd_attention_bf16 is the mma.sync baseline): T == 144 */
int msd_attention_tc5_bf16(const void* d_qkv, void* d_out, int B, int T, int heads, void* stream);
/* x [B,C,S,S] fp32 -> A [B*(S/P)^2, C*P*P] bf16, column order (c,ky,kx), token n = (S/P)*py + px. */
int msd_patchify_bf16(const float* d_x, void* d_a, int B, int C, int S, int P, void* stream);
/* head 1x1: x haloed NHWC bf16 [B,H+2,W+2,Cin] -> y NCHW fp32 [B,Cout,H,W]; w fp32 [Cout,Cin]. */
int msd_head1x1_nchw_f32(const void* d_x, const float* d_w, const float* d_b, float* d_y, int B, int H, int W, int Cin,
                         int Cout, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Weight-gradient contractions of the training step (what autograd computes for nn.Linear / nn.Conv2d in the reference's
 * Trainer.train_epoch, src/training/trainer.py:141-146 -> loss.backward()).  Both ACCUMULATE into fp32 outputs the
 * caller zeroed (the contraction is split across CTAs and merged with fp32 atomics).
 *   linear : d_dw[out_f, in_f] (row pitch ldo) += dY[rows, out_f]^T . X[rows, in_f]          (bf16 operands)
 *   conv3x3: x [B,H+2,W+2,Cin], dc [B,H+2,W+2,Cout] haloed NHWC bf16, halo rows of dc zero;
 *            d_dw[9][cinpad][ldo] (tap = ky*3+kx major) += sum_p x[p + off_tap, ci] . dc[p, co]
 * ------------------------------------------------------------------------------------------------- */
int msd_wgrad_linear_bf16(const void* d_dy, int ld_dy, const void* d_x, int ld_x, float* d_dw, int ldo, long long rows,
                          int out_f, int in_f, void* stream);
int msd_wgrad_conv3x3_bf16(const void* d_x, const void* d_dc, float* d_dw, int B, int H, int W, int Cin, int Cout, int cinpad,
                           int ldo, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Bandwidth-bound pieces of the training step (csrc/train_kernels.cu, attention_bwd.cu), exported so the tests can pin
 * each against torch autograd.  Layouts as in the forward: decoder tensors are haloed NHWC bf16 [B,H+2,W+2,C].
 * ------------------------------------------------------------------------------------------------- */
/* d_sums fp64 [C] = column sums of a bf16 [rows, C] matrix (bias gradients); msd_f64_to_f32 converts the result. */
int msd_colsum_bf16(const void* d_a, long long rows, int C, double* d_sums, void* stream);
int msd_f64_to_f32(const double* d_in, float* d_out, int n, void* stream);
/* nn.BatchNorm2d in training mode (blocks.py:69, 84, 117; decoder.py:78) over a raw conv output c (halo rows zero):
 * d_stats fp32 [4][C] = {gamma*rstd, beta - mean*gamma*rstd, mean, rstd}; running stats updated in place; d_sums fp64 [2][C] scratch */
int msd_bn_stats_bf16(const void* d_c, int B, int H, int W, int C, const float* d_gamma, const float* d_beta, float eps,
                      float momentum, float* d_running_mean, float* d_running_var, float* d_stats, double* d_sums, void* stream);
/* The same in two steps -- d_sums fp64 [2][C] = {sum c, sum c^2} over [rows, C], then statistics + running-stat update from the
 * sums (count = B*H*W) -- so that the first step can ride in the epilogue of the conv that writes c: */
int msd_bn_sums_bf16(const void* d_c, long long rows, int C, double* d_sums, void* stream);
int msd_bn_finalize(const double* d_sums, double count, int C, const float* d_gamma, const float* d_beta, float eps, float momentum,
                    float* d_running_mean, float* d_running_var, float* d_stats, void* stream);
/* nn.Conv2d(k3, pad 1) of a training forward (blocks.py:66-72): d_y = raw bf16 conv output (+ bias, no activation, zero halo) AND
 * d_sums fp64 [2][Cout] = {sum y, sum y^2} of the stored values.  The 48- and 96-channel convs of the tcgen05 backend accumulate the
 * sums in their epilogues (no second pass over d_y); other shapes run msd_bn_sums_bf16 behind the conv.  *fused_out (may be NULL)
 * tells which. */
int msd_conv3x3_stats_nhwc_bf16(const void* d_x, const void* d_w, const float* d_bias, void* d_y, double* d_sums, int B, int H, int W,
                                int Cin, int Cout, int backend, int* fused_out, void* stream);
/* y = GELU(BN(c) [+ skip]) on interior pixels, zero halo */
int msd_bn_act_fwd_bf16(const void* d_c, const void* d_skip, const float* d_stats, void* d_y, int B, int H, int W, int C, void* stream);
/* backward of the above: d_dz = grad of the pre-activation (= grad of the skip input), d_dc = grad of c (may alias d_dz),
 * d_dgamma / d_dbeta fp32 [C] (overwritten; NULL to skip), d_coef fp32 [2][C] and d_sums fp64 [2][C] scratch.
 * d_dy must have zero halo rows like every haloed tensor the library writes (dz is formed as dy * gelu'(.) on all rows). */
int msd_bn_act_bwd_bf16(const void* d_dy, const void* d_c, const void* d_skip, const float* d_stats, void* d_dz, void* d_dc,
                        float* d_dgamma, float* d_dbeta, float* d_coef, double* d_sums, int B, int H, int W, int C, void* stream);
/* d_out = GELU(d_hpre) (erf form), n bf16 elements (n % 8 == 0) */
int msd_gelu_fwd_bf16(const void* d_hpre, void* d_out, long long n, void* stream);
/* d_out = d_dh * gelu'(d_hpre), n bf16 elements (n % 8 == 0), in place allowed */
int msd_gelu_bwd_bf16(const void* d_dh, const void* d_hpre, void* d_out, long long n, void* stream);
/* the same over a [rows, N] matrix, and d_colsum fp32 [N] += column sums of the result (the bias gradient of the Linear in front of
 * the GELU); N % 256 == 0 */
int msd_gelu_bwd_colsum_bf16(const void* d_dh, const void* d_hpre, void* d_out, float* d_colsum, int rows, int N, void* stream);
/* LayerNorm backward over [M, D=768]: d_dy bf16 (or fp32 when dy_is_f32), d_x fp32 LN input; d_g fp32 [M,D] = (accumulate ?
 * d_g : 0) + dx; optional bf16 copy of d_g; d_dgamma / d_dbeta fp32 [D] ACCUMULATED into (NULL for frozen layers) */
int msd_layernorm_bwd(const void* d_dy, int dy_is_f32, const float* d_x, const float* d_gamma, float eps, float* d_g, int accumulate,
                      void* d_g_bf16, float* d_dgamma, float* d_dbeta, int M, int D, void* stream);
/* the same, and d_colsum fp32 [D] += column sums of the bf16 copy (NULL: none): that copy is the dY operand of the Linear whose output
 * this gradient belongs to (fc2 / proj / patch embedding), so the sums are that layer's bias gradient */
int msd_layernorm_bwd_colsum(const void* d_dy, int dy_is_f32, const float* d_x, const float* d_gamma, float eps, float* d_g,
                             int accumulate, void* d_g_bf16, float* d_dgamma, float* d_dbeta, float* d_colsum, int M, int D, void* stream);
/* ConvTranspose2d k2 s2 backward gather: d_dy haloed [B,2Hin+2,2Win+2,C] -> d_dys [rows_in, 4*C] (q = di*2+dj major) */
int msd_convt_unshuffle_bf16(const void* d_dy, void* d_dys, int B, int Hin, int Win, int in_padded, int C, void* stream);
/* head 1x1 backward: d_dy fp32 NCHW -> d_dh bf16 haloed [B,H+2,W+2,Cin] and d_dyt bf16 [rows,16] (dy on the pixel list) */
int msd_head1x1_bwd(const float* d_dy, const float* d_w, void* d_dh, void* d_dyt, int B, int H, int W, int Cin, int Cout, void* stream);
/* d_out[i] = sum_b d_in[b*n + i] */
int msd_batch_sum_f32(const float* d_in, float* d_out, int B, long long n, void* stream);
/* attention backward: d_dqkv bf16 [B*T, 3*heads*64] from qkv and the gradient of the attention output */
int msd_attention_bwd_bf16(const void* d_qkv, const void* d_dout, void* d_dqkv, int B, int T, int heads, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Training step of the restoration model (BASELINE config C4): forward in training mode (BatchNorm batch statistics,
 * activations kept for the backward) and the full backward pass.  Replaces `output = self.model(noisy)` +
 * `loss.backward()` in Trainer.train_epoch (src/training/trainer.py:141-146, 178-181) for the model part; the loss and
 * its gradient come from msd_loss_fwd_bwd_f32.  Weights are NOT BatchNorm-folded here.
 *   bf16 matrices as in msd_model_weights; additionally the transposed / tap-flipped copies the data-gradient GEMMs read.
 *   Gradients are fp32 device buffers in the packed layouts noted below; NULL pointers are skipped (frozen parameters).
 * ------------------------------------------------------------------------------------------------- */
typedef struct msd_bn_params {
  const float *gamma, *beta;          /* [C] */
  float *running_mean, *running_var;  /* [C], updated in place by the training forward (momentum 0.1, unbiased var) */
} msd_bn_params;

typedef struct msd_train_block {
  const void *wT_qkv, *wT_proj, *wT_fc1, *wT_fc2; /* W^T: [768,2304], [768,768], [768,3072], [3072,768] bf16 */
  int trainable;                                   /* 0 = frozen block (encoder.py:270-293): data gradients only */
} msd_train_block;

typedef struct msd_train_stage {
  const void *w_up; const float *b_up;   /* [4*Cout, Kpad(Cin)] rows (di,dj,o), no BN fold; bias [4*Cout] */
  const void *wT_up;                     /* [Cin, 4*Cout] = w_up^T (data gradient) */
  msd_bn_params bn_up;
  const void *w_conv[4];                 /* [Cout, 9*Kpad], no bias, no BN fold */
  const void *wd_conv[4];                /* data-gradient weights [Cin, 9*Kpad(Cout)] with flipped taps; for the FIRST conv of
                                            each ResidualBlock (index 0, 2) a tenth identity tap adds the skip gradient */
  msd_bn_params bn_conv[4];
} msd_train_stage;

typedef struct msd_train_weights {
  msd_model_weights enc;                 /* encoder fields as in the eval forward; decoder fields of `enc` are ignored */
  const msd_train_block* blocks_t;       /* [depth] */
  msd_train_stage stages[4];
  const void *w_head3; const float *b_head3; const void *wd_head3; msd_bn_params bn_head;
  const float *w_head1; const float *b_head1;
  float bn_eps, bn_momentum;
} msd_train_weights;

typedef struct msd_block_grads {
  float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *w_qkv, *b_qkv, *w_proj, *b_proj, *w_fc1, *b_fc1, *w_fc2, *b_fc2;
} msd_block_grads;

typedef struct msd_stage_grads {
  float *w_up;                           /* [4*Cout, Cin] rows (di,dj,o) */
  float *bn_up_g, *bn_up_b;
  float *w_conv[4];                      /* [9][Kpad(Cin)][Cout] tap-major */
  float *bn_g[4], *bn_b[4];
} msd_stage_grads;

typedef struct msd_model_grads {
  float *w_pe, *b_pe, *pe_ln_g, *pe_ln_b, *pos_embed;
  const msd_block_grads* blocks;         /* [depth] host array */
  float *norm_g, *norm_b;
  msd_stage_grads stages[4];
  float *w_head3;                        /* [9][Kpad][C] */
  float *bn_head_g, *bn_head_b;
  float *w_head1, *b_head1;              /* [out_channels, C], [out_channels] */
} msd_model_grads;

int msd_model_set_train_weights(msd_model* m, const msd_train_weights* w);
/* byte offsets of the kept activations inside the training workspace (tests / debugging); see csrc/model.cu */
int msd_model_train_layout(const msd_model* m, int B, size_t* out, int capacity);
size_t msd_model_train_workspace_bytes(const msd_model* m, int B);
/* d_x [B,C,S,S] fp32 -> d_y [B,out_channels,S,S] fp32; keeps every activation the backward needs inside d_ws */
int msd_model_forward_train(msd_model* m, const float* d_x, int B, float* d_y, void* d_ws, size_t ws_bytes, void* stream);
/* d_grad_y [B,out_channels,S,S] fp32; d_ws must be the workspace of the matching msd_model_forward_train call */
int msd_model_backward(msd_model* m, const float* d_grad_y, int B, void* d_ws, size_t ws_bytes, const msd_model_grads* grads,
                       void* stream);

/* Gradient groups for the data-parallel step ("gradients averaged over ranks, overlapped with backward").
 * msd_model_backward records one event per group when that group's gradients are complete (main and weight-gradient stream):
 *   group 0 = every decoder parameter, group 1 + j = encoder block depth-1-j (group 1 also encoder.norm), group depth+1 = patch
 *   embedding, its LayerNorm and pos_embed.  msd_model_grad_group_wait makes `stream` wait for group k of the LAST backward, so the
 *   caller can start that group's all-reduce on a communication stream while the rest of the backward is still running. */
int msd_model_grad_groups(const msd_model* m);
int msd_model_grad_group_wait(msd_model* m, int group, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Data formats either side of the path (csrc/ingest.cu).
 * msd_ingest_u16_f32: EuroSATMultiSpectral.__getitem__ preprocessing (src/data/dataset.py:160-186) for a batch of raw tiles:
 *   d_dn uint16 [B,C,h,w] -> per-image scale (max > 10: /10000, max > 1.5: /255, else 1) -> bilinear SxS (align_corners=False)
 *   -> clamp(0,1) -> (x-0.5)/0.25 -> d_out fp32 [B,C,S,S]; d_scale_ws fp32 [B] scratch.
 * msd_tile_extract_f32 / msd_tile_blend_f32: full-scene tiling (BASELINE config C5; no reference semantics):
 *   tiles[t] = scene[:, ty0[t]:+S, tx0[t]:+S];  out = sum_t w_t restored_t / sum_t w_t over scene rows [row0,row1) with the
 *   separable weight w(y,x) = r(y) r(x), r(i) = min(i+1, S-i, ramp)/ramp; gy[ny], gx[nx] = sorted tile-grid origins,
 *   tile_index[ny*nx] = slot of grid tile (r,k) in d_restored or -1.
 * ------------------------------------------------------------------------------------------------- */
int msd_ingest_u16_f32(const uint16_t* d_dn, float* d_out, float* d_scale_ws, int B, int C, int h, int w, int S, void* stream);
/* inverse of the ingest normalisation (the writer side): d_dn uint16 = round(clamp(x*0.25 + 0.5, 0, 1) * 10000) */
int msd_export_u16(const float* d_x, uint16_t* d_dn, long long numel, void* stream);
int msd_tile_extract_f32(const float* d_scene, const int* d_ty0, const int* d_tx0, float* d_tiles, int n_tiles, int C, int H, int W,
                         int S, void* stream);
int msd_tile_blend_f32(const float* d_restored, const int* d_gy, const int* d_gx, const int* d_tile_index, int ny, int nx, float* d_out,
                       int C, int H, int W, int S, int ramp, int row0, int row1, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Optimizer step.  Replaces torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW.step in
 * Trainer.train_epoch (src/training/trainer.py:188-213; optimizer built in scripts/train.py:144-160).
 * `tensors` is a HOST array of n entries whose pointers are device fp32 buffers of `numel` elements.
 *   msd_grad_sqnorm_f32: *d_sqnorm (device fp64) = sum over all tensors of sum g^2
 *   msd_adamw_step_f32 : decoupled weight decay, bias-corrected moments (step is 1-based); when max_grad_norm > 0 the gradient
 *                        is multiplied by min(1, max_grad_norm / (sqrt(*d_sqnorm) + 1e-6)) on the fly (gradients are not modified);
 *                        a non-finite *d_sqnorm (inf / nan gradients, e.g. an overflowed fp16-scaled step) skips the whole update,
 *                        as GradScaler.step does in the reference's loop (trainer.py:194-207), instead of poisoning the weights
 * ------------------------------------------------------------------------------------------------- */
typedef struct msd_adamw_tensor {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
  long long numel;
  float lr, weight_decay;
  int step;                 /* 1-based step count of THIS tensor (torch.optim.AdamW keeps one per parameter); 0 = use the call's `step` */
} msd_adamw_tensor;
int msd_grad_sqnorm_f32(const msd_adamw_tensor* tensors, int n, double* d_sqnorm, void* stream);
/* Weight repack after an optimizer step: the fp32 master weights -> the bf16 / transposed / tap-flipped operands of
 * msd_train_weights, one launch per 24 ops instead of a cast / permute / flip / pad launch per tensor.  `ops` is a HOST array.
 *   CAST            dst bf16 [d0*d1]            = src
 *   TRANSPOSE       dst bf16 [d1, d0]           = src[d0, d1]^T                       (W^T of a Linear, data gradient)
 *   CONV3           dst bf16 [d0, 9*kpad]       = src[Cout=d0, Cin=d1, 3, 3] tap-major (ky, kx, ci), ci >= Cin zero
 *   CONV3_DGRAD     dst bf16 [d1, T*kpad]       : Wd[ci, tap, co] = W[co, ci, 8 - tap]; flag != 0: T = 10, tap 9 = identity
 *   CONVT           dst bf16 [4*d1, kpad]       = src[Cin=d0, Cout=d1, 2, 2] as rows (di, dj, o), ci >= Cin zero
 *   CONVT_T         dst bf16 [d0, 4*d1]         = the transpose of the un-padded CONVT matrix
 *   REPEAT_F32      dst fp32 [flag*d0]          = src[d0] repeated flag times                (ConvTranspose bias per sub-pixel) */
enum { MSD_REPACK_CAST = 0, MSD_REPACK_TRANSPOSE = 1, MSD_REPACK_CONV3 = 2, MSD_REPACK_CONV3_DGRAD = 3, MSD_REPACK_CONVT = 4,
       MSD_REPACK_CONVT_T = 5, MSD_REPACK_REPEAT_F32 = 6 };
typedef struct msd_repack_op {
  const float* src; void* dst;
  int kind, d0, d1, kpad, flag;
} msd_repack_op;
int msd_repack_weights(const msd_repack_op* ops, int n, void* stream);
int msd_adamw_step_f32(const msd_adamw_tensor* tensors, int n, float beta1, float beta2, float eps, int step, float max_grad_norm,
                       const double* d_sqnorm, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSD_B200_H_ */
