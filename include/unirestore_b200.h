/*
 * unirestore_b200 C-ABI  --  the drop-in boundary of the B200-native UniRestore hot path.
 *
 * The reference (unirestore/UniRestore) has NO plugin/FFI layer: its operator boundary is the
 * Python nn.Module surface of src/modules/diffuie.  The host side of this
 * repo (unirestore_b200/*.py) mirrors that surface and reaches the GPU only through the symbols
 * declared here (ctypes binding: unirestore_b200/_cabi.py).
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - activations are bf16, channels-last (NHWC; a [tokens, C] matrix is the H=1 case);
 *     statistics / small vectors / latents are fp32;
 *   - the caller (PyTorch) owns every buffer; kernels never allocate and never synchronise, so
 *     every call is CUDA-graph capturable; `stream` is a cudaStream_t passed as void*;
 *   - return value 0 = success, negative = error (ur_last_error() gives the text, thread-local).
 *
 * Each entry point cites the reference interface (file:line under /root/reference/src/modules/diffuie,
 * or the diffusers symbol the reference calls there) whose arithmetic it implements.
 */
#ifndef UNIRESTORE_B200_H
#define UNIRESTORE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UR_OK 0
#define UR_ERR_ARG (-1)     /* bad shape / alignment / unsupported configuration */
#define UR_ERR_CUDA (-2)    /* CUDA runtime / driver error (launch failure, ...)  */

/* activation codes of the GEMM epilogue */
#define UR_ACT_NONE 0
#define UR_ACT_SILU 1
#define UR_ACT_GELU 2       /* exact erf GELU (nn.GELU() default; scedit.py:32, taskeditor.py:33) */
#define UR_ACT_GEGLU 3      /* out = a * gelu(g); tile columns [0,bn/2)=a, [bn/2,bn)=g (diffusers GEGLU) */
#define UR_ACT_GATE 4       /* out = a * g (SimpleGate, nafnet_arch.py:22-25), same column pairing */

#define UR_DT_BF16 0
#define UR_DT_F32 1

int ur_init(int device);
const char* ur_last_error(void);
int ur_version(void);

/* ------------------------------------------------------------------------------------------------
 * ur_conv_gemm -- tcgen05/TMEM implicit-GEMM convolution / linear layer with fused epilogue.
 *   out[b, y, x, n] = epilogue( alpha * sum_{tap, c} X[b, y*stride+dy[tap], x*stride+dx[tap], c] * W[n, tap*kc + c] )
 *   epilogue(v) = act(v + bias[n] + rowvec[b, n]) * chscale[b, n] + residual[b, y, x, n]
 * X is the channel-concatenation of up to two NHWC bf16 sources (torch.cat of base_model.py:189,197 and
 * taskeditor.py:101 is never materialised); out-of-image taps read zeros (conv padding).
 * Replaces: nn.Conv2d 3x3 / 1x1 and nn.Linear everywhere on the path -- diffusers ResnetBlock2D.conv1/conv2/
 * conv_shortcut (base_model.py:54), Downsample2D / Upsample2D (base_model.py:148,202; autoencoder.py:19,60),
 * Attention.to_q/k/v/to_out and FeedForward (base_model.py:138), CSCEAdapter.proj/tuner (scedit.py:28-38),
 * NAFBlock.conv1/3/4/5 (nafnet_arch.py:32-95), AdaNAFV2.conv_in/group_conv/pwconv (cfrm.py:18-36),
 * TaskFeatureAdapter gates / t_gate / conv_out (taskeditor.py:20-68).
 * ---------------------------------------------------------------------------------------------- */
typedef struct ur_conv_desc {
  const void* x1;      /* bf16 NHWC source 1 */
  const void* x2;      /* bf16 NHWC source 2 (channel concat) or NULL */
  int c1, c2;          /* channels taken from each source (c1 % 64 == 0 when c2 > 0; c % 8 == 0) */
  int ld1, ld2;        /* channel pitch of each source in elements (>= c) */
  int batch, hin, win; /* input extent */
  const void* w;       /* bf16 [n, ntaps*kc] (K contiguous); kc = c1+c2, or group_kc for grouped conv */
  int w_batched;       /* 1: w has a leading [batch] dimension (batched GEMM, tile never spans images) */
  int64_t w_ld;        /* row pitch of w in elements (0 = dense: ntaps*kc) */
  int64_t w_bs;        /* batch stride of w in elements (0 = dense: n*w_ld) */
  int n;               /* GEMM N (= output channels; gated acts store n/2 channels) */
  int ntaps;           /* 1..9 */
  int tap_dy[9];
  int tap_dx[9];
  int stride;          /* 1 or 2: input coordinate = out*stride + d */
  int group_kc;        /* grouped conv: input channels per group (multiple of 64); 0 = dense */
  int group_nc;        /* grouped conv: output channels per group (multiple of the N tile) */
  int hout, wout;      /* output positions computed per image */
  void* out;
  int out_dtype;       /* UR_DT_BF16 / UR_DT_F32 */
  int64_t out_sb, out_sy, out_sx; /* element strides of out along batch / y / x (channel stride 1) */
  float alpha;
  const float* bias;   /* [n] or NULL */
  const float* rowvec; /* [*, n] per-image additive vector (time embedding projection) or NULL */
  int64_t rowvec_sb;   /* batch stride of rowvec (0 = broadcast) */
  const float* chscale;/* [*, n_out] per-(image,) channel multiplier or NULL */
  int64_t chscale_sb;
  const void* residual;/* bf16, added last, or NULL */
  int64_t res_sb, res_sy, res_sx;
  int act;             /* UR_ACT_* */
  int bn;              /* N tile: 0 = auto, else one of 64/128/160/256 (gated acts: caller packs for it) */
  double* stats;       /* optional: per-(image, output channel) (sum, sumsq) of the bf16 OUTPUT are ACCUMULATED into
                          stats[(b * stats_ld + stats_off + n) * 2 + {0,1}] (caller zeroes it): the statistics pass of
                          the GroupNorm that consumes this output, fused into the epilogue (ur_chan_stats layout) */
  int stats_ld, stats_off;
  void* workspace;     /* optional caller-owned scratch (16-byte aligned) enabling split-K for problems with few output
                          tiles and a long K (the UNet's 8x8 level): fp32 partial sums, 4*batch*hout*wout*n bytes */
  int64_t workspace_bytes;
} ur_conv_desc;

int ur_conv_gemm(const ur_conv_desc* desc_host, void* stream);
/* N tile the auto heuristic picks for (n, m_tiles); weight packers for gated acts must use it. */
int ur_conv_gemm_pick_bn(int n, int gated);
/* Development switch (A/B timing): 1 routes every call to the non-persistent kernel; returns the previous value. */
int ur_debug_force_gemm_v1(int on);
/* Development: device buffer of 128 int64 that receives per-role clock64 timestamps of CTA 0 (NULL = off). */
int ur_debug_set_gemm_trace(void* buf);
/* Development: CTA-pair (tcgen05 cta_group::2) GEMM tiles: -1 auto (cost model), 0 never, 1 whenever legal. */
int ur_debug_set_gemm_pair_mode(int mode);
/* Development: 0 disables split-K (ur_conv_desc.workspace is then ignored); returns the previous value. */
int ur_debug_set_gemm_splitk(int on);
/* Development: persistent-kernel epilogue stores: 0 = st.global, 1 = one TMA store per 128-row sub-block (group
 * barrier), 2 = default: one TMA store per epilogue warp (32-row boxes, no barrier on the store path) in the general
 * kernel; launches whose epilogue is bias / residual / statistics only run the lean instantiation, whose stores are
 * issued by one store warp per epilogue group (128-row boxes, mbarrier hand-off).  0 and 1 force the general kernel.
 * Returns the previous value. */
int ur_debug_set_gemm_tma_store(int on);
int ur_debug_set_attention_trace(void* buf);   /* 64 int64 */
/* Development: 0 = head_dim 64 launches whose keys fit one 128-key tile (cross-attention on the prompt) run the general
 * kernel instead of the query-tile-loop kernel; returns the previous value. */
int ur_debug_set_attention_kv1(int on);
/* Development: attention kernel generation for head_dim 64 / 128: 1 = attention_kernel (default: two softmax threads per
 * row, two CTAs per SM), 2 = attention2_kernel (two query tiles per CTA, one softmax thread per row, single TMEM pass, P in
 * tensor memory, exp2 partly on the FMA pipe; measured equal, profiles/attention_experiments_r2.txt); returns the previous
 * value. */
int ur_debug_set_attention_impl(int impl);
/* Development: how many of every 8 exponential pairs attention2_kernel evaluates with the FMA-pipe polynomial
 * instead of MUFU.EX2 (0..3); returns the previous value. */
int ur_debug_set_attention_poly(int n);

/* ------------------------------------------------------------------------------------------------
 * Normalisation (HBM-bound, bf16 channels-last, 128-bit vectorised)
 * stats layout: double [batch][stats_ld channels][2] = (sum, sum of squares) over the pixels of one image.
 * ---------------------------------------------------------------------------------------------- */
/* Per-(image, channel) sums: first half of nn.GroupNorm (diffusers ResnetBlock2D.norm1/norm2, Attention.group_norm,
 * Transformer2DModel.norm, conv_norm_out; AdaNAFV2.group_norm cfrm.py:19), nn.InstanceNorm2d (taskeditor.py:31,40,49)
 * and nn.AdaptiveAvgPool2d(1) (nafnet_arch.py:62, cfrm.py:24,30, taskeditor.py:35,44,53). */
int ur_chan_stats(const void* x, int64_t ld, int64_t img_stride, int batch, int pixels, int channels, double* stats,
                  int stats_ld, int stats_off, int zero_first, void* stream);
/* out = [silu]( (cat(x1,x2) - mean_g) * rstd_g * gamma + beta ), group statistics from ur_chan_stats (or from the
 * ur_conv_gemm epilogue).  stats2 == NULL: stats is [batch][c1+c2][2]; else stats is [batch][c1][2] and stats2 is
 * [batch][c2][2] (each source carries the statistics its producer accumulated).
 * gamma / beta are PARAMETERS: the kernel reads them before its programmatic-dependency wait, so they must not be
 * written by a kernel launched just before on the same stream (x1, x2 and the statistics may be). */
int ur_norm_apply(const void* x1, int64_t ld1, int64_t is1, int c1, const void* x2, int64_t ld2, int64_t is2, int c2,
                  const double* stats, const double* stats2, int groups, int batch, int pixels, const float* gamma, const float* beta,
                  float eps, int silu, void* out, int64_t ldo, int64_t iso, void* stream);
/* One-launch nn.GroupNorm (+SiLU) over cat(x1, x2): a thread-block cluster per image keeps the partial statistics in
 * distributed shared memory, so there is no statistics array and no second launch (ur_groupnorm_cluster.cu).  Same
 * reference call sites as ur_norm_apply; meant for tensors that stay L2-resident between its two passes. */
int ur_group_norm(const void* x1, int64_t ld1, int64_t is1, int c1, const void* x2, int64_t ld2, int64_t is2, int c2,
                  int groups, int batch, int pixels, const float* gamma, const float* beta, float eps, int silu,
                  void* out, int64_t ldo, int64_t iso, void* stream);
/* cluster size ur_group_norm launches with on this device (16 when the non-portable size is available, else 8). */
int ur_group_norm_cluster_size(void);
int ur_debug_set_group_norm_cluster(int n);   /* development: override the cluster size (0 = probe) */
/* nn.LayerNorm over the channel dim of every token (BasicTransformerBlock.norm1-3; timm LayerNorm2d nafnet_arch.py:97-98). */
int ur_layernorm(const void* x, int64_t ldx, void* out, int64_t ldo, int64_t rows, int channels, const float* gamma,
                 const float* beta, float eps, void* stream);
/* x[b,p,c] *= scale[b,c] in place (nafnet_arch.py:122 `x * self.sca(x)`; cfrm.py:49-52). */
int ur_scale_channels(void* x, int64_t ld, int64_t img_stride, int batch, int pixels, int channels, const float* scale,
                      int scale_ld, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Attention helpers (unfused path: S = QK^T by ur_conv_gemm, softmax here, O = P V by ur_conv_gemm)
 * diffusers Attention / F.scaled_dot_product_attention (controller.py:183-185, VAE mid block, base_model.py:138).
 * ---------------------------------------------------------------------------------------------- */
/* Fused flash-style attention on tcgen05/TMEM (head_dim 64 or 128): out[b,q,h*d:(h+1)*d] = softmax(q k^T * scale) v.
 * q/k/v/out are bf16 token matrices addressed as ptr + b*bs + token*ld + h*head_dim (channel-slice views of a packed
 * qkv buffer are fine); kv_shared = 1: k/v have no batch dimension (the constant null prompt, base_model.py:221). */
int ur_attention(const void* q, int64_t ldq, int64_t q_bs, const void* k, int64_t ldk, int64_t k_bs, const void* v,
                 int64_t ldv, int64_t v_bs, void* out, int64_t ldo, int64_t out_bs, int batch, int heads, int head_dim,
                 int tq, int tk, int kv_shared, float scale, void* stream);
int ur_softmax_rows(const float* scores, int64_t ld_s, void* probs, int64_t ld_p, int64_t rows, int n_valid, int n_pad,
                    void* stream);
int ur_transpose_tokens(const void* x, int64_t ld, int64_t batch_stride, int batch, int tokens, int dim, void* out,
                        int tokens_pad, void* stream);

/* ------------------------------------------------------------------------------------------------
 * CFRM / TFA specific small kernels
 * ---------------------------------------------------------------------------------------------- */
/* NAFBlock: y = SimpleGate(dwconv3x3(x)) and GAP sums of y (nafnet_arch.py:41-49,62,115-122). weight fp32 [2c,9]. */
int ur_dwconv3x3_gate(const void* x, int batch, int h, int w, int c, const float* weight, const float* bias, void* y,
                      double* stats, void* stream);
/* y[b,n] = act_out(W[n,:] . act_in(x[b, group(n)*k : +k]) + bias[n]); acts: 0 none 1 silu 2 gelu 3 tanh.
 * in_mode 1: x is a ur_chan_stats array and the input is sum*in_scale (pooled mean).
 * TimestepEmbedding / time_emb_proj (base_model.py:104-106, controller.py:196-197), NAFBlock.sca (nafnet_arch.py:61-65). */
int ur_small_linear(const void* x, int in_mode, float in_scale, int64_t x_ld, const float* w, const float* bias,
                    float* y, int64_t y_ld, int batch, int n, int k, int groups, int act_in, int act_out, void* stream);
/* diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): out[b] = [cos(t f), sin(t f)]. */
int ur_timestep_embedding(const int64_t* timesteps, int batch, int dim, float* out, void* stream);
/* AdaNAFV2 intra-/inter-group attention -> scale[b, 4c] (cfrm.py:22-34,47-52). */
int ur_adanaf_scales(const double* stats, int pixels, int batch, int c4, int groups, const float* w_intra,
                     const float* b_intra, const float* w_inter, const float* b_inter, float* scale, void* stream);
/* TaskFeatureAdapter prompt update (taskeditor.py:78-106): pooled gate branches -> o [B,D], cond_next [B,T,D/2]. */
int ur_tfa_gates(const double* stats, int pixels, int batch, int prompt_len, int dim, const float* cond,
                 const float* w_out, const float* b_out, const float* w_pt, const float* b_pt, float* o,
                 float* cond_next, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Latent-space elementwise (fp32 NCHW latents [B,4,h,w]; *_nhwc8 = bf16 channels-last copy padded to 8 ch)
 * ---------------------------------------------------------------------------------------------- */
/* z = (mean + exp(0.5*clamp(logvar,-30,20)) * noise) * scaling_factor (autoencoder.py:152-155). */
int ur_posterior_sample(const float* moments, const float* noise, float scaling_factor, int batch, int hw, float* z,
                        void* z_nhwc8, void* stream);
/* out = a*x + b*y (DDPMScheduler.add_noise, unifie.py:88); out_nhwc8 = bf16(scale8 * out) (autoencoder.py:170). */
int ur_latent_axpby(const float* x, float a, const float* y, float b, int batch, int hw, float* out, void* out_nhwc8,
                    float scale8, void* stream);
/* DDIMScheduler.step, eta = 0 (unifie.py:150), in place on x. */
int ur_ddim_step(float* x, const float* eps, int ld_eps, float sqrt_alpha_t, float sqrt_one_minus_alpha_t,
                 float sqrt_alpha_prev, float sqrt_one_minus_alpha_prev, int clip_sample, int batch, int hw,
                 void* x_nhwc8, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Image layout (fp32 NCHW images <-> channels-last)
 * ---------------------------------------------------------------------------------------------- */
/* out bf16 [B,H,W,8] = a*img + b for the first `channels`, zeros after (autoencoder.py:151 `x*2-1`). */
int ur_image_to_nhwc8(const float* img, int64_t sb, int64_t sc, int64_t sy, int64_t sx, int batch, int channels, int h,
                      int w, float a, float b, void* out, void* stream);
/* out fp32 [B,C,h,w] = q(a*src[:, y0:y0+h, x0:x0+w, :C] + b) (autoencoder.py:175 `(x+1)/2`, crop of unifie.py:164);
 * quantize != 0: q(v) = clamp(round(255 v), 0, 255) / 255, the 8-bit quantisation the validate loop applies to every
 * prediction (eval_image_restoration.py:71), fused into the image write-out. */
int ur_nhwc_to_image(const float* src, int ld, int hs, int ws, int batch, int channels, int h, int w, int y0, int x0,
                     float a, float b, int quantize, float* out, void* stream);

/* out[r, :] = cat(x1[r, :c1], x2[r, :c2]) on bf16 rows -- torch.cat(dim=1) of base_model.py:189,197 materialised; only
 * used when source 1 of a two-source convolution is not 64-channel aligned (never on the sd-turbo shapes). */
int ur_concat_channels(const void* x1, int64_t ld1, int c1, const void* x2, int64_t ld2, int c2, int64_t rows, void* out,
                       void* stream);

/* Image pre / post around the path (unifie.py:124-134,165-168): out = reflect_pad_{bottom,right}(bicubic_resize(img,
 * (hr, wr))) on fp32 NCHW (arbitrary input strides in elements), F.interpolate(mode="bicubic", align_corners=False,
 * antialias=False) + F.pad(mode="reflect") semantics; hr == hin && wr == win skips the resize.
 * out: dense fp32 [batch, channels, hr + pad_b, wr + pad_r]; quantize != 0 applies the 8-bit quantisation of
 * eval_image_restoration.py:71 to the result (the resize back to the input resolution is the last op of the path). */
int ur_resize_pad(const float* img, int64_t sb, int64_t sc, int64_t sy, int64_t sx, int batch, int channels, int hin,
                  int win, int hr, int wr, int pad_b, int pad_r, int quantize, float* out, void* stream);

/* Metrics step right after the path (eval_image_restoration.py:71,255-313: 8-bit quantisation of the prediction,
 * skimage peak_signal_noise_ratio / structural_similarity(win 7, uniform, sample covariance, channel_axis 0)):
 * out[b] = { sum (t-p)^2 over the image, sum of the SSIM map over channels and the 3-pixel-cropped interior }, fp64.
 * pred / target: dense fp32 [batch, channels, h, w] on the device. */
int ur_image_metrics(const float* pred, const float* target, int batch, int channels, int h, int w, int quantize_pred,
                     float data_range, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNIRESTORE_B200_H */
