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
} ur_conv_desc;

int ur_conv_gemm(const ur_conv_desc* desc_host, void* stream);
/* N tile the auto heuristic picks for (n, m_tiles); weight packers for gated acts must use it. */
int ur_conv_gemm_pick_bn(int n, int gated);

#ifdef __cplusplus
}
#endif
#endif /* UNIRESTORE_B200_H */
