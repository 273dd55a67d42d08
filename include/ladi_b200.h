/* ladi_b200.h -- C ABI of libladi_b200.so: hand-written sm_100a kernels for the LaDI-VTON try-on hot path.
 *
 * The reference (miccunifi/ladi-vton) has no FFI of its own: it is pure Python and every GPU operation on its hot path is
 * a PyTorch/cuDNN/cuBLAS library call made from `StableDiffusionTryOnePipeline.__call__`
 * (/root/reference/src/vto_pipelines/tryon_pipe.py:494-765).  This header is the boundary a maintainer binds (ctypes, see
 * INTEGRATION.md) to replace those library calls.  Each entry point names the reference call-site(s) it replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are CALLER-OWNED DEVICE pointers (e.g. torch.Tensor.data_ptr());
 *  - activations are NHWC bf16 (`pitch` = elements between consecutive pixels, multiple of 8 => 16-byte aligned rows);
 *  - `stream` is a cudaStream_t passed as void*; nothing allocates, synchronises or uses an implicit stream, so every call
 *    is CUDA-graph capturable;
 *  - return 0 (LADI_OK) on success; otherwise an error code, with a thread-local message in ladi_last_error();
 *    no exceptions/aborts cross the ABI;
 *  - there is NO CPU fallback: without a CUDA device every compute entry point returns LADI_ERR_CUDA.
 */
#ifndef LADI_B200_H
#define LADI_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LADI_API __attribute__((visibility("default")))
#else
#define LADI_API
#endif

#define LADI_OK 0
#define LADI_ERR_INVALID 1
#define LADI_ERR_CUDA 2

#define LADI_ACT_NONE 0
#define LADI_ACT_SILU 1
#define LADI_ACT_GEGLU 2
#define LADI_ACT_GELU 3 /* GELU(erf): CLIP encoder-layer MLP and the adapter's projection MLP (inversion_adapter.py:12-20) */
#define LADI_ACT_RELU 4 /* warping module: ConvNet_TPS.py:32-42,93-105, unet_parts.py:15-22 */

LADI_API const char* ladi_last_error(void);
LADI_API int ladi_abi_version(void);
/* number of kernels this process has launched (or recorded into a CUDA graph under capture) through the library so far */
LADI_API long long ladi_launch_count(void);

/* ---- implicit-GEMM convolution / GEMM (tcgen05 + TMA) --------------------------------------------------------------
 * Replaces nn.Conv2d (3x3 s1 p1, 3x3 s2, 1x1) and nn.Linear of: diffusers ResnetBlock2D / Transformer2DModel /
 * Downsample2D / Upsample2D inside UNet2DConditionModel.forward (call-site tryon_pipe.py:732), the VAE Encoder/Decoder
 * (src/models/vae.py:99-119, 183-212; AutoencoderKL.py:151,163) and EMASC (src/models/emasc.py:25-40).
 *
 * out[n,y,x,:] = epilogue( sum_{tap,src,c} in_src[n, y*stride+ky-pad_lo, x*stride+kx-pad_lo, c] * W[:, k(tap,src,c)]
 *                          + sum_{sc,c} sc[n,y,x,c] * W[:, k(sc,c)] )
 * The sources are concatenated along channels (UNet skip `torch.cat([x, skip], 1)`, diffusers CrossAttnUpBlock2D) WITHOUT
 * materialising the concat; the optional `sc` sources are the ResnetBlock2D 1x1 `conv_shortcut` folded into the same
 * accumulation.  Weight layout: bf16 [c_out][k_total], K order = for tap(ky,kx) row-major: for src: channels padded to a
 * multiple of 64 (zero weights in the padding); then for sc: channels padded to 64.  k_total = 64 * (#K blocks).
 * A GEMM out[M,N] = A[M,K] W[N,K]^T is ksize=1, n=1, h_out=1, w_out=M, src_c=K.
 * epilogue: +bias[c] (or +bias[row] if bias_per_row; bias += *step_ptr * bias_step_stride when step_ptr != NULL),
 *           act (SiLU, or GEGLU on interleaved (value,gate) column pairs -> c_out/2 outputs), +residual, *row_scale[row]. */
typedef struct ladi_conv_desc {
  int n, h_out, w_out, c_out;
  int h_in, w_in;            /* 0 => h_out*stride, w_out*stride */
  int ksize, stride, pad_lo; /* ksize 1|3, stride 1|2, pad_lo = top/left padding (1; 0 for the VAE (0,1,0,1) downsample) */
  int n_src;
  const void* src[2];
  int src_c[2];
  int src_pitch[2];
  int n_sc;
  const void* sc[2];
  int sc_c[2];
  int sc_pitch[2];
  const void* weight;
  int k_total;
  int weight_pitch;          /* elements between weight rows (>= k_total, multiple of 8) */
  const float* bias;
  int bias_per_row;
  int bias_step_stride;
  const int* step_ptr;
  const void* residual;
  int residual_pitch;
  const float* row_scale;
  int act;
  void* out;
  int out_pitch;
  int out_fp32;
  int force_bn;              /* 0 = auto; else N tile in {32,64,128,160,192,256} (tests / tuning) */
  int force_direct_epilogue; /* 1 = direct-store epilogue even where the staged TMA-store epilogue applies (tests) */
  void* splitk_ws;           /* optional fp32 workspace: enables split-K for few-tile / long-K shapes (NULL = never split) */
  int64_t splitk_ws_bytes;
  int pair_mode;             /* CTA pairs (tcgen05 cta_group::2, M = 256 across the two SMs of a TPC, each staging half of B):
                                0 = library default (pairs when the shape allows; env LADI_CONV_2CTA=0 disables), 1 = force (error if the shape
                                cannot pair), 2 = never */
  /* LayerNorm folded into the GEMMs on either side of it (BasicTransformerBlock.norm1/2/3 of the UNet, call-site tryon_pipe.py:732):
   * the PRODUCER of the normalised tensor also writes, per output row and 32-column chunk, {sum, sum of squares} of what it stores
   * (rowstat_out: fp32 [rows][c_out/32][2]); the CONSUMER multiplies the RAW tensor with W' = W diag(gamma) and corrects in its epilogue:
   * out[r,c] = rstd_r * (acc[r,c] - mean_r * ln_colsum[c]) + bias[c], bias = W beta (+ the linear's own bias); the LayerNorm pass and its
   * HBM round trip disappear.  ln_stats is the producer's rowstat_out for this GEMM's A rows (normalised width = src_c[0], %% 64 == 0). */
  float* rowstat_out;
  const float* ln_stats;
  const float* ln_colsum;    /* fp32 [c_out]: row sums of the bf16 weight actually multiplied */
  float ln_eps;
  /* nearest-2x upsample fused into the 3x3 convolution that follows it (diffusers Upsample2D: F.interpolate(scale_factor=2) + conv;
   * UNet up blocks and VAE decoder vae.py:142-174): src = the HALF-resolution tensor [n, h_out/2, w_out/2, C]; each output parity
   * (y&1, x&1) is a 2x2-tap convolution of it with merged weights (weights.pack_conv_up2x: [4 parities * c_out][4 taps * Cpad]),
   * 4/9 of the multiply-adds of the materialised form and no intermediate tensor.  ksize 3, stride 1, bf16 output. */
  int up2x;
} ladi_conv_desc;
LADI_API int ladi_conv2d_bf16(const ladi_conv_desc* d, void* stream);

/* ---- fused multi-head attention, head_dim 64 (flash-style: S and O live in TMEM, online softmax in exp2) ------------
 * Replaces diffusers CrossAttention.forward for attn1 (self) and attn2 (cross, 77 text tokens) of every
 * BasicTransformerBlock in the UNet (torch SDPA / xformers in the reference, src/inference.py:143-147).
 * element (b, i, h, e) of q is q[b*q_batch_stride + i*q_pitch + h*64 + e]; same for k, v, out.  No mask. */
typedef struct ladi_attn_desc {
  int batch, heads, nq, nkv;
  const void* q; int q_pitch; int64_t q_batch_stride;
  const void* k; int k_pitch; int64_t k_batch_stride;
  const void* v; int v_pitch; int64_t v_batch_stride;
  void* out; int out_pitch; int64_t out_batch_stride;
  float scale;               /* softmax(scale * q k^T) */
  int variant;               /* 0 = auto (nkv <= 128: 8; nq >= 512: 5; else 1); 1 = one query tile per CTA; 2 / 4 / 5 / 6 = two query tiles per CTA (P in smem / in TMEM /
                                + lazy single-pass softmax / + ping-pong experiment); 8 = persistent kernel for a single K/V tile (tests / tuning) */
  void* trace;               /* optional device int64[8192]: per-phase clock64() stamps of CTA (0,0,0) (tools/attn_trace.py); NULL */
  int head_dim;              /* ladi_attention_d512_bf16 only: 0 or 512 (the VAE), or 256 (reduced-width test models) */
} ladi_attn_desc;
LADI_API int ladi_attention_bf16(const ladi_attn_desc* d, void* stream);
/* The same for ONE head of width 512: the VAE mid-block AttentionBlock (src/models/vae.py:81-90,112 encoder; :142-150,187 decoder;
 * N = h*w tokens: 3072 at 512x384, 12288 at 1024x768).  heads must be 1, pitches >= head_dim; `variant` / `trace` ignored.  Flash-style:
 * the N x N score matrix is never written (the reference's baddbmm + softmax + bmm materialises it per sample). */
LADI_API int ladi_attention_d512_bf16(const ladi_attn_desc* d, void* stream);

/* ---- normalisation ---------------------------------------------------------------------------------------------------
 * GroupNorm(+SiLU) over the channel-concat of up to two NHWC sources (diffusers ResnetBlock2D.norm1/norm2,
 * Transformer2DModel.norm, conv_norm_out; VAE conv_norm_out vae.py:115-116,200-201), two passes:
 *   stats: partial sums per (image, pixel chunk, group) -> ws[n][chunks][groups][2] fp32 (chunks returned by the query);
 *   apply: y = (x-mean)*rstd*gamma+beta, optional SiLU, optional "+ add" AFTER the activation (vae.py:204-205), bf16 out
 *          of pitch out_pitch holding the concatenated, normalised tensor. */
LADI_API int ladi_groupnorm_chunks(int hw);
LADI_API int ladi_groupnorm_stats(const void* x0, int c0, int pitch0, const void* x1, int c1, int pitch1, int n, int hw, int groups,
                         float* ws, void* stream);
LADI_API int ladi_groupnorm_apply(const void* x0, int c0, int pitch0, const void* x1, int c1, int pitch1, int n, int hw, int groups,
                         const float* ws, const float* gamma, const float* beta, float eps, int silu, const void* add,
                         int add_pitch, void* out, int out_pitch, void* stream);
/* LayerNorm over the last dim (BasicTransformerBlock.norm1/2/3; inversion adapter LayerNorms). */
LADI_API int ladi_layernorm(const void* x, int x_pitch, int rows, int c, const float* gamma, const float* beta, float eps, void* out,
                   int out_pitch, void* stream);
/* Row softmax fp32 -> bf16 (VAE mid-block AttentionBlock: softmax in fp32, Appendix A.5). */
LADI_API int ladi_softmax_rows(const float* s, int rows, int cols, int s_pitch, float scale, void* out, int out_pitch, void* stream);

/* CLS-row attention of the inversion adapter's CLIP ViT-H encoder layer (16 heads, head_dim 80, 257 tokens): only token 0 of the
 * layer output is consumed (/root/reference/src/models/inversion_adapter.py:22-28), so only the CLS query is attended.
 * q0 [batch, heads*head_dim] bf16, kv [batch, tokens, >= 2*heads*head_dim] bf16 (K columns then V columns) -> out [batch, heads*head_dim]. */
LADI_API int ladi_cls_attention(const void* q0, int q_pitch, const void* kv, int kv_pitch, int batch, int tokens, int heads, int head_dim,
                       float scale, void* out, int out_pitch, void* stream);

/* ---- pointwise glue ------------------------------------------------------------------------------------------------- */
/* out = a + b (bf16, equal pitch-free dense [count]); VAE decoder `sample += int_feat` (vae.py:193). */
LADI_API int ladi_add_bf16(const void* a, const void* b, void* out, int64_t count, void* stream);
/* nearest 2x upsample NHWC (diffusers Upsample2D F.interpolate(scale_factor=2, mode="nearest")). */
LADI_API int ladi_upsample2x_nhwc(const void* x, int n, int h, int w, int c, void* out, void* stream);
/* NCHW fp32 [n,c,h*f,w*f] -> NHWC bf16 [n,h,w,pitch] channels [c_off, c_off+c): nearest sampling every f-th pixel
 * (F.interpolate default mode, tryon_pipe.py:434-436), times `scale`, optionally gated by (gate[n,0,..] < 0.5) -- the
 * `masked_image = image * (mask < 0.5)` of diffusers prepare_mask_and_masked_image (call-site tryon_pipe.py:630). */
LADI_API int ladi_nchw_f32_to_nhwc_bf16(const float* x, int n, int c, int h, int w, int f, float scale, const float* gate, void* out,
                               int out_pitch, int c_off, void* stream);
/* NHWC (bf16 or fp32) -> NCHW fp32, channels [c_off, c_off+c) of an NHWC tensor of given pitch. */
LADI_API int ladi_nhwc_to_nchw_f32(const void* x, int x_is_fp32, int n, int c, int h, int w, int x_pitch, int c_off, float* out,
                          void* stream);
/* DiagonalGaussianDistribution.sample * scaling_factor (vae.py:330-348; tryon_pipe.py:647,462):
 * moments NHWC fp32 [n,h,w,2*cz] -> latents NCHW fp32 [n,cz,h,w] = (mean + exp(0.5*clamp(logvar,-30,20))*noise) * scale. */
LADI_API int ladi_posterior_sample(const float* moments, int m_pitch, const float* noise_nchw, int n, int cz, int h, int w, float scale,
                          float* out_nchw, void* stream);
/* mask (1 - m) rows for EMASC mask_features (src/utils/data_utils.py:9-14): mask NCHW fp32 [n,1,H,W] nearest-resized by
 * integer factor f to [n, H/f, W/f] and written as 1-m (fp32 per output pixel). */
LADI_API int ladi_inv_mask_rows(const float* mask, int n, int H, int W, int f, float* out, void* stream);
/* bilinear /8 downsample, align_corners=False, no antialias (tryon_pipe.py:632-634), NCHW fp32 -> NCHW fp32. */
LADI_API int ladi_bilinear_down8(const float* x, int n, int c, int H, int W, float* out, void* stream);
/* One DDIM step with classifier-free guidance and re-assembly of the next UNet input (tryon_pipe.py:715,735-741 +
 * DDIMScheduler.step): eps NHWC fp32 [cfg?2B:B, h, w, eps_pitch] (first 4 channels), latents NCHW fp32 [B,4,h,w] updated
 * in place; unet_in NHWC bf16 [cfg?2B:B, h, w, in_pitch] channels 0..3 rewritten for both halves.  coef = device table
 * [steps][8] = {1/sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev-sigma_t^2), sigma_t, 0, 0, 0}, indexed by *step_ptr, which is
 * then incremented by the last block when advance != 0.  noise: NCHW fp32 [B,4,h,w] variance noise of the eta > 0 (stochastic)
 * DDIM update, x_prev += sigma_t * noise; NULL for eta = 0. */
LADI_API int ladi_ddim_cfg_step(const float* eps, int eps_pitch, float* latents, void* unet_in, int in_pitch, int B, int h, int w,
                       int cfg, float guidance, const float* coef, int* step_ptr, int advance, const float* noise, void* stream);
/* diffusers prepare_mask_and_masked_image, tensor branch (called at tryon_pipe.py:630), without host synchronisation: range checks
 * image in [-1,1] / mask in [0,1] recorded in flags[0] / flags[1] (device int32[2], OR-ed; the pipeline reads them once, with the result)
 * and the IN-PLACE binarisation of the caller's mask at 0.5.  image fp32 [n_image], mask fp32 [n_mask], both dense. */
LADI_API int ladi_check_binarise(const float* image, long long n_image, float* mask, long long n_mask, int* flags, void* stream);
/* (x/2+0.5).clamp(0,1) NHWC bf16/fp32 [n,h,w,pitch] (first 3 channels) -> NHWC fp32 [n,h,w,3] (tryon_pipe.py:356-358). */
LADI_API int ladi_image_out(const void* x, int x_is_fp32, int n, int h, int w, int x_pitch, float* out, void* stream);
/* The same clamp followed by numpy_to_pil's (x * 255).round().astype(uint8) (DiffusionPipeline.numpy_to_pil, called at
 * tryon_pipe.py:760; round-half-to-even like numpy): NHWC uint8 [n,h,w,3], so output_type="pil" moves a quarter of the bytes to the host. */
LADI_API int ladi_image_out_u8(const void* x, int x_is_fp32, int n, int h, int w, int x_pitch, unsigned char* out, void* stream);
/* Sinusoidal timestep embedding of the UNet (diffusers get_timestep_embedding, flip_sin_to_cos, the input of time_embedding.linear_1)
 * for n timesteps (device fp32 [n]), written as a (hi, lo) pair of bf16 column blocks: out bf16 [n, 2*k_pad], hi at columns [0, channels),
 * lo = bf16(e - hi) at [k_pad, k_pad + channels), zeros elsewhere (k_pad = channels rounded up to 64: the K layout of the packed [W | W] weight). */
LADI_API int ladi_timestep_embedding(const float* timesteps, int n, int channels, int k_pad, void* out, void* stream);
/* ---- dataset tensorisation edge: src/utils/posemap.py:6-35 kpoint_to_heatmap for n_maps key-points
 * (keypoints fp32 [n_maps,2] = (x, y) in pixels; the reference calls it per sample and joint with sigma 9, src/dataset/vitonhd.py:277-287):
 * out fp32 [n_maps,h,w] = exp(-|p - k|^2 / sigma^2) / (max + eps), all zeros for a key-point with no coordinate > 0. */
LADI_API int ladi_pose_heatmaps(const float* keypoints, int n_maps, int h, int w, float sigma, float* out, void* stream);

/* ---- text / vision conditioning front-end (not on the per-step path) ------------------------------------
 * Exact softmax attention for short sequences and any head width that is a multiple of 8 (<= 128), optional causal mask:
 * the CLIP text transformer driven by src/utils/encode_text_word_embedding.py:41-54 (77 causal tokens, 16 x 64) and the CLIP ViT-H
 * vision tower called at src/inference.py:269-273 (257 tokens, 16 x 80).  q/k/v/out: bf16 [batch, n, heads*head_dim] views with
 * row pitches and batch strides in elements (multiples of 8).  nkv <= 1024; causal needs nq == nkv. */
LADI_API int ladi_attention_small(const void* q, const void* k, const void* v, void* out, int batch, int heads, int nq, int nkv,
                         int head_dim, int q_pitch, int k_pitch, int v_pitch, int out_pitch, long long q_batch_stride,
                         long long k_batch_stride, long long v_batch_stride, long long out_batch_stride, float scale, int causal,
                         void* stream);
/* out[row,:] = (src[row] >= 0 ? tok[src[row],:] : word_emb[-src[row]-1,:]) + pos[row % seq,:]  (bf16 tables, fp32 sum):
 * token_embedding + the '$' -> pseudo-word-embedding substitution + position_embedding of
 * src/utils/encode_text_word_embedding.py:27-38 (the index list `src` is built on the host from input_ids). */
LADI_API int ladi_clip_embed(const int* src, const void* tok, const void* word_emb, const void* pos, void* out, int rows, int seq, int c,
                    int out_pitch, void* stream);
/* pixels NCHW fp32 [n,ch,h,w] -> bf16 rows [n*(h/patch)*(w/patch), k_pad], column (c*patch+ky)*patch+kx, zero padding columns:
 * the im2col of CLIPVisionEmbeddings.patch_embedding (14x14 stride-14 convolution, no bias), followed by one ladi_conv2d_bf16 GEMM. */
LADI_API int ladi_patchify(const float* pixels, void* out, int n, int ch, int h, int w, int patch, int k_pad, void* stream);
/* x[b,0,:] = cls + pos[0,:]; x[b,1+i,:] = patch[b*n_patches+i,:] + pos[1+i,:]  (CLIPVisionEmbeddings.forward), all bf16. */
LADI_API int ladi_vit_assemble(const void* patch, int patch_pitch, const void* cls, const void* pos, void* out, int n, int n_patches, int c,
                      void* stream);

/* ---- cloth-warping front-end: ConvNet_TPS + refinement U-Net, src/inference.py:236-266 -----------------
 * torchvision resize(x, (oh, ow), BILINEAR, antialias=True) of an NCHW fp32 tensor, written as channels [c_off, c_off+c) of an NHWC
 * bf16 tensor (inference.py:239-247: cloth / im_mask / pose_map -> 256x192; the agnostic concat is built in place). */
LADI_API int ladi_resize_aa(const float* x, int n, int c, int h, int w, int oh, int ow, void* out, int out_pitch, int c_off, void* stream);
/* src/inference.py:265-271: torchvision resize((cloth + 1) / 2, (oh, ow), antialias=True).clamp(0, 1) followed by the CLIP image
 * processor's per-channel (v - mean) / std (AutoProcessor call, :267): x NCHW fp32 in [-1,1] -> out NCHW fp32 [n,c,oh,ow], the
 * `pixel_values` of the vision tower.  quantise != 0: v = floor(v * 255) / 255 first (a processor that round-trips through uint8). */
LADI_API int ladi_clip_preprocess(const float* x, int n, int c, int h, int w, int oh, int ow, const float* mean, const float* stdev,
                         int quantise, float* out, void* stream);
/* NHWC bf16 [n,h,w,c] -> [n,h/2,w/2,4c], channel (sy*2+sx)*c+ch <- pixel (2y+sy, 2x+sx): the 4x4 stride-2 pad-1 convolutions of
 * FeatureExtraction / FeatureRegression (ConvNet_TPS.py:31-38,93-97) become 3x3 stride-1 pad-1 ladi_conv2d_bf16 calls. */
LADI_API int ladi_space_to_depth2(const void* x, int n, int h, int w, int c, int x_pitch, void* out, int out_pitch, void* stream);
/* in place x[row,c] = x[row,c]*scale[c] + shift[c]: eval-mode BatchNorm2d that follows a ReLU (ConvNet_TPS.py:33,39,41). */
LADI_API int ladi_channel_affine(void* x, long long rows, int c, int pitch, const float* scale, const float* shift, void* stream);
/* FeatureL2Norm (ConvNet_TPS.py:58-66), in place over the channel dimension of NHWC bf16 rows. */
LADI_API int ladi_l2norm_channels(void* x, long long rows, int c, int pitch, void* stream);
/* FeatureCorrelation (ConvNet_TPS.py:69-81): out NHWC bf16 [n,h,w,h*w], out[b,yB,xB,xA*h+yA] = <B[b,yB,xB,:], A[b,yA,xA,:]>. */
LADI_API int ladi_feature_correlation(const void* feat_a, const void* feat_b, int n, int h, int w, int c, void* out, int out_pitch, void* stream);
/* points = tanh(theta) (ConvNet_TPS.py:122); grid = target_coordinate_repr @ (inverse_kernel @ [points; 0]) (TPSGridGen.forward,
 * :183-193).  theta fp32 [n, >= 2*n_ctrl]; inverse_kernel [(n_ctrl+3)^2]; target_coordinate_repr [n_points, n_ctrl+3];
 * points fp32 [n, n_ctrl, 2]; grid fp32 [n, n_points, 2]. */
LADI_API int ladi_tps_grid(const float* theta, int theta_pitch, const float* inverse_kernel, const float* target_coordinate_repr, int n, int n_ctrl,
                  int n_points, float* points, float* grid, void* stream);
/* inference.py:252-257: antialiased-bilinear resize of low_grid fp32 [n,gh,gw,2] to (h,w) fused with F.grid_sample(cloth, grid,
 * bilinear, padding_mode='border', align_corners=False); cloth NCHW fp32 [n,c,h,w] -> channels [c_off,c_off+c) of NHWC bf16 out. */
LADI_API int ladi_warp_grid_sample(const float* low_grid, int gh, int gw, const float* cloth, int n, int c, int h, int w, void* out, int out_pitch,
                          int c_off, void* stream);
/* nn.MaxPool2d(2) / nn.Upsample(scale_factor=2, bilinear, align_corners=True) on NHWC bf16 (unet_parts.py:31-34,47). */
LADI_API int ladi_maxpool2_nhwc(const void* x, int n, int h, int w, int c, void* out, void* stream);
LADI_API int ladi_upsample2x_bilinear_ac(const void* x, int n, int h, int w, int c, void* out, void* stream);
/* NHWC fp32 [n,h,w,pitch] (first c channels) -> NCHW fp32 clamped to [lo,hi] (inference.py:262 warped_cloth.clamp(-1,1)). */
LADI_API int ladi_nhwc_f32_to_nchw_clamp(const float* x, int n, int c, int h, int w, int x_pitch, float lo, float hi, float* out, void* stream);

/* ================================================================================================================================
 * Module-level ABI: an opaque per-device engine handle + one entry point per module of the try-on path.
 * What a non-Python host binds to run the reference's objects without re-writing their forward passes:
 *   unet(x, t, encoder_hidden_states).sample           tryon_pipe.py:732 (diffusers UNet2DConditionModel.forward)   -> ladi_unet_forward
 *   vae.encode(x) -> (posterior moments, 6 skips)      tryon_pipe.py:640,457; AutoencoderKL.py:145-157; vae.py:99-119 -> ladi_vae_encode
 *   vae.decode(z, intermediate_features, int_layers)   tryon_pipe.py:352-353; AutoencoderKL.py:159-188; vae.py:183-212 -> ladi_vae_decode_emasc
 *   emasc(features) + mask_features(features, mask)    tryon_pipe.py:684-685; emasc.py:37-40; data_utils.py:4-16     -> ladi_emasc_forward
 *   inversion_adapter(clip_features)                   src/inference.py:276; inversion_adapter.py:22-28              -> ladi_inversion_adapter_forward
 *   the denoising loop                                 tryon_pipe.py:713-747                                         -> ladi_denoise_loop
 * Rules: all pointers are caller-owned device pointers; activations NHWC bf16 (channel pitch = multiple of 8); the caller supplies the
 * activation workspace (ladi_workspace_bytes) and the stream; no call allocates, synchronises or uses an implicit stream, so every call is
 * CUDA-graph capturable.  ladi_engine_create allocates the two scratch buffers all calls share (64 MB split-K partials, 16 MB GroupNorm
 * statistics) and ladi_engine_destroy frees them.  One handle per (device, stream); calls on one handle are not concurrent. */
typedef struct ladi_engine ladi_engine;

/* one packed weight: name (the key the Python packer uses, e.g. "down_blocks.0.resnets.0.w1"), device pointer, rows x cols
 * (bf16 matrices: cols = K = row pitch, layouts as documented at ladi_conv_desc; fp32 vectors: rows = 1) */
typedef struct ladi_weight {
  const char* name;
  const void* ptr;
  int rows, cols;
} ladi_weight;

typedef struct ladi_engine_config {
  /* UNet2DConditionModel (SD-2-inpainting layout, hubconf.py:30-33) */
  int unet_channels[4];        /* block_out_channels */
  int unet_heads[4];           /* attention_head_dim = head COUNT per level in SD-2 configs (head width is always 64) */
  int unet_down_attn[4];       /* down block i is a CrossAttnDownBlock2D */
  int unet_up_attn[4];         /* up block i is a CrossAttnUpBlock2D */
  int unet_layers_per_block;
  int unet_in_channels, unet_out_channels;
  float unet_norm_eps;
  int norm_groups;             /* GroupNorm groups (32), UNet and VAE */
  int fuse_upsample;           /* 1: upsampler weights are packed for the sub-pixel form (ladi_conv_desc.up2x) */
  /* AutoencoderKL (LaDI-VTON fork, src/models/AutoencoderKL.py) */
  int vae_channels[4];
  int vae_layers_per_block, vae_latent_channels, vae_in_channels, vae_out_channels;
  /* EMASC (hubconf.py:40-53) */
  int emasc_scales;
  int emasc_in[8], emasc_out[8], emasc_stride[8];   /* per scale: channels in / out, resolution divisor of the feature (1, 1, 2, 4, 8) */
  /* InversionAdapter (hubconf.py:16-27) */
  int adapter_dim, adapter_heads, adapter_mlp, adapter_hidden, adapter_out;
  int plan_only;               /* 1: no device: the handle only answers ladi_workspace_bytes / ladi_engine_trace (tests, tooling) */
} ladi_engine_config;

LADI_API int ladi_engine_create(const ladi_engine_config* cfg, const ladi_weight* table, int n_weights, ladi_engine** out);
LADI_API int ladi_engine_destroy(ladi_engine* engine);
#define LADI_Q_TEMB_TOTAL 0 /* columns of the per-step bias table (sum of the resnets' widths) */
#define LADI_Q_KV_TOTAL 1   /* columns of the text K/V buffer (sum over cross-attention layers of 2C) */
#define LADI_Q_IN_PITCH 2   /* channel pitch of the UNet input buffer (in_channels rounded up to 8) */
LADI_API int ladi_engine_query(const ladi_engine* engine, int what);

#define LADI_MODULE_UNET 0        /* (batch = UNet batch, height = latent h, width = latent w) */
#define LADI_MODULE_VAE_ENCODE 1  /* (batch, image height, image width) */
#define LADI_MODULE_VAE_DECODE 2  /* (batch, latent h, latent w) */
#define LADI_MODULE_EMASC 3       /* (batch, image height, image width) */
#define LADI_MODULE_ADAPTER 4     /* (batch, height = tokens, width ignored) */
#define LADI_MODULE_UNET_PLAN 5   /* ladi_unet_plan_steps: (batch = number of timesteps, height / width ignored) */
/* bytes of caller workspace one call of `module` needs at this shape (a walk of the module body with the allocator only); -1 on error */
LADI_API int64_t ladi_workspace_bytes(ladi_engine* engine, int module, int batch, int height, int width);
/* the launch sequence of the last ladi_workspace_bytes walk, one op per line (tooling / the CPU test that pins it to the Python sequencing) */
LADI_API const char* ladi_engine_trace(const ladi_engine* engine);

/* eps = unet(x, t, ctx):  x_in NHWC bf16 [batch, h, w, in_pitch] (first in_channels valid; channel order latents4, mask1, masked4,
 * pose18, cloth4, tryon_pipe.py:724-726); step_ptr device int32[2] = {row of `steps` to use, 0}; steps fp32 [n_steps][temb_total] =
 * conv1.bias + time_emb_proj(silu(time_embedding(t_s))) per resnet (step-invariant, built once per call); ctx_kv bf16
 * [batch, ctx_tokens, kv_total] = the text context already projected by every cross-attention layer's to_k / to_v;
 * eps_out NHWC fp32 [batch, h, w, 4]. */
LADI_API int ladi_unet_forward(ladi_engine* engine, const void* x_in, const int* step_ptr, const float* steps, const void* ctx_kv, int batch, int lat_h,
                               int lat_w, int ctx_tokens, void* eps_out, void* workspace, int64_t workspace_bytes, void* stream);
/* The two step-invariant tables ladi_unet_forward reads (the reference recomputes both inside every forward, tryon_pipe.py:732):
 *   steps  fp32 [n][temb_total] = conv1.bias + time_emb_proj(silu(time_embedding(t_s))) for every resnet, from the n timesteps (device fp32 [n],
 *          the scheduler's `timesteps`, tryon_pipe.py:650-651); needs ladi_workspace_bytes(engine, LADI_MODULE_UNET_PLAN, n, 0, 0) of workspace;
 *   ctx_kv bf16 [rows = batch * ctx_tokens][kv_total] = the text context (bf16 [rows][ctx_dim]) through every cross-attention layer's to_k / to_v. */
LADI_API int ladi_unet_plan_steps(ladi_engine* engine, const float* timesteps, int n, float* steps_out, void* workspace, int64_t workspace_bytes, void* stream);
LADI_API int ladi_unet_plan_context(ladi_engine* engine, const void* ctx, int rows, int ctx_dim, void* ctx_kv_out, void* stream);
/* x NHWC bf16 [batch, H, W, 8] (3 valid channels) -> posterior moments NHWC fp32 [batch, H/8, W/8, 2*latent] (quant_conv folded in) and the
 * retained encoder features (vae.py:100-109): skips[1] (= skips[2]) [batch,H,W,C0], skips[3] [batch,H/2,W/2,C0], skips[4] [.., /4, C1],
 * skips[5] [.., /8, C2]; skips[0] is unused (the input itself); a null entry (or skips == NULL) keeps that feature in the workspace. */
LADI_API int ladi_vae_encode(ladi_engine* engine, const void* x_nhwc8, int batch, int height, int width, float* moments, void* const* skips,
                             void* workspace, int64_t workspace_bytes, void* stream);
/* z NHWC bf16 [batch, h, w, 8] (latents / scaling_factor, 4 valid channels) + the EMASC outputs feats[0..n_feats) in the reference's list
 * order with their encoder layer indices int_layers[] (vae.py:183-212; n_feats = 0: plain decode) -> image NHWC fp32 [batch, 8h, 8w, 4]. */
LADI_API int ladi_vae_decode_emasc(ladi_engine* engine, const void* z_nhwc8, int batch, int lat_h, int lat_w, const void* const* feats, int n_feats,
                                   const int* int_layers, float* image_out, void* workspace, int64_t workspace_bytes, void* stream);
/* per scale i: outs[i] = conv3x3(silu(conv3x3(feats[i]))) * inv_masks[i] (fp32 per output pixel = 1 - nearest-resized mask; NULL: no
 * masking); feats / outs NHWC bf16 at resolution (height, width) / emasc_stride[i]. */
LADI_API int ladi_emasc_forward(ladi_engine* engine, const void* const* feats, const float* const* inv_masks, int batch, int height, int width,
                                void* const* outs, void* workspace, int64_t workspace_bytes, void* stream);
/* feats bf16 [batch, tokens, adapter_dim] (CLIP ViT-H last_hidden_state) -> out bf16 [batch, adapter_out] */
LADI_API int ladi_inversion_adapter_forward(ladi_engine* engine, const void* feats, int batch, int tokens, void* out, void* workspace,
                                            int64_t workspace_bytes, void* stream);
/* n_steps x (ladi_unet_forward + ladi_ddim_cfg_step) enqueued on `stream`: unet_in NHWC bf16 [cfg ? 2*batch : batch, h, w, in_pitch] with
 * the static channels already written, latents NCHW fp32 [batch,4,h,w], step_ptr device int32[2] zeroed by the caller (advances on the
 * device: no host sync in the loop), coef fp32 [n_steps][8] (see ladi_ddim_cfg_step), eps_scratch fp32 [cfg ? 2*batch : batch, h, w, 4]. */
LADI_API int ladi_denoise_loop(ladi_engine* engine, void* unet_in, float* latents, int* step_ptr, const float* steps, const float* coef, const void* ctx_kv,
                               int batch, int lat_h, int lat_w, int ctx_tokens, int cfg, float guidance, int n_steps, float* eps_scratch, void* workspace,
                               int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif
