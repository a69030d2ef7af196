/*
 * ttvdm.h — C ABI of libttvdm_sm100.so: the hand-written sm_100a kernels behind the This&That / SVD
 * denoising hot path (UNetSpatioTemporalConditionModel + GestureNet ControlNetModel + Euler sampler) and, as the
 * scope table's "next" rows, the gesture rasteriser, the VAE either side of the loop and the CLIP conditioning builder.
 *
 * The reference (Kiteretsu77/This_and_That_VDM) is 100 % Python and has no FFI of its own: every entry
 * point below replaces a *library-dispatched torch op class* on the hot path. Each declaration cites the
 * reference call site(s) it stands in for (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a raw DEVICE pointer owned by the caller; the library never allocates or frees
 *     caller memory and never synchronises (all entry points are CUDA-graph capturable);
 *   - `stream` is a cudaStream_t passed as void*;
 *   - activations are bf16, channels-last: an image batch is [N, H, W, C] == a token matrix [N*H*W, C];
 *     video rows are ordered (b, f, s) exactly like the reference's `[batch*frames, h*w, C]` tokens;
 *   - return value: 0 on success, negative ttvdm_status otherwise; ttvdm_last_error() gives the text;
 *   - no C++ exceptions cross this boundary.
 */
#ifndef TTVDM_H_
#define TTVDM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  TTVDM_OK = 0,
  TTVDM_ERR_SHAPE = -1, /* bad shape / alignment / unsupported configuration */
  TTVDM_ERR_CUDA = -2,  /* CUDA runtime or driver error (see ttvdm_last_error) */
  TTVDM_ERR_ARCH = -3   /* device is not sm_100 */
} ttvdm_status;

/* Process-wide immutable state (driver entry points, SM count). Safe to call repeatedly. */
int ttvdm_init(int device);
/* Drops that state again (the library owns no device memory, streams or events); a later call re-initialises. */
int ttvdm_destroy(void);
/* Copies the last error message of the calling thread into buf (NUL terminated). */
int ttvdm_last_error(char* buf, size_t n);
/* Library / ABI version, bumped when a struct below changes. */
int ttvdm_abi_version(void);
/* Number of kernel launches this library has enqueued since load (bench.py's `gpu_launches`). */
uint64_t ttvdm_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * K1/K2/K3/K4/K15 — one tcgen05 (TMA -> smem -> UMMA -> TMEM) GEMM family with fused epilogue.
 *
 *   out = act( s0 * (A (*) W^T + bias + rowvec[row / rows_per_vec]) + s1 * res1 + s2 * res2 )
 *
 * A-operand addressing modes (the implicit-GEMM part is done by TMA coordinates, never materialised):
 *   TTVDM_A_LINEAR : A = [M, k1] (optionally concatenated with a2 = [M, k2] along K)
 *                    replaces nn.Linear q/k/v/o, proj_in/proj_out, FeedForward
 *                    (svd/diffusion_arch/transformer_temporal.py:235,272,326,373 + diffusers Attention /
 *                    FeedForward built at :240,:253), 1x1 conv_shortcut of ResnetBlock2D and the 13 zero
 *                    convs (svd/temporal_controlnet.py:253-297, 616-622)
 *   TTVDM_A_CONV3X3: A = NHWC [n_img, H, W, k1], 3x3, pad 1, stride 1; W = [N, 9*k1] (tap-major, then
 *                    channel) replaces nn.Conv2d 3x3: conv_in svd/unet_spatio_temporal_condition.py:133,
 *                    conv_in_concat svd/temporal_controlnet.py:203, ResnetBlock2D conv1/conv2, Upsample2D
 *                    conv, conv_out svd/unet_spatio_temporal_condition.py:247
 *   TTVDM_A_TCONV3 : A = [n_img(=B), H(=F), W(=S), k1], 3 taps over F, zero padded; W = [N, 3*k1]
 *                    replaces nn.Conv3d (3,1,1) of TemporalResnetBlock (diffusers; built at
 *                    svd/diffusion_arch/unet_3d_blocks.py:1891,1995,2094,2212,2311)
 * geglu = 1: weight rows are interleaved (hidden_j, gate_j) and the epilogue writes
 *            out[:, j] = (acc_2j + b_2j) * gelu_erf(acc_2j+1 + b_2j+1)  (N/2 output columns)
 *            replaces diffusers GEGLU (Linear(C, 8C) -> chunk -> h * F.gelu(gate)).
 * ------------------------------------------------------------------------------------------------ */
enum { TTVDM_A_LINEAR = 0, TTVDM_A_CONV3X3 = 1, TTVDM_A_TCONV3 = 2 };

typedef struct {
  int mode;          /* TTVDM_A_* */
  const void* a;     /* bf16 */
  const void* a2;    /* bf16 or NULL (LINEAR mode only) */
  int k1, k2;        /* channels of a / a2; multiples of 64 unless a2 == NULL (then k1 % 8 == 0) */
  int lda, lda2;     /* row (pixel) stride of a / a2 in elements */
  int n_img, H, W;   /* CONV3X3: images, height, width. TCONV3: B, F, S. LINEAR: ignored */
  const void* w;     /* bf16 [N, taps*(k1+k2)] */
  int M, N;          /* rows of the output, output features (before GEGLU halving) */
  const float* bias;   /* [N] or NULL */
  const float* rowvec; /* [M / rows_per_vec, ldrv] or NULL (timestep-embedding shift, one vector per video) */
  int rows_per_vec; int ldrv;
  float s0;
  const void* res1; int ldr1; float s1; /* bf16 [M, *] or NULL */
  const void* res2; int ldr2; float s2;
  int geglu;
  void* out; int ldo; int out_fp32;     /* bf16 (default) or fp32 output, row stride ldo elements */
  int act;                              /* 0 none, 1 SiLU (TimestepEmbedding MLPs) */
  /* ---- normalisation fusions (ABI 3). All optional (NULL = off); bf16 output, N % 64 == 0, no GEGLU for the *_out ones.
   * gn_stats_out : double [M / gn_rows_per_inst][N / 2][2] — the epilogue ADDS, per group instance and channel PAIR, the
   *                sum and sum of squares of the bf16 values it stores (caller zeroes the buffer). Sums are kept in
   *                shared memory per CTA — which then walks a CONTIGUOUS range of tiles — and reach the buffer as a
   *                few fp64 atomics per CTA and instance, not per tile. The
   *                GroupNorm that consumes `out` (ttvdm_groupnorm.pstats1/2) then needs no statistics pass: north-star
   *                "GroupNorm fused into the preceding conv's epilogue" (diffusers ResnetBlock2D / TemporalResnetBlock
   *                norm1/norm2 via svd/diffusion_arch/unet_3d_blocks.py:1891-2311, TransformerSpatioTemporalModel.norm
   *                svd/diffusion_arch/transformer_temporal.py:234,323, conv_norm_out svd/unet_spatio_temporal_condition.py:244).
   * row_sums_out : float [N / 32][M][2] — per row and 32-column chunk, sum / sum of squares of the bf16 values stored
   *                (plain stores, exactly one writer per slot: no zeroing, no atomics, bit-reproducible): the LayerNorm
   *                statistics of the consumer (BasicTransformerBlock / TemporalBasicTransformerBlock norm1-3, norm_in).
   * ln_rowsums   : float [K / 32][M][2] of the A operand (the producer's row_sums_out; K = k1 + k2). When set, A is the UN-normalised
   *                tensor, W must have the LayerNorm gain folded in (W' = W diag(gamma)), `bias` the folded bias
   *                (b + W beta), ln_colsum[n] = sum_k W'[n, k], and the epilogue computes
   *                    LN(A) W^T + b  =  rstd_r * (acc_rn + prevec[(r / prevec_rows) % prevec_mod][n] - mean_r * ln_colsum[n]) + bias[n]
   *                so the LayerNorm kernel and its round trip through HBM disappear. ln_sum_add / ln_sq_add (per prevec
   *                row: [prevec_mod][2] floats or NULL) are added to the row sums first (frame positional embedding,
   *                svd/diffusion_arch/transformer_temporal.py:356).
   */
  void* gn_stats_out; int gn_rows_per_inst;
  float* row_sums_out;
  /* row sums are taken over  stored_bf16(out)[r, n] + rs_addvec[(r / rs_add_rows) % rs_add_mod][n]  (fp32 [rs_add_mod, ld_rs_add]
   * or NULL): the consumer normalises `out + frame positional embedding` without that sum ever being materialised */
  const float* rs_addvec; int rs_add_rows; int rs_add_mod; int ld_rs_add;
  const float* ln_rowsums; const float* ln_colsum; float ln_eps;
  const float* prevec; int prevec_rows; int prevec_mod; int ldpv;
  const float* ln_row_add;              /* [prevec_mod][2] or NULL */
  /* ABI 4. CONV3X3 only: 0 / 1 = stride 1; 2 = stride 2 with padding 1 (diffusers Downsample2D conv): H, W are the OUTPUT
   * height / width (M = n_img*H*W), A is the [n_img, 2H, 2W, k1] input, read through a TMA box with element stride 2 —
   * no im2col pass. */
  int conv_stride;
  /* ABI 4. CONV3X3 only: conv_taps = 0 / 9: the 3 x 3 taps (dy, dx in -1..1). conv_taps = 4: a 2 x 2 window whose first
   * tap sits at (conv_dy0, conv_dx0) relative to the output pixel, W = [N, 4*k1] — one of the four parity convolutions
   * that nearest-x2 upsampling followed by a 3 x 3 convolution (diffusers Upsample2D) decomposes into on the LOW-resolution
   * input (16 instead of 36 tap-pixels); ttvdm_interleave2x puts the four results into place. */
  int conv_taps, conv_dy0, conv_dx0;
} ttvdm_gemm_params;

int ttvdm_gemm(const ttvdm_gemm_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K5 — spatial self-attention, flash style on tcgen05 (S and O accumulators in TMEM, online softmax).
 * q/k/v: bf16 token matrices with row stride ld* (so a fused [M, 3C] qkv buffer works), rows ordered
 * (image, token); head h uses columns [h*64, h*64+64). softmax(q k^T * scale) v, no mask.
 * Replaces F.scaled_dot_product_attention in diffusers AttnProcessor2_0 for BasicTransformerBlock.attn1.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  const void* q; const void* k; const void* v; void* out;
  int ldq, ldk, ldv, ldo;
  int n_img, heads, seq; /* head_dim fixed at 64 */
  float scale;
} ttvdm_attn_params;
int ttvdm_attn_spatial(const ttvdm_attn_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K6/K8 — cross-attention against a short, precomputed context (L <= 128 keys). Units with >= 512 rows per
 * context (and context stride <= 2) run the tcgen05 flash kernel in cross mode: the context is one ragged KV tile,
 * the rows that read it are gathered by a 3-D TMA box with an element stride (csrc/attn_flash.cu); smaller units and
 * n_ctx > 2 run a warp-level mma.sync kernel with K/V of one (context, head) pinned in shared memory
 * (csrc/attn_cross.cu). q, kc, vc, out must be 16-byte aligned.
 * q: [rows, heads*64] (row stride ldq). kc/vc: [n_ctx, L, heads*64] bf16 — K/V of the constant context,
 * projected ONCE per video (reference recomputes them per frame and per pixel,
 * svd/unet_spatio_temporal_condition.py:452, svd/diffusion_arch/transformer_temporal.py:316-319).
 * Context of row r:  ctx = (r / ctx_div) % ctx_mod
 *   spatial  (BasicTransformerBlock.attn2):          ctx_div = F*S, ctx_mod = B         -> ctx = b
 *   temporal (TemporalBasicTransformerBlock.attn2):  the reference quirk — temporal row (b, s) reads
 *            context ((b*S + s) mod B) (svd/diffusion_arch/transformer_temporal.py:310-319). The caller
 *            passes quirk_S = S, quirk_B = B (+ global batch offset) and the kernel derives it per row.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  const void* q; const void* kc; const void* vc; void* out;
  int ldq, ldo;
  int rows, heads, L;
  int F, S;          /* rows are ordered (b, f, s); rows = B_local*F*S */
  int n_ctx;         /* contexts held in kc/vc (global batch B) */
  int temporal;      /* 0: ctx = b_global ; 1: ctx = (b_global*S + s) mod n_ctx */
  int batch_offset;  /* global index of the first local batch element (batch sharding) */
  float scale;
  int head_dim;      /* ABI 4: 0 or 64 (default), or 128 (warp-level kernel only; the reference UNet's class-default heads) */
} ttvdm_xattn_params;
int ttvdm_attn_cross(const ttvdm_xattn_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K7 — temporal self-attention over the F (<= 32) frames of each (b, s, head); rows ordered (b, f, s),
 * i.e. the kernel walks frames with stride S*ld instead of materialising the reference's
 * `(b f) s c -> (b s) f c` permute (diffusers TemporalBasicTransformerBlock.attn1).
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  const void* q; const void* k; const void* v; void* out;
  int ldq, ldk, ldv, ldo;
  int B, F, S, heads;
  float scale;
  int head_dim;      /* ABI 4: 0 or 64 (default), or 128 */
} ttvdm_tattn_params;
int ttvdm_attn_temporal(const ttvdm_tattn_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K9/K10 — GroupNorm(32) statistics + apply (+SiLU), channels-last, optional 2-source channel concat
 * (the up-blocks' torch.cat([hidden, skip], dim=1), svd/diffusion_arch/unet_3d_blocks.py:2242,2352).
 * A "group instance" spans rows_per_inst consecutive rows: H*W for the per-frame 4-D norm
 * (ResnetBlock2D.norm1/2, TransformerSpatioTemporalModel.norm, conv_norm_out) and F*H*W for the 5-D
 * norm of TemporalResnetBlock (stats over all frames of one video).
 * stats: caller-provided workspace of n_inst*32*2 doubles (sum, sum of squares), zeroed by the call.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  const void* x1; int c1; int ld1;
  const void* x2; int c2; int ld2;  /* NULL / 0 when no concat */
  int rows; int rows_per_inst;
  float eps;
  void* stats;               /* workspace, n_inst*32*2*sizeof(double) bytes */
  const float* gamma; const float* beta; /* [c1+c2] */
  int silu;
  void* out; int ldo;        /* bf16 [rows, c1+c2] */
  /* ABI 3: per-(instance, channel pair) sums produced by the epilogue of the GEMM that wrote x1 / x2
   * (ttvdm_gemm_params.gn_stats_out, same rows_per_inst): double [n_inst][c/2][2] or NULL. A source with pstats is not
   * read by the statistics pass; with all sources covered that pass is not launched at all. */
  const void* pstats1; const void* pstats2;
} ttvdm_groupnorm_params;
int ttvdm_groupnorm(const ttvdm_groupnorm_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K11 — LayerNorm over C (eps 1e-5) with optional fused "+ frame positional embedding":
 *   x' = x + addvec[(row / S) % F]      (svd/diffusion_arch/transformer_temporal.py:356 `hidden_states_mix + emb`)
 *   sum_out = x' (optional), out = LN(x') * gamma + beta
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  const void* x; int ldx; int rows; int C;
  const float* addvec; int F; int S; /* addvec fp32 [F, C] or NULL */
  void* sum_out; int ldsum;          /* bf16 or NULL */
  const float* gamma; const float* beta; float eps;
  void* out; int ldo;
} ttvdm_layernorm_params;
int ttvdm_layernorm(const ttvdm_layernorm_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Layout helpers around the 3x3 convs.
 *  im2col_s2 : Downsample2D = Conv2d(C, C, 3, stride 2, pad 1): gathers [n,Ho,Wo,9*C] patches so that
 *              the conv becomes a LINEAR GEMM (3 calls per network).
 *  upsample2x: Upsample2D's F.interpolate(scale_factor=2, mode="nearest") in NHWC.
 * ------------------------------------------------------------------------------------------------ */
int ttvdm_im2col_s2(const void* x, void* out, int n_img, int H, int W, int C, void* stream);
int ttvdm_upsample2x(const void* x, void* out, int n_img, int H, int W, int C, void* stream);
/* parts: bf16 [4][n_img, H, W, C], parity p = 2*py + px  ->  out bf16 [n_img, 2H, 2W, C], out[n, 2y+py, 2x+px] = parts[p][n, y, x] */
int ttvdm_interleave2x(const void* parts, void* out, int n_img, int H, int W, int C, void* stream);
/* K13 helper: out[i, :] = [cos(t_i * w_j) | sin(t_i * w_j)], w_j = exp(-ln(1e4) * j / (dim/2)) — diffusers
 * Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0); t fp32 [n] on device, out bf16 [n, dim]. */
int ttvdm_sinusoid(const float* t, void* out, int n, int dim, void* stream);
/* out[r, :] = a[r, :] + scale * b[r, :]  (U4: ControlNet residual merge, svd/unet_spatio_temporal_condition.py:485-502) */
int ttvdm_axpy(const void* a, const void* b, void* out, float scale, size_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K14 — sampler glue (svd/pipeline_stable_video_diffusion_controlnet.py:626-635, 697-709 and diffusers
 * EulerDiscreteScheduler.scale_model_input / step, v-prediction).
 *  prepare: model_in[b, f, h, w, 0:64] (bf16, channel-padded NHWC) =
 *           [ latents[f, 0:4] / sqrt(sigma^2+1) | image_latents[b, 0:4] | cond[f, 0:4] (or 0) | 0... ]
 *           latents fp32 [F,4,h,w] (reference NCHW state), image_latents fp32 [B,4,h,w], cond fp32 [F,4,h,w].
 *  step   : eps = uncond + g[f] * (cond - uncond); x0 = eps * (-sigma/sqrt(sigma^2+1)) + x/(sigma^2+1);
 *           x += (x - x0)/sigma * (sigma_next - sigma)           (all fp32, in place on latents)
 *           eps_u / eps_c: fp32 [F*h*w, 4] channels-last network outputs of the two CFG halves.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  const float* latents; const float* image_latents; const float* cond; /* cond may be NULL */
  void* model_in; int c_pad; /* bf16 [B_local, F, h, w, c_pad] */
  int B_local; int batch_offset; int F; int h; int w;
  float sigma;
} ttvdm_prepare_params;
int ttvdm_sampler_prepare(const ttvdm_prepare_params* p, void* stream);

typedef struct {
  float* latents;                       /* fp32 [F,4,h,w], updated in place */
  const float* eps_u; const float* eps_c; int ld_eps; /* fp32 channels-last [F*h*w, ld_eps] */
  const float* guidance;                /* fp32 [F] */
  int F; int h; int w;
  float sigma; float sigma_next;
} ttvdm_euler_params;
int ttvdm_sampler_euler_step(const ttvdm_euler_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Gesture rasteriser ("next" row of the scope table): the [F, 3, H, W] fp32 this/that condition of
 * data_loader/video_this_that_dataset.py:28-130 (get_thisthat_sam; duplicated in app.py:282-328) — per point a
 * 21x21 coloured square on a 255 image of the ORIGINAL frame size, cv2.filter2D with the 99x99 Gaussian
 * (sigma 10, BORDER_REFLECT_101), cv2.resize INTER_CUBIC to (W, H), optional np.fliplr, / 255; frames without a
 * point are zeros; later points overwrite earlier ones on the same frame. Points in data.txt order (point 0 is
 * drawn [0,0,255], the others [0,255,0]; BGR order is kept, as in the reference).
 * scratch: device, >= n_points * (H + W) floats. out: device fp32 [F, 3, H, W], fully written.
 * ------------------------------------------------------------------------------------------------ */
#define TTVDM_GESTURE_MAX_POINTS 16
typedef struct {
  int n_points;
  int frame_idx[TTVDM_GESTURE_MAX_POINTS], vertical[TTVDM_GESTURE_MAX_POINTS], horizontal[TTVDM_GESTURE_MAX_POINTS];
  int org_h, org_w;  /* size of the original frame (im_0.jpg) the coordinates refer to */
  int H, W, F;
  int dilate, flip;
  void* scratch;
  void* out;
} ttvdm_gesture_params;
int ttvdm_gesture_raster(const ttvdm_gesture_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * VAE either side of the loop ("next" row of the scope table): diffusers AutoencoderKLTemporalDecoder as the reference
 * calls it — vae.encode(image).latent_dist.mode() (svd/pipeline_stable_video_diffusion_controlnet.py:199, :652) and
 * vae.decode(latents[i:i+chunk], num_frames=...) (:257-283). Its convolutions, GroupNorms, linears and the 3-tap
 * temporal convolutions run on ttvdm_gemm / ttvdm_groupnorm / ttvdm_upsample2x above; the three entry points below
 * are what only the VAE needs.
 *
 *  softmax_rows     : the mid-block Attention (ONE head of 512 dims over all H*W/64 tokens of a frame) is
 *                     scores = ttvdm_gemm(Q, K) (fp32, scaled), P = softmax_rows(scores), out = ttvdm_gemm(P, V^T).
 *                     x fp32 [rows, cols] (row stride ldx) -> out bf16 [rows, cols_out] (row stride ldo);
 *                     columns [cols, cols_out) are written as 0 (K padding of the following GEMM). causal = P > 0 (CLIP
 *                     text tower): rows come in blocks of P queries (one block per head); row r is query r % P and only
 *                     attends to columns 0..r % P, the rest are written as 0. causal = 0: no mask.
 *  im2col_s2_pad01  : Downsample2D(padding=0) of the encoder = F.pad(x, (0,1,0,1)) + Conv2d(C, C, 3, stride 2):
 *                     gathers [n, H/2, W/2, 9*C] patches (tap-major, then channel) for a LINEAR GEMM.
 *  vae_time_conv_out: TemporalDecoder.time_conv_out = Conv3d(3, 3, (3,1,1), padding (1,0,0)) over the frames of each
 *                     video, fused with the channels-last -> NCHW conversion of the decoded frames.
 *                     x: device fp32 [(b, f, s), ldx] (3 real channels per row); w / bias: HOST pointers to the 27
 *                     weights [co][ci][t] and 3 biases (they travel in the kernel's parameter space);
 *                     out: device fp32 [B*F, 3, S].
 * ------------------------------------------------------------------------------------------------ */
int ttvdm_softmax_rows(const float* x, int ldx, void* out, int ldo, int rows, int cols, int cols_out, int causal,
                       void* stream);
int ttvdm_im2col_s2_pad01(const void* x, void* out, int n_img, int H, int W, int C, void* stream);
int ttvdm_vae_time_conv_out(const float* x, int ldx, const float* w, const float* bias, float* out, int B, int F, int S,
                            void* stream);

/* ------------------------------------------------------------------------------------------------
 * Conditioning builder ("next" row #2): encode_clip (svd/pipeline_stable_video_diffusion_controlnet.py:130-188) — CLIP
 * image tower (transformers CLIPVisionModelWithProjection, `self.image_encoder(image).image_embeds`, :155), CLIP text
 * tower (CLIPTextModel, `text_encoder(prompt)[0]`, :166), concat, a fresh nn.LayerNorm((78, 1024)) (:172-173) and the
 * CFG zero stack. Linears / attention GEMMs run on ttvdm_gemm, LayerNorms on ttvdm_layernorm, softmax on
 * ttvdm_softmax_rows (causal for the text tower); the two entry points below are what only this row needs.
 *  act_inplace    : the towers' MLP activation on a bf16 buffer; kind 2 = GELU (erf form, `hidden_act: "gelu"` of the
 *                   ViT-H / SD-2.1 text checkpoints), kind 3 = quick GELU x*sigmoid(1.702x) (OpenAI CLIP checkpoints).
 *  layernorm_flat : out[r, :] = (x[r, :] - mean_r) / sqrt(var_r + eps) over ALL n elements of row r (normalized_shape
 *                   = (78, 1024), no affine: the reference's LayerNorm is freshly constructed). fp32 in / out.
 * ------------------------------------------------------------------------------------------------ */
int ttvdm_act_inplace(void* x, size_t n, int kind, void* stream);
int ttvdm_layernorm_flat(const float* x, float* out, int rows, size_t n, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Weight repack entry points (run once per model load). Sources are PyTorch-layout parameters exactly as a
 * diffusers state dict holds them (svd/unet_spatio_temporal_condition.py / svd/temporal_controlnet.py modules), in fp32,
 * fp16 or bf16; outputs are the operands ttvdm_gemm consumes.
 *  pack_conv_weight : w [cout, cin, taps] (Conv2d 3x3: taps = 9 row-major (ky, kx); Conv3d (3,1,1): taps = 3; 1x1: taps = 1)
 *                     -> out bf16 [cout, taps * cin_pad], tap-major then channel, channels zero-padded to cin_pad.
 *  pack_linear      : w [N, K] (+ bias [N]) with an optional LayerNorm(gamma, beta) folded in front of it:
 *                        out_w[row, k]   = bf16(w[n, k] * gamma[k])
 *                        out_colsum[row] = sum_k out_w[row, k]                       (fp32; NULL to skip)
 *                        out_bias[row]   = bias[n] + sum_k w[n, k] * beta[k]         (fp32; NULL to skip)
 *                     row = out_row0 + n, or with geglu = 1 the (hidden_j, gate_j) interleave the GEGLU epilogue expects
 *                     (row = out_row0 + 2n for n < N/2, out_row0 + 2(n - N/2) + 1 otherwise). ldo = row stride of out_w.
 *                     out_row0 lets to_q | to_k | to_v land in one fused [3C, C] operand.
 *  pack_vector      : cast n elements to fp32 (biases, norm affine parameters).
 * ------------------------------------------------------------------------------------------------ */
enum { TTVDM_DT_F32 = 0, TTVDM_DT_F16 = 1, TTVDM_DT_BF16 = 2 };
int ttvdm_pack_conv_weight(const void* w, int src_dtype, int cout, int cin, int taps, int cin_pad, void* out, void* stream);
typedef struct {
  const void* w; const void* bias;          /* [N, K], [N] or NULL */
  const void* gamma; const void* beta;      /* [K] or NULL: LayerNorm in front of the linear */
  int src_dtype;                            /* TTVDM_DT_* of w / bias / gamma / beta */
  int N, K;
  int geglu; int out_row0; int ldo;
  void* out_w; float* out_bias; float* out_colsum;
} ttvdm_pack_linear_params;
int ttvdm_pack_linear(const ttvdm_pack_linear_params* p, void* stream);
int ttvdm_pack_vector(const void* src, int src_dtype, size_t n, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Workspace queries. No entry point allocates; the only caller-provided scratch areas are the ones below (all others
 * need none, which the two *_workspace_bytes queries state explicitly).
 * ------------------------------------------------------------------------------------------------ */
size_t ttvdm_groupnorm_workspace_bytes(int rows, int rows_per_inst);      /* ttvdm_groupnorm_params.stats */
size_t ttvdm_gemm_gn_stats_bytes(int M, int N, int gn_rows_per_inst);     /* ttvdm_gemm_params.gn_stats_out */
size_t ttvdm_gemm_row_sums_bytes(int M, int N);                           /* ttvdm_gemm_params.row_sums_out */
size_t ttvdm_gemm_workspace_bytes(const ttvdm_gemm_params* p);            /* 0 */
size_t ttvdm_attn_workspace_bytes(const ttvdm_attn_params* p);            /* 0 */
size_t ttvdm_gesture_scratch_bytes(int n_points, int H, int W);           /* ttvdm_gesture_params.scratch */

#ifdef __cplusplus
}
#endif
#endif /* TTVDM_H_ */
