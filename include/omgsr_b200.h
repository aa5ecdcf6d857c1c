/*
 * omgsr_b200 — C ABI of the B200 (sm_100a) kernels behind OMGSR's one-step SR inference path.
 *
 * The reference (wuer5/OMGSR) has no FFI of its own: its operator surface for this path is the Python
 * module API of infer/omgsr_s_infer_model.py, infer/omgsr_f_infer_model.py, infer/vaehook.py and
 * infer/wavelet_color_fix.py, underneath which diffusers/PyTorch dispatch to cuDNN, cuBLAS and SDPA.
 * Each entry point below names the reference call site whose library kernels it replaces (paths are
 * relative to the reference tree). omgsr_b200/capi.py is the ctypes binding; INTEGRATION.md shows the
 * stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative OMGSR_ERR_* code; the message is available from
 *     omgsr_last_error() (thread-local). Nothing throws, nothing exits, there is no CPU fallback.
 *   - all pointers are DEVICE pointers unless stated; the library allocates no device memory. The caller
 *     (PyTorch's caching allocator in this repo) owns every buffer.
 *   - every function only enqueues work on `stream` (a cudaStream_t passed as void*): no hidden
 *     synchronisation, so call sequences are CUDA-graph capturable.
 *   - activations are NHWC ("channels last") bf16 with an explicit channel stride `ld*` in elements so a
 *     layer can read or write a channel slice of a wider (concatenated) buffer. Strides are multiples of 8.
 *   - weights are bf16 [Cout, R*S*Cin] with k = (r*S + s)*Cin + cin; biases / norm affine vectors are fp32.
 */
#ifndef OMGSR_B200_H
#define OMGSR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define OMGSR_OK 0
#define OMGSR_ERR_SHAPE (-1)
#define OMGSR_ERR_ALIGN (-2)
#define OMGSR_ERR_ARCH (-3)
#define OMGSR_ERR_CUDA (-4)

/* epilogue activations of omgsr_gemm_bf16 / omgsr_conv2d_nhwc_bf16 */
#define OMGSR_ACT_NONE 0
#define OMGSR_ACT_SILU 1
#define OMGSR_ACT_GELU_ERF 2
#define OMGSR_ACT_GELU_TANH 3
#define OMGSR_ACT_GEGLU 4 /* out[:, j] = v[:, j] * gelu_erf(g[:, j]); weight rows interleaved per tile by the packer */

/* flags */
#define OMGSR_OUT_F32 1      /* store fp32 instead of bf16 */
#define OMGSR_B_MN_MAJOR 2   /* batched GEMM: B operand given as [K, N] (row-major V), not [N, K] */
#define OMGSR_CONV_NO_HALO 4     /* conv2d: force the generic implicit-GEMM kernel (testing / A-B timing) */
#define OMGSR_RES_F32 8      /* GEMMs: `res` is fp32 (with OMGSR_OUT_F32: an fp32 residual stream, N % 32 == 0) */

int omgsr_version(void);
const char* omgsr_last_error(void);
/* frees per-process caches (function attributes are process-wide; nothing device-side is owned). */
int omgsr_destroy(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches). */
long long omgsr_launch_count(void);
/* name of the kernel variant that served the most recent GEMM / convolution call of this process (diagnostics:
 * bench.py names the kernel its roofline figure was measured on). */
const char* omgsr_last_variant(void);

/* ---- contractions (tcgen05 + TMEM + TMA) ------------------------------------------------------ */

/* D[M,N] = epi(A[M,K] * Wt[N,K]^T). Replaces torch.nn.Linear under the UNet transformer blocks and
 * the FLUX DiT (reference call sites infer/omgsr_s_infer_model.py:75-79, infer/omgsr_f_infer_model.py:191-200).
 * epi: +bias[N] -> act -> *gate[m / gate_rows, N] -> +res[M,N].  K % 64 == 0. */
int omgsr_gemm_bf16(const void* A, const void* Wt, const float* bias, const void* res, const float* gate, void* D,
                    long long M, int N, int K, long long lda, long long ldd, long long ldr, int act, int gate_rows,
                    int flags, int force_bn, void* stream);

/* Row-batched projection whose q / k head columns leave the epilogue RMS-normalised and rotated: for output columns in
 * [q0, q1) (weight wq[128]) and [k0, k1) (weight wk[128]), per 128-column head
 *     y = rope(x / sqrt(mean(x^2) + eps) * w),  rope: (y[2i], y[2i+1]) -> (y[2i] c - y[2i+1] s, y[2i+1] c + y[2i] s)
 * with c = cos_t[pos, :], s = sin_t[pos, :] ([positions, 128] fp32, frequencies repeated twice) and pos = pos_offset + m
 * (row m of every batch entry); all other columns get +bias and the activation on [act_c0, act_c1) as in
 * omgsr_gemm_batched_bf16. Replaces the Linear -> RMSNorm -> apply_rotary_emb chain of diffusers' FLUX attention
 * processors (reached from infer/omgsr_f_infer_model.py:191-200): the fused q|k|v projections of the double blocks and the
 * q|mlp|k|v projection of the single blocks. N % 128 == 0, ranges multiples of 128, bf16 output rows 32-byte aligned. */
int omgsr_gemm_rope_bf16(const void* A, const void* Wt, const float* bias, void* D, int batch, int M, int N, int K,
                         long long lda, long long a_batch, long long ldd, long long d_batch, int act, int act_c0,
                         int act_c1, const float* wq, const float* wk, int q0, int q1, int k0, int k1,
                         const float* cos_t, const float* sin_t, int pos_offset, float eps, int force_bn, void* stream);

/* 1: route the 256-wide GEMMs / convolutions to the CTA-pair kernel (tcgen05 cta_group::2, M = 256 per cluster of
 * two CTAs, each CTA streaming half of the weight tile); 0: single-CTA kernel. Process-wide switch. */
int omgsr_set_gemm_2cta(int on);

/* Schedule of the 3x3 input-halo convolution (16 < Cout <= 128): -1 automatic (rolling input rows, each loaded and
 * GroupNorm-transformed once, for the fused input GroupNorm variants; row-shared 2-row tiles otherwise), 0 never the
 * rolling-row kernel, 1 the rolling-row kernel for every eligible convolution, -2 back to the start-up default
 * (OMGSR_HALO_ROLLING). Test / A-B timing only. */
int omgsr_set_halo_rolling(int mode);
/* 1 if omgsr_conv2d_gn_nhwc_bf16 can take a 3x3 / s1 / p1 convolution with 256 output channels on this device (two
 * 128-channel passes of the rolling-row kernel): asked by the engines before they leave a GroupNorm + SiLU
 * (diffusers ResnetBlock2D norm -> silu -> conv, task queue infer/vaehook.py:248-276) to the convolution after it. */
int omgsr_conv_gn_wide_fusable(int B, int OH, int OW, int Cin, int Cout);

/* Row-batched linear layer: D[b, m, :] = epi(A[b, m, :] * Wt^T) for b < batch, with independent batch strides
 * for A / D / res (tokens of one image may live inside a longer joint [text; image] sequence), a per-batch
 * gate vector gate[b * gate_ld + n], and the activation restricted to output columns [act_c0, act_c1)
 * (act_c1 <= act_c0: all columns). Used by the FLUX DiT blocks (fused q|mlp|k|v projection of the single
 * blocks, gated residual projections). */
int omgsr_gemm_batched_bf16(const void* A, const void* Wt, const float* bias, const void* res, const float* gate,
                            void* D, int batch, int M, int N, int K, long long lda, long long a_batch, long long ldd,
                            long long d_batch, long long ldr, long long r_batch, long long gate_ld, int act,
                            int act_c0, int act_c1, int flags, int force_bn, void* stream);

/* Batched attention-style GEMM over (batch b, head h):
 *   D[b,h,m,n] = alpha-free sum_k A[b,m,h,k] * Bm[b,n,h,k]          (flags & OMGSR_B_MN_MAJOR == 0; Q K^T)
 *   D[b,m,h,n] =            sum_k A[b,h,m,k] * Bm[b,k,h,n]          (flags & OMGSR_B_MN_MAJOR;      P V)
 * Used for the materialised-score attention path (VAE mid-block attention d=512,
 * infer/vaehook.py:137-171, and as the bring-up path for UNet/DiT attention).
 * strides are in elements: a_* for A (row, head, batch), b_* for Bm, d_* for D. b_valid_rows = number of rows of
 * Bm that exist per (b,h) (N for Q K^T; the un-padded kv length for P V, where K is the padded width of P):
 * rows beyond it are read as zero. */
int omgsr_bgemm_bf16(const void* A, const void* Bm, void* D, int batch, int heads, int M, int N, int K,
                     long long a_row, long long a_head, long long a_batch, long long b_row, long long b_head,
                     long long b_batch, long long d_row, long long d_head, long long d_batch, int b_valid_rows,
                     int flags, int force_bn, void* stream);

/* Fused flash attention (TMA + tcgen05, online softmax, scores never leave the SM):
 *   out[b,m,h,:] = softmax_n(scale * q[b,m,h,:] . k[b,n,h,:]) v[b,n,h,:],  D in {64, 128}.
 * Replaces F.scaled_dot_product_attention under the UNet self/cross attention (d=64) and the FLUX joint
 * attention (d=128) (infer/omgsr_s_infer_model.py:75-79, infer/omgsr_f_infer_model.py:191-200).
 * Strides in elements (row = token, head, batch); D is contiguous. */
int omgsr_flash_attn_bf16(const void* q, const void* k, const void* v, void* out, int B, int H, int M, int N, int D,
                          long long q_row, long long q_head, long long q_batch, long long k_row, long long k_head,
                          long long k_batch, long long v_row, long long v_head, long long v_batch, long long o_row,
                          long long o_head, long long o_batch, float scale, void* stream);

/* A-B timing switch for the attention kernels: 0 = automatic, 1 = always the single-tile kernel, 2 = the 64-key
 * ping-pong kernel instead of the 128-key ones, 3 = the natural-layout 128-key kernel without the early S release,
 * 4 = the transposed-O 128-key kernel at d = 128, 5 = the natural-layout kernel that releases S from registers (the
 * automatic choice). Test/bench only. */
int omgsr_flash_attn_set_variant(int variant);

/* NHWC implicit-GEMM convolution, ksize in {1,3}, stride in {1,2}; `upsample2x` fuses a nearest 2x
 * upsample in front of a 3x3/p1 conv (weights must be packed by pack_upsample_weight: 4 x [Cout, 4*Cin]).
 * Replaces cuDNN under diffusers ResnetBlock2D/Downsample2D/Upsample2D in the VAE and UNet
 * (infer/omgsr_s_infer_model.py:84,173; infer/vaehook.py:248-329 lists the same layers). Cin % 64 == 0.
 * epi: +bias -> act -> +res. gn_sums (nullable): fp32 [B, gn_groups, 2] = (sum, sum of squares) of the STORED bf16
 * output per (image, group), produced by the epilogue, i.e. the statistics pass of the GroupNorm that follows the conv
 * costs no extra trip over HBM (needs >= 128 output pixels per image and Cout/gn_groups in {4, 8, 16}; otherwise
 * OMGSR_ERR_SHAPE). The sums are bit-reproducible: every epilogue warp writes its partial sums to its own slot of
 * gn_scratch (caller-provided fp32 workspace of >= omgsr_gn_scratch_floats(B, gn_groups, upsample2x ? 4 : 1) floats;
 * the call zeroes it) and a fixed-order reduction kernel folds the slots into gn_sums - no floating-point atomics. */
int omgsr_conv2d_nhwc_bf16(const void* x, const void* w, const float* bias, const void* res, void* y, int B, int H,
                           int W, int Cin, int Cout, int ksize, int stride, int pad_t, int pad_l, int pad_b,
                           int pad_r, int upsample2x, int act, long long ldx, long long ldy, long long ldr,
                           float* gn_sums, int gn_groups, float* gn_scratch, long long gn_scratch_floats, int flags,
                           int force_bn, void* stream);
/* workspace size (in floats) for the deterministic GroupNorm statistics of `launches` kernel launches over B images
 * with G groups: omgsr_conv2d_nhwc_bf16 / omgsr_conv2d_gn_nhwc_bf16 (launches = 4 with upsample2x, else 1) and
 * omgsr_groupnorm_stats (launches = 1). Host-only; depends on the SM count of the current device. */
long long omgsr_gn_scratch_floats(int B, int G, int launches);
/* GroupNorm + SiLU of the INPUT fused into a 3x3 / stride 1 / pad 1 convolution: y = conv(silu(a[b,c] * x + s[b,c])) with
 * gn_in_coef [B, Cin] float2 from omgsr_groupnorm_coef (the diffusers ResnetBlock2D pattern norm -> SiLU -> conv,
 * infer/vaehook.py:384-413 custom_group_norm + F.silu before each conv task). The normalised tensor is never written:
 * the transform runs on the operand tile in shared memory. Only on the CTA-pair input-halo path: W >= 96,
 * Cin % 64 == 0, 16 < Cout <= 128, at least two 2x128-pixel tiles; anything else returns OMGSR_ERR_SHAPE (the caller
 * then uses omgsr_groupnorm_apply + omgsr_conv2d_nhwc_bf16). Other arguments as omgsr_conv2d_nhwc_bf16. */
int omgsr_conv2d_gn_nhwc_bf16(const void* x, const float* gn_in_coef, const void* w, const float* bias, const void* res,
                              void* y, int B, int H, int W, int Cin, int Cout, long long ldx, long long ldy,
                              long long ldr, float* gn_sums, int gn_groups, float* gn_scratch,
                              long long gn_scratch_floats, void* stream);

/* ---- thin convolutions (CUDA cores; too few channels for the tensor pipe) --------------------- */

/* conv_in: x NCHW (bf16 if !x_f32 else fp32) [B,Cin,H,W], Cin <= 16, ksize 1|3 'same' -> y NHWC bf16 [B,H,W,Cout].
 * w: fp32 [ksize*ksize*Cin, Cout]. (VAE encoder.conv_in, decoder.conv_in, unet.conv_in) */
int omgsr_conv_in_nchw(const void* x, int x_f32, const float* w, const float* bias, void* y, int B, int Cin, int H,
                       int W, int Cout, int ksize, long long ldy, void* stream);
/* im2col of a thin NCHW input for a 3x3/p1 conv: y[b*H*W + pixel, k] (bf16, row length Kpad, Kpad % 8 == 0) with
 * k = (r*3+s)*Cin + ci and zero fill for k >= 9*Cin; the conv is then omgsr_gemm_bf16 on y with weights packed
 * [Cout, Kpad] (large images: the 27-tap conv_in of the VAE encoder runs on the tensor cores this way). */
int omgsr_im2col3x3_nchw(const void* x, int x_f32, void* y, int B, int Cin, int H, int W, int Kpad, void* stream);
/* conv_out: x NHWC bf16 [B,H,W,Cin] -> y NCHW [B,Cout,H,W] (bf16 or fp32), Cout <= 4, 3x3 p1, optional clamp
 * to [-1,1] (infer/omgsr_s_infer_model.py:85). w: bf16 [Cout, 9*Cin]. (VAE decoder.conv_out) */
int omgsr_conv_out_nchw(const void* x, const void* w, const float* bias, void* y, int y_f32, int B, int H, int W,
                        int Cin, int Cout, int clamp, long long ldx, void* stream);

/* ---- normalisation / glue (HBM-bound) --------------------------------------------------------- */

/* GroupNorm statistics: stats[B, G, 2] = (sum, sum of squares) over (HW, C/G). Bit-reproducible: per-block partial
 * sums go to `scratch` (>= omgsr_gn_scratch_floats(B, G, 1) floats, caller-provided) and are reduced in a fixed order. */
int omgsr_groupnorm_stats(const void* x, long long ldx, float* stats, int B, long long HW, int C, int G,
                          float* scratch, long long scratch_floats, void* stream);
/* y = (x - mean) * rstd * gamma + beta, optional SiLU. (diffusers GroupNorm+SiLU pairs: VAE eps 1e-6,
 * UNet eps 1e-5; infer/vaehook.py:384-413 restates the same op as custom_group_norm.)
 * stats_mode 0: stats[B,G,2] = (sum, sumsq) from omgsr_groupnorm_stats on this tensor.
 * stats_mode 1: stats[stats_mod,G,2] = (mean, var) from omgsr_groupnorm_combine, row = batch index % stats_mod.
 * coef_scratch: caller-provided B*C float2 (8*B*C bytes) for the per-(image, channel) affine coefficients. */
/* only the coefficient stage of omgsr_groupnorm_apply: coef[B, C] float2 = (rstd * gamma, beta - mean * rstd * gamma),
 * for a consumer that applies the normalisation itself (omgsr_conv2d_gn_nhwc_bf16). Same stats_mode convention. */
int omgsr_groupnorm_coef(const float* stats, const float* gamma, const float* beta, void* coef, int B, long long HW,
                         int C, int G, float eps, int stats_mode, int stats_mod, void* stream);
int omgsr_groupnorm_apply(const void* x, long long ldx, void* y, long long ldy, const float* stats,
                          const float* gamma, const float* beta, void* coef_scratch, int B, long long HW, int C,
                          int G, float eps, int silu, int stats_mode, int stats_mod, void* stream);
/* Cross-tile GroupNorm statistics of the tiled VAE (GroupNormParam.summary, infer/vaehook.py:489-508).
 * sums[entries,G,2] = per-(tile,image) (sum, sumsq) with image = entry % n_img; pixels[entries] = pixels of
 * that tile (fp32, device). out[n_img,G,2] = pixel-weighted mean of the per-tile (mean, biased var).
 * wsum_out (nullable, [n_img]): when given, out holds the UN-normalised numerators sum_e pixels_e*(mean_e, var_e) and
 * wsum_out the weights sum_e pixels_e, so partial results of several GPUs can be all-reduced before dividing. */
int omgsr_groupnorm_combine(const float* sums, const float* pixels, int entries, int n_img, int C, int G, float* out,
                            float* wsum_out, void* stream);
/* LayerNorm over the last dim with y = xhat * g + b where g/b are fp32 vectors of length C taken from
 * gamma + (row / rows_per_batch) * gb_stride (gb_stride = 0: shared affine; >0: per-image AdaLN modulation,
 * caller passes g = 1 + scale, b = shift). gamma/beta may be null (no affine). Input row r lives at
 * x + r*ldx, or at x + (r / rows_per_batch)*x_batch_stride + (r % rows_per_batch)*ldx when x_batch_stride > 0
 * (a token range of a longer joint sequence); the output is dense bf16 [rows, C]. ldx = 0 means C. x is bf16, or fp32
 * when x_f32 (the FLUX engine keeps its residual stream in fp32).
 * (UNet BasicTransformerBlock norms, eps 1e-5; FLUX AdaLayerNormZero / Continuous, eps 1e-6) */
int omgsr_layernorm(const void* x, int x_f32, void* y, const float* gamma, const float* beta, long long rows, int C,
                    float eps, long long rows_per_batch, long long gb_stride, long long ldx, long long x_batch_stride,
                    void* stream);
/* softmax over rows of fp32 scores: P[r, j] = softmax_j(scale * S[r, j]) for j < n, P[r, j] = 0 for
 * n <= j < n_pad. */
int omgsr_softmax_rows(const float* S, long long ld_s, void* P, long long ld_p, long long rows, int n, int n_pad,
                       float scale, void* stream);
/* The same with a causal mask (CLIP text encoders, transformers modeling_clip.py `_create_4d_causal_attention_mask`
 * as used by `CLIPTextModel` in infer/infer_omgsr_s.py:19-45 and `FluxPipeline.encode_prompt`, infer_omgsr_f.py:36-43):
 * rows are laid out [..., queries]; query i attends to keys 0..i, P[r, j] = 0 for j > i. */
int omgsr_softmax_rows_causal(const float* S, long long ld_s, void* P, long long ld_p, long long rows, int n, int n_pad,
                              float scale, int queries, void* stream);
/* Token + position embedding of the text encoders (CLIPTextEmbeddings: token_embedding(ids) + position_embedding;
 * pos may be null): out[r, :] = tok[ids[r], :] + pos[r % T, :], bf16 tables [vocab, C] / [T, C], int32 ids, C % 8 == 0. */
int omgsr_embed_tokens(const int* ids, const void* tok, const void* pos, void* out, long long rows, int T, int C,
                       int vocab, void* stream);
/* Row softmax with an additive bias (T5Attention of `T5EncoderModel` in FluxPipeline.encode_prompt,
 * infer_omgsr_f.py:36-43: softmax(q k^T + position_bias), no 1/sqrt(d)): P[r, j] = softmax_j(scale * S[r, j] +
 * bias[(r % bias_rows) * ld_b + j]); bias fp32 [bias_rows = heads * queries, >= n]. */
int omgsr_softmax_rows_bias(const float* S, long long ld_s, void* P, long long ld_p, long long rows, int n, int n_pad,
                            float scale, const float* bias, long long bias_rows, long long ld_b, void* stream);
/* T5LayerNorm: y[r, :] = x[r, :] * rsqrt(mean(x[r, :]^2) + eps) * weight (no mean subtraction, no bias); x bf16 or
 * fp32 rows at pitch ldx (0 = C), y dense bf16. C % 8 == 0. */
int omgsr_rmsnorm(const void* x, int x_f32, void* y, const float* weight, long long rows, int C, float eps,
                  long long ldx, void* stream);
/* out = a * b elementwise, dense bf16, n % 8 == 0 (T5DenseGatedActDense: gelu_new(wi_0 x) * wi_1 x). */
int omgsr_mul_bf16(const void* a, const void* b, void* out, long long n, void* stream);
/* FLUX q/k RMSNorm (per head, d = 128, learned weight, fp32 statistics) + rotary embedding, in place on
 * x[b, r, h, 128] = x + b*batch_stride + r*ld + h*128 (bf16). Heads [0, heads_per_weight) use weight[0],
 * the next heads_per_weight use weight[1], ... so q and k of a fused qkv row take one launch.
 * cos/sin: fp32 [positions, 128] (repeat-interleaved pairs); row r uses position pos_offset + r.
 * (diffusers FluxAttnProcessor2_0 under infer/omgsr_f_infer_model.py:191-200) */
int omgsr_rmsnorm_rope(void* x, long long ld, long long batch_stride, const float* weight, const float* cos_t,
                       const float* sin_t, long long rows_per_batch, int batch, int heads, int heads_per_weight,
                       int pos_offset, float eps, void* stream);

/* ---- latent-sized elementwise steps ------------------------------------------------------------ */

/* VAE posterior sample: moments NHWC fp32 [B,h,w,ldm] (first 2*Cz channels = mean | logvar),
 * eps NCHW fp32 [B,Cz,h,w] -> z NCHW (bf16/fp32) = ((mean + exp(0.5*clamp(logvar,-30,20))*eps) - shift) * scale.
 * (infer/omgsr_s_infer_model.py:173, infer/omgsr_f_infer_model.py:15-18) */
int omgsr_posterior_sample(const float* moments, long long ldm, const float* eps, void* z, int z_f32, int B, int Cz,
                           int h, int w, float shift, float scale, void* stream);
/* generic latent axpby: out = a*x + b*y + c (x, y, out same dtype: bf16 or fp32), n elements.
 * DDPM x0 step (infer/omgsr_s_infer_model.py:80-82), Euler step (infer/omgsr_f_infer_model.py:202),
 * un-scale (infer/omgsr_f_infer_model.py:210) and the training-time mid-timestep mix
 * z_t = (1-sigma) z + sigma eps (train/train_omgsr_f.py:537-538) are all this op. */
int omgsr_latent_axpby(const void* x, const void* y, void* out, int is_f32, long long n, float a, float b, float c,
                       void* stream);
/* x0 prediction fused with un-scale and SD post_quant_conv (1x1, CzxCz):
 *   z0 = (z - s1*eps_hat)/s2 / scale + shift ; out[b,:,y,x] = Wpq * z0[b,:,y,x] + bpq   (Wpq null -> identity)
 * z NCHW (bf16/fp32), eps_hat NHWC fp32 (ld = lde), out NCHW bf16. */
int omgsr_x0_unscale(const void* z, int z_f32, const float* eps_hat, long long lde, void* out, const float* Wpq,
                     const float* bpq, int B, int Cz, int h, int w, float s1, float s2, float scale, float shift,
                     void* stream);
/* FLUX 2x2 pack / unpack between NCHW latents [B,C,h,w] and token rows [B,(h/2)(w/2),4C] (bf16).
 * (infer/omgsr_f_infer_model.py:21-41) */
int omgsr_pack_latents(const void* z, void* tokens, int B, int C, int h, int w, long long ld_tok, void* stream);
int omgsr_unpack_latents(const void* tokens, void* z, int B, int C, int h, int w, long long ld_tok, void* stream);
/* Gaussian-weighted blend of per-tile denoiser outputs into the full latent
 * (infer/omgsr_s_infer_model.py:137-161, infer/omgsr_f_infer_model.py:283-314): out = sum_t w*tile_t / sum_t w.
 * tiles: fp32|bf16, logical [ntiles,B,C,T,T] addressed through tile_strides (HOST array of 5 element
 * strides: tile, batch, channel, y, x); out: fp32 logical [B,C,h,w] through out_strides (HOST array of 4);
 * ofs_xy: DEVICE int [ntiles,2] = (x, y) tile offsets; weights: fp32 [T,T]. */
int omgsr_tile_blend(const void* tiles, int tiles_f32, const float* weights, const int* ofs_xy, int ntiles,
                     float* out, int B, int C, int h, int w, int T, const long long* tile_strides,
                     const long long* out_strides, void* stream);

/* ---- colour fix (infer/wavelet_color_fix.py) ---------------------------------------------------- */

/* out = content + LP5(style - content), LP5 = 5-stage a-trous [1,2,1]x[1,2,1]/16 cascade with replicate
 * padding (identical to wavelet_reconstruction, infer/wavelet_color_fix.py:77-125). fp32 NCHW [B,3,H,W];
 * tmp0/tmp1 are caller-provided scratch of the same size. */
int omgsr_wavelet_fix(const float* content, const float* style, float* out, float* tmp0, float* tmp1, int B, int C,
                      int H, int W, void* stream);
/* one a-trous stage (wavelet_blur, infer/wavelet_color_fix.py:77-96): 3x3 binomial kernel at dilation `radius`,
 * replicate padding, depthwise. fp32 NCHW. */
int omgsr_atrous_blur(const float* in, float* out, int B, int C, int H, int W, int radius, void* stream);
/* AdaIN: out = (content - mean_c)/std_c * std_s + mean_s per (b,c), unbiased var + 1e-5
 * (infer/wavelet_color_fix.py:44-74). scratch: B*C*4 doubles (32*B*C bytes). */
int omgsr_adain_fix(const float* content, const float* style, float* out, void* scratch, int B, int C, int H, int W,
                    void* stream);

/* ---- layout ---------------------------------------------------------------------------------------- */
int omgsr_nchw_to_nhwc(const void* x, int x_f32, void* y, int B, int C, int H, int W, long long ldy, void* stream);
/* clamp != 0: clamp to [-1, 1] on the way out (infer/omgsr_s_infer_model.py:85) */
int omgsr_nhwc_to_nchw(const void* x, long long ldx, void* y, int y_f32, int B, int C, int H, int W, int clamp,
                       void* stream);

/* ---- driver-side image pre/post-processing --------------------------------
 * The reference prepares its input on the CPU with PIL (infer/infer_omgsr_s.py:76-90):
 *   img.resize((w*upscale, h*upscale))  [Pillow default filter = BICUBIC]; img.resize((w - w%8, h - h%8), LANCZOS);
 *   to_tensor(img).to(weight_dtype) * 2 - 1
 * and converts the result back with clip(x*0.5+0.5, 0, 1).float() -> ToPILImage (:96-98). These entries restate
 * that chain for 8-bit RGB on the device, bit-exactly (Pillow's resampler is fixed-point integer arithmetic). */
#define OMGSR_RESIZE_BICUBIC 0
#define OMGSR_RESIZE_LANCZOS 1
/* HOST-only helper: Pillow's per-axis tables (precompute_coeffs + normalize_coeffs_8bpc). bounds: [out_size][2] =
 * (first tap, tap count); coeffs: [out_size][ksize] 22-bit fixed point. bounds == NULL: only *ksize is returned. */
int omgsr_resample_coeffs(int in_size, int out_size, int filter, int* bounds, int* coeffs, int max_ksize, int* ksize);
/* device bytes omgsr_resize_u8 needs (horizontal-pass intermediate + the four tables) */
long long omgsr_resize_u8_workspace(int src_h, int src_w, int dst_h, int dst_w, int filter);
/* src: [src_h, src_w, 3] u8 (device, dense) -> dst: [dst_h, dst_w, 3] u8; == PIL.Image.resize(..., filter) for RGB */
int omgsr_resize_u8(const void* src, int src_h, int src_w, void* dst, int dst_h, int dst_w, int filter,
                    void* workspace, long long workspace_bytes, void* stream);
/* [H, W, 3] u8 -> [1, 3, H, W] bf16 (or fp32): to_tensor(img).to(dtype) * 2 - 1 with the reference's rounding order */
int omgsr_u8_to_input(const void* hwc_u8, void* nchw, int out_f32, int H, int W, void* stream);
/* [3, H, W] bf16|fp32 in [-1, 1] -> [H, W, 3] u8: clip(x*0.5+0.5, 0, 1).float().mul(255).byte() (truncating) */
int omgsr_output_to_u8(const void* nchw, int in_f32, void* hwc_u8, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OMGSR_B200_H */
