/*
 * libvitseg — C ABI of the B200 (sm_100a) kernels behind the ViT-segmentation hot path of
 * mtumalan/VisionTransformer (model/CE and model/PAED training + inference).
 *
 * The reference has no FFI of its own: its hot path is torch.nn / transformers.ViTModel calls
 * (SURVEY.md §8b).  Each entry point below names the reference call it replaces.
 *   model/CE/classes.py:221-297   ViTSegmentationModel / LightningViTModel
 *   model/PAED/classes.py:336-369 paed_loss_multiclass_soft
 *   model/PAED/classes.py:608-701 PAEDTrainer.dice_loss / paed_loss_soft / _forward_step_paed
 *   TF = transformers/models/vit/modeling_vit.py (the un-vendored backbone the reference instantiates)
 *
 * Conventions
 *   - plain C, no C++ types; every function returns 0 on success, <0 for an invalid argument /
 *     unsupported shape, >0 for a cudaError_t.  vs_last_error() returns the message.
 *   - the caller owns all device memory; pointers are raw device pointers; nothing is retained.
 *   - all launches are asynchronous on the caller's stream (passed as void* = cudaStream_t); no host sync.
 *   - bf16 matrices are row-major with 16-byte aligned base and row stride.
 *   - there is no CPU fallback: without a CUDA device every compute entry fails.
 */
#ifndef VITSEG_H_
#define VITSEG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VS_ABI_VERSION 3

const char* vs_last_error(void);
int vs_abi_version(void);
/* number of SMs of the current device (148 on B200); <0 on error */
int vs_device_sm_count(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM on tcgen05 tensor cores: D[M,N] = epilogue( sum_k A(m,k) * B(n,k) ), bf16 operands, fp32 TMEM accumulate.
 * Replaces nn.Linear / nn.Conv2d-as-GEMM forward, dgrad and wgrad (TF:151,216-218,262,290,305;
 * model/CE/classes.py:241) and their autograd backward.
 *
 * Operand storage ("major"):
 *   a_mn_major = 0 : A is a row-major [M, K] matrix (lda >= K)           — activations in forward/dgrad
 *   a_mn_major = 1 : A is a row-major [K, M] matrix (lda >= M), i.e. A^T — dY in wgrad
 *   b_mn_major = 0 : B is a row-major [N, K] matrix (ldb >= K)           — nn.Linear weight [out,in]
 *   b_mn_major = 1 : B is a row-major [K, N] matrix (ldb >= N)           — weight in dgrad, X in wgrad
 *
 * Epilogue, applied per element v = acc (fp32) in this order:
 *   v += bias[n]                               (bias != NULL; fp32 [N])
 *   out2[m,n] = bf16(v)                        (out2 != NULL; pre-activation copy kept for backward)
 *   v = act(v)                                 (act: 0 none, 1 exact-erf GELU, 2 ReLU)
 *   v *= gelu'(aux[m,n])  | v = aux[m,n]>0?v:0 (aux_mode 1 | 2; aux bf16 [M, ldaux])
 *   v += residual[rm, n]                       (residual != NULL; fp32, ld ldr)
 *   store to out (out_dtype 0 = bf16, 1 = fp32); accumulate != 0: atomic fp32 add (split-K / grad accumulation)
 *   out_colsum[n] += sum_m bf16(out[m,n])       (out_colsum != NULL; bf16 outputs)
 * Broadcast residual (patch embedding + position embedding, TF:117-124): row_tokens = T1 > 0 reads residual row
 *   (m % T1), i.e. a [T1, N] table shared by every image; 0 = residual row m.
 * split_k: 0 = automatic (only >1 when accumulate != 0), else the number of K partitions.
 * Tiling: CTA pairs (tcgen05 cta_group::2, 256 x {256,192,128} tiles) or single CTAs (128 x {256,128}), chosen by a
 *   wave-quantisation cost model unless tile_cfg forces one.
 * ------------------------------------------------------------------------------------------------ */
typedef struct vs_gemm_desc {
  int32_t M, N, K;
  int32_t a_mn_major, b_mn_major;
  const void* A;
  int64_t lda;
  const void* B;
  int64_t ldb;
  void* out;
  int64_t ldo;
  int32_t out_dtype;
  int32_t accumulate;
  const float* bias;
  int32_t act;
  void* out2;
  int64_t ldo2;
  const void* aux;
  int64_t ldaux;
  int32_t aux_mode;
  const float* residual;
  int64_t ldr;
  int32_t row_tokens;
  int32_t split_k;
  int32_t tile_cfg; /* 0 = automatic; 1..5 force {pair 256xBN256, pair BN192, pair BN128, single-CTA BN256, BN128} */
  /* hidden-state dropout (ViTSelfOutput / ViTOutput, TF:267,310): v = keep ? v/(1-p) : 0 applied after bias/act and
   * before the residual add; fp32 outputs only.  dropout_seed is a DEVICE pointer to the per-step counter. */
  float dropout_p;
  const uint32_t* dropout_seed;
  uint32_t dropout_site;
  /* ABI 3: column sums of the (bf16-rounded) output, out_colsum[n] += sum_m out[m,n] — the bias gradient of the Linear
   * that produced this GEMM's A gradient (autograd of nn.Linear bias, TF:290).  bf16 outputs only.  Formed inside the
   * epilogue from the staged shared-memory tile when the TMA-store epilogue runs, by a following vs_colsum_bf16 launch
   * otherwise: the result is the same either way. */
  float* out_colsum;
} vs_gemm_desc;

int vs_gemm_bf16(const vs_gemm_desc* d, void* stream);

/* column sums of a bf16 [M, N] matrix into fp32 out[N] (bias gradients; autograd of nn.Linear bias).
 * accumulate != 0 adds to out. */
int vs_colsum_bf16(const void* x, int64_t ldx, int32_t M, int32_t N, float* out, int32_t accumulate, void* stream);
/* same, but the first nf columns of the matrix are still fp32 in x_f32 [M, ldf] (the dQ accumulator of the attention
 * backward): they are rounded to bf16, STORED into x[:, :nf] and summed from the rounded values — one pass instead of a
 * cast pass plus a column-sum pass. */
int vs_colsum_cast_bf16(void* x, int64_t ldx, int32_t M, int32_t N, float* out, int32_t accumulate, const float* x_f32,
                        int64_t ldf, int32_t nf, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm (TF:325-326,333,340,416,455; eps 1e-12), fp32 residual stream in, bf16 (and/or fp32) out.
 *   x fp32 [M, D]; gamma/beta fp32 [D]; y_bf16 and/or y_f32 may be NULL; mean/rstd fp32 [M] may be NULL (inference).
 * Backward: dx_out[M,D] (fp32) = dx_in (nullable: skip-connection gradient) + LN'(dy); dgamma/dbeta accumulate (+=).
 *   dy is bf16 or fp32 (dy_is_f32); dx_bf16 (nullable) receives a bf16 copy of dx_out, multiplied by the dropout
 *   mask of site `dropout_site` when dropout_p > 0 (it feeds the backward GEMMs of the projection whose output was
 *   dropped out in forward).  dbias_colsum (nullable, fp32 [D], +=) receives the column sums of that bf16 output,
 *   i.e. the bias gradient of that projection (autograd of nn.Linear bias) without a separate vs_colsum_bf16 pass.
 *   All matrix operands must be 16-byte aligned (rows are streamed with cp.async.bulk).
 * ------------------------------------------------------------------------------------------------ */
int vs_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int32_t M, int32_t D,
                     void* y_bf16, float* y_f32, float* mean, float* rstd, void* stream);
int vs_layernorm_bwd(const void* dy, int32_t dy_is_f32, const float* x, const float* gamma, const float* mean,
                     const float* rstd, const float* dx_in, int32_t M, int32_t D, float* dx_out, void* dx_bf16,
                     float* dgamma, float* dbeta, float* dbias_colsum, float dropout_p, const uint32_t* dropout_seed,
                     uint32_t dropout_site, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-head self-attention (TF:220-251 + sdpa_attention_forward): softmax(Q K^T * scale) V per (batch, head).
 *   qkv  bf16 [B, N, 3, H, 64]  (fused projection output; head_dim must be 64)
 *   ctx  bf16 [B, N, H, 64]
 *   lse  fp32 [B, H, N]  natural-log row logsumexp of the scaled scores (saved for backward; may be NULL)
 * Backward: dctx bf16 [B,N,H,64] -> dqkv bf16 [B,N,3,H,64].  delta is an fp32 [B,H,N] scratch (caller-provided).
 *   dq_accum != NULL: fp32 [B,N,H,64] scratch that RECEIVES dQ (the Q slot of dqkv is left untouched: the caller
 *     rounds dq_accum into it, e.g. with vs_colsum_cast_bf16) — any N.
 *   dq_accum == NULL: N <= 256 only; all three slots of dqkv are written as bf16 by a kernel that keeps whole
 *     (batch, head) items on one SM (no fp32 atomics, no cast pass).
 * dropout_p > 0 drops attention probabilities (attention_probs_dropout_prob, SDPA dropout_p) with a counter-based
 *   mask regenerated identically in backward.
 * ------------------------------------------------------------------------------------------------ */
int vs_attention_fwd(const void* qkv, void* ctx, float* lse, int32_t B, int32_t N, int32_t H, float scale,
                     float dropout_p, const uint32_t* dropout_seed, uint32_t dropout_site, void* stream);
int vs_attention_bwd(const void* qkv, const void* ctx, const void* dctx, const float* lse, void* dqkv,
                     float* dq_accum, float* delta, int32_t B, int32_t N, int32_t H, float scale, float dropout_p,
                     const uint32_t* dropout_seed, uint32_t dropout_site, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Embedding glue (TF:100-128, 153-167)
 *   vs_patchify: image fp32 NCHW [B,3,S,S] -> bf16 patch matrix [B*(T+1), 3*P*P], K ordered (c,ph,pw) = the
 *                row-major flattening of projection.weight [D,3,P,P].  Row b*(T+1)+1+t holds patch t; the CLS
 *                rows b*(T+1) are never written (the caller zeroes the buffer once) so that the projection GEMM,
 *                its wgrad and the token matrix share one row indexing.
 *   vs_cls_rows: x[b, 0, :] = cls[:] + pos[0, :]   (x fp32 [B, T+1, D])
 *   vs_embed_bwd: dcls[D] += sum_b dx[b,0,:];  dpos[T+1, D] += sum_b dx[b,:,:];
 *                 dbias[D] += sum_{b, t>=1} dx[b,t,:]  (patch projection bias)
 * ------------------------------------------------------------------------------------------------ */
int vs_patchify(const float* img, void* out, int32_t B, int32_t S, int32_t P, void* stream);
int vs_cls_rows(const float* cls, const float* pos, float* x, int32_t B, int32_t T1, int32_t D, void* stream);
int vs_embed_bwd(const float* dx, float* dcls, float* dpos, float* dbias, int32_t B, int32_t T1, int32_t D,
                 void* stream);

/* ------------------------------------------------------------------------------------------------
 * Segmentation head (model/CE/classes.py:240-244,250-257)
 *   vs_head_im2col : tokens bf16 [B, T+1, D] (CLS dropped) -> 3x3 zero-padded patch matrix bf16 [B*T, 9*D],
 *                    K ordered (ky,kx,c)
 *   vs_head_col2im : dcol bf16 [B*T, 9*D] -> dtokens fp32 [B, T+1, D] (CLS row zeroed)
 *   vs_conv1x1_fwd : feat bf16 [B*T, F] x w fp32 [C, F] + b -> low-res logits fp32 NCHW [B, C, g, g]
 *   vs_conv1x1_bwd : dlogits fp32 [B,C,g,g] -> dfeat bf16 [B*T,F] (masked by feat>0 = ReLU'), dw[C,F] +=, db[C] +=
 * ------------------------------------------------------------------------------------------------ */
int vs_head_im2col(const void* tokens, void* col, int32_t B, int32_t g, int32_t D, void* stream);
int vs_head_col2im(const void* dcol, float* dtokens, int32_t B, int32_t g, int32_t D, void* stream);
int vs_conv1x1_fwd(const void* feat, const float* w, const float* b, float* logits, int32_t B, int32_t g, int32_t F,
                   int32_t C, void* stream);
int vs_conv1x1_bwd(const float* dlogits, const void* feat, const float* w, void* dfeat, float* dw, float* db,
                   int32_t B, int32_t g, int32_t F, int32_t C, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Bilinear upsample (F.interpolate(mode='bilinear', align_corners=False), model/CE/classes.py:260)
 *   low fp32 [B,C,g,g] -> full fp32 [B,C,S,S];  _bwd is its adjoint (dfull -> dlow, overwrites dlow);
 *   _argmax writes uint8 [B,S,S] class ids (testViTModel.py:121-126: argmax of sigmoid(logits) == argmax logits;
 *   C == 1 thresholds logit > 0, i.e. sigmoid > 0.5).
 * ------------------------------------------------------------------------------------------------ */
int vs_upsample_bilinear_fwd(const float* low, float* full, int32_t B, int32_t C, int32_t g, int32_t S, void* stream);
int vs_upsample_bilinear_bwd(const float* dfull, float* dlow, int32_t B, int32_t C, int32_t g, int32_t S,
                             void* stream);
int vs_upsample_argmax(const float* low, uint8_t* mask, int32_t B, int32_t C, int32_t g, int32_t S, void* stream);
/* class ids -> RGB through a [C,3] uint8 palette (colored_pred = index_to_color[pred_labels],
 *   model/CE/testViTModel.py:139-143): mask uint8 [n] -> rgb uint8 [n,3]; ids >= C map to black. */
int vs_colorize_mask(const uint8_t* mask, const uint8_t* palette, uint8_t* rgb, int64_t n, int32_t C, void* stream);
/* _argmax_stats: the same class map (mask may be NULL) plus exact per-image, per-class pixel counts against int64
 *   labels [B,S,S]: counts int32 [B, NC, 3] = {intersection, predicted, target}, NC = C (2 for C == 1); overwritten.
 *   Labels outside [0, NC) are not counted as targets.  Pixel accuracy / IoU / Dice / precision / recall of
 *   model/PAED/classes.py:430-447,684-689, model/PAED/segmentation.py:38-86 and
 *   model/CE/datasetTestViTmodel.py:188-217 are functions of these counts (visiontransformer_b200/metrics.py). */
int vs_upsample_argmax_stats(const float* low, const int64_t* labels, uint8_t* mask, int32_t* counts, int32_t B,
                             int32_t C, int32_t g, int32_t S, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Data-parallel gradient all-reduce in the NVSwitch (net-new: the reference trains on one device,
 * model/CE/createViTmodel.py:72-73).  multicast_ptr = multicast (NVLS) address of n fp32 elements of a buffer that
 * every rank allocated symmetrically; on return of the kernel this rank's 1/world slice holds scale * (sum over ranks)
 * on EVERY rank (multimem.ld_reduce + multimem.st).  All ranks must call it for the same range, bracketed by
 * cross-GPU barriers (done by visiontransformer_b200/dp.py).  One 128-thread, <= 32-register CTA per SM: sized to be
 * resident next to a tcgen05 GEMM CTA instead of displacing it.
 * ------------------------------------------------------------------------------------------------ */
int vs_multimem_allreduce_f32(void* multicast_ptr, int64_t n, int32_t rank, int32_t world, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Worker-side pre-processing: Pillow's Image.resize(size, BILINEAR) — what
 * transforms.Resize((224, 224)) does to the PIL image at model/CE/testViTModel.py:92-97 — followed by ToTensor, on
 * planar uint8 [C,H,W] device images (the layout nvJPEG delivers).  Two passes as in libImaging/Resample.c: horizontal
 * into a rounded uint8 intermediate [C,H,Wout], then vertical; 22-bit fixed-point weights.  bounds [out*2] = (first
 * source index, tap count) and coef [out*ksize] int32 per output coordinate come from the host (precompute_coeffs /
 * normalize_coeffs_8bpc restated in visiontransformer_b200/worker.py).  Bit-identical to Pillow for the same pixels.
 *   vs_resample_v_u8 writes dst_f32 = value / scale (ToTensor: scale = 255, a true division like .div(255)) and/or
 *   dst_u8 (either may be NULL).
 *   vs_u8_to_f32: ToTensor alone (no resize needed).
 * ------------------------------------------------------------------------------------------------ */
int vs_resample_h_u8(const uint8_t* src, int64_t row_stride, int64_t plane_stride, int32_t C, int32_t H, int32_t W,
                     const int32_t* bounds, const int32_t* coef, int32_t ksize, int32_t Wout, uint8_t* dst, void* stream);
int vs_resample_v_u8(const uint8_t* src, int32_t C, int32_t H, int32_t Wout, const int32_t* bounds, const int32_t* coef,
                     int32_t ksize, int32_t Hout, float* dst_f32, float scale, uint8_t* dst_u8, void* stream);
int vs_u8_to_f32(const uint8_t* src, float* dst, int64_t n, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused upsample + cross-entropy (model/CE/classes.py:276-285: F.interpolate -> nn.CrossEntropyLoss, mean over
 * B*S*S, ignore_index -100), with the label resize of model/CE/classes.py:273-274 folded into the label read.
 *   labels [B, LH, LW], int64 (label_dtype 0, as the reference's dataset produces) or uint8 (label_dtype 1, 255 is NOT
 *   an ignore value: only int64 -100 is).  LH x LW != S x S: the label of output pixel (y, x) is taken at
 *   (min(floor(y * LH/S), LH-1), min(floor(x * LW/S), LW-1)), i.e. F.interpolate(mode='nearest') of the label map.
 *   loss_sum[0] += sum of per-pixel NLL, loss_sum[1] += number of non-ignored pixels   (caller zeroes)
 *   dlow (nullable) fp32 [B,C,g,g] = d(sum NLL)/d low  (caller zeroes; host scales by 1/count)
 * ------------------------------------------------------------------------------------------------ */
int vs_upsample_ce(const float* low, const void* labels, int32_t label_dtype, int32_t LH, int32_t LW, float* loss_sum,
                   float* dlow, int32_t B, int32_t C, int32_t g, int32_t S, void* stream);

/* ------------------------------------------------------------------------------------------------
 * PAED binary loss (model/PAED/classes.py:608-701): p = sigmoid(up(low)), BCE, Dice sums, Sobel edge map,
 * SDF-weighted sums.  C == 1.  mask / sdf_ext / sdf_int fp32 [B,S,S]; low fp32 [B,1,g,g].
 *   pass 1 (vs_paed_binary_stats), caller zeroes stats and keys:
 *       stats[b*8 + 0..5] += {sum bce, sum p*t, sum p, sum t, sum sdf_int*p, sum sdf_ext*edge} of image b
 *       keys[b] = max over pixels of (float_bits(edge) << 32 | ~pixel_index): the per-image max edge and its FIRST
 *                 arg-max (torch.max backward semantics)
 *   pass 2 (vs_paed_binary_bwd): coef[b*8 + 0..6] = dL/d{bce_sum, inter, psum, tsum, int_sum, ext_sum, max_edge} of
 *       image b; accumulates dlow (caller zeroes).  The scalar algebra between the passes (means, Dice ratio,
 *       abs(paed), cross-rank sums) is host code on [B,8] tensors.
 * ------------------------------------------------------------------------------------------------ */
int vs_paed_binary_stats(const float* low, const float* mask, const float* sdf_ext, const float* sdf_int,
                         float* stats, uint64_t* keys, int32_t B, int32_t g, int32_t S, void* stream);
int vs_paed_binary_bwd(const float* low, const float* mask, const float* sdf_ext, const float* sdf_int,
                       const float* coef, const uint64_t* keys, float* dlow, int32_t B, int32_t g, int32_t S,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * PAED multi-class soft loss (model/PAED/classes.py:336-369 applied to softmax(up(low)) and one-hot labels):
 *   loss = mean( m (1-p) 2 |blur(m - p)| ), blur = separable 19-tap Gaussian (sigma 3), zero padding
 *   (SURVEY Appendix D3).  labels int64 [B,S,S]; scratch fp32 buffers t1,t2,t3 of [B,C,S,S] (t3 only when dlow).
 *   loss_sum[0] += sum over all elements (host divides by B*C*S*S); dlow (nullable) += d(sum)/d low (caller zeroes).
 * ------------------------------------------------------------------------------------------------ */
int vs_paed_multiclass(const float* low, const int64_t* labels, float* t1, float* t2, float* t3, float* loss_sum,
                       float* dlow, int32_t B, int32_t C, int32_t g, int32_t S, void* stream);

/* dense form of the same loss for callers holding [B,C,S,S] one-hot mask and probability tensors (the free
 * function model/PAED/classes.py:336-369; sigma fixed at 3): loss_sum[0] += sum(penalty * |blur(msk - prob)|) with
 * penalty = 2 msk (1 - prob) when class_penalty else 1; dprob (nullable) = d(sum)/d prob (overwritten). */
int vs_paed_multiclass_dense(const float* msk, const float* prob, float* t1, float* t2, float* t3, float* loss_sum,
                             float* dprob, int32_t B, int32_t C, int32_t S, int32_t class_penalty, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dropout glue.  vs_dropout_rows: in-place dropout of an fp32 buffer of n elements (n % 4 == 0) + optional bf16 copy —
 * the embedding dropout (TF:126) in forward and the same mask on the gradient in backward.  vs_dropout_mask writes
 * the keep mask of a site as bytes (scheme 0: hidden states, dense index space; scheme 1: attention probabilities
 * viewed as rows of row_len keys); tests.
 * ------------------------------------------------------------------------------------------------ */
int vs_dropout_rows(float* x, void* x_bf16, int64_t n, float dropout_p, const uint32_t* dropout_seed,
                    uint32_t dropout_site, void* stream);
int vs_dropout_mask(uint8_t* out, int64_t n, int32_t scheme, int32_t row_len, float dropout_p,
                    const uint32_t* dropout_seed, uint32_t dropout_site, void* stream);

/* ------------------------------------------------------------------------------------------------
 * PAED signed-distance targets on the device: compute_sdf (model/PAED/segmentation.py:6-34) for a batch of masks.
 *   mask fp32 [B,S,S] (> 0.5 = object) -> sdf_ext = EDT(~mask) / max, sdf_int = EDT(mask) / max, fp32 [B,S,S], each
 *   divided by its own per-image maximum when that is > 0.  Exact Euclidean distance transform (integer squared
 *   distances, double-precision root rounded to fp32 as SciPy does); images without any zero element reproduce
 *   scipy.ndimage.distance_transform_edt's virtual zero at (row -1, col 0).
 *   workspace: vs_sdf_workspace_bytes(B, S) bytes of device memory, 4-byte aligned, caller-owned.
 * ------------------------------------------------------------------------------------------------ */
int64_t vs_sdf_workspace_bytes(int32_t B, int32_t S);
int vs_sdf_targets(const float* mask, float* sdf_ext, float* sdf_int, void* workspace, int32_t B, int32_t S,
                   void* stream);

/* ------------------------------------------------------------------------------------------------
 * Weight shadows: fp32 master -> bf16 copy (one pass), plus utility conversions.
 *   vs_cast_f32_bf16: n elements.  vs_pack_conv3x3: OIHW fp32 [O,I,3,3] -> bf16 [O, (ky,kx,I)].
 *   vs_unpack_conv3x3_grad: fp32 [O,(ky,kx,I)] -> += OIHW fp32 grad.
 *   vs_dq_finalize: fp32 [M, D] -> bf16 strided copy (dq accumulators into dqkv[:, 0:D]).
 * ------------------------------------------------------------------------------------------------ */
int vs_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream);
int vs_cast_bf16_rows(const float* src, int64_t lds, void* dst, int64_t ldd, int32_t M, int32_t D, void* stream);
int vs_pack_conv3x3(const float* w, void* out, int32_t O, int32_t I, void* stream);
int vs_unpack_conv3x3_grad(const float* g, float* dw, int32_t O, int32_t I, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused Adam / AdamW over the flat fp32 parameter arena (torch.optim.Adam / AdamW semantics, no amsgrad):
 * one pass updates param / exp_avg / exp_avg_sq, writes the bf16 weight shadow and optionally zeroes grad.
 * lr_dev / step_dev are DEVICE pointers (learning rate; 1-based step count) so captured graphs follow schedulers.
 * Elements in [skip_begin, skip_end) are left untouched (parameters that never receive a gradient: the pooler).
 * ------------------------------------------------------------------------------------------------ */
int vs_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16, int64_t n,
                 const float* lr_dev, const int32_t* step_dev, float beta1, float beta2, float eps, float weight_decay,
                 int32_t decoupled, float grad_scale, int32_t zero_grad, int64_t skip_begin, int64_t skip_end,
                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITSEG_H_ */
