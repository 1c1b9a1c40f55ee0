/*
 * vfmseg_b200 — C ABI of the B200-native (sm_100a) slide-inference hot path.
 *
 * The reference (tpy001/VFMSeg) is pure Python and has no FFI of its own; what a maintainer binds
 * instead are the torch module calls on its hot path. Every entry point below names the
 * reference call it replaces (paths relative to the reference repo). See INTEGRATION.md for the
 * ctypes stub a `rein` maintainer would add.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the parameter says "host"; the caller owns every
 *     buffer; nothing here allocates device memory or synchronises the stream;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - activations are token-major: row = (crop, token), channels contiguous;
 *   - bf16 buffers are `void*`, 16-byte aligned, leading dimensions multiples of 8 elements;
 *   - return value: 0 = ok, < 0 = error (message in vfm_last_error(), thread-local).
 */
#ifndef VFMSEG_B200_H_
#define VFMSEG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VFM_ABI_VERSION 2   /* 2: VfmBlockParams grew the folded-LayerNorm weight sets */

enum {
  VFM_OK = 0,
  VFM_ERR_INVALID = -1,   /* bad shape / alignment / null pointer */
  VFM_ERR_CUDA = -2,      /* a CUDA runtime / driver call failed  */
  VFM_ERR_WORKSPACE = -3, /* workspace too small                  */
  VFM_ERR_DEVICE = -4     /* current device is not sm_100         */
};

const char* vfm_last_error(void);
int vfm_abi_version(void);
/* Number of kernels this library has launched in this process (for bench.py's gpu_launches). */
long long vfm_launch_count(void);
/* Per-launch timing for bench.py's roofline leg: while enabled every kernel launch is bracketed by
 * CUDA events on its own stream. vfm_prof_report synchronises the device and writes one
 * "name,launches,total_ms\n" line per kernel family into buf (host), then clears the records. */
int vfm_prof_enable(int on);
int vfm_prof_report(char* buf /*host*/, size_t buf_bytes);
/* 0 when the current CUDA device can run this library (compute capability 10.x). */
int vfm_device_check(void);

/* ---------------------------------------------------------------- dense operators (tcgen05)
 * All GEMMs compute acc[M,N] = A[M,K] * W[N,K]^T with bf16 operands and fp32 accumulation in
 * TMEM; W is exactly the nn.Linear / 1x1-conv weight layout. They differ in the fused epilogue. */

/* out = bf16(acc + bias).  bias may be NULL.
 * Replaces nn.Linear qkv (+ merged peft LoRA) rein/models/backbones/dino_layers/attention.py:51,58
 * and the bias-free 1x1 fusion conv rein/models/heads/linear_head.py:36-40. */
int vfm_gemm_bias_bf16(const void* A, int lda, const void* W, int ldw, const float* bias, void* out,
                       int ldo, int M, int N, int K, void* stream);

/* out = bf16(gelu_erf(acc + bias)).  Replaces Mlp.fc1 + nn.GELU, dino_layers/mlp.py:35-36. */
int vfm_gemm_bias_gelu_bf16(const void* A, int lda, const void* W, int ldw, const float* bias,
                            void* out, int ldo, int M, int N, int K, void* stream);

/* x[M,N] (fp32, in place) += gamma * (acc + bias); when tap != NULL also stores bf16(x) into
 * tap[(row - crop - 1), tap_col0 + col] for every non-cls row (crop = row / tokens_per_crop).
 * Replaces attn.proj / mlp.fc2 + LayerScale + residual add, dino_layers/block.py:112-113,
 * layer_scale.py:27, and the feature tap dino_v2.py:261-267. */
int vfm_gemm_bias_ls_residual(const void* A, int lda, const void* W, int ldw, const float* bias,
                              const float* gamma, float* x, int ldx, void* tap, int tap_ld,
                              int tap_col0, int tokens_per_crop, int M, int N, int K, void* stream);

/* The same residual update (x bit-identical to vfm_gemm_bias_ls_residual) that also prepares the LayerNorm which always
 * follows it (dino_layers/block.py:89-90,112-113): xb[M,N] = bf16(x_new) and stats[row][N/128] = (sum, sum of squares) of
 * x_new over each 128-column slot (fp32 pairs; written once per slot, no atomics). N % 256 == 0. */
int vfm_gemm_bias_ls_residual_stats(const void* A, int lda, const void* W, int ldw, const float* bias,
                                    const float* gamma, float* x, int ldx, void* xb, int ldxb, float* stats,
                                    int M, int N, int K, void* stream);

/* out = bf16(act(Linear(LayerNorm(x)))) with the norm folded away: A = bf16(x) [M,K] and stats [M][K/128] from
 * vfm_gemm_bias_ls_residual_stats, Wf = bf16(ln_w * W), bias_f = b + W ln_b, colsum[n] = sum_k Wf[n,k];
 * out = act(rstd_r * (acc - mean_r * colsum) + bias_f), act = exact-erf GELU when gelu != 0. K % 256 == 0.
 * Replaces norm1 -> attn.qkv and norm2 -> mlp.fc1 -> GELU (dino_layers/block.py:89-90, attention.py:51, mlp.py:35-36). */
int vfm_gemm_lnfold_bf16(const void* A, int lda, const void* Wf, int ldw, const float* bias_f, const float* colsum,
                         const float* stats, float eps, int gelu, void* out, int ldo, int M, int N, int K,
                         void* stream);

/* vfm_gemm_lnfold_bf16 (no activation) followed by the rotary embedding of vfm_gemm_bias_rope_bf16: norm1 -> attn.qkv ->
 * RoPE of EVA02 (rein/models/backbones/eva_02.py:337-369) with the LayerNorm folded into the weights. */
int vfm_gemm_lnfold_rope_bf16(const void* A, int lda, const void* Wf, int ldw, const float* bias_f, const float* colsum,
                              const float* stats, float eps, void* out, int ldo, int M, int N, int K,
                              const float* cos_t, const float* sin_t, int rope_cols, int tokens_per_seq, void* stream);

/* x[crop*(patches+1) + 1 + p, :] = acc + bias + pos[1 + p, :] for GEMM row = crop*patches + p.
 * Replaces PatchEmbed.proj (conv k=s=16) + pos-embed add, patch_embed.py:75-77, dino_v2.py:219-226. */
int vfm_gemm_patch_embed(const void* A, int lda, const void* W, int ldw, const float* bias,
                         const float* pos, float* x, int patches, int M, int N, int K, void* stream);

/* ConvTranspose2d(k=2,s=2) (+ folded eval BatchNorm) + GELU as GEMM + pixel shuffle.
 * W is [4*c_out, K] with row (dy*2+dx)*c_out + co; in rows = crop*h*w + y*w + x; out is
 * token-major [n*4hw, c_out]. Replaces linear_head.py:42-48. */
int vfm_gemm_convt2x2_gelu(const void* A, int lda, const void* W, int ldw, const float* bias,
                           void* out, int c_out, int h, int w, int M, int K, void* stream);

/* conv_seg: out[crop, cls, pix] (fp32 NCHW) = acc + bias[cls]; W is [32, K] zero-padded past
 * num_classes (<= 32). Replaces mmseg BaseDecodeHead.cls_seg as used at linear_head.py:68. */
int vfm_gemm_cls_nchw(const void* A, int lda, const void* W, int ldw, const float* bias, float* out,
                      int num_classes, int pix_per_crop, int M, int K, void* stream);

/* out(fp32) = acc (+ bias): plain GEMM used by the parity tests. */
int vfm_gemm_f32(const void* A, int lda, const void* W, int ldw, const float* bias, float* out,
                 int ldo, int M, int N, int K, void* stream);

/* softmax(Q K^T) V per (sequence, head), head_dim 64, scale already folded into Q.
 * qkv: [n_seq*seq_len, 3*heads*64] bf16 as emitted by the qkv GEMM; out: [n_seq*seq_len, heads*64].
 * Replaces dino_layers/attention.py:58-66 (and the xformers branch :79-84). */
int vfm_attention_fwd(const void* qkv, void* out, int n_seq, int seq_len, int heads, void* stream);
/* Same, choosing how a sequence is tiled: mode 0 = automatic, 1 = tensor-core tiles over every token, 2 = token 0 of
 * each sequence (the ViT cls token, dino_v2.py:225) is split off and handled on the CUDA cores so that 1 + 64k
 * tokens need no partial tile; 3 = the serial kernel (one score buffer, four CTAs per SM; an alternative kept for comparison);
 * 4 = the ping-pong kernel (256-query units, 128-key tiles, one persistent CTA per SM, attention_pp_sm100.cuh), 5 = the same
 * with token 0 split off; 6 / 7 = modes 4 / 5 with the softmax warps working per 64-key half (attention_ph_sm100.cuh; an experiment,
 * measured slower). Mode 0 picks 5 for ViT windows (seq_len = 1 + a multiple of 256, >= 513), whatever n_seq is, and 1
 * otherwise. The result is the same softmax attention in every mode. */
int vfm_attention_fwd_ex(const void* qkv, void* out, int n_seq, int seq_len, int heads, int mode, void* stream);
/* Cross attention: q [n_seq*q_len, >= heads*64] (row pitch q_ld), kv [n_seq*kv_len, 2*heads*64] packed (k | v),
 * out [n_seq*q_len, heads*64] (row pitch out_ld). Replaces the attention core of
 * rein/models/heads/Transformer.py:113-136 (CrossAttention._forward with a context). */
int vfm_attention_cross(const void* q, int q_ld, const void* kv, int kv_ld, void* out, int out_ld, int n_seq, int q_len,
                        int kv_len, int heads, void* stream);

/* ---------------------------------------------------------------- memory-bound operators */

typedef struct {
  float mean[3];    /* per OUTPUT channel (RGB after the flip), in 0..255 units */
  float inv_std[3];
  int flip;         /* 1: stored channel order is BGR (mmseg bgr_to_rgb=True)   */
} VfmPixelNorm;

/* Gathers 16x16 patches of every crop window into the patch-embed GEMM operand
 * out[(crop*gh*gw + gy*gw + gx), c*256 + py*16 + px] (bf16). crops: device int[4*n_crops] =
 * {image, y1, x1, 0}. img is fp32 normalised NCHW (is_u8 = 0) or uint8 NCHW (is_u8 = 1, then
 * nrm applies mmseg SegDataPreProcessor: lora_dinov2_linear.py:13-21).
 * Replaces the crop slicing of slide_inference (Ms_VFM_encoder_decoder.py:433-442) and the
 * unfold implicit in patch_embed.py:75. */
int vfm_patch_gather(const void* img, int is_u8, const VfmPixelNorm* nrm /*host*/, int img_h,
                     int img_w, const int* crops, int n_crops, int gh, int gw, void* out,
                     void* stream);

/* x[crop*tokens, :] = cls_token + pos[0, :].  dino_v2.py:225-226. */
int vfm_cls_rows(float* x, const float* cls_token, const float* pos, int n_crops, int tokens, int C,
                 void* stream);

/* LayerNorm rows of fp32 x[M,C] -> bf16 out.  C in {128..1024, multiple of 128}. block.py:63,75. */
int vfm_layernorm(const float* x, const float* gamma, const float* beta, void* out, int M, int C,
                  float eps, void* stream);

/* Same, and additionally stores bf16(x) of every non-cls row into tap[(row - crop - 1), tap_col0 + c] (the feature tap
 * of dino_v2.py:261-267, taken where the residual stream is read anyway). out may be NULL (tap only). */
int vfm_layernorm_tap(const float* x, const float* gamma, const float* beta, void* out, int M, int C,
                      float eps, void* tap, int tap_ld, int tap_col0, int tokens_per_crop, void* stream);

/* GroupNorm(groups, C) (+ReLU) over bf16 token-major [n_crops*P, C]. linear_head.py:36-40. */
int vfm_groupnorm_relu(const void* in, void* out, const float* gamma, const float* beta,
                       int n_crops, int P, int C, int groups, float eps, int relu, void* stream);

/* Bilinear-resize every crop's low-res logits to the crop size, accumulate overlapping windows
 * in row-major crop order, divide by the cover count, argmax.  lowres: fp32
 * [n_img*n_crops, nc, lh, lw]; boxes: device int[2*n_crops] = {y1, x1}; labels: uint8 [n_img,H,W];
 * logits_out: optional fp32 [n_img, nc, H, W].
 * Replaces mmseg predict_by_feat resize + slide_inference pad/add/count/divide
 * (Ms_VFM_encoder_decoder.py:453-461) + postprocess_result argmax.
 * Kernel choice (all bit-identical): exact x4 windows, W % 4 == 0, nc <= 19 and aligned outputs run the class-major tile
 * kernel; other shapes a per-pixel gather. Environment VFM_MERGE_MODE=1 forces the round-1 window-major tile kernel (A/B
 * runs); the same variable makes vfm_ms_merge_argmax use its per-pixel kernel. */
int vfm_slide_merge_argmax(const float* lowres, const int* boxes, int n_crops, int nc, int crop_h,
                           int crop_w, int lh, int lw, int H, int W, int n_img, uint8_t* labels,
                           float* logits_out, void* stream);

/* Second pass of the flip test-time augmentation fused into the merge: lowres / boxes are the windows of the MIRRORED image,
 * a = slide(img) (fp32 [n_img, nc, H, W], e.g. logits_out of vfm_slide_merge_argmax on the first pass); writes
 * labels = argmax((a + flip(slide(flip(img)))) / 2) and, when logits_out != NULL (may alias a), those logits — the same bits as
 * vfm_slide_merge_argmax + vfm_tta_flip_mean_argmax without materialising the second pass's logits.
 * rein/models/segmentors/hrda_encoder_decoder.py:196-229 (scales = [1]). Needs W % 4 == 0, x4 windows, nc <= 19. */
int vfm_slide_merge_flip_argmax(const float* lowres, const int* boxes, int n_crops, int nc, int crop_h, int crop_w, int lh,
                                int lw, int H, int W, int n_img, const float* a, uint8_t* labels, float* logits_out,
                                void* stream);

/* Horizontal-flip test-time augmentation, combining step: a = slide(img), b = slide(flip(img, x)), both fp32
 * [n_img, nc, H, W]; logits = (a + mirror_x(b)) / 2 (optional output, may alias a), labels = argmax, first maximum
 * wins. Bit-identical to res = 0; res += a; res += flip(b); res / 2.
 * Replaces rein/models/segmentors/hrda_encoder_decoder.py:196-229 (scales = [1], flip = True) + the
 * postprocess_result argmax. */
int vfm_tta_flip_mean_argmax(const float* a, const float* b, int n_img, int nc, int H, int W,
                             uint8_t* labels, float* logits_out, void* stream);

/* cm[(min(label, nc)) * nc + pred] += 1 over pixels with label != ignore_index; cm is int64
 * [(nc+1) * nc], accumulated (not zeroed). area_intersect = diag, area_pred = column sums over all
 * nc+1 rows, area_label = row sums of the first nc rows.
 * Replaces mmseg IoUMetric.intersect_and_union as called at rein/dg_metrics.py:50-52. */
int vfm_confusion_matrix(const uint8_t* pred, const uint8_t* label, long long n, int nc,
                         int ignore_index, long long* cm, void* stream);

/* ---------------------------------------------------------------- fused drivers */

typedef struct {
  const float* ln1_w; const float* ln1_b;
  const void* qkv_w;  const float* qkv_b;   /* LoRA merged, q rows pre-scaled by head_dim^-0.5 */
  const void* proj_w; const float* proj_b;
  const float* ls1;                          /* LayerScale gamma (ones when absent) */
  const float* ln2_w; const float* ln2_b;
  const void* fc1_w;  const float* fc1_b;
  const void* fc2_w;  const float* fc2_b;
  const float* ls2;
  /* Optional (all three of a set, or null): the Linear that follows a LayerNorm with the norm folded in —
   * wf = bf16(ln_w[k] * W[n, k]), bf = b + W ln_b (fp32), cs[n] = sum_k wf[n, k] (fp32, of the rounded weights).
   * When present (and embed_dim % 256 == 0) vfm_vit_forward skips the LayerNorm pass: the preceding residual GEMM emits
   * bf16(x) and the row statistics (vfm_gemm_bias_ls_residual_stats), this GEMM applies them (vfm_gemm_lnfold_bf16). */
  const void* qkv_wf; const float* qkv_bf; const float* qkv_cs;   /* norm1 -> attn.qkv */
  const void* fc1_wf; const float* fc1_bf; const float* fc1_cs;   /* norm2 -> mlp.fc1 */
} VfmBlockParams;

typedef struct {
  int embed_dim, depth, heads, mlp_hidden, n_taps;
  int tap_blocks[8];            /* out_indices, ascending */
  float ln_eps;
  const void* patch_w;          /* [embed_dim, 768] bf16 */
  const float* patch_b;
  const float* cls_token;       /* [embed_dim] */
  const float* pos_embed;       /* [1 + gh*gw, embed_dim] for the grid of this call */
  const VfmBlockParams* blocks; /* host array [depth] */
} VfmVitParams;

size_t vfm_vit_workspace_bytes(const VfmVitParams* p /*host*/, int n_crops, int gh, int gw);

/* DinoVisionTransformer.forward_features over a batch of crop windows
 * (rein/models/backbones/dino_v2.py:252-268) with peft LoRA merged. Writes the four feature
 * taps token-major, cls dropped: taps[(crop*gh*gw + p), t*embed_dim + c] (bf16). */
int vfm_vit_forward(const VfmVitParams* p /*host*/, const void* img, int is_u8,
                    const VfmPixelNorm* nrm /*host*/, int img_h, int img_w, const int* crops,
                    int n_crops, int gh, int gw, void* taps, void* workspace,
                    size_t workspace_bytes, void* stream);

typedef struct {
  int in_channels;   /* n_taps * embed_dim */
  int mid_channels;  /* embed_dim */
  int groups, num_classes;
  float gn_eps;
  const void* fusion_w;                 /* [mid, in] bf16, no bias */
  const float* gn_w; const float* gn_b;
  const void* up1_w; const float* up1_b; /* [4*mid/2, mid], eval BatchNorm folded in */
  const void* up2_w; const float* up2_b; /* [4*mid/4, mid/2] */
  const void* cls_w; const float* cls_b; /* [32, mid/4] zero padded; [num_classes] */
} VfmLinearHeadParams;

size_t vfm_linear_head_workspace_bytes(const VfmLinearHeadParams* p /*host*/, int n_crops, int gh, int gw);

/* LinearHead.forward (rein/models/heads/linear_head.py:50-70) on token-major taps;
 * lowres: fp32 [n_crops, num_classes, 4*gh, 4*gw]. */
int vfm_linear_head_forward(const VfmLinearHeadParams* p /*host*/, const void* taps, int n_crops,
                            int gh, int gw, float* lowres, void* workspace, size_t workspace_bytes,
                            void* stream);

/* ---------------------------------------------------------------- coarse-to-fine path (MsVFMEncoderDecoder + VFMHead)
 * Stage 0 of ms_inference yields coarse logits low0 [n_img, nc, lh, lw] (fp32) for the whole down-scaled image; the
 * reference upsamples them to the full image and slices "context" windows (Ms_VFM_encoder_decoder.py:417,443). These
 * entry points sample low0 with the same bilinear (align_corners=False) arithmetic instead of storing that field. */

/* resize(inputs, size=(h, w), bilinear) of the network input, Ms_VFM_encoder_decoder.py:413, fused with mmseg
 * SegDataPreProcessor when the input is uint8 (configs/_base_/models/lora_dinov2_ms_masked.py:5-13). out: fp32 [B,3,h,w]. */
int vfm_image_resize_norm(const void* img, int is_u8, const VfmPixelNorm* nrm, int B, int H, int W, float* out, int h, int w,
                          void* stream);
/* counts[b * n_crops + k] = #pixels of window k of image b with max softmax(context) > thr
 * (Ms_VFM_encoder_decoder.py:446-448; the host compares count / (crop_h*crop_w) with test_cfg.conf). boxes: {y1, x1}. */
int vfm_ms_confidence(const float* low0, const int* boxes, int n_crops, int nc, int crop_h, int crop_w, int lh, int lw, int H,
                      int W, int n_img, float thr, int* counts, void* stream);
/* GEMM operand of VFMHead.seg_logits_embed[0] (Conv2d(nc, C/4, 2, 2), VFMHead.py:38-39) for the refined windows:
 * context window -> resize to (ctx_h, ctx_w) (VFMHead.py:63-67) -> 2x2 patches; out bf16 [n_ref*ctx_h/2*ctx_w/2, kpad],
 * column = cin*4 + dy*2 + dx, zero padded. crops: n_ref x {image, y1, x1, 0}. */
int vfm_ms_context_im2col(const float* low0, const int* crops, int n_ref, int nc, int crop_h, int crop_w, int lh, int lw, int H,
                          int W, int ctx_h, int ctx_w, void* out, int kpad, void* stream);
/* GEMM operand of a Conv2d(C, C', 2, 2) over token-major bf16 activations [n*h*w, C] (VFMHead.py:42):
 * out [n*h/2*w/2, 4C], column = (dy*2+dx)*C + c. */
int vfm_space_to_depth2(const void* in, void* out, int n, int h, int w, int C, void* stream);
/* GroupNorm(groups, C) + activation over token-major bf16 [n*P, C], any channels-per-group; act 0 none / 1 ReLU /
 * 2 exact-erf GELU; out bf16, or fp32 when out_f32 (VFMHead.py:30-32,40-47; Transformer.py:91-92,247). */
int vfm_groupnorm_act(const void* in, void* out, int out_f32, const float* gamma, const float* beta, int n, int P, int C, int groups,
                      float eps, int act, void* stream);
/* GEGLU, Transformer.py:52-59: in bf16 [M, 2I] = (x | gate) -> out bf16 [M, I] = x * gelu(gate). */
int vfm_geglu(const void* in, void* out, long long M, int I, void* stream);
int vfm_cast_f32_bf16(const float* in, void* out, long long n, void* stream);
/* Stage-1 merge of ms_inference (Ms_VFM_encoder_decoder.py:449-461) + argmax: windows with ref_index >= 0 contribute
 * refined[ref_index] ([nc, rh, rw] resized to the window), the others the context value; sum / count; first max wins. */
int vfm_ms_merge_argmax(const float* low0, const float* refined, const int* ref_index, const int* boxes, int n_crops, int nc,
                        int crop_h, int crop_w, int lh, int lw, int rh, int rw, int H, int W, int n_img, uint8_t* labels,
                        float* logits_out, void* stream);

/* ---- SAM ViT backbone (BASELINE config 5), rein/models/backbones/sam_vit.py ---------------------------------------- */
/* Same as vfm_gemm_patch_embed / vfm_layernorm_tap with cls_rows = 0: no cls row in x / pos / the taps (SAMViT has no cls
 * token, sam_vit.py:123-132). cls_rows = 1 is the behaviour of the plain entry points. */
int vfm_gemm_patch_embed_ex(const void* A, int lda, const void* W, int ldw, const float* bias, const float* pos, float* x,
                            int patches, int cls_rows, int M, int N, int K, void* stream);
int vfm_layernorm_tap_ex(const float* x, const float* gamma, const float* beta, void* out, int M, int C, float eps,
                         void* tap, int tap_ld, int tap_col0, int tokens_per_crop, int cls_rows, void* stream);
/* Same, with window_partition (sam_vit.py:292-316) folded into the store: out_map[input row] = output row (NULL: identity).
 * The padding rows of the destination are never written: the caller zero-fills the buffer once and reuses it. */
int vfm_layernorm_tap_map(const float* x, const float* gamma, const float* beta, void* out, int M, int C, float eps,
                          void* tap, int tap_ld, int tap_col0, int tokens_per_crop, int cls_rows, const int* out_map, void* stream);

/* rel[seq][head][token][0..k_h) = q . Rh[qh, kh, :], rel[..][k_h + kw] = q . Rw[qw, kw, :] with the UNSCALED q of the packed
 * qkv buffer (bf16 [n_seq * q_h * q_w, 3 * heads * head_dim]); Rh fp32 [q_h][k_h][head_dim], Rw fp32 [q_w][k_w][head_dim] are
 * get_rel_pos's gathers (sam_vit.py:358-388). Replaces the two einsums of add_decomposed_rel_pos, sam_vit.py:417-421. */
int vfm_relpos_terms(const void* qkv, const float* Rh, const float* Rw, float* rel, int n_seq, int heads, int head_dim,
                     int q_h, int q_w, int k_h, int k_w, void* stream);

/* dst[i, :] = map[i] >= 0 ? src[map[i], :] : 0 for bf16 rows of C elements (C % 8 == 0): window_partition with its zero
 * padding and window_unpartition, sam_vit.py:292-346. */
int vfm_rows_gather(const void* src, void* dst, const int* map, long long n_rows, int C, void* stream);

/* softmax(scale * q k^T + rel[.., kh(k)] + rel[.., k_h + kw(k)]) v per (sequence, head), key k = kh * k_w + kw; head_dim 64
 * or 80; rel may be NULL (no bias). Replaces Attention.forward's core, sam_vit.py:272-287. */
int vfm_attention_relpos(const void* qkv, const float* rel, void* out, int n_seq, int seq_len, int heads, int head_dim,
                         int k_h, int k_w, float scale, void* stream);
/* The same attention for windows whose keys fit one score tile (seq_len <= 208, head_dim 80) on tcgen05: the bias is added by
 * the tensor core ([rel_h | rel_w] x one-hot key columns as extra k-steps of the score MMA); arguments as
 * vfm_attention_relpos_ex with rel == NULL (g_col0 < 0: no bias). sam_vit.py:272-287,391-428. */
int vfm_attention_window_tc(const void* qkv, int ld, int g_col0, void* out, int n_seq, int seq_len, int heads, int head_dim,
                            int k_h, int k_w, float scale, void* stream);
/* Same, with window_unpartition (sam_vit.py:335-346) folded into the store: out_map[window-order row] = output row, < 0 for the
 * padding rows, which are not written (NULL: identity). */
int vfm_attention_window_tc_map(const void* qkv, int ld, int g_col0, void* out, const int* out_map, int n_seq, int seq_len, int heads,
                                int head_dim, int k_h, int k_w, float scale, void* stream);
/* The same attention over a whole token grid (key-tile loop, online softmax) on tcgen05, head_dim 80, k_h and k_w rounded up
 * to 16 summing to at most 128 bias columns (grids up to 64 x 64). onehot: bf16 [onehot_rows >= seq_len, 64 * ceil((bh + bw) / 64)],
 * row k = 1.0 at column kh(k) and at column bh + kw(k) (bh = k_h rounded up to 16) — a constant of the grid the caller builds
 * once. sam_vit.py:272-287,391-428. */
int vfm_attention_global_tc(const void* qkv, int ld, int g_col0, const void* onehot, int onehot_rows, void* out, int n_seq,
                            int seq_len, int heads, int head_dim, int k_h, int k_w, float scale, void* stream);
/* Same with an explicit row pitch ld (elements) of the qkv buffer and, when g_col0 >= 0 (then rel must be NULL), the bias
 * taken from table terms stored in the qkv rows: G_h[head][r] = q . T_h[r] (r in [0, 2 k_h - 1)) at column
 * g_col0 + head * (2 k_h - 1) + r, all heads' G_w behind them; rel_h[q, kh] = G_h[qh - kh + k_h - 1] (get_rel_pos,
 * sam_vit.py:382-388). G is linear in the block input: the engine emits it from the qkv GEMM (weights T . W_q). */
int vfm_attention_relpos_ex(const void* qkv, int ld, int g_col0, const float* rel, void* out, int n_seq, int seq_len, int heads,
                            int head_dim, int k_h, int k_w, float scale, void* stream);

/* qkv Linear + bias with the 2-D rotary embedding of vfm_rope_qk applied in the GEMM epilogue to columns [0, rope_cols) (the q
 * and k thirds) of every non-cls row, on the fp32 accumulator. eva_02.py:337-369. */
int vfm_gemm_bias_rope_bf16(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo, int M,
                            int N, int K, const float* cos_t, const float* sin_t, int rope_cols, int tokens_per_seq, void* stream);
/* ---------------------------------------------------------------- EVA02 backbone (rein/models/backbones/eva_02.py)
 * In-place 2-D rotary embedding of the q and k thirds of packed qkv activations [M, 3C] bf16 (VisionRotaryEmbeddingFast
 * :119-160 applied at :362-369); token 0 of every sequence (cls) is left untouched. cos/sin: fp32 [tokens_per_seq-1, 64]. */
int vfm_rope_qk(void* qkv, long long M, int C, int heads, int tokens_per_seq, const float* cos_t, const float* sin_t, void* stream);
/* SwiGLU.forward between its GEMMs (:234-239): in bf16 [M, 2*Hp] = (w1 x | w2 x) -> LayerNorm_H(silu(x1) * x2) -> out bf16
 * [M, Hp]; H = real hidden width (2730), Hp = padded width (multiple of 8), padding columns written as zeros. */
int vfm_swiglu_layernorm(const void* in, void* out, const float* gamma, const float* beta, long long M, int H, int Hp, float eps,
                         void* stream);


/* EVA2.forward_features over a batch of crop windows as ONE call (rein/models/backbones/eva_02.py:816-849, blocks :486-493,
 * attention :331-381, SwiGLU :234-241) with peft LoRA merged: the launch sequence vfmseg_b200/eva_engine.py used to issue from
 * Python (patch gather -> patch-embed GEMM + pos_embed -> cls rows -> per block: [LayerNorm | folded] qkv GEMM + RoPE ->
 * attention -> proj + residual [+ statistics] -> [LayerNorm | folded] w1|w2 GEMM -> SwiGLU + LayerNorm -> w3 + residual
 * [+ statistics]), same kernels in the same order, so the taps are bit-identical to the per-operator calls. */
typedef struct {
  const float* ln1_w; const float* ln1_b;
  const void* qkv_w; const float* qkv_b;     /* [3C, C] bf16 = q (pre-scaled) | k | v; bias = q_bias * scale | 0 | v_bias */
  const void* proj_w; const float* proj_b;   /* [C, C] */
  const float* ln2_w; const float* ln2_b;
  const void* w12; const float* b12;         /* [2 * hidden_pad, C] = w1 | w2, rows beyond `hidden` zero */
  const float* ffn_ln_w; const float* ffn_ln_b;
  const void* w3; const float* b3;           /* [C, hidden_pad], columns beyond `hidden` zero */
  /* LayerNorm folded into the Linear behind it (see VfmBlockParams); all NULL = not folded */
  const void* qkv_wf; const float* qkv_bf; const float* qkv_cs;   /* norm1 -> q|k|v */
  const void* w12_wf; const float* w12_bf; const float* w12_cs;   /* norm2 -> w1|w2 */
} VfmEvaBlockParams;

typedef struct {
  int embed_dim, depth, heads, hidden, hidden_pad, n_taps;
  int grid;                     /* windows are grid x grid patches (pos_embed / RoPE tables are fixed to it, eva_02.py:690-697) */
  int tap_blocks[8];            /* out_indices, ascending */
  float ln_eps;
  const void* patch_w;          /* [embed_dim, 768] bf16 */
  const float* patch_b;
  const float* cls_token;       /* [embed_dim] */
  const float* pos_embed;       /* [1 + grid*grid, embed_dim] */
  const float* rope_cos;        /* [grid*grid, 64] */
  const float* rope_sin;
  const float* ones;            /* [embed_dim] of 1.0f (no LayerScale in EVA02) */
  const VfmEvaBlockParams* blocks; /* host array [depth] */
} VfmEvaParams;

size_t vfm_eva_workspace_bytes(const VfmEvaParams* p /*host*/, int n_crops);
/* taps[(crop*grid*grid + p), t*embed_dim + c] (bf16): the un-normalised residual stream after the blocks in tap_blocks. */
int vfm_eva_forward(const VfmEvaParams* p /*host*/, const void* img, int is_u8, const VfmPixelNorm* nrm /*host*/, int img_h,
                    int img_w, const int* crops, int n_crops, void* taps, void* workspace, size_t workspace_bytes,
                    void* stream);


/* SAMViT.forward over a batch of crop windows as ONE call (rein/models/backbones/sam_vit.py:123-147, blocks :201-217, attention
 * :272-287 with add_decomposed_rel_pos :391-428, window_partition / window_unpartition :292-346): the launch sequence
 * vfmseg_b200/sam_engine.py used to issue from Python, same kernels in the same order (bit-identical taps). */
typedef struct {
  int window;                                 /* 0: global attention over the grid; else the window size of this block */
  int qkv_n;                                  /* rows of qkv_w: 3C (+ heads * 2 * (2 size - 1) table-term rows, padded to 32) */
  const float* ln1_w; const float* ln1_b;
  const void* qkv_w; const float* qkv_b;      /* [qkv_n, C] bf16: q | k | v | G_h | G_w (see vfm_attention_relpos_ex) */
  const void* proj_w; const float* proj_b;
  const float* ln2_w; const float* ln2_b;
  const void* lin1_w; const float* lin1_b;    /* [hidden, C] */
  const void* lin2_w; const float* lin2_b;    /* [C, hidden] */
  const void* lin1_wf; const float* lin1_bf; const float* lin1_cs;   /* norm2 folded into lin1, or NULL */
} VfmSamBlockParams;

typedef struct {
  int embed_dim, depth, heads, head_dim, hidden, n_taps, grid, use_rel_pos;
  int tap_blocks[8];              /* out_indices, ascending */
  float ln_eps;
  const void* patch_w;            /* [embed_dim, 768] bf16 */
  const float* patch_b;
  const float* pos_embed;         /* [grid*grid, embed_dim] */
  const float* ones;              /* [embed_dim] of 1.0f (no LayerScale) */
  /* window maps for the n_crops of this call (device int32): part[window-order row] = token row or -1 (zero padding),
   * unpart[token row] = window-order row; win_rows = number of window-order rows; win_buf = bf16 [win_rows, embed_dim], zero-filled
   * ONCE by the caller (the padding rows are never written). All unused when no block is windowed. */
  const int* part; const int* unpart; int win_rows; void* win_buf;
  const void* onehot; int onehot_rows;   /* vfm_attention_global_tc's key matrix for the grid, or NULL */
  const VfmSamBlockParams* blocks;       /* host array [depth] */
} VfmSamParams;

size_t vfm_sam_workspace_bytes(const VfmSamParams* p /*host*/, int n_crops);
/* taps[(crop*grid*grid + p), t*embed_dim + c] (bf16): the raw block outputs at tap_blocks. */
int vfm_sam_forward(const VfmSamParams* p /*host*/, const void* img, int is_u8, const VfmPixelNorm* nrm /*host*/, int img_h,
                    int img_w, const int* crops, int n_crops, void* taps, void* workspace, size_t workspace_bytes,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VFMSEG_B200_H_ */
