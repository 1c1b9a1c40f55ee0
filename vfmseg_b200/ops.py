"""Tensor-level wrappers over the C ABI: each takes torch CUDA tensors (device memory + the current
stream are all torch contributes) and calls one entry point of libvfmseg_b200.so.

These are the per-operator building blocks the parity tests exercise; the product path uses the
fused drivers in `engine.py`.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _C


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "vfmseg_b200 ops need contiguous CUDA tensors"
    return t.data_ptr()


def _bf16(t):
    assert t.dtype == torch.bfloat16
    return _ptr(t)


def _f32(t):
    if t is None:
        return None
    assert t.dtype == torch.float32
    return _ptr(t)


def gemm_f32(a, w, bias=None):
    """fp32 out = a @ w.T (+ bias); a [M,K] bf16, w [N,K] bf16."""
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    _C.call("vfm_gemm_f32", _bf16(a), K, _bf16(w), K, _f32(bias), _f32(out), N, M, N, K, _stream())
    return out


def gemm_bias_bf16(a, w, bias=None, out=None):
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    _C.call("vfm_gemm_bias_bf16", _bf16(a), a.stride(0), _bf16(w), K, _f32(bias), _bf16(out), out.stride(0),
            M, N, K, _stream())
    return out


def gemm_bias_gelu_bf16(a, w, bias):
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    _C.call("vfm_gemm_bias_gelu_bf16", _bf16(a), K, _bf16(w), K, _f32(bias), _bf16(out), N, M, N, K, _stream())
    return out


def gemm_bias_ls_residual_(x, a, w, bias, gamma, tap=None, tap_col0=0, tokens_per_crop=1):
    """x (fp32 [M,N]) += gamma * (a @ w.T + bias) in place; optional bf16 tap snapshot."""
    M, K = a.shape
    N = w.shape[0]
    _C.call("vfm_gemm_bias_ls_residual", _bf16(a), K, _bf16(w), K, _f32(bias), _f32(gamma), _f32(x), N,
            _bf16(tap) if tap is not None else None, tap.shape[1] if tap is not None else 0, tap_col0,
            tokens_per_crop, M, N, K, _stream())
    return x


def gemm_bias_ls_residual_stats_(x, a, w, bias, gamma):
    """x (fp32 [M,N]) += gamma * (a @ w.T + bias) in place (same bits as gemm_bias_ls_residual_); also returns
    xb = bf16(x) and stats [M, N/128, 2] = (sum, sum of squares) of every row over each 128-column slot."""
    M, K = a.shape
    N = w.shape[0]
    xb = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    stats = torch.empty(M, N // 128, 2, device=a.device, dtype=torch.float32)
    _C.call("vfm_gemm_bias_ls_residual_stats", _bf16(a), K, _bf16(w), K, _f32(bias), _f32(gamma), _f32(x), N,
            _bf16(xb), N, _f32(stats), M, N, K, _stream())
    return xb, stats


def fold_layernorm(w, b, ln_w, ln_b):
    """Linear(LayerNorm(x)) with the norm folded into the Linear (see vfm_gemm_lnfold_bf16): returns
    (wf bf16 [N,K] = ln_w * w, bias_f fp32 [N] = b + w @ ln_b, colsum fp32 [N] = row sums of the ROUNDED wf)."""
    w32, g, beta = w.float(), ln_w.float(), ln_b.float()
    wf = (w32 * g[None, :]).to(torch.bfloat16)
    bias_f = (b.float() if b is not None else torch.zeros(w32.shape[0], device=w32.device)) + w32.double().matmul(beta.double()).float()
    colsum = wf.double().sum(dim=1).float()
    return wf.contiguous(), bias_f.contiguous(), colsum.contiguous()


def gemm_lnfold_bf16(xb, stats, wf, bias_f, colsum, eps, gelu=False):
    """bf16(act(Linear(LayerNorm(x)))) from xb = bf16(x), the row statistics and the folded weights."""
    M, K = xb.shape
    N = wf.shape[0]
    out = torch.empty(M, N, device=xb.device, dtype=torch.bfloat16)
    _C.call("vfm_gemm_lnfold_bf16", _bf16(xb), K, _bf16(wf), K, _f32(bias_f), _f32(colsum), _f32(stats), float(eps),
            int(bool(gelu)), _bf16(out), N, M, N, K, _stream())
    return out


def gemm_lnfold_rope_bf16(xb, stats, wf, bias_f, colsum, eps, cos_t, sin_t, rope_cols, tokens_per_seq):
    """gemm_lnfold_bf16 + the RoPE epilogue of gemm_bias_rope_bf16 (EVA02 norm1 -> qkv -> rope)."""
    M, K = xb.shape
    N = wf.shape[0]
    out = torch.empty(M, N, device=xb.device, dtype=torch.bfloat16)
    _C.call("vfm_gemm_lnfold_rope_bf16", _bf16(xb), K, _bf16(wf), K, _f32(bias_f), _f32(colsum), _f32(stats), float(eps),
            _bf16(out), N, M, N, K, _f32(cos_t), _f32(sin_t), rope_cols, tokens_per_seq, _stream())
    return out


def gemm_patch_embed(a, w, bias, pos, n_crops, patches):
    M, K = a.shape
    N = w.shape[0]
    x = torch.empty(n_crops * (patches + 1), N, device=a.device, dtype=torch.float32)
    _C.call("vfm_gemm_patch_embed", _bf16(a), K, _bf16(w), K, _f32(bias), _f32(pos), _f32(x), patches, M, N, K,
            _stream())
    return x


def gemm_convt2x2_gelu(a, w, bias, c_out, h, w_):
    M, K = a.shape
    out = torch.empty(4 * M, c_out, device=a.device, dtype=torch.bfloat16)
    _C.call("vfm_gemm_convt2x2_gelu", _bf16(a), K, _bf16(w), K, _f32(bias), _bf16(out), c_out, h, w_, M, K, _stream())
    return out


def gemm_cls_nchw(a, w32, bias, num_classes, pix_per_crop):
    M, K = a.shape
    assert w32.shape[0] == 32
    out = torch.empty(M // pix_per_crop, num_classes, pix_per_crop, device=a.device, dtype=torch.float32)
    _C.call("vfm_gemm_cls_nchw", _bf16(a), K, _bf16(w32), K, _f32(bias), _f32(out), num_classes, pix_per_crop, M, K,
            _stream())
    return out


def attention_fwd(qkv, n_seq, seq_len, heads, mode=0):
    """mode 0 = automatic tiling, 1 = tensor tiles over every token, 2 = token 0 split off (cls token)."""
    out = torch.empty(n_seq * seq_len, heads * 64, device=qkv.device, dtype=torch.bfloat16)
    _C.call("vfm_attention_fwd_ex", _bf16(qkv), _bf16(out), n_seq, seq_len, heads, int(mode), _stream())
    return out


def attention_cross(q, kv, n_seq, q_len, kv_len, heads):
    """q [n_seq*q_len, heads*64], kv [n_seq*kv_len, 2*heads*64] (k | v) -> [n_seq*q_len, heads*64]."""
    out = torch.empty(n_seq * q_len, heads * 64, device=q.device, dtype=torch.bfloat16)
    _C.call("vfm_attention_cross", _bf16(q), q.stride(0), _bf16(kv), kv.stride(0), _bf16(out), out.stride(0), n_seq, q_len,
            kv_len, heads, _stream())
    return out


def pixel_norm(mean, std, flip) -> _C.VfmPixelNorm:
    n = _C.VfmPixelNorm()
    for i in range(3):
        n.mean[i] = float(mean[i])
        n.inv_std[i] = 1.0 / float(std[i])
    n.flip = int(bool(flip))
    return n


def patch_gather(img, crops, gh, gw, norm: _C.VfmPixelNorm | None = None):
    """img: [B,3,H,W] fp32 (normalised) or uint8; crops: int32 [n,4] = (image, y1, x1, 0)."""
    is_u8 = img.dtype == torch.uint8
    assert is_u8 or img.dtype == torch.float32
    assert crops.dtype == torch.int32 and crops.shape[1] == 4
    n = crops.shape[0]
    out = torch.empty(n * gh * gw, 768, device=img.device, dtype=torch.bfloat16)
    _C.call("vfm_patch_gather", _ptr(img), int(is_u8), C.byref(norm) if norm is not None else None,
            img.shape[2], img.shape[3], _ptr(crops), n, gh, gw, _bf16(out), _stream())
    return out


def cls_rows_(x, cls_token, pos, n_crops, tokens):
    _C.call("vfm_cls_rows", _f32(x), _f32(cls_token), _f32(pos), n_crops, tokens, x.shape[1], _stream())
    return x


def layernorm(x, gamma, beta, eps):
    M, Cc = x.shape
    out = torch.empty(M, Cc, device=x.device, dtype=torch.bfloat16)
    _C.call("vfm_layernorm", _f32(x), _f32(gamma), _f32(beta), _bf16(out), M, Cc, float(eps), _stream())
    return out


def groupnorm_relu(x, gamma, beta, n_crops, groups, eps, relu=True):
    R, Cc = x.shape
    out = torch.empty_like(x)
    _C.call("vfm_groupnorm_relu", _bf16(x), _bf16(out), _f32(gamma), _f32(beta), n_crops, R // n_crops, Cc, groups,
            float(eps), int(relu), _stream())
    return out


def slide_merge_argmax(lowres, boxes, n_img, crop_hw, out_hw, want_logits=False):
    """lowres: fp32 [n_img*n_crops, nc, lh, lw]; boxes: int32 [n_crops, 2] = (y1, x1)."""
    n_crops = boxes.shape[0]
    _, nc, lh, lw = lowres.shape
    H, W = out_hw
    labels = torch.empty(n_img, H, W, device=lowres.device, dtype=torch.uint8)
    logits = torch.empty(n_img, nc, H, W, device=lowres.device, dtype=torch.float32) if want_logits else None
    _C.call("vfm_slide_merge_argmax", _f32(lowres), _ptr(boxes), n_crops, nc, crop_hw[0], crop_hw[1], lh, lw, H, W,
            n_img, _ptr(labels), _f32(logits), _stream())
    return labels, logits


def slide_merge_flip_argmax(lowres, boxes, n_img, crop_hw, a, want_logits=False):
    """Second pass of the flip test-time augmentation fused into the merge: `lowres` are the window logits of the MIRRORED images,
    `a` = slide(img) (fp32 [B, nc, H, W]). Returns (uint8 labels, logits or None) of (a + flip(slide(flip(img)))) / 2; the logits are
    written over `a`. Same bits as slide_merge_argmax + tta_flip_mean_argmax (hrda_encoder_decoder.py:196-229)."""
    n_crops = boxes.shape[0]
    _, nc, lh, lw = lowres.shape
    B, _, H, W = a.shape
    assert B == n_img and a.dtype == torch.float32 and a.is_contiguous()
    labels = torch.empty(n_img, H, W, device=lowres.device, dtype=torch.uint8)
    _C.call("vfm_slide_merge_flip_argmax", _f32(lowres), _ptr(boxes), n_crops, nc, crop_hw[0], crop_hw[1], lh, lw, H, W, n_img,
            _f32(a), _ptr(labels), _f32(a) if want_logits else None, _stream())
    return labels, (a if want_logits else None)


def flip_merge_supported(lowres, crop_hw, W) -> bool:
    """Shapes vfm_slide_merge_flip_argmax takes (the tiled merge kernel: x4 windows, W % 4 == 0, at most 19 classes)."""
    _, nc, lh, lw = lowres.shape
    return W % 4 == 0 and nc <= 19 and crop_hw[0] == 4 * lh and crop_hw[1] == 4 * lw


def tta_flip_mean_argmax(a, b, want_logits=False):
    """a = slide(img), b = slide(flip(img, x)): fp32 [B, nc, H, W]. Returns (uint8 labels [B,H,W], logits or None) of
    (a + flip(b, x)) / 2; the logits are written over `a` (hrda_encoder_decoder.py:196-229, scales = [1])."""
    assert a.shape == b.shape and a.dtype == b.dtype == torch.float32 and a.is_contiguous() and b.is_contiguous()
    B, nc, H, W = a.shape
    labels = torch.empty(B, H, W, device=a.device, dtype=torch.uint8)
    _C.call("vfm_tta_flip_mean_argmax", _f32(a), _f32(b), B, nc, H, W, _ptr(labels), _f32(a) if want_logits else None, _stream())
    return labels, (a if want_logits else None)


def confusion_matrix_(cm, pred, label, num_classes, ignore_index=255):
    """cm: int64 [(nc+1), nc], accumulated in place."""
    assert cm.dtype == torch.int64 and cm.numel() == (num_classes + 1) * num_classes
    assert pred.dtype == torch.uint8 and label.dtype == torch.uint8 and pred.numel() == label.numel()
    _C.call("vfm_confusion_matrix", _ptr(pred), _ptr(label), pred.numel(), num_classes, ignore_index, _ptr(cm),
            _stream())
    return cm


def resize_bilinear(low, size):
    """Bilinear (align_corners=False) resize of fp32 [B,nc,h,w] logits to `size`, through the merge
    kernel with one window covering the whole output (mmseg predict_by_feat / resize)."""
    B = low.shape[0]
    boxes = torch.zeros(1, 2, dtype=torch.int32, device=low.device)
    _, logits = slide_merge_argmax(low.contiguous(), boxes, B, tuple(size), tuple(size), want_logits=True)
    return logits


def resize_argmax(logits, size):
    """mmseg postprocess_result's `resize(seg_logits, size=ori_shape, bilinear, align_corners=False)` + argmax in one
    kernel: fp32 [B,nc,h,w] -> (uint8 labels [B,H,W], fp32 logits [B,nc,H,W]); any scale, up or down."""
    B = logits.shape[0]
    boxes = torch.zeros(1, 2, dtype=torch.int32, device=logits.device)
    return slide_merge_argmax(logits.contiguous(), boxes, B, tuple(size), tuple(size), want_logits=True)


# ------------------------------------------------------------------ coarse-to-fine path (config 3)
def image_resize_norm(img, size, norm: _C.VfmPixelNorm | None = None):
    """Bilinear (align_corners=False) resize of [B,3,H,W] uint8 (normalised on the fly) or fp32 input -> fp32 [B,3,h,w]."""
    is_u8 = img.dtype == torch.uint8
    assert is_u8 or img.dtype == torch.float32
    B, _, H, W = img.shape
    out = torch.empty(B, 3, size[0], size[1], device=img.device, dtype=torch.float32)
    _C.call("vfm_image_resize_norm", _ptr(img), int(is_u8), C.byref(norm) if norm is not None else None, B, H, W, _f32(out),
            size[0], size[1], _stream())
    return out


def ms_confidence(low0, boxes, crop_hw, out_hw, thr):
    """int32 [n_img, n_crops]: pixels per window whose max softmax of the upsampled coarse logits exceeds thr."""
    n_img, nc, lh, lw = low0.shape
    n_crops = boxes.shape[0]
    counts = torch.empty(n_img, n_crops, device=low0.device, dtype=torch.int32)
    _C.call("vfm_ms_confidence", _f32(low0), _ptr(boxes), n_crops, nc, crop_hw[0], crop_hw[1], lh, lw, out_hw[0], out_hw[1],
            n_img, float(thr), _ptr(counts), _stream())
    return counts


def ms_context_im2col(low0, crops, crop_hw, out_hw, ctx_hw, kpad):
    n_img, nc, lh, lw = low0.shape
    n_ref = crops.shape[0]
    out = torch.empty(n_ref * (ctx_hw[0] // 2) * (ctx_hw[1] // 2), kpad, device=low0.device, dtype=torch.bfloat16)
    _C.call("vfm_ms_context_im2col", _f32(low0), _ptr(crops), n_ref, nc, crop_hw[0], crop_hw[1], lh, lw, out_hw[0], out_hw[1],
            ctx_hw[0], ctx_hw[1], _bf16(out), kpad, _stream())
    return out


def space_to_depth2(x, n, h, w):
    Cc = x.shape[1]
    out = torch.empty(n * (h // 2) * (w // 2), 4 * Cc, device=x.device, dtype=torch.bfloat16)
    _C.call("vfm_space_to_depth2", _bf16(x), _bf16(out), n, h, w, Cc, _stream())
    return out


def groupnorm_act(x, gamma, beta, n, groups, eps, act=0, out_f32=False):
    """act: 0 none, 1 ReLU, 2 GELU(erf). x bf16 [n*P, C]."""
    R, Cc = x.shape
    out = torch.empty(R, Cc, device=x.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    _C.call("vfm_groupnorm_act", _bf16(x), _ptr(out), int(out_f32), _f32(gamma), _f32(beta), n, R // n, Cc, groups, float(eps),
            int(act), _stream())
    return out


def geglu(x):
    M, I2 = x.shape
    out = torch.empty(M, I2 // 2, device=x.device, dtype=torch.bfloat16)
    _C.call("vfm_geglu", _bf16(x), _bf16(out), M, I2 // 2, _stream())
    return out


def cast_f32_bf16(x):
    out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    _C.call("vfm_cast_f32_bf16", _f32(x), _bf16(out), x.numel(), _stream())
    return out


def ms_merge_argmax(low0, refined, ref_index, boxes, crop_hw, out_hw, want_logits=False):
    """low0 fp32 [n_img,nc,lh,lw]; refined fp32 [n_ref,nc,rh,rw] (or None); ref_index int32 [n_img,n_crops]; boxes int32 [n_crops,2]."""
    n_img, nc, lh, lw = low0.shape
    n_crops = boxes.shape[0]
    H, W = out_hw
    rh, rw = (refined.shape[2], refined.shape[3]) if refined is not None else (1, 1)
    labels = torch.empty(n_img, H, W, device=low0.device, dtype=torch.uint8)
    logits = torch.empty(n_img, nc, H, W, device=low0.device, dtype=torch.float32) if want_logits else None
    _C.call("vfm_ms_merge_argmax", _f32(low0), _f32(refined), _ptr(ref_index), _ptr(boxes), n_crops, nc, crop_hw[0], crop_hw[1],
            lh, lw, rh, rw, H, W, n_img, _ptr(labels), _f32(logits), _stream())
    return labels, logits


# ------------------------------------------------------------------ EVA02 (config 4)
def rope_qk_(qkv, heads, tokens_per_seq, cos_t, sin_t):
    """In-place RoPE on the q and k thirds of packed qkv [M, 3C] bf16; token 0 of each sequence is skipped."""
    M, C3 = qkv.shape
    _C.call("vfm_rope_qk", _bf16(qkv), M, C3 // 3, heads, tokens_per_seq, _f32(cos_t), _f32(sin_t), _stream())
    return qkv


def swiglu_layernorm(x12, gamma, beta, H, eps):
    """x12 bf16 [M, 2*Hp] = (w1 x | w2 x) -> LayerNorm over H of silu(x1) * x2 -> bf16 [M, Hp] (padding columns zero)."""
    M, Hp2 = x12.shape
    out = torch.empty(M, Hp2 // 2, device=x12.device, dtype=torch.bfloat16)
    _C.call("vfm_swiglu_layernorm", _bf16(x12), _bf16(out), _f32(gamma), _f32(beta), M, H, Hp2 // 2, float(eps), _stream())
    return out


def layernorm_tap(x, gamma, beta, eps, tap=None, tap_col0=0, tokens_per_crop=1, want_out=True):
    """LayerNorm (fp32 [M,C] -> bf16) that can also snapshot the un-normalised rows, cls rows dropped, into columns
    [tap_col0, tap_col0 + C) of `tap` (bf16 [n*(tokens-1), n_taps*C])."""
    M, Cc = x.shape
    out = torch.empty(M, Cc, device=x.device, dtype=torch.bfloat16) if want_out else None
    _C.call("vfm_layernorm_tap", _f32(x), _f32(gamma), _f32(beta), _bf16(out) if out is not None else None, M, Cc, float(eps),
            _bf16(tap) if tap is not None else None, tap.shape[1] if tap is not None else 0, tap_col0, tokens_per_crop, _stream())
    return out


# ------------------------------------------------------------------ SAM ViT backbone (BASELINE config 5)
def gemm_patch_embed_nocls(a, w, bias, pos, n_crops, patches):
    """PatchEmbed + pos_embed add for a backbone without a cls token (sam_vit.py:126-132): x fp32 [n*patches, N]."""
    M, K = a.shape
    N = w.shape[0]
    x = torch.empty(n_crops * patches, N, device=a.device, dtype=torch.float32)
    _C.call("vfm_gemm_patch_embed_ex", _bf16(a), K, _bf16(w), K, _f32(bias), _f32(pos), _f32(x), patches, 0, M, N, K, _stream())
    return x


def layernorm_tap_nocls(x, gamma, beta, eps, tap=None, tap_col0=0, want_out=True):
    """layernorm_tap for token streams without cls rows: tap row = x row."""
    M, Cc = x.shape
    out = torch.empty(M, Cc, device=x.device, dtype=torch.bfloat16) if want_out else None
    _C.call("vfm_layernorm_tap_ex", _f32(x), _f32(gamma), _f32(beta), _bf16(out) if out is not None else None, M, Cc, float(eps),
            _bf16(tap) if tap is not None else None, tap.shape[1] if tap is not None else 0, tap_col0, 1, 0, _stream())
    return out


def relpos_terms(qkv, Rh, Rw, n_seq, heads, head_dim):
    """rel fp32 [n_seq, heads, q_h*q_w, k_h+k_w]: q . Rh[qh, kh] | q . Rw[qw, kw] on the unscaled q of the packed qkv buffer."""
    q_h, k_h, _ = Rh.shape
    q_w, k_w, _ = Rw.shape
    assert qkv.shape == (n_seq * q_h * q_w, 3 * heads * head_dim)
    rel = torch.empty(n_seq, heads, q_h * q_w, k_h + k_w, device=qkv.device, dtype=torch.float32)
    _C.call("vfm_relpos_terms", _bf16(qkv), _f32(Rh), _f32(Rw), _f32(rel), n_seq, heads, head_dim, q_h, q_w, k_h, k_w, _stream())
    return rel


def rows_gather(src, row_map):
    """dst[i] = src[row_map[i]] (zeros where row_map[i] < 0); src bf16 [*, C], row_map int32 [n]."""
    assert row_map.dtype == torch.int32 and src.shape[1] % 8 == 0
    dst = torch.empty(row_map.numel(), src.shape[1], device=src.device, dtype=torch.bfloat16)
    _C.call("vfm_rows_gather", _bf16(src), _bf16(dst), _ptr(row_map), row_map.numel(), src.shape[1], _stream())
    return dst


def attention_relpos(qkv, rel, n_seq, seq_len, heads, head_dim, k_h, k_w, scale):
    """softmax(scale q k^T + decomposed rel-pos bias) v on the packed qkv buffer; rel from relpos_terms or None."""
    assert qkv.shape == (n_seq * seq_len, 3 * heads * head_dim)
    out = torch.empty(n_seq * seq_len, heads * head_dim, device=qkv.device, dtype=torch.bfloat16)
    _C.call("vfm_attention_relpos", _bf16(qkv), _f32(rel) if rel is not None else None, _bf16(out), n_seq, seq_len, heads,
            head_dim, k_h, k_w, float(scale), _stream())
    return out


def attention_relpos_terms(qkv, n_seq, seq_len, heads, head_dim, k_h, k_w, scale, g_col0):
    """Same, with the bias read from the table terms G_h | G_w that the qkv GEMM wrote behind q | k | v (columns from g_col0
    on: heads x (2 k_h - 1) then heads x (2 k_w - 1)); see vfm_attention_relpos_ex."""
    assert qkv.shape[0] == n_seq * seq_len and qkv.stride(0) == qkv.shape[1]
    out = torch.empty(n_seq * seq_len, heads * head_dim, device=qkv.device, dtype=torch.bfloat16)
    _C.call("vfm_attention_relpos_ex", _bf16(qkv), qkv.shape[1], g_col0, None, _bf16(out), n_seq, seq_len, heads, head_dim,
            k_h, k_w, float(scale), _stream())
    return out


def attention_window_tc(qkv, n_seq, seq_len, heads, head_dim, k_h, k_w, scale, g_col0):
    """tcgen05 window attention (seq_len <= 208, head_dim 80) with the rel-pos bias added by the tensor core; same
    operands as attention_relpos_terms."""
    assert qkv.shape[0] == n_seq * seq_len and qkv.stride(0) == qkv.shape[1]
    out = torch.empty(n_seq * seq_len, heads * head_dim, device=qkv.device, dtype=torch.bfloat16)
    _C.call("vfm_attention_window_tc", _bf16(qkv), qkv.shape[1], g_col0, _bf16(out), n_seq, seq_len, heads, head_dim,
            k_h, k_w, float(scale), _stream())
    return out


def relpos_onehot(k_h, k_w, device):
    """The one-hot key matrix of vfm_attention_global_tc for a k_h x k_w grid: bf16 [k_h*k_w rounded up to 64, 64*NA]."""
    bh, bw = (k_h + 15) // 16 * 16, (k_w + 15) // 16 * 16
    cols = (bh + bw + 63) // 64 * 64
    n = k_h * k_w
    e = torch.zeros((n + 63) // 64 * 64, cols, dtype=torch.bfloat16)
    idx = torch.arange(n)
    e[idx, idx // k_w] = 1
    e[idx, bh + idx % k_w] = 1
    return e.to(device)


def attention_global_tc(qkv, onehot, n_seq, seq_len, heads, head_dim, k_h, k_w, scale, g_col0):
    """tcgen05 attention over a whole token grid with the rel-pos bias added by the tensor core (head_dim 80)."""
    assert qkv.shape[0] == n_seq * seq_len and qkv.stride(0) == qkv.shape[1] and onehot.is_contiguous()
    out = torch.empty(n_seq * seq_len, heads * head_dim, device=qkv.device, dtype=torch.bfloat16)
    _C.call("vfm_attention_global_tc", _bf16(qkv), qkv.shape[1], g_col0, _bf16(onehot), onehot.shape[0], _bf16(out), n_seq, seq_len,
            heads, head_dim, k_h, k_w, float(scale), _stream())
    return out


def layernorm_tap_nocls_into(x, gamma, beta, eps, out, out_map, tap=None, tap_col0=0):
    """layernorm_tap_nocls writing row i of the result to row out_map[i] of `out` (window order; padding rows untouched)."""
    M, Cc = x.shape
    assert out_map.dtype == torch.int32 and out_map.numel() == M and out.shape[1] == Cc
    _C.call("vfm_layernorm_tap_map", _f32(x), _f32(gamma), _f32(beta), _bf16(out), M, Cc, float(eps),
            _bf16(tap) if tap is not None else None, tap.shape[1] if tap is not None else 0, tap_col0, 1, 0, _ptr(out_map), _stream())
    return out


def attention_window_tc_into(qkv, n_seq, seq_len, heads, head_dim, k_h, k_w, scale, g_col0, out_map, n_out_rows):
    """attention_window_tc storing window-order row i at output row out_map[i] (< 0: padding, dropped): token order."""
    assert qkv.shape[0] == n_seq * seq_len and qkv.stride(0) == qkv.shape[1] and out_map.numel() == n_seq * seq_len
    out = torch.empty(n_out_rows, heads * head_dim, device=qkv.device, dtype=torch.bfloat16)
    _C.call("vfm_attention_window_tc_map", _bf16(qkv), qkv.shape[1], g_col0, _bf16(out), _ptr(out_map), n_seq, seq_len, heads,
            head_dim, k_h, k_w, float(scale), _stream())
    return out


def gemm_bias_rope_bf16(a, w, bias, cos_t, sin_t, rope_cols, tokens_per_seq):
    """qkv GEMM with RoPE on columns [0, rope_cols) of the non-cls rows applied in the epilogue (EVA02)."""
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    _C.call("vfm_gemm_bias_rope_bf16", _bf16(a), a.stride(0), _bf16(w), K, _f32(bias), _bf16(out), N, M, N, K,
            _f32(cos_t), _f32(sin_t), rope_cols, tokens_per_seq, _stream())
    return out
