"""ORACLE — TEST INFRASTRUCTURE ONLY. Never imported by the product path (vfmseg_b200/).

Plain-PyTorch fp32 CPU restatement of the reference's slide-inference hot path, written as free
functions over a state dict that uses the reference's own key names. Each function cites the
reference lines it follows (paths relative to /root/reference). Third-party semantics that are
not vendored in the reference (mmseg 1.2.2, mmcv 2.1.0, peft 0.10.0) are restated from their
published behaviour and marked [3P]; where the reference holds a verbatim in-repo copy, that copy
is cited instead.

Pinning: this restatement is validated against the reference's OWN modules (imported unmodified
through oracle/ref_shim.py) by oracle/make_golden.py and tests/test_oracle_vs_reference.py; the
reference ships no tests or golden vectors of its own, and the [3P] pieces cannot
be run here (packages absent), so parity is pinned for the in-repo code and "unpinned" at the
third-party boundary.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- slide grid
def slide_boxes(H: int, W: int, crop: Sequence[int], stride: Sequence[int]) -> List[Tuple[int, int, int, int]]:
    """(y1, y2, x1, x2) per window, row-major. rein/models/segmentors/Ms_VFM_encoder_decoder.py:424-441
    (verbatim in-repo copy of mmseg EncoderDecoder.slide_inference [3P])."""
    h_crop, w_crop = crop
    h_stride, w_stride = stride
    h_grids = max(H - h_crop + h_stride - 1, 0) // h_stride + 1
    w_grids = max(W - w_crop + w_stride - 1, 0) // w_stride + 1
    boxes = []
    for h_idx in range(h_grids):
        for w_idx in range(w_grids):
            y1 = h_idx * h_stride
            x1 = w_idx * w_stride
            y2 = min(y1 + h_crop, H)
            x2 = min(x1 + w_crop, W)
            y1 = max(y2 - h_crop, 0)
            x1 = max(x2 - w_crop, 0)
            boxes.append((y1, y2, x1, x2))
    return boxes


# --------------------------------------------------------------------------- preprocessing
def preprocess(img_u8_bgr: Tensor, mean, std, bgr_to_rgb=True) -> Tensor:
    """mmseg SegDataPreProcessor.forward [3P], configured at configs/_base_/models/lora_dinov2_linear.py:13-21:
    channel flip, .float(), (x - mean) / std. img: uint8 [B,3,H,W]."""
    x = img_u8_bgr
    if bgr_to_rgb:
        x = x[:, [2, 1, 0]]
    x = x.float()
    m = torch.tensor(mean, dtype=torch.float32).view(1, 3, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(1, 3, 1, 1)
    return (x - m) / s


# --------------------------------------------------------------------------- backbone
def interpolate_pos_encoding(pos_embed: Tensor, npatch: int, w: int, h: int, patch: int) -> Tensor:
    """rein/models/backbones/dino_v2.py:184-215. NB the caller unpacks `B, nc, w, h = x.shape`
    (:218), so `w` is the image HEIGHT and `h` the WIDTH; kept as in the reference."""
    N = pos_embed.shape[1] - 1
    if npatch == N and w == h:
        return pos_embed
    pos_embed = pos_embed.float()
    class_pos = pos_embed[:, 0]
    patch_pos = pos_embed[:, 1:]
    dim = pos_embed.shape[-1]
    w0 = w // patch
    h0 = h // patch
    w0, h0 = w0 + 0.1, h0 + 0.1
    s = int(math.sqrt(N))
    patch_pos = F.interpolate(patch_pos.reshape(1, s, s, dim).permute(0, 3, 1, 2),
                              scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)), mode="bicubic")
    assert int(w0) == patch_pos.shape[-2] and int(h0) == patch_pos.shape[-1]
    patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat((class_pos.unsqueeze(0), patch_pos), dim=1)


def _linear(x, sd, name):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _qkv(x, sd, pre, lora_scale):
    """nn.Linear qkv, optionally peft-wrapped: base(x) + lora_B(lora_A(x)) * (alpha / r)
    (peft 0.10.0 lora.Linear.forward [3P]; dropout is identity in eval)."""
    if pre + ".base_layer.weight" in sd:
        y = F.linear(x, sd[pre + ".base_layer.weight"], sd.get(pre + ".base_layer.bias"))
        a = sd[pre + ".lora_A.default.weight"]
        b = sd[pre + ".lora_B.default.weight"]
        return y + F.linear(F.linear(x, a), b) * lora_scale
    return _linear(x, sd, pre)


def attention(x: Tensor, sd, pre: str, num_heads: int, lora_scale: float) -> Tensor:
    """rein/models/backbones/dino_layers/attention.py:56-69 (the branch MemEffAttention.forward takes
    when xformers is absent, :73-77)."""
    B, N, C = x.shape
    qkv = _qkv(x, sd, pre + ".qkv", lora_scale).reshape(B, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4)
    scale = (C // num_heads) ** -0.5
    q, k, v = qkv[0] * scale, qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return _linear(x, sd, pre + ".proj")


def block(x: Tensor, sd, pre: str, num_heads: int, lora_scale: float, eps: float = 1e-6) -> Tensor:
    """rein/models/backbones/dino_layers/block.py:89-114, eval branch (:111-113); LayerScale
    layer_scale.py:27; Mlp mlp.py:34-40 (exact-erf GELU); LayerNorm eps 1e-6 dino_v2.py:104."""
    C = x.shape[-1]
    h = F.layer_norm(x, (C,), sd[pre + ".norm1.weight"], sd[pre + ".norm1.bias"], eps)
    h = attention(h, sd, pre + ".attn", num_heads, lora_scale)
    if pre + ".ls1.gamma" in sd:
        h = h * sd[pre + ".ls1.gamma"]
    x = x + h
    h = F.layer_norm(x, (C,), sd[pre + ".norm2.weight"], sd[pre + ".norm2.bias"], eps)
    h = _linear(F.gelu(_linear(h, sd, pre + ".mlp.fc1")), sd, pre + ".mlp.fc2")
    if pre + ".ls2.gamma" in sd:
        h = h * sd[pre + ".ls2.gamma"]
    return x + h


def dino_forward(x: Tensor, sd: Dict[str, Tensor], *, depth: int, num_heads: int, patch: int = 16,
                 out_indices=(7, 11, 15, 23), lora_scale: float = 1.0, return_tokens: bool = False):
    """DinoVisionTransformer.forward_features, rein/models/backbones/dino_v2.py:252-268, with
    prepare_tokens_with_masks :217-228 and PatchEmbed.forward patch_embed.py:68-81.
    `sd` holds un-prefixed backbone keys (cls_token, pos_embed, patch_embed.proj.*, blocks.N.*)."""
    B, _, h, w = x.shape
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat((sd["cls_token"].expand(B, -1, -1), t), dim=1)
    # dino_v2.py:218 names the spatial dims (w, h) = (H, W)
    t = t + interpolate_pos_encoding(sd["pos_embed"], t.shape[1] - 1, h, w, patch)
    outs, toks = [], []
    for i in range(depth):
        t = block(t, sd, f"blocks.{i}", num_heads, lora_scale)
        if i in out_indices:
            toks.append(t)
            outs.append(t[:, 1:, :].permute(0, 2, 1).reshape(B, -1, h // patch, w // patch).contiguous())
    return (outs, toks) if return_tokens else outs


# --------------------------------------------------------------------------- LinearHead
def linear_head_forward(feats: Sequence[Tensor], sd: Dict[str, Tensor], *, groups: int = 32) -> Tensor:
    """LinearHead.forward, rein/models/heads/linear_head.py:50-70: cat -> fusion_conv (mmcv ConvModule
    [3P]: 1x1 conv without bias -> GroupNorm(eps 1e-5) -> ReLU, :36-40) -> ConvT(k2,s2) ->
    SyncBatchNorm (eval: running stats) -> GELU -> ConvT(k2,s2) -> GELU (:42-48) -> cls_seg =
    conv_seg(dropout(x)) (mmseg BaseDecodeHead [3P]; Dropout2d is identity in eval).
    `sd` holds un-prefixed decode_head keys."""
    x = torch.cat(list(feats), dim=1)
    x = F.conv2d(x, sd["fusion_conv.conv.weight"])
    x = F.relu(F.group_norm(x, groups, sd["fusion_conv.gn.weight"], sd["fusion_conv.gn.bias"], 1e-5))
    x = F.conv_transpose2d(x, sd["output_upscaling.0.weight"], sd["output_upscaling.0.bias"], stride=2)
    x = F.batch_norm(x, sd["output_upscaling.1.running_mean"], sd["output_upscaling.1.running_var"],
                     sd["output_upscaling.1.weight"], sd["output_upscaling.1.bias"], False, 0.0, 1e-5)
    x = F.gelu(x)
    x = F.conv_transpose2d(x, sd["output_upscaling.3.weight"], sd["output_upscaling.3.bias"], stride=2)
    x = F.gelu(x)
    return F.conv2d(x, sd["conv_seg.weight"], sd["conv_seg.bias"])


# --------------------------------------------------------------------------- segmentor
def split_state_dict(sd: Dict[str, Tensor]):
    """Full LoraBackboneEncoderDecoder state dict -> (backbone, head) dicts with the prefixes
    'backbone.base_model.model.' (peft naming [3P], Lora_encoder_decoder.py:24,36) / 'decode_head.' removed."""
    bb, hd = {}, {}
    for k, v in sd.items():
        if k.startswith("backbone.base_model.model."):
            bb[k[len("backbone.base_model.model."):]] = v
        elif k.startswith("backbone."):
            bb[k[len("backbone."):]] = v
        elif k.startswith("decode_head."):
            hd[k[len("decode_head."):]] = v
    return bb, hd


def encode_decode(x: Tensor, sd, cfg) -> Tensor:
    """mmseg EncoderDecoder.encode_decode [3P] = decode_head.predict(extract_feat(x)): head forward then
    predict_by_feat's bilinear resize (align_corners=False) to the input size."""
    bb, hd = sd
    feats = dino_forward(x, bb, depth=cfg["depth"], num_heads=cfg["num_heads"], patch=cfg.get("patch", 16),
                         out_indices=cfg["out_indices"], lora_scale=cfg.get("lora_scale", 1.0))
    low = linear_head_forward(feats, hd, groups=cfg.get("groups", 32))
    return F.interpolate(low, size=x.shape[2:], mode="bilinear", align_corners=False), low


def slide_inference(inputs: Tensor, sd, cfg, *, crop, stride, return_lowres: bool = False):
    """mmseg EncoderDecoder.slide_inference [3P]; in-repo verbatim copy
    rein/models/segmentors/Ms_VFM_encoder_decoder.py:424-461 (grid, F.pad add, count_mat, divide)."""
    B, _, H, W = inputs.shape
    preds = None
    count = inputs.new_zeros((B, 1, H, W))
    lows = []
    for (y1, y2, x1, x2) in slide_boxes(H, W, crop, stride):
        logit, low = encode_decode(inputs[:, :, y1:y2, x1:x2], sd, cfg)
        lows.append(low)
        if preds is None:
            preds = inputs.new_zeros((B, logit.shape[1], H, W))
        preds += F.pad(logit, (int(x1), int(W - x2), int(y1), int(H - y2)))
        count[:, :, y1:y2, x1:x2] += 1
    assert (count == 0).sum() == 0
    out = preds / count
    return (out, lows) if return_lowres else out


def tta_flip_combine(slide_fn, inputs: Tensor, flip: bool = True) -> Tensor:
    """rein/models/segmentors/hrda_encoder_decoder.py:196-229 with `test_time_aug` on and scales = [1] (:198; the resizes
    by scale_factor 1 are identities): res = zeros; res += slide(img); if flip: res += flip(slide(flip(img, [3])), [3]);
    res / (2 if flip else 1). `slide_fn(img) -> [B,C,H,W]` is the plain slide inference."""
    a = slide_fn(inputs)
    res = torch.zeros_like(a)
    res += a
    if flip:
        res += torch.flip(slide_fn(torch.flip(inputs, [3])), [3])
        return res / 2
    return res / 1


def whole_inference(inputs: Tensor, sd, cfg) -> Tensor:
    """mmseg EncoderDecoder.whole_inference [3P] = encode_decode on the full image."""
    return encode_decode(inputs, sd, cfg)[0]


def postprocess(seg_logits: Tensor) -> Tensor:
    """mmseg BaseSegmentor.postprocess_result [3P] with ori_shape == img_shape and no padding:
    argmax over classes (first maximum wins), int64 [B,1,H,W]."""
    return seg_logits.argmax(dim=1, keepdim=True)


# --------------------------------------------------------------------------- coarse-to-fine (config 3)
def split_ms_state_dict(sd: Dict[str, Tensor]):
    """MsVFMEncoderDecoder state dict -> (backbone, decode_head, aux_decoder) dicts; prefixes
    'backbone.model.base_model.model.' (LoRABackbone, lora_backbone.py:23), 'decode_head.', 'aux_decoder.'
    (Ms_VFM_encoder_decoder.py:112)."""
    bb, hd, aux = {}, {}, {}
    for k, v in sd.items():
        for pre, dst in (("backbone.model.base_model.model.", bb), ("decode_head.", hd), ("aux_decoder.", aux)):
            if k.startswith(pre):
                dst[k[len(pre):]] = v
    return bb, hd, aux


def _cross_attention(x: Tensor, context: Tensor, sd, pre: str, heads: int) -> Tensor:
    """CrossAttention._forward, rein/models/heads/Transformer.py:113-136 (mask=None; MemEffAttention.forward :140-142
    falls back to it when xformers is absent)."""
    q = F.linear(x, sd[pre + ".to_q.weight"])
    k = F.linear(context, sd[pre + ".to_k.weight"])
    v = F.linear(context, sd[pre + ".to_v.weight"])
    B, N, inner = q.shape
    d = inner // heads
    q, k, v = (t.reshape(B, t.shape[1], heads, d).permute(0, 2, 1, 3) for t in (q, k, v))
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * d ** -0.5
    out = torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), v)
    out = out.permute(0, 2, 1, 3).reshape(B, N, inner)
    return F.linear(out, sd[pre + ".to_out.0.weight"], sd[pre + ".to_out.0.bias"])


def transformer_decoder_forward(query: Tensor, img_feats: Tensor, sd, *, heads: int, depth: int, pre: str = "transformer_decoder") -> Tensor:
    """MaskTransformerDecoder.forward with mask_enable=False, Transformer.py:270-283 (= TransformerDecoder.forward
    :242-252): GroupNorm(32, eps 1e-6) of the query (:91-92), tokens, `depth` BasicTransformerBlocks (:173-177: self
    attention, cross attention on img_feats, GEGLU feed-forward :52-79; nn.LayerNorm eps 1e-5), back to NCHW."""
    B, C, h, w = img_feats.shape
    x = F.group_norm(query, 32, sd[pre + ".norm.weight"], sd[pre + ".norm.bias"], 1e-6)
    x = x.flatten(2).transpose(1, 2)
    ctx = img_feats.flatten(2).transpose(1, 2)
    for i in range(depth):
        b = f"{pre}.transformer_blocks.{i}"
        ln = lambda t, n: F.layer_norm(t, (C,), sd[f"{b}.{n}.weight"], sd[f"{b}.{n}.bias"], 1e-5)
        x = _cross_attention(ln(x, "norm1"), ln(x, "norm1"), sd, b + ".attn1", heads) + x
        x = _cross_attention(ln(x, "norm2"), ctx, sd, b + ".attn2", heads) + x
        u = F.linear(ln(x, "norm3"), sd[b + ".ff.net.0.proj.weight"], sd[b + ".ff.net.0.proj.bias"])
        a, gate = u.chunk(2, dim=-1)
        x = F.linear(a * F.gelu(gate), sd[b + ".ff.net.2.weight"], sd[b + ".ff.net.2.bias"]) + x
    return x.transpose(1, 2).reshape(B, C, h, w)


def vfm_head_forward(feats: Sequence[Tensor], seg_logits: Tensor, sd, *, heads: int, depth: int) -> Tensor:
    """VFMHead.forward, rein/models/heads/VFMHead.py:61-89: context resized to 4x the feature grid (:63-67), fuse_conv
    (1x1 conv + GroupNorm(32) + GELU, :28-33), seg_logits_embed (two k2s2 convs + GN + GELU, 1x1 conv + GN, :38-49),
    decoder(query=img_feats, img_feats=seg_logits_embed) (:82), cls_seg (:87). `sd` holds un-prefixed aux_decoder keys."""
    h, w = feats[0].shape[2:]
    ctx = F.interpolate(seg_logits, size=(4 * h, 4 * w), mode="bilinear", align_corners=False)
    f = F.conv2d(torch.cat(list(feats), dim=1), sd["fuse_conv.0.weight"], sd["fuse_conv.0.bias"])
    f = F.gelu(F.group_norm(f, 32, sd["fuse_conv.1.weight"], sd["fuse_conv.1.bias"], 1e-5))
    e = F.conv2d(ctx, sd["seg_logits_embed.0.weight"], sd["seg_logits_embed.0.bias"], stride=2)
    e = F.gelu(F.group_norm(e, 32, sd["seg_logits_embed.1.weight"], sd["seg_logits_embed.1.bias"], 1e-5))
    e = F.conv2d(e, sd["seg_logits_embed.3.weight"], sd["seg_logits_embed.3.bias"], stride=2)
    e = F.gelu(F.group_norm(e, 32, sd["seg_logits_embed.4.weight"], sd["seg_logits_embed.4.bias"], 1e-5))
    e = F.conv2d(e, sd["seg_logits_embed.6.weight"], sd["seg_logits_embed.6.bias"])
    e = F.group_norm(e, 32, sd["seg_logits_embed.7.weight"], sd["seg_logits_embed.7.bias"], 1e-5)
    out = transformer_decoder_forward(f, e, sd, heads=heads, depth=depth)
    return F.conv2d(out, sd["conv_seg.weight"], sd["conv_seg.bias"])


def ms_inference(inputs: Tensor, sd3, cfg, *, crop, stride, threshold: float, conf: float, lr_size=(512, 1024),
                 return_info: bool = False):
    """MsVFMEncoderDecoder.ms_inference, rein/models/segmentors/Ms_VFM_encoder_decoder.py:400-466.
    stage 0 (:413,420): inputs resized to `lr_size` (hard-coded (512, 1024) there), whole_inference, whose
    predict_by_feat [3P] resizes the head output to metas['img_shape'] = the ORIGINAL image size; the :417 resize at
    stage 1 is then the identity. stage 1 (:424-461): per window, context = crop of those logits; refine with the aux
    decoder when mean(max softmax(context) > threshold) < conf (batch mean, :446-449), else reuse the context."""
    bb, hd, aux = sd3
    B, _, H, W = inputs.shape
    vit = dict(depth=cfg["depth"], num_heads=cfg["num_heads"], patch=cfg.get("patch", 16), out_indices=cfg["out_indices"],
               lora_scale=cfg.get("lora_scale", 1.0))
    lr = F.interpolate(inputs, size=tuple(lr_size), mode="bilinear", align_corners=False)
    low0 = linear_head_forward(dino_forward(lr, bb, **vit), hd, groups=cfg.get("groups", 32))
    seg = F.interpolate(low0, size=(H, W), mode="bilinear", align_corners=False)
    preds = inputs.new_zeros((B, seg.shape[1], H, W))
    count = inputs.new_zeros((B, 1, H, W))
    fracs, refined = [], []
    for (y1, y2, x1, x2) in slide_boxes(H, W, crop, stride):
        context = seg[:, :, y1:y2, x1:x2]
        confidence = (torch.softmax(context, dim=1).max(dim=1)[0] > threshold).float().mean().item()
        fracs.append(confidence)
        if confidence < conf:
            feats = dino_forward(inputs[:, :, y1:y2, x1:x2], bb, **vit)
            logit = vfm_head_forward(feats, context, aux, heads=cfg["aux_heads"], depth=cfg["aux_depth"])
            refined.append(True)
        else:
            logit = context
            refined.append(False)
        logit = F.interpolate(logit, size=(y2 - y1, x2 - x1), mode="bilinear", align_corners=False)
        preds += F.pad(logit, (int(x1), int(W - x2), int(y1), int(H - y2)))
        count[:, :, y1:y2, x1:x2] += 1
    assert (count == 0).sum() == 0
    out = preds / count
    return (out, dict(low0=low0, fracs=fracs, refined=refined)) if return_info else out


# --------------------------------------------------------------------------- EVA02 backbone (config 4)
def eva_rope_tables(head_dim: int, pt_seq_len: int, ft_seq_len: int):
    """VisionRotaryEmbeddingFast.__init__, rein/models/backbones/eva_02.py:119-156: dim = head_dim // 2 per axis,
    freqs_for='lang' (theta 10000), positions t = arange(ft) / ft * pt, each frequency repeated twice, (h | w) halves."""
    dim = head_dim // 2
    freqs = 1.0 / (10000 ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    t = torch.arange(ft_seq_len) / ft_seq_len * pt_seq_len
    f = torch.einsum("i,f->if", t, freqs).repeat_interleave(2, dim=-1)
    f2 = torch.cat((f[:, None, :].expand(ft_seq_len, ft_seq_len, dim), f[None, :, :].expand(ft_seq_len, ft_seq_len, dim)), dim=-1)
    return f2.cos().reshape(-1, 2 * dim), f2.sin().reshape(-1, 2 * dim)


def _rotate_half(x: Tensor) -> Tensor:
    """eva_02.py:54-58: (x0, x1) -> (-x1, x0) on interleaved pairs."""
    x = x.reshape(*x.shape[:-1], -1, 2)
    return torch.stack((-x[..., 1], x[..., 0]), dim=-1).reshape(*x.shape[:-2], -1)


def eva_forward(x: Tensor, sd: Dict[str, Tensor], *, depth: int, num_heads: int, patch: int = 16, out_indices=(7, 11, 15, 23),
                lora_scale: float = 1.0, pt_hw_seq_len: int = 16, ln_eps: float = 1e-5):
    """EVA2.forward_features, rein/models/backbones/eva_02.py:816-849, with Block.forward :486-493 (init_values=None),
    Attention.forward :331-381 (subln, rope, xattn) and SwiGLU.forward :234-241. `sd` holds un-prefixed backbone keys in
    peft layout. peft [3P]: `q_proj.weight` of a wrapped layer is the BASE weight, and the reference calls F.linear on it
    (:337-339), so only attn.proj — called as a module — carries its LoRA update. LayerNorm eps is nn.LayerNorm's 1e-5:
    EVA2 does not forward its configured norm_layer to Block (:705-727)."""
    B, _, H, W = x.shape
    C = sd["cls_token"].shape[-1]
    d = C // num_heads
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch).flatten(2).transpose(1, 2)
    t = torch.cat((sd["cls_token"].expand(B, -1, -1), t), dim=1) + sd["pos_embed"]            # no interpolation (:825-826)
    cos, sin = eva_rope_tables(d, pt_hw_seq_len, H // patch)
    outs = []
    for i in range(depth):
        p = f"blocks.{i}."
        h = F.layer_norm(t, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], ln_eps)
        q = F.linear(h, sd[p + "attn.q_proj.base_layer.weight"], sd[p + "attn.q_bias"])
        k = F.linear(h, sd[p + "attn.k_proj.base_layer.weight"])
        v = F.linear(h, sd[p + "attn.v_proj.base_layer.weight"], sd[p + "attn.v_bias"])
        q, k, v = (u.reshape(B, -1, num_heads, d).permute(0, 2, 1, 3) for u in (q, k, v))
        q = torch.cat((q[:, :, :1], q[:, :, 1:] * cos + _rotate_half(q[:, :, 1:]) * sin), dim=2)
        k = torch.cat((k[:, :, :1], k[:, :, 1:] * cos + _rotate_half(k[:, :, 1:]) * sin), dim=2)
        a = ((q @ k.transpose(-1, -2)) * d ** -0.5).softmax(dim=-1) @ v                       # xformers [3P] default scale
        a = a.permute(0, 2, 1, 3).reshape(B, -1, C)
        wp = sd[p + "attn.proj.base_layer.weight"] + lora_scale * (sd[p + "attn.proj.lora_B.default.weight"] @ sd[p + "attn.proj.lora_A.default.weight"])
        t = t + F.linear(a, wp, sd[p + "attn.proj.base_layer.bias"])
        h = F.layer_norm(t, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], ln_eps)
        hid = F.silu(F.linear(h, sd[p + "mlp.w1.weight"], sd[p + "mlp.w1.bias"])) * F.linear(h, sd[p + "mlp.w2.weight"], sd[p + "mlp.w2.bias"])
        hid = F.layer_norm(hid, (hid.shape[-1],), sd[p + "mlp.ffn_ln.weight"], sd[p + "mlp.ffn_ln.bias"], ln_eps)
        t = t + F.linear(hid, sd[p + "mlp.w3.weight"], sd[p + "mlp.w3.bias"])
        if i in out_indices:
            outs.append(t[:, 1:, :].permute(0, 2, 1).reshape(B, -1, H // patch, W // patch).contiguous())
    return outs


def eva_slide_inference(inputs: Tensor, bb, hd, cfg, *, crop, stride) -> Tensor:
    """slide_inference (same loop as above) with the EVA02 backbone + LinearHead."""
    B, _, H, W = inputs.shape
    preds, count = None, inputs.new_zeros((B, 1, H, W))
    for (y1, y2, x1, x2) in slide_boxes(H, W, crop, stride):
        feats = eva_forward(inputs[:, :, y1:y2, x1:x2], bb, depth=cfg["depth"], num_heads=cfg["num_heads"],
                            out_indices=cfg["out_indices"], lora_scale=cfg["lora_scale"])
        logit = F.interpolate(linear_head_forward(feats, hd, groups=cfg.get("groups", 32)), size=(y2 - y1, x2 - x1), mode="bilinear",
                              align_corners=False)
        if preds is None:
            preds = inputs.new_zeros((B, logit.shape[1], H, W))
        preds += F.pad(logit, (int(x1), int(W - x2), int(y1), int(H - y2)))
        count[:, :, y1:y2, x1:x2] += 1
    return preds / count


# --------------------------------------------------------------------------- SAM ViT backbone (BASELINE config 5)
def sam_rel_pos_table(q_size: int, k_size: int, rel_pos: Tensor) -> Tensor:
    """get_rel_pos, rein/models/backbones/sam_vit.py:358-388, for self-attention on one grid (q_size == k_size = n, the only
    way Attention.forward calls it, :278-280): the stored (L, C) table is brought to 2n - 1 entries by linear interpolation
    when L differs (global blocks are built with 4n - 1 entries, :248-254) and entry [q, k] is row q - k + n - 1.
    Returns [n, n, C]."""
    assert q_size == k_size, "the reference only ever passes equal sizes"
    n = q_size
    table = rel_pos
    if table.shape[0] != 2 * n - 1:
        table = F.interpolate(table.t()[None], size=2 * n - 1, mode="linear")[0].t()
    offset = torch.arange(n)[:, None] - torch.arange(n)[None, :] + (n - 1)
    return table[offset]


def sam_attention(x: Tensor, sd, pre: str, num_heads: int, lora_scale: float) -> Tensor:
    """Attention.forward, sam_vit.py:263-289 on [B, H, W, C] with add_decomposed_rel_pos :391-428: the bias uses the
    UNSCALED q (rel_h = q . Rh[qh, kh], rel_w = q . Rw[qw, kw]), added to (q * scale) k^T. qkv carries LoRA (peft [3P])."""
    B, H, W, C = x.shape
    d = C // num_heads
    w = sd[pre + "qkv.base_layer.weight"] + lora_scale * (sd[pre + "qkv.lora_B.default.weight"] @ sd[pre + "qkv.lora_A.default.weight"])
    qkv = F.linear(x, w, sd[pre + "qkv.base_layer.bias"]).reshape(B, H * W, 3, num_heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.reshape(3, B * num_heads, H * W, d).unbind(0)
    attn = (q * d ** -0.5) @ k.transpose(-2, -1)
    Rh = sam_rel_pos_table(H, H, sd[pre + "rel_pos_h"])
    Rw = sam_rel_pos_table(W, W, sd[pre + "rel_pos_w"])
    rq = q.reshape(B * num_heads, H, W, d)
    rel_h = torch.einsum("bhwc,hkc->bhwk", rq, Rh)
    rel_w = torch.einsum("bhwc,wkc->bhwk", rq, Rw)
    attn = (attn.view(-1, H, W, H, W) + rel_h[:, :, :, :, None] + rel_w[:, :, :, None, :]).view(-1, H * W, H * W)
    o = (attn.softmax(dim=-1) @ v).view(B, num_heads, H, W, d).permute(0, 2, 3, 1, 4).reshape(B, H, W, C)
    return F.linear(o, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def sam_forward(x: Tensor, sd: Dict[str, Tensor], *, depth: int, num_heads: int, window_size: int, global_attn_indexes,
                out_indices, patch: int = 16, lora_scale: float = 1.0, ln_eps: float = 1e-6):
    """SAMViT.forward, sam_vit.py:123-147 with Block.forward :201-217 (window_partition / window_unpartition :292-346:
    zero padding AFTER norm1 to a multiple of the window, so padded tokens enter attention as q = k = v = bias) and
    MLPBlock :17-29 (exact-erf GELU). pos_embed is added without interpolation (:131-132): the input must have the
    size the model was built for. Returns the raw block outputs at out_indices as [B, C, h, w]."""
    B = x.shape[0]
    C = sd["pos_embed"].shape[-1]
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch).permute(0, 2, 3, 1)
    t = t + sd["pos_embed"]
    H, W = t.shape[1], t.shape[2]
    outs = []
    for i in range(depth):
        p = f"blocks.{i}."
        h = F.layer_norm(t, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], ln_eps)
        ws = 0 if i in global_attn_indexes else window_size
        if ws > 0:
            ph, pw = (ws - H % ws) % ws, (ws - W % ws) % ws
            h = F.pad(h, (0, 0, 0, pw, 0, ph))
            Hp, Wp = H + ph, W + pw
            h = h.view(B, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, C)
        h = sam_attention(h, sd, p + "attn.", num_heads, lora_scale)
        if ws > 0:
            h = h.view(B, Hp // ws, Wp // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)[:, :H, :W, :]
        t = t + h
        h = F.layer_norm(t, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], ln_eps)
        t = t + F.linear(F.gelu(F.linear(h, sd[p + "mlp.lin1.weight"], sd[p + "mlp.lin1.bias"])), sd[p + "mlp.lin2.weight"], sd[p + "mlp.lin2.bias"])
        if i in out_indices:
            outs.append(t.permute(0, 3, 1, 2).contiguous())
    return outs


def sam_slide_inference(inputs: Tensor, bb, hd, cfg, *, crop, stride) -> Tensor:
    """slide_inference (same loop as above) with the SAM ViT backbone + LinearHead."""
    B, _, H, W = inputs.shape
    preds, count = None, inputs.new_zeros((B, 1, H, W))
    for (y1, y2, x1, x2) in slide_boxes(H, W, crop, stride):
        feats = sam_forward(inputs[:, :, y1:y2, x1:x2], bb, depth=cfg["depth"], num_heads=cfg["num_heads"],
                            window_size=cfg["window_size"], global_attn_indexes=cfg["global_attn_indexes"],
                            out_indices=cfg["out_indices"], lora_scale=cfg["lora_scale"])
        logit = F.interpolate(linear_head_forward(feats, hd, groups=cfg.get("groups", 32)), size=(y2 - y1, x2 - x1), mode="bilinear",
                              align_corners=False)
        if preds is None:
            preds = inputs.new_zeros((B, logit.shape[1], H, W))
        preds += F.pad(logit, (int(x1), int(W - x2), int(y1), int(H - y2)))
        count[:, :, y1:y2, x1:x2] += 1
    return preds / count


# --------------------------------------------------------------------------- metric (integer/np)
def intersect_and_union(pred: Tensor, label: Tensor, num_classes: int, ignore_index: int):
    """mmseg IoUMetric.intersect_and_union [3P] as called from rein/dg_metrics.py:50-52:
    three float32 histc over the non-ignored pixels."""
    mask = label != ignore_index
    pred = pred[mask]
    label = label[mask]
    inter = pred[pred == label]
    area_i = torch.histc(inter.float(), bins=num_classes, min=0, max=num_classes - 1).cpu()
    area_p = torch.histc(pred.float(), bins=num_classes, min=0, max=num_classes - 1).cpu()
    area_l = torch.histc(label.float(), bins=num_classes, min=0, max=num_classes - 1).cpu()
    return area_i, area_p + area_l - area_i, area_p, area_l


def confusion_matrix_np(pred: np.ndarray, label: np.ndarray, num_classes: int, ignore_index: int) -> np.ndarray:
    """Integer restatement: cm[(min(label, nc)), pred] over label != ignore. diag / column sums /
    row sums reproduce intersect_and_union exactly (int64 instead of float32 counts)."""
    pred = pred.reshape(-1).astype(np.int64)
    label = label.reshape(-1).astype(np.int64)
    keep = label != ignore_index
    p, l = pred[keep], np.minimum(label[keep], num_classes)
    ok = p < num_classes
    cm = np.bincount(l[ok] * num_classes + p[ok], minlength=(num_classes + 1) * num_classes)
    return cm.reshape(num_classes + 1, num_classes)


def areas_from_confusion(cm: np.ndarray, num_classes: int):
    inter = np.diag(cm[:num_classes]).astype(np.int64)
    pred = cm.sum(0).astype(np.int64)
    label = cm[:num_classes].sum(1).astype(np.int64)
    return inter, pred + label - inter, pred, label


def total_area_to_metrics(inter, union, pred, label) -> Dict[str, float]:
    """mmseg IoUMetric.compute_metrics / total_area_to_metrics [3P] for metrics=['mIoU'], in the
    reference's float32 tensor arithmetic: aAcc = sum(I)/sum(L); IoU = I/U; Acc = I/L;
    np.round(np.nanmean(v) * 100, 2)."""
    f = lambda a: torch.as_tensor(np.asarray(a)).to(torch.float32)
    inter, union, label = f(inter), f(union), f(label)
    aacc = (inter.sum() / label.sum()).numpy()
    iou = (inter / union).numpy()
    acc = (inter / label).numpy()
    return {"aAcc": float(np.round(np.nanmean(aacc) * 100, 2)), "mIoU": float(np.round(np.nanmean(iou) * 100, 2)),
            "mAcc": float(np.round(np.nanmean(acc) * 100, 2))}
