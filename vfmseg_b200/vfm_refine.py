"""Engine side of the coarse-to-fine path (BASELINE config 3): weight packing and the launch sequence of VFMHead
("MGRNet", rein/models/heads/VFMHead.py:61-89) with its transformer decoder (rein/models/heads/Transformer.py:
113-136 attention, :52-79 GEGLU feed-forward, :158-177 block, :228-283 decoder), and MsVFMEncoderDecoder.ms_inference
(rein/models/segmentors/Ms_VFM_encoder_decoder.py:400-466).

Token-major everywhere: a [n, C, h, w] map is the bf16 matrix [n*h*w, C]; 1x1 convs and k=2,s=2 convs are GEMMs
(tcgen05), attention is the fused kernel, the rest are the memory-bound kernels of csrc/ms_refine.cuh.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _C, ops


@dataclass
class VfmHeadSpec:
    in_channels: int      # 4 taps x embed_dim
    channels: int         # 256
    num_classes: int
    n_heads: int
    d_head: int
    depth: int
    groups: int = 32


class PackedVfmHead:
    """Device-resident weights of VFMHead in GEMM layout. State-dict keys are the reference's
    (fuse_conv.{0,1}, seg_logits_embed.{0,1,3,4,6,7}, transformer_decoder.*, conv_seg)."""

    def __init__(self, sd: Dict[str, torch.Tensor], spec: VfmHeadSpec, device):
        self.spec = spec
        if spec.d_head != 64:
            raise ValueError("vfmseg_b200 attention kernel needs d_head 64")
        if spec.num_classes > 32:
            raise ValueError("vfmseg_b200 head kernels support num_classes <= 32")
        keep: List[torch.Tensor] = []

        def dev(t, dtype):
            t = t.detach().to(device=device, dtype=dtype).contiguous()
            keep.append(t)
            return t

        C, nc = spec.channels, spec.num_classes
        inner = spec.n_heads * spec.d_head
        scale = spec.d_head ** -0.5            # Transformer.py:100,121 (0.125: exact in bf16)
        f32, bf = torch.float32, torch.bfloat16
        self.fuse_w = dev(sd["fuse_conv.0.weight"].reshape(C, -1), bf)
        self.fuse_b = dev(sd["fuse_conv.0.bias"], f32)
        self.fuse_gn = (dev(sd["fuse_conv.1.weight"], f32), dev(sd["fuse_conv.1.bias"], f32))
        # seg_logits_embed[0]: Conv2d(nc, C/4, 2, 2); operand columns = cin*4 + dy*2 + dx = the flattened weight, padded to 8
        self.kpad = (4 * nc + 7) // 8 * 8
        w = torch.zeros(C // 4, self.kpad)
        w[:, :4 * nc] = sd["seg_logits_embed.0.weight"].float().reshape(C // 4, -1)
        self.emb1_w, self.emb1_b = dev(w, bf), dev(sd["seg_logits_embed.0.bias"], f32)
        self.emb1_gn = (dev(sd["seg_logits_embed.1.weight"], f32), dev(sd["seg_logits_embed.1.bias"], f32))
        # seg_logits_embed[3]: Conv2d(C/4, C/2, 2, 2); space_to_depth2 columns = (dy*2+dx)*Cin + c
        self.emb2_w = dev(sd["seg_logits_embed.3.weight"].float().permute(0, 2, 3, 1).reshape(C // 2, -1), bf)
        self.emb2_b = dev(sd["seg_logits_embed.3.bias"], f32)
        self.emb2_gn = (dev(sd["seg_logits_embed.4.weight"], f32), dev(sd["seg_logits_embed.4.bias"], f32))
        self.emb3_w = dev(sd["seg_logits_embed.6.weight"].reshape(C, -1), bf)
        self.emb3_b = dev(sd["seg_logits_embed.6.bias"], f32)
        self.emb3_gn = (dev(sd["seg_logits_embed.7.weight"], f32), dev(sd["seg_logits_embed.7.bias"], f32))
        t = "transformer_decoder."
        self.dec_norm = (dev(sd[t + "norm.weight"], f32), dev(sd[t + "norm.bias"], f32))
        self.ones = dev(torch.ones(C), f32)
        self.blocks = []
        for i in range(spec.depth):
            p = f"{t}transformer_blocks.{i}."
            blk = dict(
                n1=(dev(sd[p + "norm1.weight"], f32), dev(sd[p + "norm1.bias"], f32)),
                n2=(dev(sd[p + "norm2.weight"], f32), dev(sd[p + "norm2.bias"], f32)),
                n3=(dev(sd[p + "norm3.weight"], f32), dev(sd[p + "norm3.bias"], f32)),
                qkv1=dev(torch.cat([sd[p + "attn1.to_q.weight"].float() * scale, sd[p + "attn1.to_k.weight"].float(),
                                    sd[p + "attn1.to_v.weight"].float()], 0), bf),
                out1_w=dev(sd[p + "attn1.to_out.0.weight"], bf), out1_b=dev(sd[p + "attn1.to_out.0.bias"], f32),
                q2=dev(sd[p + "attn2.to_q.weight"].float() * scale, bf),
                kv2=dev(torch.cat([sd[p + "attn2.to_k.weight"].float(), sd[p + "attn2.to_v.weight"].float()], 0), bf),
                out2_w=dev(sd[p + "attn2.to_out.0.weight"], bf), out2_b=dev(sd[p + "attn2.to_out.0.bias"], f32),
                ff_in_w=dev(sd[p + "ff.net.0.proj.weight"], bf), ff_in_b=dev(sd[p + "ff.net.0.proj.bias"], f32),
                ff_out_w=dev(sd[p + "ff.net.2.weight"], bf), ff_out_b=dev(sd[p + "ff.net.2.bias"], f32),
            )
            assert blk["qkv1"].shape == (3 * inner, C)
            self.blocks.append(blk)
        wc = torch.zeros(32, C)
        wc[:nc] = sd["conv_seg.weight"].float().reshape(nc, -1)
        self.cls_w, self.cls_b = dev(wc, bf), dev(sd["conv_seg.bias"], f32)
        self._keep = keep


def context_tokens(head: PackedVfmHead, low0: torch.Tensor, crops: torch.Tensor, crop_hw, out_hw, gh: int, gw: int) -> torch.Tensor:
    """VFMHead.seg_logits_embed (VFMHead.py:38-49,63-71) for the windows in `crops` (int32 [n,4] = image, y1, x1, 0):
    context window of the upsampled coarse logits -> (4gh, 4gw) -> conv k2s2 + GN + GELU -> conv k2s2 + GN + GELU ->
    conv 1x1 + GN. Returns bf16 [n*gh*gw, C]."""
    s, n = head.spec, crops.shape[0]
    a1 = ops.ms_context_im2col(low0, crops, crop_hw, out_hw, (4 * gh, 4 * gw), head.kpad)
    e = ops.gemm_bias_bf16(a1, head.emb1_w, head.emb1_b)
    e = ops.groupnorm_act(e, *head.emb1_gn, n, s.groups, 1e-5, act=2)
    e = ops.space_to_depth2(e, n, 2 * gh, 2 * gw)
    e = ops.gemm_bias_bf16(e, head.emb2_w, head.emb2_b)
    e = ops.groupnorm_act(e, *head.emb2_gn, n, s.groups, 1e-5, act=2)
    e = ops.gemm_bias_bf16(e, head.emb3_w, head.emb3_b)
    return ops.groupnorm_act(e, *head.emb3_gn, n, s.groups, 1e-5, act=0)


def vfm_head_forward(head: PackedVfmHead, taps: torch.Tensor, ctx: torch.Tensor, n: int, gh: int, gw: int) -> torch.Tensor:
    """VFMHead.forward after the context embedding (VFMHead.py:69-87): taps bf16 [n*gh*gw, in_channels] (token-major,
    the K-concatenation of the four feature maps), ctx bf16 [n*gh*gw, C] -> fp32 [n, num_classes, gh, gw]."""
    s = head.spec
    P = gh * gw
    f = ops.gemm_bias_bf16(taps, head.fuse_w, head.fuse_b)                          # fuse_conv[0], VFMHead.py:29
    f = ops.groupnorm_act(f, *head.fuse_gn, n, s.groups, 1e-5, act=2)               # fuse_conv[1:], :30-32
    x = ops.groupnorm_act(f, *head.dec_norm, n, s.groups, 1e-6, act=0, out_f32=True)   # Transformer.py:91-92,276
    for b in head.blocks:                                                           # BasicTransformerBlock._forward :173-177
        h = ops.layernorm(x, *b["n1"], 1e-5)
        a = ops.attention_fwd(ops.gemm_bias_bf16(h, b["qkv1"]), n, P, s.n_heads)    # attn1: self-attention :113-136
        ops.gemm_bias_ls_residual_(x, a, b["out1_w"], b["out1_b"], head.ones)
        h = ops.layernorm(x, *b["n2"], 1e-5)
        q = ops.gemm_bias_bf16(h, b["q2"])
        kv = ops.gemm_bias_bf16(ctx, b["kv2"])                                      # attn2: keys/values from the context
        a = ops.attention_cross(q, kv, n, P, P, s.n_heads)
        ops.gemm_bias_ls_residual_(x, a, b["out2_w"], b["out2_b"], head.ones)
        h = ops.layernorm(x, *b["n3"], 1e-5)
        u = ops.geglu(ops.gemm_bias_bf16(h, b["ff_in_w"], b["ff_in_b"]))            # GEGLU :52-59
        ops.gemm_bias_ls_residual_(x, u, b["ff_out_w"], b["ff_out_b"], head.ones)
    out = ops.gemm_cls_nchw(ops.cast_f32_bf16(x), head.cls_w, head.cls_b, s.num_classes, P)   # cls_seg, VFMHead.py:87
    return out.view(n, s.num_classes, gh, gw)


def ms_slide(engine, aux: PackedVfmHead, img: torch.Tensor, crop_size, stride, *, threshold: float, conf: float,
             lr_size: Tuple[int, int] = (512, 1024), want_logits: bool = False, gate: str = "image"):
    """MsVFMEncoderDecoder.ms_inference (Ms_VFM_encoder_decoder.py:400-466) over [B,3,H,W] uint8 or fp32 input.

    stage 0: whole-image pass at `lr_size` (hard-coded (512, 1024) in the reference, :413) -> coarse logits low0
    stage 1: per window, refine with the aux decoder unless mean(max softmax(context) > threshold) >= conf (:446-452)
    gate='image': every image decides for itself (the reference with its batch_size=1 test loop);
    gate='batch': the reference's literal semantics for B > 1, one decision per window from the batch mean (:448).
    Returns (labels uint8 [B,H,W], logits fp32 [B,nc,H,W] or None, info dict).
    """
    B, _, H, W = img.shape
    ps = engine.vit.spec.patch_size
    if crop_size[0] % ps or crop_size[1] % ps or lr_size[0] % ps or lr_size[1] % ps:
        raise _C.VfmError("crop_size and lr_size must be multiples of the patch size")
    from .engine import slide_boxes
    is_u8 = img.dtype == torch.uint8
    lr = ops.image_resize_norm(img, lr_size, engine.pixel_norm if is_u8 else None)
    crops0 = torch.tensor([(b, 0, 0, 0) for b in range(B)], dtype=torch.int32, device=img.device)
    low0 = engine.crops_lowres(lr, crops0, tuple(lr_size))                 # [B, nc, lr/4]: whole_inference, :420
    boxes = slide_boxes(H, W, crop_size, stride)
    crops, bx = engine._crop_table(B, boxes)
    nk = len(boxes)
    counts = ops.ms_confidence(low0, bx, tuple(crop_size), (H, W), threshold)
    frac = counts.float() / float(crop_size[0] * crop_size[1])
    if gate == "batch":
        frac = frac.mean(0, keepdim=True).expand(B, nk)
    elif gate != "image":
        raise ValueError("gate must be 'image' or 'batch'")
    need = (frac < conf).reshape(-1)
    sel = torch.nonzero(need).reshape(-1)          # the one device->host sync of the path (sizes the refinement GEMMs)
    n_ref = int(sel.numel())
    ref_index = torch.full((B * nk,), -1, dtype=torch.int32, device=img.device)
    refined = None
    gh, gw = crop_size[0] // ps, crop_size[1] // ps
    if n_ref:
        ref_index[sel] = torch.arange(n_ref, dtype=torch.int32, device=img.device)
        crops_ref = crops[sel].contiguous()
        refined = torch.empty(n_ref, aux.spec.num_classes, gh, gw, dtype=torch.float32, device=img.device)
        for s0 in range(0, n_ref, engine.max_crops_per_pass):
            s1 = min(s0 + engine.max_crops_per_pass, n_ref)
            cr = crops_ref[s0:s1].contiguous()
            taps = engine.backbone_taps(img, cr, gh, gw)
            ctx = context_tokens(aux, low0, cr, tuple(crop_size), (H, W), gh, gw)
            refined[s0:s1] = vfm_head_forward(aux, taps, ctx, s1 - s0, gh, gw)
    labels, logits = ops.ms_merge_argmax(low0, refined, ref_index, bx, tuple(crop_size), (H, W), want_logits=want_logits)
    return labels, logits, dict(low0=low0, refined=refined, ref_index=ref_index.view(B, nk), counts=counts, n_refined=n_ref)
