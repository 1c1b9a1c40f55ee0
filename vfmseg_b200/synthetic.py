"""Deterministic synthetic weights, images and labels (there are no checkpoints or datasets here).

The reference's default initialisation would make parity vacuous — cls_token / pos_embed are zeros
(rein/models/backbones/dino_v2.py:123-126), LayerScale gamma is 1e-5
(configs/_base_/models/lora_dinov2_linear.py:31) and peft's LoRA B is zeros — so every tensor is
drawn from a distribution that actually exercises the kernels. The state dict uses
the reference's key names, so the same dict loads into the reference modules (oracle) and into
vfmseg_b200's registered classes.
"""
from __future__ import annotations

import math
from typing import Dict

import torch


def model_config(embed_dim=1024, depth=24, num_heads=16, img_size=512, out_indices=(7, 11, 15, 23), num_classes=19,
                 crop_size=(512, 512), stride=(341, 341), lora_r=32, lora_alpha=32, mode="slide", gn_groups=32) -> dict:
    """A config dict shaped like configs/_base_/models/lora_dinov2_linear.py:4-55 (same `type=` names)."""
    return dict(
        type="LoraBackboneEncoderDecoder",
        checkpoint=None,
        Lora_config=dict(r=lora_r, lora_alpha=lora_alpha, target_modules=["qkv"], lora_dropout=0.1),
        data_preprocessor=dict(type="SegDataPreProcessor", mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375],
                               size=tuple(crop_size), bgr_to_rgb=True, pad_val=0, seg_pad_val=255),
        backbone=dict(type="DinoVisionTransformer", patch_size=16, embed_dim=embed_dim, depth=depth, num_heads=num_heads,
                      mlp_ratio=4, img_size=img_size, ffn_layer="mlp", init_values=1e-05, block_chunks=0, qkv_bias=True,
                      proj_bias=True, ffn_bias=True, out_indices=list(out_indices)),
        decode_head=dict(type="LinearHead", in_channels=[embed_dim] * len(out_indices), in_index=list(range(len(out_indices))),
                         channels=embed_dim // 4, dropout_ratio=0.1, num_classes=num_classes,
                         norm_cfg=dict(type="GN", num_groups=gn_groups), align_corners=False,
                         loss_decode=dict(type="CrossEntropyLoss", use_sigmoid=False, loss_weight=1.0)),
        train_cfg=dict(),
        test_cfg=dict(mode=mode, stride=list(stride), crop_size=list(crop_size)),
    )


def tiny_config(**kw) -> dict:
    """Small model for fast CPU oracle runs: dim 256 (4 heads), depth 4, 64x64 crops."""
    d = dict(embed_dim=256, depth=4, num_heads=4, img_size=64, out_indices=(0, 1, 2, 3), crop_size=(64, 64),
             stride=(43, 43), lora_r=8, lora_alpha=16, gn_groups=32)
    d.update(kw)
    return model_config(**d)


def synthetic_state_dict(cfg: dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Full segmentor state dict (keys as in the reference: 'backbone.base_model.model.*' because the
    backbone is peft-wrapped inside the segmentor, Lora_encoder_decoder.py:24; 'decode_head.*')."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    bb, hd, lc = cfg["backbone"], cfg["decode_head"], cfg["Lora_config"]
    C, depth, P = bb["embed_dim"], bb["depth"], bb["patch_size"]
    hidden = int(C * bb["mlp_ratio"])
    n_pos = (bb["img_size"] // P) ** 2
    r = lc["r"]

    def normal(*shape, std=0.02):
        return torch.randn(*shape, generator=g) * std

    def uniform(*shape, lo=0.0, hi=1.0):
        return torch.rand(*shape, generator=g) * (hi - lo) + lo

    sd: Dict[str, torch.Tensor] = {}
    p = "backbone.base_model.model."
    sd[p + "cls_token"] = normal(1, 1, C, std=0.1)
    sd[p + "pos_embed"] = normal(1, 1 + n_pos, C, std=0.1)
    sd[p + "mask_token"] = torch.zeros(1, C)
    sd[p + "patch_embed.proj.weight"] = normal(C, 3, P, P)
    sd[p + "patch_embed.proj.bias"] = normal(C)
    for i in range(depth):
        b = f"{p}blocks.{i}."
        sd[b + "norm1.weight"] = uniform(C, lo=0.5, hi=1.5)
        sd[b + "norm1.bias"] = normal(C, std=0.1)
        sd[b + "attn.qkv.base_layer.weight"] = normal(3 * C, C)
        # queries/keys a little hotter than 0.02 so the softmax is not uniform
        sd[b + "attn.qkv.base_layer.weight"][: 2 * C] *= 2.0
        sd[b + "attn.qkv.base_layer.bias"] = normal(3 * C)
        bound = 1.0 / math.sqrt(C)  # kaiming_uniform(a=sqrt(5)) bound, as peft initialises lora_A
        sd[b + "attn.qkv.lora_A.default.weight"] = uniform(r, C, lo=-bound, hi=bound)
        sd[b + "attn.qkv.lora_B.default.weight"] = normal(3 * C, r)
        sd[b + "attn.proj.weight"] = normal(C, C)
        sd[b + "attn.proj.bias"] = normal(C)
        sd[b + "ls1.gamma"] = uniform(C, lo=0.05, hi=0.5)
        sd[b + "norm2.weight"] = uniform(C, lo=0.5, hi=1.5)
        sd[b + "norm2.bias"] = normal(C, std=0.1)
        sd[b + "mlp.fc1.weight"] = normal(hidden, C)
        sd[b + "mlp.fc1.bias"] = normal(hidden)
        sd[b + "mlp.fc2.weight"] = normal(C, hidden)
        sd[b + "mlp.fc2.bias"] = normal(C)
        sd[b + "ls2.gamma"] = uniform(C, lo=0.05, hi=0.5)
    sd[p + "norm.weight"] = torch.ones(C)
    sd[p + "norm.bias"] = torch.zeros(C)

    h = "decode_head."
    n_in = sum(hd["in_channels"])
    mid = hd["in_channels"][0]
    ch = hd["channels"]
    nc = hd["num_classes"]
    assert ch == mid // 4, "LinearHead: conv_seg input channels (channels) must equal in_channels[0] // 4"
    sd[h + "fusion_conv.conv.weight"] = normal(mid, n_in, 1, 1, std=n_in ** -0.5)
    sd[h + "fusion_conv.gn.weight"] = uniform(mid, lo=0.5, hi=1.5)
    sd[h + "fusion_conv.gn.bias"] = normal(mid, std=0.1)
    sd[h + "output_upscaling.0.weight"] = normal(mid, mid // 2, 2, 2, std=mid ** -0.5)
    sd[h + "output_upscaling.0.bias"] = normal(mid // 2)
    sd[h + "output_upscaling.1.weight"] = uniform(mid // 2, lo=0.5, hi=1.5)
    sd[h + "output_upscaling.1.bias"] = normal(mid // 2, std=0.1)
    sd[h + "output_upscaling.1.running_mean"] = normal(mid // 2, std=0.1)
    sd[h + "output_upscaling.1.running_var"] = uniform(mid // 2, lo=0.5, hi=2.0)
    sd[h + "output_upscaling.1.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    sd[h + "output_upscaling.3.weight"] = normal(mid // 2, mid // 4, 2, 2, std=(mid // 2) ** -0.5)
    sd[h + "output_upscaling.3.bias"] = normal(mid // 4)
    sd[h + "conv_seg.weight"] = normal(nc, ch, 1, 1, std=2.0 * ch ** -0.5)
    sd[h + "conv_seg.bias"] = normal(nc, std=0.1)
    return sd


def backbone_checkpoint_from(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The 'checkpoints/dinov2_converted.pth'-style dict the reference's __init__ loads
    (Lora_encoder_decoder.py:28-36): plain backbone keys, no LoRA tensors."""
    out = {}
    p = "backbone.base_model.model."
    for k, v in sd.items():
        if k.startswith(p) and "lora_" not in k:
            out[k[len(p):].replace(".base_layer", "")] = v
    return out


def synthetic_images(n: int, H: int, W: int, seed: int = 1234) -> torch.Tensor:
    """uint8 BGR [n,3,H,W]: uniform noise low-pass filtered (8x8 average, bilinear back up) plus a
    little per-pixel noise, so neighbouring windows differ smoothly and pixels are not all alike."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    hs, ws = max(H // 8, 1), max(W // 8, 1)
    coarse = torch.rand(n, 3, hs, ws, generator=g) * 255.0
    img = torch.nn.functional.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=False)
    img = img + (torch.rand(n, 3, H, W, generator=g) - 0.5) * 32.0
    return img.clamp_(0, 255).round_().to(torch.uint8)


def region_images(n: int, H: int, W: int, seed: int = 11, cell: int = 128, num_classes: int = 19):
    """The "trained-network-like" input recipe: uint8 BGR [n,3,H,W] images made of cell x cell blocks, each block one of
    `num_classes` palette colours (+-8 levels of per-pixel noise), and the planted uint8 label map [n,H,W] (block -> palette
    index). Together with a classifier fitted on the features of exactly these images (oracle/probe.py; the fitted
    `conv_seg` travels in tests/golden/*_probe.npz) the head's logits get what a trained network's have: one class well
    ahead inside every region, near-ties only along region boundaries. With the plain random classifier the top-2 margin is
    below the bf16 logit error on ~0.5 % of the pixels whatever the input is (19 near-iid logits), and no implementation in
    bf16 — the reference itself under CPU autocast included — can agree with the fp32 argmax on 99.9 % of them."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    palette = torch.randint(0, 256, (num_classes, 3), generator=g).float()
    hs, ws = -(-H // cell), -(-W // cell)
    lab = torch.randint(0, num_classes, (n, hs, ws), generator=g)
    img = palette[lab].permute(0, 3, 1, 2)
    img = img.repeat_interleave(cell, 2).repeat_interleave(cell, 3)[:, :, :H, :W]
    img = img + (torch.rand(n, 3, H, W, generator=g) - 0.5) * 16.0
    planted = lab.repeat_interleave(cell, 1).repeat_interleave(cell, 2)[:, :H, :W].contiguous()
    return img.clamp_(0, 255).round_().to(torch.uint8), planted.to(torch.uint8)


def with_probe_classifier(sd: Dict[str, torch.Tensor], weight, bias, prefix: str = "decode_head.") -> Dict[str, torch.Tensor]:
    """Copy of `sd` with the head's 1x1 classifier replaced by a fitted probe (weight [nc, ch], bias [nc])."""
    out = dict(sd)
    w = torch.as_tensor(weight, dtype=torch.float32)
    out[prefix + "conv_seg.weight"] = w.reshape(w.shape[0], w.shape[1], 1, 1).clone()
    out[prefix + "conv_seg.bias"] = torch.as_tensor(bias, dtype=torch.float32).clone()
    return out


def synthetic_labels(n: int, H: int, W: int, num_classes: int = 19, seed: int = 4321, ignore_frac: float = 0.05) -> torch.Tensor:
    """uint8 [n,H,W]: piecewise-constant class map (16x16 blocks) with ~5 % ignore (255) pixels."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    hs, ws = (H + 15) // 16, (W + 15) // 16
    blocks = torch.randint(0, num_classes, (n, hs, ws), generator=g)
    lab = blocks.repeat_interleave(16, 1).repeat_interleave(16, 2)[:, :H, :W].contiguous()
    ign = torch.rand(n, H, W, generator=g) < ignore_frac
    lab[ign] = 255
    return lab.to(torch.uint8)


# ------------------------------------------------------------------ coarse-to-fine configs (BASELINE config 3)
def ms_model_config(embed_dim=1024, depth=24, num_heads=16, img_size=512, out_indices=(7, 11, 15, 23), num_classes=19,
                    crop_size=(512, 512), stride=(320, 320), lora_r=32, lora_alpha=32, threshold=0.968, conf=0.8,
                    aux_channels=256, aux_heads=8, aux_depth=3, mode="ms_slide_inference") -> dict:
    """A config dict shaped like configs/_base_/models/lora_dinov2_ms_masked.py:3-87 (same `type=` names). `train_cfg`
    carries the log_config entry tools/test.py:118-119 injects (read by Ms_VFM_encoder_decoder.py:104)."""
    n = len(out_indices)
    head_common = dict(in_channels=[embed_dim] * n, in_index=list(range(n)), dropout_ratio=0.1, num_classes=num_classes,
                       norm_cfg=dict(type="GN", num_groups=32), align_corners=False,
                       loss_decode=dict(type="CrossEntropyLoss", use_sigmoid=False, loss_weight=1.0))
    return dict(
        type="MsVFMEncoderDecoder",
        data_preprocessor=dict(type="SegDataPreProcessor", mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375],
                               size=(1024, 1024), bgr_to_rgb=True, pad_val=0, seg_pad_val=255),
        backbone=dict(type="LoRABackbone",
                      backbone=dict(type="DinoVisionTransformer", patch_size=16, embed_dim=embed_dim, depth=depth,
                                    num_heads=num_heads, mlp_ratio=4, img_size=img_size, ffn_layer="mlp", init_values=1e-05,
                                    block_chunks=0, qkv_bias=True, proj_bias=True, ffn_bias=True, out_indices=list(out_indices)),
                      checkpoint=None,
                      Lora_config=dict(r=lora_r, lora_alpha=lora_alpha, target_modules=["qkv"], lora_dropout=0.1)),
        decode_head=dict(type="LinearHead", channels=embed_dim // 4, **head_common),
        aux_head=dict(type="VFMHead",
                      transformer=dict(type="MaskTransformerDecoder", query_dim=aux_channels, n_heads=aux_heads, d_head=64,
                                       depth=aux_depth, dropout=0.1, mask_ratio=0.2),
                      channels=aux_channels, **head_common),
        detail_loss=1.0, scales=[1, 0.5], hr_crop_size=tuple(crop_size), feature_scale=0.5, crop_coord_divisible=32,
        train_cfg=dict(log_config=dict(img_interval=100)),
        test_cfg=dict(mode=mode, threadshod=threshold, conf=conf, lr_img_size=(512, 1024), stride=list(stride),
                      crop_size=list(crop_size)),
    )


def tiny_ms_config(**kw) -> dict:
    d = dict(embed_dim=256, depth=4, num_heads=4, img_size=64, out_indices=(0, 1, 2, 3), crop_size=(64, 64), stride=(43, 43),
             lora_r=8, lora_alpha=16, aux_depth=2)
    d.update(kw)
    return ms_model_config(**d)


def synthetic_ms_state_dict(cfg: dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    """State dict of MsVFMEncoderDecoder with the reference's key names: 'backbone.model.base_model.model.*'
    (LoRABackbone, lora_backbone.py:23), 'decode_head.*', 'aux_decoder.*' (Ms_VFM_encoder_decoder.py:112)."""
    base_cfg = dict(backbone=cfg["backbone"]["backbone"], decode_head=cfg["decode_head"], Lora_config=cfg["backbone"]["Lora_config"])
    base = synthetic_state_dict(base_cfg, seed=seed)
    sd: Dict[str, torch.Tensor] = {}
    for k, v in base.items():
        sd[k.replace("backbone.base_model.model.", "backbone.model.base_model.model.", 1)] = v
    g = torch.Generator(device="cpu").manual_seed(seed + 1000)

    def normal(*shape, std=0.02):
        return torch.randn(*shape, generator=g) * std

    def uniform(*shape, lo=0.0, hi=1.0):
        return torch.rand(*shape, generator=g) * (hi - lo) + lo

    def gn(prefix, c):
        sd[prefix + ".weight"] = uniform(c, lo=0.5, hi=1.5)
        sd[prefix + ".bias"] = normal(c, std=0.1)

    ah = cfg["aux_head"]
    tr = ah["transformer"]
    C, nc = ah["channels"], ah["num_classes"]
    n_in = sum(ah["in_channels"])
    inner = tr["n_heads"] * tr["d_head"]
    a = "aux_decoder."
    sd[a + "fuse_conv.0.weight"] = normal(C, n_in, 1, 1, std=n_in ** -0.5)
    sd[a + "fuse_conv.0.bias"] = normal(C, std=0.1)
    gn(a + "fuse_conv.1", C)
    sd[a + "seg_logits_embed.0.weight"] = normal(C // 4, nc, 2, 2, std=(4 * nc) ** -0.5)
    sd[a + "seg_logits_embed.0.bias"] = normal(C // 4, std=0.1)
    gn(a + "seg_logits_embed.1", C // 4)
    sd[a + "seg_logits_embed.3.weight"] = normal(C // 2, C // 4, 2, 2, std=C ** -0.5)
    sd[a + "seg_logits_embed.3.bias"] = normal(C // 2, std=0.1)
    gn(a + "seg_logits_embed.4", C // 2)
    sd[a + "seg_logits_embed.6.weight"] = normal(C, C // 2, 1, 1, std=(C // 2) ** -0.5)
    sd[a + "seg_logits_embed.6.bias"] = normal(C, std=0.1)
    gn(a + "seg_logits_embed.7", C)
    t = a + "transformer_decoder."
    gn(t + "norm", C)
    sd[t + "mask_token"] = normal(1, C, 1, 1, std=1.0)
    for i in range(tr["depth"]):
        b = f"{t}transformer_blocks.{i}."
        for att, cdim in (("attn1", C), ("attn2", C)):
            sd[b + att + ".to_q.weight"] = normal(inner, C, std=2.0 * C ** -0.5)    # hot enough for a non-uniform softmax
            sd[b + att + ".to_k.weight"] = normal(inner, cdim, std=2.0 * cdim ** -0.5)
            sd[b + att + ".to_v.weight"] = normal(inner, cdim, std=cdim ** -0.5)
            sd[b + att + ".to_out.0.weight"] = normal(C, inner, std=0.5 * inner ** -0.5)
            sd[b + att + ".to_out.0.bias"] = normal(C, std=0.02)
        sd[b + "ff.net.0.proj.weight"] = normal(8 * C, C, std=C ** -0.5)
        sd[b + "ff.net.0.proj.bias"] = normal(8 * C, std=0.02)
        sd[b + "ff.net.2.weight"] = normal(C, 4 * C, std=0.5 * (4 * C) ** -0.5)
        sd[b + "ff.net.2.bias"] = normal(C, std=0.02)
        for n_ in ("norm1", "norm2", "norm3"):
            gn(b + n_, C)
    sd[a + "conv_seg.weight"] = normal(nc, C, 1, 1, std=2.0 * C ** -0.5)
    sd[a + "conv_seg.bias"] = normal(nc, std=0.1)
    return sd


def ms_backbone_checkpoint_from(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Plain backbone checkpoint (what LoRABackbone.__init__ loads, lora_backbone.py:27-35) from an MsVFM state dict."""
    out = {}
    p = "backbone.model.base_model.model."
    for k, v in sd.items():
        if k.startswith(p) and "lora_" not in k:
            out[k[len(p):].replace(".base_layer", "")] = v
    return out


# ------------------------------------------------------------------ EVA02 configs (BASELINE config 4)
def eva_model_config(embed_dim=1024, depth=24, num_heads=16, img_size=512, out_indices=(7, 11, 15, 23), num_classes=19,
                     crop_size=(512, 512), stride=(320, 320), lora_r=32, lora_alpha=32, mode="slide") -> dict:
    """A config dict shaped like configs/_base_/models/lora_eva02_linear.py:3-71 (same `type=` names)."""
    n = len(out_indices)
    return dict(
        type="EncoderDecoder",
        data_preprocessor=dict(type="SegDataPreProcessor", mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375],
                               size=tuple(crop_size), bgr_to_rgb=True, pad_val=0, seg_pad_val=255),
        backbone=dict(type="LoRABackbone",
                      Lora_config=dict(lora_alpha=lora_alpha, lora_dropout=0.1, r=lora_r,
                                       target_modules=["q_proj", "k_proj", "v_proj", "attn.proj"]),
                      checkpoint=None,
                      backbone=dict(type="EVA2", depth=depth, drop_path_rate=0.1, embed_dim=embed_dim, img_size=img_size, in_chans=3,
                                    init_values=None, intp_freq=True, mlp_ratio=2.6666666666666665, naiveswiglu=True,
                                    norm_layer=dict(eps=1e-06, requires_grad=True, type="LN"), num_heads=num_heads,
                                    out_indices=list(out_indices), patch_size=16, pt_hw_seq_len=16, qkv_bias=True, rope=True,
                                    subln=True, use_abs_pos_emb=True, use_checkpoint=False, use_rel_pos_bias=False,
                                    use_shared_rel_pos_bias=False, xattn=True)),
        decode_head=dict(type="LinearHead", in_channels=[embed_dim] * n, in_index=list(range(n)), channels=embed_dim // 4,
                         dropout_ratio=0.1, num_classes=num_classes, norm_cfg=dict(type="GN", num_groups=32), align_corners=False,
                         loss_decode=dict(type="CrossEntropyLoss", use_sigmoid=False, loss_weight=1.0)),
        train_cfg=dict(),
        test_cfg=dict(mode=mode, stride=list(stride), crop_size=list(crop_size)),
    )


def tiny_eva_config(**kw) -> dict:
    d = dict(embed_dim=256, depth=4, num_heads=4, img_size=64, out_indices=(0, 1, 2, 3), crop_size=(64, 64), stride=(43, 43),
             lora_r=8, lora_alpha=16)
    d.update(kw)
    return eva_model_config(**d)


def synthetic_eva_state_dict(cfg: dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    """State dict of EncoderDecoder(LoRABackbone(EVA2), LinearHead) with the reference's key names
    ('backbone.model.base_model.model.*' with peft's base_layer / lora_A / lora_B for q_proj, k_proj, v_proj, attn.proj)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    bb, lc = cfg["backbone"]["backbone"], cfg["backbone"]["Lora_config"]
    C, depth, P = bb["embed_dim"], bb["depth"], bb["patch_size"]
    hidden = int(C * bb["mlp_ratio"])
    n_pos = (bb["img_size"] // P) ** 2
    r = lc["r"]

    def normal(*shape, std=0.02):
        return torch.randn(*shape, generator=g) * std

    def uniform(*shape, lo=0.0, hi=1.0):
        return torch.rand(*shape, generator=g) * (hi - lo) + lo

    sd: Dict[str, torch.Tensor] = {}
    p = "backbone.model.base_model.model."
    sd[p + "cls_token"] = normal(1, 1, C, std=0.1)
    sd[p + "pos_embed"] = normal(1, 1 + n_pos, C, std=0.1)
    sd[p + "patch_embed.proj.weight"] = normal(C, 3, P, P)
    sd[p + "patch_embed.proj.bias"] = normal(C)
    bound = 1.0 / math.sqrt(C)

    def lora_linear(key, n_out, n_in, std, bias):
        sd[key + ".base_layer.weight"] = normal(n_out, n_in, std=std)
        if bias:
            sd[key + ".base_layer.bias"] = normal(n_out)
        sd[key + ".lora_A.default.weight"] = uniform(r, n_in, lo=-bound, hi=bound)
        sd[key + ".lora_B.default.weight"] = normal(n_out, r)

    for i in range(depth):
        b = f"{p}blocks.{i}."
        for nm in ("norm1", "norm2"):
            sd[b + nm + ".weight"] = uniform(C, lo=0.5, hi=1.5)
            sd[b + nm + ".bias"] = normal(C, std=0.1)
        lora_linear(b + "attn.q_proj", C, C, 0.04, False)     # hotter q/k: a non-uniform softmax
        lora_linear(b + "attn.k_proj", C, C, 0.04, False)
        lora_linear(b + "attn.v_proj", C, C, 0.02, False)
        sd[b + "attn.q_bias"] = normal(C)
        sd[b + "attn.v_bias"] = normal(C)
        lora_linear(b + "attn.proj", C, C, 0.01, True)        # no LayerScale in EVA02: keep the residual branches moderate
        sd[b + "mlp.w1.weight"] = normal(hidden, C, std=0.03)
        sd[b + "mlp.w1.bias"] = normal(hidden)
        sd[b + "mlp.w2.weight"] = normal(hidden, C, std=0.03)
        sd[b + "mlp.w2.bias"] = normal(hidden)
        sd[b + "mlp.ffn_ln.weight"] = uniform(hidden, lo=0.5, hi=1.5)
        sd[b + "mlp.ffn_ln.bias"] = normal(hidden, std=0.1)
        sd[b + "mlp.w3.weight"] = normal(C, hidden, std=0.005)
        sd[b + "mlp.w3.bias"] = normal(C)
    lin_cfg = dict(backbone=dict(embed_dim=C, depth=0, patch_size=P, mlp_ratio=4, img_size=bb["img_size"]),
                   decode_head=cfg["decode_head"], Lora_config=dict(r=r))
    for k, v in synthetic_state_dict(lin_cfg, seed=seed + 7).items():
        if k.startswith("decode_head."):
            sd[k] = v
    return sd


# ------------------------------------------------------------------ SAM ViT configs (BASELINE config 5)
def sam_model_config(embed_dim=1280, depth=32, num_heads=16, img_size=512, out_indices=(7, 15, 23, 31),
                     global_attn_indexes=(7, 15, 23, 31), window_size=14, num_classes=19, crop_size=(512, 512), stride=(320, 320),
                     lora_r=32, lora_alpha=32, mode="slide") -> dict:
    """A config dict shaped like configs/_base_/models/lora_sam_linear.py:3-55 (same `type=` names)."""
    n = len(out_indices)
    return dict(
        type="EncoderDecoder",
        data_preprocessor=dict(type="SegDataPreProcessor", mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375],
                               size=tuple(crop_size), bgr_to_rgb=True, pad_val=0, seg_pad_val=255),
        backbone=dict(type="LoRABackbone",
                      backbone=dict(type="SAMViT", img_size=img_size, embed_dim=embed_dim, depth=depth, num_heads=num_heads,
                                    global_attn_indexes=list(global_attn_indexes), out_indices=list(out_indices),
                                    window_size=window_size, use_rel_pos=True),
                      checkpoint=None,
                      Lora_config=dict(r=lora_r, lora_alpha=lora_alpha, target_modules=["qkv"], lora_dropout=0.1)),
        decode_head=dict(type="LinearHead", in_channels=[embed_dim] * n, in_index=list(range(n)), channels=embed_dim // 4,
                         dropout_ratio=0.1, num_classes=num_classes, norm_cfg=dict(type="GN", num_groups=32), align_corners=False,
                         loss_decode=dict(type="CrossEntropyLoss", use_sigmoid=False, loss_weight=1.0)),
        train_cfg=dict(),
        test_cfg=dict(mode=mode, stride=list(stride), crop_size=list(crop_size)),
    )


def tiny_sam_config(**kw) -> dict:
    """8 heads x 80 (640 is the smallest width with head_dim 80 that the head's ConvT GEMMs accept: C/4 % 32 == 0), 16x16
    tokens: windowed blocks pad 16 -> 28 (four 14x14 windows, three of them with pad tokens), global blocks interpolate
    their 63-entry rel-pos tables to 31."""
    d = dict(embed_dim=640, depth=4, num_heads=8, img_size=256, out_indices=(0, 1, 2, 3), global_attn_indexes=(1, 3),
             crop_size=(256, 256), stride=(171, 171), lora_r=8, lora_alpha=16)
    d.update(kw)
    return sam_model_config(**d)


def synthetic_sam_state_dict(cfg: dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    """State dict of EncoderDecoder(LoRABackbone(SAMViT), LinearHead) with the reference's key names
    ('backbone.model.base_model.model.*', peft layout for attn.qkv; rein/models/backbones/sam_vit.py:97-121,229-254)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    bb, lc = cfg["backbone"]["backbone"], cfg["backbone"]["Lora_config"]
    C, depth, P, heads = bb["embed_dim"], bb["depth"], 16, bb["num_heads"]
    grid, ws, d = bb["img_size"] // P, bb["window_size"], bb["embed_dim"] // bb["num_heads"]
    r = lc["r"]

    def normal(*shape, std=0.02):
        return torch.randn(*shape, generator=g) * std

    def uniform(*shape, lo=0.0, hi=1.0):
        return torch.rand(*shape, generator=g) * (hi - lo) + lo

    sd: Dict[str, torch.Tensor] = {}
    p = "backbone.model.base_model.model."
    sd[p + "pos_embed"] = normal(1, grid, grid, C, std=0.1)
    sd[p + "patch_embed.proj.weight"] = normal(C, 3, P, P)
    sd[p + "patch_embed.proj.bias"] = normal(C)
    bound = 1.0 / math.sqrt(C)
    for i in range(depth):
        b = f"{p}blocks.{i}."
        for nm in ("norm1", "norm2"):
            sd[b + nm + ".weight"] = uniform(C, lo=0.5, hi=1.5)
            sd[b + nm + ".bias"] = normal(C, std=0.1)
        sd[b + "attn.qkv.base_layer.weight"] = normal(3 * C, C, std=0.04)
        sd[b + "attn.qkv.base_layer.bias"] = normal(3 * C, std=0.1)          # pad tokens enter attention as the bias
        sd[b + "attn.qkv.lora_A.default.weight"] = uniform(r, C, lo=-bound, hi=bound)
        sd[b + "attn.qkv.lora_B.default.weight"] = normal(3 * C, r)
        n_rel = (4 * grid - 1) if i in bb["global_attn_indexes"] else (2 * ws - 1)
        sd[b + "attn.rel_pos_h"] = normal(n_rel, d, std=0.1)
        sd[b + "attn.rel_pos_w"] = normal(n_rel, d, std=0.1)
        sd[b + "attn.proj.weight"] = normal(C, C, std=0.01)                   # no LayerScale: keep the branches moderate
        sd[b + "attn.proj.bias"] = normal(C)
        sd[b + "mlp.lin1.weight"] = normal(4 * C, C, std=0.03)
        sd[b + "mlp.lin1.bias"] = normal(4 * C)
        sd[b + "mlp.lin2.weight"] = normal(C, 4 * C, std=0.005)
        sd[b + "mlp.lin2.bias"] = normal(C)
    lin_cfg = dict(backbone=dict(embed_dim=C, depth=0, patch_size=P, mlp_ratio=4, img_size=bb["img_size"]),
                   decode_head=cfg["decode_head"], Lora_config=dict(r=r))
    for k, v in synthetic_state_dict(lin_cfg, seed=seed + 7).items():
        if k.startswith("decode_head."):
            sd[k] = v
    return sd
