"""`EVA2` — registered backbone with the reference's constructor kwargs, parameter names and forward contract
(rein/models/backbones/eva_02.py:614-853: RoPE, sub-LN q/k/v projections with q/v bias only, SwiGLU feed-forward with an
inner LayerNorm, fixed-grid pos_embed), executed by the sm_100a engine (eva_engine.py). nn.Modules are parameter
containers; their names make reference state dicts (and peft-wrapped ones) load unchanged."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from ..engine import SlideEngine
from ..eva_engine import EvaSpec, PackedEva
from ..registry import BACKBONES


class _Attention(nn.Module):  # eva_02.py:245-330 with subln=True, no relative position bias
    def __init__(self, dim, qkv_bias):
        super().__init__()
        self.q_proj = nn.Linear(dim, dim, bias=False)
        self.k_proj = nn.Linear(dim, dim, bias=False)
        self.v_proj = nn.Linear(dim, dim, bias=False)
        if qkv_bias:
            self.q_bias = nn.Parameter(torch.zeros(dim))
            self.v_bias = nn.Parameter(torch.zeros(dim))
        self.proj = nn.Linear(dim, dim)


class _SwiGLU(nn.Module):  # eva_02.py:204-232 with subln=True
    def __init__(self, dim, hidden):
        super().__init__()
        self.w1 = nn.Linear(dim, hidden)
        self.w2 = nn.Linear(dim, hidden)
        self.ffn_ln = nn.LayerNorm(hidden)
        self.w3 = nn.Linear(hidden, dim)


class _Block(nn.Module):  # eva_02.py:410-484 (init_values=None: no gamma_1/gamma_2)
    def __init__(self, dim, hidden, qkv_bias):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = _Attention(dim, qkv_bias)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _SwiGLU(dim, hidden)


class _PatchEmbed(nn.Module):  # eva_02.py:496-524
    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__()
        self.patch_shape = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.patch_shape[0] * self.patch_shape[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


@BACKBONES.register_module()
class EVA2(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=80, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4 * 2 / 3, qkv_bias=False, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0,
                 hybrid_backbone=None, norm_layer=None, init_values=None, use_checkpoint=False, use_abs_pos_emb=True,
                 use_rel_pos_bias=False, use_shared_rel_pos_bias=False, out_indices=[3, 5, 7, 11], subln=True, xattn=True,
                 naiveswiglu=True, rope=True, pt_hw_seq_len=16, intp_freq=True, pretrained=None, **unused):
        super().__init__()
        unsupported = dict(hybrid_backbone=hybrid_backbone, init_values=init_values, use_rel_pos_bias=use_rel_pos_bias,
                           use_shared_rel_pos_bias=use_shared_rel_pos_bias, qk_scale=qk_scale)
        bad = [k for k, v in unsupported.items() if v]
        if bad or not (subln and naiveswiglu and rope and use_abs_pos_emb and intp_freq) or in_chans != 3:
            raise NotImplementedError(f"vfmseg_b200 EVA2 covers the shipped configuration (configs/_base_/models/lora_eva02_linear.py:"
                                      f"25-50: subln, naiveswiglu, rope, abs pos-embed, no rel-pos bias / LayerScale); got {bad}")
        self.embed_dim = self.num_features = embed_dim
        self.depth, self.num_heads, self.patch_size = depth, num_heads, patch_size
        self.hidden = int(embed_dim * mlp_ratio)
        self.out_indices = list(out_indices)
        self.pt_hw_seq_len = pt_hw_seq_len
        self.patch_embed = _PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
        self.blocks = nn.ModuleList([_Block(embed_dim, self.hidden, qkv_bias) for _ in range(depth)])
        self._lora_scale = 1.0
        self._engine: Optional[SlideEngine] = None
        self.register_load_state_dict_post_hook(lambda m, _k: m.invalidate())

    def invalidate(self):
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def spec(self) -> EvaSpec:
        return EvaSpec(self.embed_dim, self.depth, self.num_heads, self.hidden, self.patch_size, tuple(self.out_indices),
                       self.patch_embed.patch_shape[0], self.pt_hw_seq_len)

    def packed(self, device) -> PackedEva:
        return PackedEva(dict(self.state_dict()), self.spec(), self._lora_scale, device)

    def engine(self) -> SlideEngine:
        dev = self.cls_token.device
        if dev.type != "cuda":
            raise RuntimeError("vfmseg_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        if self._engine is None or self._engine.device != dev:
            self._engine = SlideEngine(self.packed(dev), None)
        return self._engine

    def forward_features(self, x: torch.Tensor) -> List[torch.Tensor]:
        """[B,3,H,W] fp32 (H = W = img_size) -> tuple of [B, C, H/16, W/16] fp32 maps (eva_02.py:816-849)."""
        B, _, H, W = x.shape
        gh, gw = H // self.patch_size, W // self.patch_size
        crops = torch.tensor([(b, 0, 0, 0) for b in range(B)], dtype=torch.int32, device=x.device)
        taps = self.engine().backbone_taps(x.contiguous().float(), crops, gh, gw)
        t = taps.view(B, gh, gw, len(self.out_indices), self.embed_dim)
        return tuple(t[:, :, :, i].permute(0, 3, 1, 2).float().contiguous() for i in range(len(self.out_indices)))

    def forward(self, x):
        return self.forward_features(x)
