"""`SAMViT` — registered backbone with the reference's constructor kwargs, parameter names and forward contract
(rein/models/backbones/sam_vit.py:53-147: ViTDet-style windowed / global attention with decomposed relative position
embeddings, no cls token), executed by the sm_100a engine (sam_engine.py). nn.Modules are parameter containers; their
names make reference state dicts (and peft-wrapped ones) load unchanged."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from ..engine import SlideEngine
from ..registry import BACKBONES
from ..sam_engine import PackedSam, SamSpec


class _Attention(nn.Module):  # sam_vit.py:220-261
    def __init__(self, dim, num_heads, qkv_bias, use_rel_pos, size, global_attn):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        if use_rel_pos:
            n = (4 if global_attn else 2) * size - 1          # :248-260
            self.rel_pos_h = nn.Parameter(torch.zeros(n, dim // num_heads))
            self.rel_pos_w = nn.Parameter(torch.zeros(n, dim // num_heads))


class _Mlp(nn.Module):  # sam_vit.py:17-29
    def __init__(self, dim, hidden):
        super().__init__()
        self.lin1 = nn.Linear(dim, hidden)
        self.lin2 = nn.Linear(hidden, dim)


class _Block(nn.Module):  # sam_vit.py:150-199
    def __init__(self, dim, num_heads, hidden, qkv_bias, use_rel_pos, window_size, grid):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim, num_heads, qkv_bias, use_rel_pos, grid if window_size == 0 else window_size, window_size == 0)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, hidden)
        self.window_size = window_size


class _PatchEmbed(nn.Module):  # sam_vit.py:431-464
    def __init__(self, patch_size, in_chans, embed_dim):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


@BACKBONES.register_module()
class SAMViT(nn.Module):
    def __init__(self, img_size=1024, out_indices=[3, 5, 7, 11], patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4.0, qkv_bias=True, norm_layer=None, act_layer=None, use_abs_pos=True, use_rel_pos=False,
                 rel_pos_zero_init=True, window_size=0, global_attn_indexes=(), init_cfg=None, **unused):
        super().__init__()
        if not use_abs_pos or in_chans != 3:
            raise NotImplementedError("vfmseg_b200 SAMViT covers the shipped configuration (configs/_base_/models/lora_sam_linear.py:"
                                      "16-26: absolute pos-embed, 3 input channels)")
        self.img_size, self.patch_size = img_size, patch_size
        self.embed_dim, self.depth, self.num_heads = embed_dim, depth, num_heads
        self.hidden = int(embed_dim * mlp_ratio)
        self.out_indices = list(out_indices)
        self.window_size = window_size
        self.global_attn_indexes = tuple(global_attn_indexes)
        self.use_rel_pos = use_rel_pos
        grid = img_size // patch_size
        self.patch_embed = _PatchEmbed(patch_size, in_chans, embed_dim)
        self.pos_embed = nn.Parameter(torch.zeros(1, grid, grid, embed_dim))
        self.blocks = nn.ModuleList([
            _Block(embed_dim, num_heads, self.hidden, qkv_bias, use_rel_pos, window_size if i not in self.global_attn_indexes else 0, grid)
            for i in range(depth)])
        self._lora_scale = 1.0
        self._engine: Optional[SlideEngine] = None
        self.register_load_state_dict_post_hook(lambda m, _k: m.invalidate())

    def invalidate(self):
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def spec(self) -> SamSpec:
        return SamSpec(self.embed_dim, self.depth, self.num_heads, self.hidden, self.patch_size, tuple(self.out_indices),
                       self.img_size // self.patch_size, self.window_size, self.global_attn_indexes, self.use_rel_pos)

    def packed(self, device) -> PackedSam:
        return PackedSam(dict(self.state_dict()), self.spec(), self._lora_scale, device)

    def engine(self) -> SlideEngine:
        dev = self.pos_embed.device
        if dev.type != "cuda":
            raise RuntimeError("vfmseg_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        if self._engine is None or self._engine.device != dev:
            self._engine = SlideEngine(self.packed(dev), None)
        return self._engine

    def forward_features(self, x: torch.Tensor) -> List[torch.Tensor]:
        """[B,3,H,W] fp32 (H = W = img_size) -> tuple of [B, C, H/16, W/16] fp32 maps (sam_vit.py:133-147)."""
        B, _, H, W = x.shape
        gh, gw = H // self.patch_size, W // self.patch_size
        crops = torch.tensor([(b, 0, 0, 0) for b in range(B)], dtype=torch.int32, device=x.device)
        taps = self.engine().backbone_taps(x.contiguous().float(), crops, gh, gw)
        t = taps.view(B, gh, gw, len(self.out_indices), self.embed_dim)
        return tuple(t[:, :, :, i].permute(0, 3, 1, 2).float().contiguous() for i in range(len(self.out_indices)))

    def forward(self, x):
        return self.forward_features(x)
