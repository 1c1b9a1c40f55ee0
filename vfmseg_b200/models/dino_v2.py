"""`DinoVisionTransformer` — registered backbone with the reference's constructor kwargs, parameter
names and forward contract (rein/models/backbones/dino_v2.py:55-182, 252-268, 332-333), executed by
the sm_100a engine instead of torch modules.

The nn.Modules below are parameter containers only (their names make reference state dicts load
unchanged); `forward` hands device pointers to libvfmseg_b200.so. There is no torch fallback.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from ..engine import PackedVit, SlideEngine, VitSpec
from ..registry import BACKBONES


class _Attention(nn.Module):  # dino_layers/attention.py:36-54
    def __init__(self, dim, num_heads, qkv_bias, proj_bias):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim, bias=proj_bias)


class _LayerScale(nn.Module):  # dino_layers/layer_scale.py:15-24
    def __init__(self, dim, init_values):
        super().__init__()
        self.gamma = nn.Parameter(init_values * torch.ones(dim))


class _Mlp(nn.Module):  # dino_layers/mlp.py:16-32
    def __init__(self, dim, hidden, bias):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden, bias=bias)
        self.fc2 = nn.Linear(hidden, dim, bias=bias)


class _Block(nn.Module):  # dino_layers/block.py:43-87
    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias, proj_bias, ffn_bias, init_values):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim, num_heads, qkv_bias, proj_bias)
        if init_values:
            self.ls1 = _LayerScale(dim, init_values)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio), ffn_bias)
        if init_values:
            self.ls2 = _LayerScale(dim, init_values)


class _PatchEmbed(nn.Module):  # dino_layers/patch_embed.py:25-66
    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__()
        self.patch_size = (patch_size, patch_size)
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


@BACKBONES.register_module()
class DinoVisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                 qkv_bias=True, ffn_bias=True, proj_bias=True, drop_path_rate=0.0, drop_path_uniform=False,
                 init_values=None, ffn_layer="mlp", block_chunks=1, out_indices=[7, 11, 15, 23], init_cfg=None,
                 resize_feat=False, **unused):
        super().__init__()
        if ffn_layer != "mlp":
            raise NotImplementedError("vfmseg_b200 DinoVisionTransformer: only ffn_layer='mlp' (every DINOv2 config of the reference)")
        if block_chunks not in (0, None):
            raise NotImplementedError("block_chunks > 0 is an FSDP wrapping hint; the reference configs set 0")
        if in_chans != 3:
            raise NotImplementedError("in_chans must be 3")
        if resize_feat:
            raise NotImplementedError("resize_feat=True is not used by the LoRA slide-inference configs")
        self.out_indices = list(out_indices)
        self.num_features = self.embed_dim = embed_dim
        self.num_tokens = 1
        self.n_blocks = depth
        self.num_heads = num_heads
        self.patch_size = patch_size
        self.mlp_ratio = mlp_ratio
        self.patch_embed = _PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
        self.blocks = nn.ModuleList([
            _Block(embed_dim, num_heads, mlp_ratio, qkv_bias, proj_bias, ffn_bias, init_values) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)   # present in checkpoints, unused on this path (:258-268)
        self.mask_token = nn.Parameter(torch.zeros(1, embed_dim))
        self._lora_scale = 1.0
        self._engine: Optional[SlideEngine] = None
        self.register_load_state_dict_post_hook(lambda m, _k: m.invalidate())

    # ------------------------------------------------------------------ engine plumbing
    def invalidate(self):
        self._engine = None

    def spec(self) -> VitSpec:
        return VitSpec(self.embed_dim, self.n_blocks, self.num_heads, int(self.embed_dim * self.mlp_ratio), self.patch_size,
                       tuple(self.out_indices))

    def packed(self, device) -> PackedVit:
        sd = {k: v for k, v in self.state_dict().items()}
        return PackedVit(sd, self.spec(), self._lora_scale, device)

    def engine(self) -> SlideEngine:
        dev = self.cls_token.device
        if dev.type != "cuda":
            raise RuntimeError("vfmseg_b200 runs on CUDA (sm_100a) only: move the model with .cuda(); there is no CPU path")
        if self._engine is None or self._engine.device != dev:
            self._engine = SlideEngine(self.packed(dev), None)
        return self._engine

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    # ------------------------------------------------------------------ reference contract
    def forward_features(self, x: torch.Tensor, masks=None) -> List[torch.Tensor]:
        """[B,3,H,W] fp32 -> list of len(out_indices) tensors [B, C, H/16, W/16] (fp32), the
        un-normalised residual stream after the tapped blocks with the cls token dropped."""
        if masks is not None:
            raise NotImplementedError("masks are a training-time feature")
        B, _, H, W = x.shape
        ps = self.patch_size
        assert H % ps == 0 and W % ps == 0, f"input {H}x{W} is not a multiple of the patch size {ps}"
        gh, gw = H // ps, W // ps
        eng = self.engine()
        crops = torch.tensor([(b, 0, 0, 0) for b in range(B)], dtype=torch.int32, device=x.device)
        taps = eng.backbone_taps(x.contiguous().float(), crops, gh, gw)
        return self.taps_to_maps(taps, B, gh, gw)

    def taps_to_maps(self, taps: torch.Tensor, B: int, gh: int, gw: int) -> List[torch.Tensor]:
        Cc = self.embed_dim
        t = taps.view(B, gh, gw, len(self.out_indices), Cc)
        return [t[:, :, :, i].permute(0, 3, 1, 2).float().contiguous() for i in range(len(self.out_indices))]

    def forward(self, *args, **kwargs):
        return self.forward_features(*args, **kwargs)
