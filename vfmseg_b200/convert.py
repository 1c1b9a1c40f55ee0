"""Offline checkpoint plumbing either side of the hot path. Weight-only preprocessing that runs once
on the host before the engine packs weights; plain torch CPU ops.

* convert_dinov2_state_dict : tools/convert_models/convert_dinov2.py:34-68 — the public DINOv2 ViT-L/14 checkpoint becomes
  the patch-16, 32x32-grid model every reference config runs: bicubic resample of the 14x14 patch-embed kernel to 16x16
  and of the 37x37 position grid to (crop / 16)^2.
* merge_backbone_checkpoint : rein/hooks/load_backbone_hook.py:11-22 — a converted backbone state dict is merged under
  the 'backbone.' prefix into the checkpoint being loaded (trained checkpoints only hold the LoRA / head tensors).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F


def convert_dinov2_state_dict(weight: Dict[str, torch.Tensor], kernel: int = 16, crop_size: Tuple[int, int] = (512, 512)) -> Dict[str, torch.Tensor]:
    out = dict(weight)
    k = "patch_embed.proj.weight"
    out[k] = F.interpolate(weight[k].float(), size=(kernel, kernel), mode="bicubic", align_corners=False)     # :34-44
    pe = weight["pos_embed"]
    pos_cls, pos_tokens = pe[:, :1, :], pe[:, 1:, :]                                                           # :47-68
    dim = pos_tokens.shape[-1]
    orig = int(pos_tokens.shape[-2] ** 0.5)
    grid = tuple(L // kernel for L in crop_size)
    r = F.interpolate(pos_tokens.reshape(-1, orig, orig, dim).permute(0, 3, 1, 2), size=grid, mode="bicubic", align_corners=False)
    out["pos_embed"] = torch.cat((pos_cls, r.permute(0, 2, 3, 1).reshape(-1, grid[0] * grid[1], dim)), dim=1)
    return out


def merge_backbone_checkpoint(checkpoint: dict, backbone_state_dict: Dict[str, torch.Tensor]) -> dict:
    """In place, like LoadBackboneHook.after_load_checkpoint; returns the checkpoint."""
    target = checkpoint["state_dict"] if "state_dict" in checkpoint else checkpoint
    target.update({f"backbone.{k}": v for k, v in backbone_state_dict.items()})
    return checkpoint
