"""Offline checkpoint plumbing either side of the hot path. Weight-only preprocessing that runs once
on the host before the engine packs weights; plain torch CPU ops.

* convert_dinov2_state_dict : tools/convert_models/convert_dinov2.py:34-68 — the public DINOv2 ViT-L/14 checkpoint becomes
  the patch-16, 32x32-grid model every reference config runs: bicubic resample of the 14x14 patch-embed kernel to 16x16
  and of the 37x37 position grid to (crop / 16)^2.
* convert_sam_state_dict : tools/convert_models/convert_sam.py:21-73 — the `image_encoder.` sub-tree of a segment-anything
  checkpoint, patch-embed kernel resampled to `kernel`, the [1, H, W, C] position grid to (crop / kernel).
* convert_eva2_state_dict : tools/convert_models/convert_eva2_512x512.py:6-117 — EVA02 psz14 checkpoint: `rope` buffers dropped,
  patch-embed kernel to 16x16, position grid (cls token kept) to 32x32 = 1024 patches.
* generate_full_weights : tools/generate_full_weights.py:6-46 — DINOv2-L backbone (37x37 grid -> 32x32, 14 -> 16 kernel) merged
  under `backbone.` into a trained Rein / LoRA head checkpoint.
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


def convert_sam_state_dict(weight: Dict[str, torch.Tensor], kernel: int = 16, crop_size: Tuple[int, int] = (512, 512)) -> Dict[str, torch.Tensor]:
    """`weight`: a segment-anything checkpoint (keys `image_encoder.*`, others are dropped)."""
    out = {k.replace("image_encoder.", ""): v for k, v in weight.items() if "image_encoder." in k}          # :21-22,31
    if len(out) <= 10:                                                                                        # :32-35
        raise KeyError(f"the checkpoint holds {len(out)} image_encoder tensors: {list(out)}")
    k = "patch_embed.proj.weight"
    out[k] = F.interpolate(out[k].float(), size=(kernel, kernel), mode="bicubic", align_corners=False)        # :39-49
    grid = tuple(L // kernel for L in crop_size)                                                              # :60
    pe = out["pos_embed"]                                                                                     # [1, H, W, C]
    out["pos_embed"] = F.interpolate(pe.permute(0, 3, 1, 2), size=grid, mode="bicubic", align_corners=False).permute(0, 2, 3, 1)   # :61-68
    return out


def _eva2_resize_pos(tokens: torch.Tensor, new_size: int) -> torch.Tensor:
    """[n, 1 + s*s, C] -> [n, 1 + new_size^2, C]; the class token is kept, only the position tokens are interpolated."""
    c = tokens.shape[-1]
    orig = int((tokens.shape[-2] - 1) ** 0.5)
    extra, pos = tokens[:, :1], tokens[:, 1:]
    pos = F.interpolate(pos.reshape(-1, orig, orig, c).permute(0, 3, 1, 2).float(), size=(new_size, new_size), mode="bicubic",
                        align_corners=False)
    return torch.cat((extra, pos.permute(0, 2, 3, 1).flatten(1, 2)), dim=1)


def convert_eva2_state_dict(checkpoint: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = dict(checkpoint["model"]) if "model" in checkpoint else dict(checkpoint)                            # :92-93
    for k in [k for k in out if "rope" in k]:                                                                 # :96-102
        out.pop(k)
    k = "patch_embed.proj.weight"
    out[k] = F.interpolate(out[k].float(), size=(16, 16), mode="bicubic", align_corners=False)                # :105-110
    new_size = int(1024 ** 0.5)                                                                               # :10,16 (num_patches = 1024 whatever new_size says)
    if "pos_embed" in out:                                                                                    # :7-38  [1, 1 + s*s, C]
        out["pos_embed"] = _eva2_resize_pos(out["pos_embed"], new_size)
    if "positional_embedding" in out:                                                                         # :39-71 [1 + s*s, C]
        out["positional_embedding"] = _eva2_resize_pos(out["positional_embedding"][None], new_size).squeeze(0)
    return out


def generate_full_weights(backbone: Dict[str, torch.Tensor], rein_head: dict) -> dict:
    """tools/generate_full_weights.py: the DINOv2-L/14 backbone resampled exactly as that script does (37x37 -> 32x32 position
    grid, 14x14 -> 16x16 kernel) and merged into the head checkpoint's state_dict under `backbone.`."""
    w = dict(backbone)
    pe = w["pos_embed"]                                                                                       # :13-29
    w["pos_embed"] = torch.cat((pe[:, :1, :],
                                F.interpolate(pe[:, 1:, :].reshape(1, 37, 37, 1024).permute(0, 3, 1, 2), size=(32, 32), mode="bicubic",
                                              align_corners=False).permute(0, 2, 3, 1).reshape(1, 1024, 1024)), dim=1)
    w["patch_embed.proj.weight"] = F.interpolate(w["patch_embed.proj.weight"].float(), size=(16, 16), mode="bicubic", align_corners=False)   # :30-35
    rein_head["state_dict"].update({f"backbone.{k}": v for k, v in w.items()})                               # :45
    return rein_head


def merge_backbone_checkpoint(checkpoint: dict, backbone_state_dict: Dict[str, torch.Tensor]) -> dict:
    """In place, like LoadBackboneHook.after_load_checkpoint; returns the checkpoint."""
    target = checkpoint["state_dict"] if "state_dict" in checkpoint else checkpoint
    target.update({f"backbone.{k}": v for k, v in backbone_state_dict.items()})
    return checkpoint
