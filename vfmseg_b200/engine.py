"""Host-side engine of the slide-inference hot path: packs weights for the CUDA library, owns the
device workspaces and drives the fused C-ABI entry points.

    images ──patch gather──► ViT blocks (tcgen05 GEMMs + fused attention) ──► taps
          ──LinearHead (GEMMs + GroupNorm)──► low-res crop logits
          ──slide_merge_argmax──► label map (+ optional fp32 logits) ──confusion_matrix──► int64 cm

Everything numerically relevant happens in libvfmseg_b200.so; torch supplies device memory,
the current stream and (once per weight load / input shape) constant folding of weights:
LoRA merge W + (alpha/r)·B·A (peft lora.Linear semantics, used at
rein/models/segmentors/Lora_encoder_decoder.py:24), the exact 2^-3 q-scale fold
(dino_layers/attention.py:60), eval-mode BatchNorm fold (heads/linear_head.py:44) and the bicubic
pos-embed resample for non-square inputs (backbones/dino_v2.py:184-215).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import _C, ops


def slide_boxes(H: int, W: int, crop: Sequence[int], stride: Sequence[int]) -> List[Tuple[int, int]]:
    """Top-left (y1, x1) of every window, row-major — the grid of mmseg slide_inference as copied in
    rein/models/segmentors/Ms_VFM_encoder_decoder.py:424-441."""
    h_crop, w_crop = crop
    h_stride, w_stride = stride
    h_grids = max(H - h_crop + h_stride - 1, 0) // h_stride + 1
    w_grids = max(W - w_crop + w_stride - 1, 0) // w_stride + 1
    out = []
    for hi in range(h_grids):
        for wi in range(w_grids):
            y2 = min(hi * h_stride + h_crop, H)
            x2 = min(wi * w_stride + w_crop, W)
            out.append((max(y2 - h_crop, 0), max(x2 - w_crop, 0)))
    return out


@dataclass
class VitSpec:
    embed_dim: int
    depth: int
    num_heads: int
    mlp_hidden: int
    patch_size: int
    out_indices: Tuple[int, ...]
    ln_eps: float = 1e-6


@dataclass
class HeadSpec:
    in_channels: int
    mid_channels: int
    groups: int
    num_classes: int
    gn_eps: float = 1e-5


def _interp_pos_embed(pos_embed: torch.Tensor, gh: int, gw: int) -> torch.Tensor:
    """pos_embed [1, 1+N, C] -> [1+gh*gw, C] with the reference's exact F.interpolate call
    (dino_v2.py:184-215; there `w` is the image height and `h` the width)."""
    N = pos_embed.shape[1] - 1
    s = int(math.sqrt(N))
    if gh * gw == N and gh == gw:
        return pos_embed[0].float().contiguous()
    pe = pos_embed.float()
    cls_pos, patch_pos = pe[:, 0], pe[:, 1:]
    dim = pe.shape[-1]
    w0, h0 = gh + 0.1, gw + 0.1
    patch_pos = F.interpolate(patch_pos.reshape(1, s, s, dim).permute(0, 3, 1, 2),
                              scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)), mode="bicubic")
    assert int(w0) == patch_pos.shape[-2] and int(h0) == patch_pos.shape[-1]
    patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(-1, dim)
    return torch.cat((cls_pos, patch_pos), dim=0).contiguous()


class PackedVit:
    """Device-resident bf16/fp32 weights of the backbone in the layout the kernels read."""

    def __init__(self, sd: Dict[str, torch.Tensor], spec: VitSpec, lora_scale: float, device):
        self.spec = spec
        self.device = device
        Cc = spec.embed_dim
        keep: List[torch.Tensor] = []

        def dev(t, dtype):
            t = t.detach().to(device=device, dtype=dtype).contiguous()
            keep.append(t)
            return t

        head_dim = Cc // spec.num_heads
        if head_dim != 64:
            raise ValueError(f"vfmseg_b200 attention kernel needs head_dim 64, got {head_dim}")
        if spec.patch_size != 16:
            raise ValueError("vfmseg_b200 patch gather is built for patch_size 16 (every reference config uses 16)")
        scale = head_dim ** -0.5
        self.patch_w = dev(sd["patch_embed.proj.weight"].reshape(Cc, -1), torch.bfloat16)
        self.patch_b = dev(sd["patch_embed.proj.bias"], torch.float32)
        self.cls_token = dev(sd["cls_token"].reshape(-1), torch.float32)
        self.pos_embed_src = sd["pos_embed"].detach().float().cpu()
        self._pos_cache: Dict[Tuple[int, int], torch.Tensor] = {}
        ones = dev(torch.ones(Cc), torch.float32)
        self.blocks = (_C.VfmBlockParams * spec.depth)()
        for i in range(spec.depth):
            p = f"blocks.{i}."
            if p + "attn.qkv.base_layer.weight" in sd:  # peft-wrapped
                w = sd[p + "attn.qkv.base_layer.weight"].float()
                b = sd.get(p + "attn.qkv.base_layer.bias")
                a_, b_ = p + "attn.qkv.lora_A.default.weight", p + "attn.qkv.lora_B.default.weight"
                if a_ in sd:
                    w = w + lora_scale * (sd[b_].float() @ sd[a_].float())
            else:
                w = sd[p + "attn.qkv.weight"].float()
                b = sd.get(p + "attn.qkv.bias")
            b = torch.zeros(3 * Cc) if b is None else b.float().clone()
            w = w.clone()
            w[:Cc] *= scale
            b[:Cc] *= scale
            blk = self.blocks[i]
            blk.ln1_w = dev(sd[p + "norm1.weight"], torch.float32).data_ptr()
            blk.ln1_b = dev(sd[p + "norm1.bias"], torch.float32).data_ptr()
            blk.qkv_w = dev(w, torch.bfloat16).data_ptr()
            blk.qkv_b = dev(b, torch.float32).data_ptr()
            blk.proj_w = dev(sd[p + "attn.proj.weight"], torch.bfloat16).data_ptr()
            blk.proj_b = dev(self._bias(sd, p + "attn.proj.bias", Cc), torch.float32).data_ptr()
            blk.ls1 = dev(sd[p + "ls1.gamma"], torch.float32).data_ptr() if p + "ls1.gamma" in sd else ones.data_ptr()
            blk.ln2_w = dev(sd[p + "norm2.weight"], torch.float32).data_ptr()
            blk.ln2_b = dev(sd[p + "norm2.bias"], torch.float32).data_ptr()
            blk.fc1_w = dev(sd[p + "mlp.fc1.weight"], torch.bfloat16).data_ptr()
            blk.fc1_b = dev(self._bias(sd, p + "mlp.fc1.bias", spec.mlp_hidden), torch.float32).data_ptr()
            blk.fc2_w = dev(sd[p + "mlp.fc2.weight"], torch.bfloat16).data_ptr()
            blk.fc2_b = dev(self._bias(sd, p + "mlp.fc2.bias", Cc), torch.float32).data_ptr()
            blk.ls2 = dev(sd[p + "ls2.gamma"], torch.float32).data_ptr() if p + "ls2.gamma" in sd else ones.data_ptr()
            if Cc % 256 == 0:
                # LayerNorm folded into the Linear that follows it (include/vfmseg_b200.h, VfmBlockParams): norm1 -> qkv, norm2 -> fc1
                from .ops import fold_layernorm
                wf, bf, cs = fold_layernorm(w, b, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
                blk.qkv_wf, blk.qkv_bf, blk.qkv_cs = dev(wf, torch.bfloat16).data_ptr(), dev(bf, torch.float32).data_ptr(), dev(cs, torch.float32).data_ptr()
                wf, bf, cs = fold_layernorm(sd[p + "mlp.fc1.weight"], self._bias(sd, p + "mlp.fc1.bias", spec.mlp_hidden),
                                            sd[p + "norm2.weight"], sd[p + "norm2.bias"])
                blk.fc1_wf, blk.fc1_bf, blk.fc1_cs = dev(wf, torch.bfloat16).data_ptr(), dev(bf, torch.float32).data_ptr(), dev(cs, torch.float32).data_ptr()
        self._keep = keep

    @staticmethod
    def _bias(sd, key, n):
        return sd[key] if key in sd else torch.zeros(n)

    def pos_embed(self, gh: int, gw: int) -> torch.Tensor:
        k = (gh, gw)
        if k not in self._pos_cache:
            self._pos_cache[k] = _interp_pos_embed(self.pos_embed_src, gh, gw).to(self.device)
        return self._pos_cache[k]

    def params(self, gh: int, gw: int) -> _C.VfmVitParams:
        s = self.spec
        p = _C.VfmVitParams()
        p.embed_dim, p.depth, p.heads, p.mlp_hidden, p.n_taps = s.embed_dim, s.depth, s.num_heads, s.mlp_hidden, len(s.out_indices)
        for i, t in enumerate(sorted(s.out_indices)):
            p.tap_blocks[i] = t
        p.ln_eps = s.ln_eps
        p.patch_w, p.patch_b = self.patch_w.data_ptr(), self.patch_b.data_ptr()
        p.cls_token = self.cls_token.data_ptr()
        p.pos_embed = self.pos_embed(gh, gw).data_ptr()
        p.blocks = C.cast(self.blocks, C.POINTER(_C.VfmBlockParams))
        return p


class PackedLinearHead:
    def __init__(self, sd: Dict[str, torch.Tensor], spec: HeadSpec, device):
        self.spec = spec
        keep: List[torch.Tensor] = []

        def dev(t, dtype):
            t = t.detach().to(device=device, dtype=dtype).contiguous()
            keep.append(t)
            return t

        mid, nc = spec.mid_channels, spec.num_classes
        if nc > 32:
            raise ValueError("vfmseg_b200 head kernels support num_classes <= 32")
        self.fusion_w = dev(sd["fusion_conv.conv.weight"].reshape(mid, -1), torch.bfloat16)
        self.gn_w = dev(sd["fusion_conv.gn.weight"], torch.float32)
        self.gn_b = dev(sd["fusion_conv.gn.bias"], torch.float32)
        # ConvT(mid -> mid/2) with eval BatchNorm folded: y = (convT(x) + b - mean) * s + beta, s = w / sqrt(var + eps)
        w1 = sd["output_upscaling.0.weight"].float()          # [mid, mid/2, 2, 2]
        b1 = sd["output_upscaling.0.bias"].float()
        s = sd["output_upscaling.1.weight"].float() / torch.sqrt(sd["output_upscaling.1.running_var"].float() + 1e-5)
        w1 = w1 * s.view(1, -1, 1, 1)
        b1 = (b1 - sd["output_upscaling.1.running_mean"].float()) * s + sd["output_upscaling.1.bias"].float()
        self.up1_w = dev(w1.permute(2, 3, 1, 0).reshape(4 * (mid // 2), mid), torch.bfloat16)  # row (dy*2+dx)*Cout + co
        self.up1_b = dev(b1.repeat(4), torch.float32)
        w2 = sd["output_upscaling.3.weight"].float()          # [mid/2, mid/4, 2, 2]
        self.up2_w = dev(w2.permute(2, 3, 1, 0).reshape(4 * (mid // 4), mid // 2), torch.bfloat16)
        self.up2_b = dev(sd["output_upscaling.3.bias"].float().repeat(4), torch.float32)
        wc = torch.zeros(32, mid // 4)
        wc[:nc] = sd["conv_seg.weight"].float().reshape(nc, -1)
        self.cls_w = dev(wc, torch.bfloat16)
        self.cls_b = dev(sd["conv_seg.bias"], torch.float32)
        self._keep = keep

    def params(self) -> _C.VfmLinearHeadParams:
        s = self.spec
        p = _C.VfmLinearHeadParams()
        p.in_channels, p.mid_channels, p.groups, p.num_classes, p.gn_eps = s.in_channels, s.mid_channels, s.groups, s.num_classes, s.gn_eps
        p.fusion_w, p.gn_w, p.gn_b = self.fusion_w.data_ptr(), self.gn_w.data_ptr(), self.gn_b.data_ptr()
        p.up1_w, p.up1_b = self.up1_w.data_ptr(), self.up1_b.data_ptr()
        p.up2_w, p.up2_b = self.up2_w.data_ptr(), self.up2_b.data_ptr()
        p.cls_w, p.cls_b = self.cls_w.data_ptr(), self.cls_b.data_ptr()
        return p


def linear_head_lowres(head: PackedLinearHead, taps: torch.Tensor, n: int, gh: int, gw: int, ws_cache: Dict[str, torch.Tensor],
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LinearHead on token-major taps (bf16 [n*gh*gw, in_channels]) -> fp32 [n, num_classes, 4gh, 4gw]."""
    lib = _C.load()
    p = head.params()
    need = lib.vfm_linear_head_workspace_bytes(C.byref(p), n, gh, gw)
    ws = ws_cache.get("head")
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=taps.device)
        ws_cache["head"] = ws
    if out is None:
        out = torch.empty(n, head.spec.num_classes, 4 * gh, 4 * gw, dtype=torch.float32, device=taps.device)
    _C.call("vfm_linear_head_forward", C.byref(p), taps.data_ptr(), n, gh, gw, out.data_ptr(), ws.data_ptr(),
            ws.numel(), torch.cuda.current_stream().cuda_stream)
    return out


class SlideEngine:
    """Runs backbone + head over batches of crop windows and merges them; one instance per device."""

    def __init__(self, vit: PackedVit, head: Optional[PackedLinearHead], max_crops_per_pass: int = 36):
        _C.check(_C.load().vfm_device_check())
        self.vit = vit
        self.head = head
        self.device = vit.device
        self.max_crops_per_pass = max_crops_per_pass
        self._ws: Dict[str, torch.Tensor] = {}
        self._tables: Dict[tuple, tuple] = {}
        self.pixel_norm: Optional[_C.VfmPixelNorm] = None

    def set_pixel_norm(self, mean, std, bgr_to_rgb=True):
        self.pixel_norm = ops.pixel_norm(mean, std, bgr_to_rgb)

    # ------------------------------------------------------------------ buffers
    def _buf(self, name: str, nbytes: int) -> torch.Tensor:
        t = self._ws.get(name)
        if t is None or t.numel() < nbytes:
            t = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws[name] = t
        return t

    def _crop_table(self, n_img: int, boxes: Sequence[Tuple[int, int]]):
        key = (n_img, tuple(boxes))
        if key not in self._tables:
            rows = [(b, y1, x1, 0) for b in range(n_img) for (y1, x1) in boxes]
            crops = torch.tensor(rows, dtype=torch.int32, device=self.device)
            bx = torch.tensor(list(boxes), dtype=torch.int32, device=self.device)
            self._tables[key] = (crops, bx)
        return self._tables[key]

    # ------------------------------------------------------------------ stages
    def backbone_taps(self, img: torch.Tensor, crops: torch.Tensor, gh: int, gw: int) -> torch.Tensor:
        """Feature taps (token-major, cls dropped) for the crop windows listed in `crops`
        (int32 [n,4] = image, y1, x1, 0): bf16 [n*gh*gw, n_taps*C]."""
        if hasattr(self.vit, "forward_taps"):      # EVA02 / SAM ViT: their own fused C drivers (vfm_eva_forward / vfm_sam_forward)
            return self.vit.forward_taps(img, crops, gh, gw, self.pixel_norm)
        lib = _C.load()
        n = crops.shape[0]
        s = self.vit.spec
        p = self.vit.params(gh, gw)
        need = lib.vfm_vit_workspace_bytes(C.byref(p), n, gh, gw)
        ws = self._buf("vit", need)
        taps = torch.empty(n * gh * gw, len(s.out_indices) * s.embed_dim, dtype=torch.bfloat16, device=self.device)
        is_u8 = img.dtype == torch.uint8
        if is_u8 and self.pixel_norm is None:
            raise _C.VfmError("uint8 input needs set_pixel_norm() (SegDataPreProcessor mean/std)")
        if not is_u8 and img.dtype != torch.float32:
            raise _C.VfmError(f"input must be uint8 or float32, got {img.dtype}")
        assert img.is_contiguous() and img.dim() == 4 and img.shape[1] == 3
        _C.call("vfm_vit_forward", C.byref(p), img.data_ptr(), int(is_u8),
                C.byref(self.pixel_norm) if self.pixel_norm is not None else None, img.shape[2], img.shape[3],
                crops.data_ptr(), n, gh, gw, taps.data_ptr(), ws.data_ptr(), ws.numel(),
                torch.cuda.current_stream().cuda_stream)
        return taps

    def head_lowres(self, taps: torch.Tensor, n: int, gh: int, gw: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return linear_head_lowres(self.head, taps, n, gh, gw, self._ws, out)

    def crops_lowres(self, img: torch.Tensor, crops: torch.Tensor, crop_hw: Tuple[int, int]) -> torch.Tensor:
        """Low-res logits of every listed window, processed in passes of <= max_crops_per_pass."""
        ps = self.vit.spec.patch_size
        gh, gw = crop_hw[0] // ps, crop_hw[1] // ps
        n = crops.shape[0]
        out = torch.empty(n, self.head.spec.num_classes, 4 * gh, 4 * gw, dtype=torch.float32, device=self.device)
        for s0 in range(0, n, self.max_crops_per_pass):
            s1 = min(s0 + self.max_crops_per_pass, n)
            taps = self.backbone_taps(img, crops[s0:s1], gh, gw)
            self.head_lowres(taps, s1 - s0, gh, gw, out=out[s0:s1])
        return out

    # ------------------------------------------------------------------ public paths
    def slide(self, img: torch.Tensor, crop_size, stride, want_logits: bool = False):
        """Slide inference over [B,3,H,W] (uint8 raw or fp32 normalised): uint8 labels [B,H,W] and,
        on request, the fp32 merged logits [B,nc,H,W] the reference's slide_inference returns."""
        B, _, H, W = img.shape
        ps = self.vit.spec.patch_size
        if crop_size[0] % ps or crop_size[1] % ps:
            raise _C.VfmError("crop_size must be a multiple of the patch size")
        if H < crop_size[0] or W < crop_size[1]:
            raise _C.VfmError(f"image {H}x{W} smaller than the crop window {tuple(crop_size)}")
        boxes = slide_boxes(H, W, crop_size, stride)
        crops, bx = self._crop_table(B, boxes)
        low = self.crops_lowres(img, crops, tuple(crop_size))
        labels, logits = ops.slide_merge_argmax(low, bx, B, tuple(crop_size), (H, W), want_logits=want_logits)
        return labels, logits, low

    def slide_flip_tta(self, img: torch.Tensor, crop_size, stride, want_logits: bool = False):
        """Slide inference with the horizontal-flip test-time augmentation (hrda_encoder_decoder.py:196-229, scales = [1]):
        (slide(img) + flip(slide(flip(img)))) / 2 -> argmax. The first pass materialises its merged logits; the second pass's
        merge kernel reads them mirrored and writes the labels (and, on request, the averaged logits over the first pass's
        buffer), so the second logit volume and the combining pass over both never exist."""
        B, _, H, W = img.shape
        _, a, _ = self.slide(img, crop_size, stride, want_logits=True)
        boxes = slide_boxes(H, W, crop_size, stride)
        crops, bx = self._crop_table(B, boxes)
        low = self.crops_lowres(torch.flip(img, [3]), crops, tuple(crop_size))
        if ops.flip_merge_supported(low, tuple(crop_size), W):
            return ops.slide_merge_flip_argmax(low, bx, B, tuple(crop_size), a, want_logits=want_logits)
        _, b = ops.slide_merge_argmax(low, bx, B, tuple(crop_size), (H, W), want_logits=True)
        return ops.tta_flip_mean_argmax(a, b, want_logits=want_logits)

    def whole(self, img: torch.Tensor, want_logits: bool = False):
        """Whole-image inference (mmseg whole_inference): one window covering the image."""
        B, _, H, W = img.shape
        return self.slide(img, (H, W), (H, W), want_logits=want_logits)

    def confusion(self, cm: torch.Tensor, labels: torch.Tensor, gt: torch.Tensor, ignore_index: int = 255):
        return ops.confusion_matrix_(cm, labels.reshape(-1), gt.reshape(-1), self.head.spec.num_classes, ignore_index)
