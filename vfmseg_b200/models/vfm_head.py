"""`VFMHead` ("MGRNet" refinement head) and its `TransformerDecoder` / `MaskTransformerDecoder` — registered with the
reference's constructor kwargs, parameter names and forward contracts (rein/models/heads/VFMHead.py:12-89,
rein/models/heads/Transformer.py:62-79,94-136,158-177,228-283), executed by the sm_100a engine (vfm_refine.py).
The nn.Modules are parameter containers whose names make reference checkpoints load unchanged."""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn

from ..registry import MODELS
from ..vfm_refine import PackedVfmHead, VfmHeadSpec, vfm_head_forward


class _GEGLU(nn.Module):  # Transformer.py:52-59
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class _FeedForward(nn.Module):  # Transformer.py:62-79 with glu=True: net = [GEGLU, Dropout, Linear]
    def __init__(self, dim, mult=4):
        super().__init__()
        inner = int(dim * mult)
        self.net = nn.Sequential(_GEGLU(dim, inner), nn.Identity(), nn.Linear(inner, dim))


class _CrossAttention(nn.Module):  # Transformer.py:94-111
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64):
        super().__init__()
        inner = dim_head * heads
        context_dim = query_dim if context_dim is None else context_dim
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Identity())


class _BasicTransformerBlock(nn.Module):  # Transformer.py:158-177
    def __init__(self, query_dim, n_heads, d_head, context_dim=None):
        super().__init__()
        self.attn1 = _CrossAttention(query_dim, None, n_heads, d_head)
        self.ff = _FeedForward(query_dim)
        self.attn2 = _CrossAttention(query_dim, context_dim, n_heads, d_head)
        self.norm1 = nn.LayerNorm(query_dim)
        self.norm2 = nn.LayerNorm(query_dim)
        self.norm3 = nn.LayerNorm(query_dim)


@MODELS.register_module()
class TransformerDecoder(nn.Module):
    """Transformer.py:228-252. forward(query, img_feats): `query` is normalised (GroupNorm eps 1e-6) and refined
    against `img_feats` as cross-attention context; executed inside VFMHead on the device."""

    def __init__(self, query_dim, img_feat_dim, n_heads, d_head, depth=1, dropout=0.0):
        super().__init__()
        self.in_channels = query_dim
        self.n_heads, self.d_head, self.depth = n_heads, d_head, depth
        self.norm = nn.GroupNorm(32, query_dim, eps=1e-6, affine=True)
        self.transformer_blocks = nn.ModuleList(
            [_BasicTransformerBlock(query_dim, n_heads, d_head, context_dim=img_feat_dim) for _ in range(depth)])


@MODELS.register_module()
class MaskTransformerDecoder(TransformerDecoder):
    """Transformer.py:254-283. `mask_enable` replaces a random 20 % of the query features by `mask_token`
    (a training-time augmentation, :263-268); ms_inference switches it off (Ms_VFM_encoder_decoder.py:422-423),
    which is the only state this implementation executes."""

    def __init__(self, mask_ratio, **kwargs):
        super().__init__(**kwargs)
        self.mask_ratio = mask_ratio
        self.mask_token = nn.Parameter(torch.randn(1, self.in_channels, 1, 1))
        self.mask_enable = True


@MODELS.register_module()
class VFMHead(nn.Module):
    def __init__(self, transformer, interpolate_mode="bilinear", *, in_channels, channels, num_classes, in_index=-1,
                 out_channels=None, dropout_ratio=0.1, norm_cfg=None, act_cfg=dict(type="ReLU"), align_corners=False,
                 loss_decode=None, ignore_index=255, input_transform="multiple_select", init_cfg=None, **unused):
        super().__init__()
        if not isinstance(in_channels, (list, tuple)):
            raise TypeError("VFMHead expects a list of in_channels (multiple_select)")
        self.in_channels = list(in_channels)
        self.in_index = list(in_index)
        assert len(self.in_channels) == len(self.in_index)
        self.channels = channels
        self.num_classes = num_classes
        self.out_channels = out_channels or num_classes
        self.dropout_ratio = dropout_ratio
        self.norm_cfg = norm_cfg
        self.align_corners = align_corners
        self.ignore_index = ignore_index
        self.interpolate_mode = interpolate_mode
        transformer = dict(transformer)
        transformer["img_feat_dim"] = channels                # VFMHead.py:22
        self.query_dim = transformer["query_dim"]
        if self.query_dim != channels:
            raise ValueError("VFMHead: the decoder refines the fused image features, so query_dim must equal channels")
        c = channels
        self.fuse_conv = nn.Sequential(nn.Conv2d(self.in_channels[0] * len(self.in_channels), c, kernel_size=1),
                                       nn.GroupNorm(32, c), nn.GELU())
        self.seg_logits_embed = nn.Sequential(
            nn.Conv2d(self.out_channels, c // 4, kernel_size=2, stride=2), nn.GroupNorm(32, c // 4), nn.GELU(),
            nn.Conv2d(c // 4, c // 2, kernel_size=2, stride=2), nn.GroupNorm(32, c // 2), nn.GELU(),
            nn.Conv2d(c // 2, c, kernel_size=1), nn.GroupNorm(32, c))
        self.transformer_decoder = MODELS.build(transformer)
        self.conv_seg = nn.Conv2d(channels, self.out_channels, kernel_size=1)
        self._packed: Optional[PackedVfmHead] = None
        self.register_load_state_dict_post_hook(lambda m, _k: m.invalidate())

    def invalidate(self):
        self._packed = None

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def spec(self) -> VfmHeadSpec:
        t = self.transformer_decoder
        return VfmHeadSpec(sum(self.in_channels), self.channels, self.out_channels, t.n_heads, t.d_head, t.depth)

    def packed(self) -> PackedVfmHead:
        dev = self.conv_seg.weight.device
        if dev.type != "cuda":
            raise RuntimeError("vfmseg_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        if self._packed is None:
            self._packed = PackedVfmHead(dict(self.state_dict()), self.spec(), dev)
        return self._packed

    def _transform_inputs(self, inputs):
        return [inputs[i] for i in self.in_index]

    def forward(self, inputs: Sequence[torch.Tensor], seg_logits: torch.Tensor, query=None) -> torch.Tensor:
        """VFMHead.py:61-89: list of [B,C,h,w] feature maps + context logits [B,nc,Hc,Wc] -> [B, nc, h, w] fp32."""
        from .. import ops
        from ..vfm_refine import context_tokens
        if getattr(self.transformer_decoder, "mask_enable", False):
            raise NotImplementedError("mask_enable=True draws a random feature mask (training augmentation); "
                                      "ms_inference runs with mask_enable=False (Ms_VFM_encoder_decoder.py:422-423)")
        xs = self._transform_inputs(inputs)
        B, _, h, w = xs[0].shape
        taps = torch.cat([x.permute(0, 2, 3, 1) for x in xs], dim=-1).reshape(B * h * w, -1).to(torch.bfloat16).contiguous()
        head = self.packed()
        # the context is given at window resolution: treat it as "coarse logits" of a window-sized image
        Hc, Wc = seg_logits.shape[2:]
        crops = torch.tensor([(b, 0, 0, 0) for b in range(B)], dtype=torch.int32, device=taps.device)
        ctx = context_tokens(head, seg_logits.float().contiguous(), crops, (Hc, Wc), (Hc, Wc), h, w)
        return vfm_head_forward(head, taps, ctx, B, h, w)
