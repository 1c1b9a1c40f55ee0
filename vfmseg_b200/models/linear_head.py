"""`LinearHead` — registered decode head with the reference's kwargs, parameter names and forward
contract (rein/models/heads/linear_head.py:13-70; mmseg BaseDecodeHead attributes), executed by the
sm_100a engine. nn.Modules are parameter containers; see engine.PackedLinearHead for the folds."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from ..engine import HeadSpec, PackedLinearHead, linear_head_lowres
from ..registry import MODELS


class _ConvModule(nn.Module):  # mmcv ConvModule naming: .conv (no bias when a norm follows), .gn
    def __init__(self, cin, cout, groups):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size=1, bias=False)
        self.gn = nn.GroupNorm(groups, cout)


@MODELS.register_module()
class LinearHead(nn.Module):
    def __init__(self, interpolate_mode="bilinear", *, in_channels, channels, num_classes, in_index=-1, out_channels=None,
                 dropout_ratio=0.1, norm_cfg=None, act_cfg=dict(type="ReLU"), align_corners=False, loss_decode=None,
                 ignore_index=255, input_transform="multiple_select", init_cfg=None, **unused):
        super().__init__()
        if not isinstance(in_channels, (list, tuple)):
            raise TypeError("LinearHead expects a list of in_channels (multiple_select)")
        if norm_cfg is None or norm_cfg.get("type") != "GN":
            raise NotImplementedError("LinearHead on this path is built with norm_cfg=dict(type='GN', num_groups=...)")
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
        self._channels = c = self.in_channels[0]
        if channels != c // 4:
            raise ValueError(f"conv_seg consumes {c // 4} channels (two halving transposed convs); got channels={channels}")
        self.fusion_conv = _ConvModule(c * len(self.in_channels), c, norm_cfg["num_groups"])
        self.output_upscaling = nn.Sequential(
            nn.ConvTranspose2d(c, c // 2, kernel_size=2, stride=2),
            nn.BatchNorm2d(c // 2),      # SyncBatchNorm in the reference; same parameters/buffers, eval semantics
            nn.GELU(),
            nn.ConvTranspose2d(c // 2, c // 4, kernel_size=2, stride=2),
            nn.GELU(),
        )
        self.conv_seg = nn.Conv2d(channels, self.out_channels, kernel_size=1)
        self._packed: Optional[PackedLinearHead] = None
        self._ws = {}
        self.register_load_state_dict_post_hook(lambda m, _k: m.invalidate())

    def invalidate(self):
        self._packed = None

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def spec(self) -> HeadSpec:
        return HeadSpec(sum(self.in_channels), self._channels, self.norm_cfg["num_groups"], self.out_channels)

    def packed(self) -> PackedLinearHead:
        dev = self.conv_seg.weight.device
        if dev.type != "cuda":
            raise RuntimeError("vfmseg_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        if self._packed is None:
            self._packed = PackedLinearHead(dict(self.state_dict()), self.spec(), dev)
        return self._packed

    def _transform_inputs(self, inputs):
        return [inputs[i] for i in self.in_index]

    def forward(self, inputs: Sequence[torch.Tensor]) -> torch.Tensor:
        """list of [B,C,h,w] maps -> [B, num_classes, 4h, 4w] fp32 logits."""
        xs = self._transform_inputs(inputs)
        B, _, h, w = xs[0].shape
        taps = torch.cat([x.permute(0, 2, 3, 1) for x in xs], dim=-1).reshape(B * h * w, -1).to(torch.bfloat16).contiguous()
        return linear_head_lowres(self.packed(), taps, B, h, w, self._ws)

    def predict(self, inputs, batch_img_metas, test_cfg=None):
        from ..ops import resize_bilinear
        low = self.forward(inputs)
        m = batch_img_metas[0]
        size = m["img_shape"] if isinstance(m["img_shape"], torch.Size) else m.get("pad_shape", m["img_shape"])[:2]
        return resize_bilinear(low, tuple(size))
