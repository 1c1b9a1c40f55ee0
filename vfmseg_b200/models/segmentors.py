"""`LoraBackboneEncoderDecoder` — registered segmentor mirroring
rein/models/segmentors/Lora_encoder_decoder.py:12-36 on top of mmseg's EncoderDecoder contract
(predict / inference / slide_inference / whole_inference / encode_decode / extract_feat;
SURVEY.md Appendix A1-A3), with the whole slide loop executed by the sm_100a engine:

  reference, per image : 18 x (crop slice -> backbone -> head -> resize -> F.pad add) -> divide -> argmax
  here                 : one gather of all windows -> batched backbone/head -> one merge+argmax kernel
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from ..engine import PackedLinearHead, PackedVit, SlideEngine
from ..registry import MODELS, ConfigDict
from ..structures import PixelData, SegDataSample
from .lora import LoraConfig, get_peft_model


@MODELS.register_module()
class SegDataPreProcessor(nn.Module):
    """mmseg SegDataPreProcessor (test mode) as configured at configs/_base_/models/lora_dinov2_linear.py:13-21.
    Stacks the uint8 images and moves them to the device; the channel flip and (x-mean)/std are applied
    inside the patch-gather kernel, so no normalised fp32 image is ever materialised."""

    def __init__(self, mean=None, std=None, size=None, size_divisor=None, pad_val=0, seg_pad_val=255, bgr_to_rgb=False,
                 rgb_to_bgr=False, batch_augments=None, test_cfg=None, **unused):
        super().__init__()
        assert not (bgr_to_rgb and rgb_to_bgr)
        self.mean = list(mean) if mean is not None else [0.0, 0.0, 0.0]
        self.std = list(std) if std is not None else [1.0, 1.0, 1.0]
        self.channel_conversion = bool(bgr_to_rgb or rgb_to_bgr)
        self.register_buffer("_dummy", torch.zeros(0), persistent=False)  # tracks the module's device

    def forward(self, data: dict, training: bool = False) -> dict:
        assert not training, "vfmseg_b200 covers the inference path only"
        inputs = data["inputs"]
        if isinstance(inputs, (list, tuple)):
            inputs = torch.stack(list(inputs), dim=0)
        if inputs.dtype != torch.uint8:
            # mmseg normalises whatever dtype arrives; here mean / std / channel flip live in the patch-gather kernel and are
            # applied to uint8 pixels only, so a float image would silently skip them
            raise TypeError(f"SegDataPreProcessor expects uint8 images (got {inputs.dtype}); already-normalised fp32 tensors go to "
                            "predict() / inference() directly")
        return dict(inputs=inputs.to(self._dummy.device, non_blocking=True).contiguous(), data_samples=data.get("data_samples"))


@MODELS.register_module()
class LoraBackboneEncoderDecoder(nn.Module):
    def __init__(self, checkpoint=None, Lora_config=None, *, backbone, decode_head, neck=None, auxiliary_head=None,
                 train_cfg=None, test_cfg=None, data_preprocessor=None, pretrained=None, init_cfg=None,
                 max_crops_per_pass: int = 36):
        super().__init__()
        if neck is not None or auxiliary_head is not None:
            raise NotImplementedError("neck / auxiliary_head are not part of the LoRA slide-inference configs")
        self.backbone = MODELS.build(backbone)
        self.decode_head = MODELS.build(decode_head)
        self.align_corners = self.decode_head.align_corners
        self.num_classes = self.decode_head.num_classes
        self.out_channels = self.decode_head.out_channels
        self.train_cfg = ConfigDict(train_cfg or {})
        self.test_cfg = ConfigDict(test_cfg or {})
        self.data_preprocessor = MODELS.build(data_preprocessor) if isinstance(data_preprocessor, dict) else data_preprocessor
        self.max_crops_per_pass = max_crops_per_pass
        if Lora_config is not None:
            self.Lora_config = LoraConfig(r=Lora_config["r"], lora_alpha=Lora_config["lora_alpha"],
                                          target_modules=Lora_config["target_modules"],
                                          lora_dropout=Lora_config.get("lora_dropout", 0.0), bias="none")
            self.backbone = get_peft_model(self.backbone, self.Lora_config)   # Lora_encoder_decoder.py:24
            if checkpoint is not None:
                self.load_pretrained_backbone(checkpoint, Lora_config["target_modules"])
        self._engine: Optional[SlideEngine] = None
        self.register_load_state_dict_post_hook(lambda m, _k: m.invalidate())

    # ------------------------------------------------------------------ weights
    def load_pretrained_backbone(self, checkpoint, target_modules):
        """Lora_encoder_decoder.py:28-36: plain backbone checkpoint, target modules renamed to '<t>.base_layer'."""
        original = torch.load(checkpoint, map_location="cpu") if isinstance(checkpoint, str) else checkpoint
        new = {}
        for name, weight in original.items():
            for t in target_modules:
                if t in name:
                    name = name.replace(t, t + ".base_layer")
                new[name] = weight
        self.inner_backbone.load_state_dict(new, strict=False)
        self.invalidate()

    @property
    def inner_backbone(self):
        b = self.backbone
        if hasattr(b, "model") and hasattr(b.model, "base_model"):   # LoRABackbone (lora_backbone.py:23)
            return b.model.base_model.model
        return b.base_model.model if hasattr(b, "base_model") else b

    def invalidate(self):
        self._engine = None

    def _weights_version(self) -> int:
        """Changes whenever any parameter or buffer is written in place (load_state_dict on a SUBMODULE, optimizer steps,
        manual edits): the packed bf16 copies inside the cached engine must not outlive the weights they were made from."""
        v = 0
        for t in self.parameters():
            v += t._version
        for t in self.buffers():
            v += t._version
        return v

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def engine(self) -> SlideEngine:
        ver = self._weights_version()
        if self._engine is not None and getattr(self, "_engine_version", None) != ver:
            self._engine = None
            for m in self.modules():          # per-module packed caches (backbone.packed(), decode_head.packed(), ...)
                if m is not self and hasattr(m, "invalidate"):
                    m.invalidate()
        if self._engine is None:
            n_taps = len(getattr(self.inner_backbone, "out_indices", [])) or None
            in_index = getattr(self.decode_head, "in_index", None)
            if in_index is not None and n_taps is not None and list(in_index) != list(range(n_taps)):
                # the fused path feeds the backbone taps to the fusion conv in out_indices order; mmseg's
                # BaseDecodeHead._transform_inputs would select / permute them by in_index
                raise NotImplementedError(f"decode_head.in_index={list(in_index)} must be {list(range(n_taps))} for the fused slide path")
            self._engine_version = ver
            bb = self.inner_backbone
            dev = next(bb.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("vfmseg_b200 runs on CUDA (sm_100a) only: call .cuda(); there is no CPU path")
            eng = SlideEngine(bb.packed(dev), self.decode_head.packed(), self.max_crops_per_pass)
            if self.data_preprocessor is not None:
                eng.set_pixel_norm(self.data_preprocessor.mean, self.data_preprocessor.std, self.data_preprocessor.channel_conversion)
            self._engine = eng
        return self._engine

    # ------------------------------------------------------------------ mmseg EncoderDecoder contract
    def extract_feat(self, inputs: torch.Tensor) -> List[torch.Tensor]:
        return self.backbone(inputs)

    def encode_decode(self, inputs: torch.Tensor, batch_img_metas=None) -> torch.Tensor:
        """Whole-window forward: logits resized to the input size."""
        _, logits, _ = self.engine().whole(self._as_input(inputs), want_logits=True)
        return logits

    def _slide(self, x: torch.Tensor, want_logits: bool):
        """Slide inference, optionally with the horizontal-flip test-time augmentation of
        rein/models/segmentors/hrda_encoder_decoder.py:196-229 (`test_cfg.test_time_aug` and `test_cfg.flip`, :114-115;
        scales = [1]): logits = (slide(img) + flip(slide(flip(img)))) / 2; the second pass's merge kernel reads the first pass's logits
        mirrored and arg-maxes the mean (SlideEngine.slide_flip_tta), so only one logit volume is ever materialised."""
        eng = self.engine()
        crop, stride = self.test_cfg.crop_size, self.test_cfg.stride
        if not (self.test_cfg.get("test_time_aug", False) and self.test_cfg.get("flip", False)):
            labels, logits, _ = eng.slide(x, crop, stride, want_logits=want_logits)
            return labels, logits
        return eng.slide_flip_tta(x, crop, stride, want_logits=want_logits)

    def slide_inference(self, inputs: torch.Tensor, batch_img_metas=None) -> torch.Tensor:
        return self._slide(self._as_input(inputs), True)[1]

    def whole_inference(self, inputs: torch.Tensor, batch_img_metas=None) -> torch.Tensor:
        return self.encode_decode(inputs, batch_img_metas)

    def inference(self, inputs: torch.Tensor, batch_img_metas=None) -> torch.Tensor:
        mode = self.test_cfg.get("mode", "whole")
        assert mode in ("slide", "whole"), f'Only "slide" or "whole" test mode are supported, but got {mode}.'
        return self.slide_inference(inputs, batch_img_metas) if mode == "slide" else self.whole_inference(inputs, batch_img_metas)

    @staticmethod
    def _as_input(inputs: torch.Tensor) -> torch.Tensor:
        if inputs.dtype not in (torch.uint8, torch.float32):
            inputs = inputs.float()
        return inputs.contiguous()

    def predict_labels(self, inputs: torch.Tensor, want_logits: bool = False):
        """Throughput path: uint8 label maps [B,H,W] (argmax fused into the merge kernel); the 159 MB/image
        fp32 logits are only produced when asked for."""
        mode = self.test_cfg.get("mode", "whole")
        eng = self.engine()
        x = self._as_input(inputs)
        if mode == "slide":
            labels, logits = self._slide(x, want_logits)
        else:
            labels, logits, _ = eng.whole(x, want_logits=want_logits)
        return labels, logits

    def postprocess_result(self, seg_logits: torch.Tensor, data_samples: Optional[Sequence[SegDataSample]] = None,
                           labels: Optional[torch.Tensor] = None) -> List[SegDataSample]:
        """mmseg BaseSegmentor.postprocess_result for C > 1: per sample strip `img_padding_size` /
        `padding_size` (left, right, top, bottom), undo a test-time flip, resize the logits to `ori_shape` (bilinear,
        align_corners of the head) and arg-max. The resize + argmax is one kernel (ops.resize_argmax); samples whose
        `ori_shape` equals the network size — Cityscapes — keep the label map the merge kernel already produced.
        The reference's cross-domain test pipelines need the resize: BDD100K 1280x720 is fed as 1820x1024
        (configs/_base_/datasets/bdd100k_1024x1024.py:15), Mapillary likewise."""
        from .. import ops
        B, C, H, W = seg_logits.shape
        if data_samples is None:
            data_samples = [SegDataSample(metainfo=dict(ori_shape=(H, W), img_shape=(H, W))) for _ in range(B)]
        if self.align_corners:
            raise NotImplementedError("postprocess_result: align_corners=True heads are not part of the reference configs")
        for i, s in enumerate(data_samples):
            meta = getattr(s, "metainfo", {}) or {}
            pad = meta.get("img_padding_size", meta.get("padding_size", [0, 0, 0, 0]))
            left, right, top, bottom = [int(v) for v in pad]
            ori = tuple(int(v) for v in tuple(meta.get("ori_shape", (H, W)))[:2])
            flip = meta.get("flip", None)
            if not (left or right or top or bottom or flip) and ori == (H, W):
                lg = seg_logits[i]
                lab = labels[i:i + 1] if labels is not None else lg.argmax(0, keepdim=True)
            else:
                lg = seg_logits[i:i + 1, :, top:H - bottom, left:W - right]
                if flip:
                    direction = meta.get("flip_direction", None)
                    assert direction in ("horizontal", "vertical")
                    lg = lg.flip(dims=(3,)) if direction == "horizontal" else lg.flip(dims=(2,))
                if tuple(lg.shape[2:]) == ori:
                    lg = lg[0]
                    lab = lg.argmax(0, keepdim=True)
                else:
                    lab, lg = ops.resize_argmax(lg, ori)
                    lg = lg[0]
            s.set_data({"seg_logits": PixelData(data=lg), "pred_sem_seg": PixelData(data=lab.long())})
        return list(data_samples)

    def predict(self, inputs: torch.Tensor, data_samples: Optional[Sequence[SegDataSample]] = None) -> List[SegDataSample]:
        """mmseg BaseSegmentor.predict: inference at the network size, then postprocess_result per sample."""
        labels, logits = self.predict_labels(inputs, want_logits=True)
        return self.postprocess_result(logits, data_samples, labels=labels)

    def forward(self, inputs, data_samples=None, mode="predict"):
        if mode == "predict":
            return self.predict(inputs, data_samples)
        if mode == "tensor":
            return self.inference(inputs, None)
        raise NotImplementedError("mode='loss' (training) is out of scope for vfmseg_b200")

    def test_step(self, data: dict):
        data = self.data_preprocessor(data, False) if self.data_preprocessor is not None else data
        return self.predict(data["inputs"], data.get("data_samples"))


@MODELS.register_module()
class EncoderDecoder(LoraBackboneEncoderDecoder):
    """mmseg's plain EncoderDecoder type name, as used by the LoRABackbone configs (configs/_base_/models/
    lora_eva02_linear.py:4, lora_sam_*.py): the LoRA wrapping lives in the backbone entry, not in the segmentor."""

    def __init__(self, backbone, decode_head, neck=None, auxiliary_head=None, train_cfg=None, test_cfg=None,
                 data_preprocessor=None, pretrained=None, init_cfg=None, max_crops_per_pass: int = 36):
        super().__init__(checkpoint=None, Lora_config=None, backbone=backbone, decode_head=decode_head, neck=neck,
                         auxiliary_head=auxiliary_head, train_cfg=train_cfg, test_cfg=test_cfg,
                         data_preprocessor=data_preprocessor, pretrained=pretrained, init_cfg=init_cfg,
                         max_crops_per_pass=max_crops_per_pass)


@MODELS.register_module()
class MsVFMEncoderDecoder(LoraBackboneEncoderDecoder):
    """rein/models/segmentors/Ms_VFM_encoder_decoder.py:62-130 (constructor) and :278-332, :400-466 (inference):
    coarse whole-image pass + confidence-gated per-window refinement by the aux decoder (VFMHead).
    Test modes: 'ms_slide_inference' (the shipped config, configs/_base_/models/lora_dinov2_ms_masked.py:79-86),
    'hr_slide_inference' (= plain slide) and 'lr_slide_inference' (slide at half resolution, logits upsampled x2).
    'msfull_slide_inference' leaves the decoder's random feature masking on at test time (:296-329 never clears
    mask_enable), so it has no deterministic result to match and is not implemented."""

    MODES = ("lr_slide_inference", "hr_slide_inference", "msfull_slide_inference", "ms_slide_inference")

    def __init__(self, backbone, decode_head, aux_head, neck=None, auxiliary_head=None, train_cfg=None, test_cfg=None,
                 pretrained=None, init_cfg=None, scales=[1], hr_crop_size=None, crop_coord_divisible=1, feature_scale=1,
                 data_preprocessor=None, debug=False, debug_interval=100, detail_loss=1.0, max_crops_per_pass: int = 36):
        super().__init__(checkpoint=None, Lora_config=None, backbone=backbone, decode_head=decode_head, neck=neck,
                         auxiliary_head=auxiliary_head, train_cfg=train_cfg, test_cfg=test_cfg,
                         data_preprocessor=data_preprocessor, pretrained=pretrained, init_cfg=init_cfg,
                         max_crops_per_pass=max_crops_per_pass)
        self.scales = sorted(scales)
        self.feature_scale = feature_scale
        self.crop_size = hr_crop_size
        self.crop_coord_divisible = crop_coord_divisible
        self.debug = debug
        self.debug_interval = (self.train_cfg.get("log_config") or {}).get("img_interval", debug_interval)   # :104
        self.aux_decoder = MODELS.build(aux_head)                                                              # :112
        self.detail_loss = detail_loss

    # ------------------------------------------------------------------ inference
    def _ms(self, inputs: torch.Tensor, want_logits: bool, gate: str):
        from ..vfm_refine import ms_slide
        dec = self.aux_decoder.transformer_decoder
        had = getattr(dec, "mask_enable", None)
        if had is not None:
            dec.mask_enable = False          # :422-423
        try:
            return ms_slide(self.engine(), self.aux_decoder.packed(), self._as_input(inputs), self.test_cfg.crop_size,
                            self.test_cfg.stride, threshold=self.test_cfg.get("threadshod", 1.0),
                            conf=self.test_cfg.get("conf", 1.0), lr_size=(512, 1024), want_logits=want_logits, gate=gate)
        finally:
            if had is not None:
                dec.mask_enable = True       # :463-464

    def ms_inference(self, inputs: torch.Tensor, batch_img_metas=None) -> torch.Tensor:
        return self._ms(inputs, True, "batch")[1]

    def _lr_slide(self, inputs: torch.Tensor, want_logits: bool):
        from .. import ops
        eng = self.engine()
        x = self._as_input(inputs)
        B, _, H, W = x.shape
        h, w = H // 2, W // 2      # resize(scale_factor=0.5), :283
        lr = ops.image_resize_norm(x, (h, w), eng.pixel_norm if x.dtype == torch.uint8 else None)
        _, lr_logits, _ = eng.slide(lr, self.test_cfg.crop_size, self.test_cfg.stride, want_logits=True)
        boxes = torch.zeros(1, 2, dtype=torch.int32, device=x.device)   # resize(scale_factor=2) + argmax, :285
        return ops.slide_merge_argmax(lr_logits, boxes, B, (2 * h, 2 * w), (2 * h, 2 * w), want_logits=want_logits)

    def inference(self, inputs: torch.Tensor, batch_img_metas=None) -> torch.Tensor:
        return self._run(inputs, True, "batch")[1]

    def _run(self, inputs, want_logits: bool, gate: str):
        mode = self.test_cfg.get("mode", "lr_slide_inference")
        assert mode in self.MODES, mode
        if mode == "ms_slide_inference":
            labels, logits, _ = self._ms(inputs, want_logits, gate)
            return labels, logits
        if mode == "hr_slide_inference":
            labels, logits, _ = self.engine().slide(self._as_input(inputs), self.test_cfg.crop_size, self.test_cfg.stride,
                                                    want_logits=want_logits)
            return labels, logits
        if mode == "lr_slide_inference":
            return self._lr_slide(inputs, want_logits)
        raise NotImplementedError("msfull_slide_inference keeps the decoder's random feature mask on at test time")

    def predict_labels(self, inputs: torch.Tensor, want_logits: bool = False):
        """Throughput path; every image gates its own windows (the reference's batch_size=1 test loop)."""
        return self._run(inputs, want_logits, "image")
