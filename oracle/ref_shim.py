"""ORACLE — TEST INFRASTRUCTURE ONLY.

Imports the reference's OWN source files, unmodified, from /root/reference so they can be run on
CPU as the parity oracle and golden-vector generator. The reference depends on packages that are
not installed in this image (mmseg, mmengine, mmcv, peft, timm, xformers, matplotlib); this module
installs minimal stand-ins for exactly the symbols the hot-path files import:

  * arithmetic-carrying stand-ins are real restatements of the pinned third-party versions
    (mmsegmentation 1.2.2 EncoderDecoder / BaseDecodeHead / IoUMetric, mmcv 2.1.0 ConvModule,
    peft 0.10.0 LoRA Linear) — SURVEY.md Appendix A; marked [3P] below;
  * everything else is inert.

/root/reference does not exist on the GPU box: only oracle/make_golden.py and the CPU tests that
are skipped when the tree is absent use this module.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import math
import os
import sys
import types
from collections import OrderedDict
from pathlib import Path

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = Path(os.environ.get("VFMSEG_REFERENCE", "/root/reference"))


def available() -> bool:
    return (REFERENCE_ROOT / "rein" / "models" / "backbones" / "dino_v2.py").exists()


# ----------------------------------------------------------------------------- registry [3P mmengine]
class Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg, **default_args):
        if isinstance(cfg, nn.Module):
            return cfg
        cfg = dict(cfg)
        for k, v in default_args.items():
            cfg.setdefault(k, v)
        typ = cfg.pop("type")
        cls = self.module_dict[typ] if isinstance(typ, str) else typ
        return cls(**cfg)


MODELS = Registry("model")
METRICS = Registry("metric")


class ConfigDict(dict):
    """Attribute access like mmengine ConfigDict (test_cfg.stride etc.)."""
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


# ----------------------------------------------------------------------------- mmengine stand-ins
class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg

    def init_weights(self):
        pass


class _Logger:
    def info(self, *a, **k):
        pass
    warning = error = debug = info

    @classmethod
    def get_current_instance(cls):
        return cls()

    @classmethod
    def get_instance(cls, *a, **k):
        return cls()


def print_log(*a, **k):
    pass


# ----------------------------------------------------------------------------- mmcv ConvModule [3P 2.1.0]
def build_norm_layer(cfg, num_features, postfix=""):
    cfg = dict(cfg)
    typ = cfg.pop("type")
    cfg.pop("requires_grad", None)
    if typ == "GN":
        return "gn" + str(postfix), nn.GroupNorm(num_channels=num_features, **cfg)
    if typ in ("BN", "BN2d"):
        return "bn" + str(postfix), nn.BatchNorm2d(num_features, **cfg)
    if typ == "SyncBN":
        return "bn" + str(postfix), nn.SyncBatchNorm(num_features, **cfg)
    if typ == "LN":
        return "ln" + str(postfix), nn.LayerNorm(num_features, **cfg)
    raise KeyError(typ)


class ConvModule(nn.Module):
    """conv -> norm -> act; bias='auto' => no conv bias when a norm follows; default act ReLU."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias="auto", conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU"), inplace=True, **kw):
        super().__init__()
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == "auto":
            bias = not self.with_norm
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        if self.with_norm:
            self.norm_name, norm = build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        if self.with_activation:
            t = act_cfg["type"]
            self.activate = {"ReLU": nn.ReLU(inplace=inplace), "GELU": nn.GELU()}[t]

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = getattr(self, self.norm_name)(x)
        if self.with_activation:
            x = self.activate(x)
        return x


# ----------------------------------------------------------------------------- mmseg stand-ins [3P 1.2.2]
def resize(input, size=None, scale_factor=None, mode="nearest", align_corners=None, warning=True):
    return F.interpolate(input, size, scale_factor, mode, align_corners)


def add_prefix(inputs, prefix):
    return {f"{prefix}.{k}": v for k, v in inputs.items()}


def accuracy(pred, target, **k):
    return torch.zeros(())


class PixelData:
    def __init__(self, data=None):
        self.data = data


class SegDataSample:
    def __init__(self, metainfo=None):
        self.metainfo = dict(metainfo or {})

    def set_data(self, d):
        for k, v in d.items():
            setattr(self, k, v)

    def set_metainfo(self, d):
        self.metainfo.update(d)


class BaseDecodeHead(BaseModule):
    """mmseg BaseDecodeHead: conv_seg, dropout, _transform_inputs, cls_seg, predict, predict_by_feat."""

    def __init__(self, in_channels, channels, *, num_classes, out_channels=None, threshold=None, dropout_ratio=0.1,
                 conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU"), in_index=-1, input_transform=None,
                 loss_decode=None, ignore_index=255, sampler=None, align_corners=False, init_cfg=None):
        super().__init__(init_cfg)
        self.in_channels = in_channels
        self.in_index = in_index
        self.input_transform = input_transform
        self.channels = channels
        self.num_classes = num_classes
        self.out_channels = out_channels or num_classes
        self.dropout_ratio = dropout_ratio
        self.norm_cfg = norm_cfg
        self.act_cfg = act_cfg
        self.conv_cfg = conv_cfg
        self.ignore_index = ignore_index
        self.align_corners = align_corners
        self.conv_seg = nn.Conv2d(channels, self.out_channels, kernel_size=1)
        self.dropout = nn.Dropout2d(dropout_ratio) if dropout_ratio > 0 else None

    def _transform_inputs(self, inputs):
        if self.input_transform == "multiple_select":
            return [inputs[i] for i in self.in_index]
        if self.input_transform == "resize_concat":
            inputs = [inputs[i] for i in self.in_index]
            up = [resize(x, size=inputs[0].shape[2:], mode="bilinear", align_corners=self.align_corners) for x in inputs]
            return torch.cat(up, dim=1)
        return inputs[self.in_index]

    def cls_seg(self, feat):
        if self.dropout is not None:
            feat = self.dropout(feat)
        return self.conv_seg(feat)

    def predict(self, inputs, batch_img_metas, test_cfg):
        return self.predict_by_feat(self.forward(inputs), batch_img_metas)

    def predict_by_feat(self, seg_logits, batch_img_metas):
        if isinstance(batch_img_metas[0]["img_shape"], torch.Size):
            size = batch_img_metas[0]["img_shape"]
        elif "pad_shape" in batch_img_metas[0]:
            size = batch_img_metas[0]["pad_shape"][:2]
        else:
            size = batch_img_metas[0]["img_shape"]
        return resize(seg_logits, size=size, mode="bilinear", align_corners=self.align_corners)


class EncoderDecoder(BaseModule):
    """mmseg EncoderDecoder (inference side only)."""

    def __init__(self, backbone, decode_head, neck=None, auxiliary_head=None, train_cfg=None, test_cfg=None,
                 data_preprocessor=None, pretrained=None, init_cfg=None):
        super().__init__(init_cfg)
        self.backbone = MODELS.build(backbone)
        self.decode_head = MODELS.build(decode_head)
        self.align_corners = self.decode_head.align_corners
        self.num_classes = self.decode_head.num_classes
        self.out_channels = self.decode_head.out_channels
        self.train_cfg = ConfigDict(train_cfg or {})
        self.test_cfg = ConfigDict(test_cfg or {})
        self.data_preprocessor_cfg = data_preprocessor
        if isinstance(data_preprocessor, dict) and "mean" in data_preprocessor:
            # mmengine BaseModel builds the preprocessor; MsVFMEncoderDecoder.__init__ reads its .mean / .std (:106-107)
            self.data_preprocessor = types.SimpleNamespace(
                mean=torch.tensor(data_preprocessor["mean"], dtype=torch.float32).view(-1, 1, 1),
                std=torch.tensor(data_preprocessor["std"], dtype=torch.float32).view(-1, 1, 1))

    def extract_feat(self, inputs):
        return self.backbone(inputs)

    def encode_decode(self, inputs, batch_img_metas):
        x = self.extract_feat(inputs)
        return self.decode_head.predict(x, batch_img_metas, self.test_cfg)

    def slide_inference(self, inputs, batch_img_metas):
        h_stride, w_stride = self.test_cfg.stride
        h_crop, w_crop = self.test_cfg.crop_size
        batch_size, _, h_img, w_img = inputs.size()
        out_channels = self.out_channels
        h_grids = max(h_img - h_crop + h_stride - 1, 0) // h_stride + 1
        w_grids = max(w_img - w_crop + w_stride - 1, 0) // w_stride + 1
        preds = inputs.new_zeros((batch_size, out_channels, h_img, w_img))
        count_mat = inputs.new_zeros((batch_size, 1, h_img, w_img))
        for h_idx in range(h_grids):
            for w_idx in range(w_grids):
                y1 = h_idx * h_stride
                x1 = w_idx * w_stride
                y2 = min(y1 + h_crop, h_img)
                x2 = min(x1 + w_crop, w_img)
                y1 = max(y2 - h_crop, 0)
                x1 = max(x2 - w_crop, 0)
                crop_img = inputs[:, :, y1:y2, x1:x2]
                batch_img_metas[0]["img_shape"] = crop_img.shape[2:]
                crop_seg_logit = self.encode_decode(crop_img, batch_img_metas)
                preds += F.pad(crop_seg_logit, (int(x1), int(preds.shape[3] - x2), int(y1), int(preds.shape[2] - y2)))
                count_mat[:, :, y1:y2, x1:x2] += 1
        assert (count_mat == 0).sum() == 0
        return preds / count_mat

    def whole_inference(self, inputs, batch_img_metas):
        return self.encode_decode(inputs, batch_img_metas)

    def inference(self, inputs, batch_img_metas):
        assert self.test_cfg.get("mode", "whole") in ["slide", "whole"]
        if self.test_cfg.mode == "slide":
            return self.slide_inference(inputs, batch_img_metas)
        return self.whole_inference(inputs, batch_img_metas)

    def predict(self, inputs, data_samples=None):
        if data_samples is not None:
            batch_img_metas = [d.metainfo for d in data_samples]
        else:
            batch_img_metas = [dict(ori_shape=inputs.shape[2:], img_shape=inputs.shape[2:], pad_shape=inputs.shape[2:],
                                    padding_size=[0, 0, 0, 0])] * inputs.shape[0]
        seg_logits = self.inference(inputs, batch_img_metas)
        return self.postprocess_result(seg_logits, data_samples)

    def postprocess_result(self, seg_logits, data_samples=None):
        batch_size, C, H, W = seg_logits.shape
        if data_samples is None:
            data_samples = [SegDataSample() for _ in range(batch_size)]
            only_prediction = True
        else:
            only_prediction = False
        for i in range(batch_size):
            if not only_prediction:
                meta = data_samples[i].metainfo
                pl, pr, pt, pb = meta.get("padding_size", [0] * 4)
                i_seg = seg_logits[i:i + 1, :, pt:H - pb, pl:W - pr]
                i_seg = resize(i_seg, size=meta["ori_shape"], mode="bilinear", align_corners=self.align_corners,
                               warning=False).squeeze(0)
            else:
                i_seg = seg_logits[i]
            i_pred = i_seg.argmax(dim=0, keepdim=True)
            data_samples[i].set_data({"seg_logits": PixelData(data=i_seg), "pred_sem_seg": PixelData(data=i_pred)})
        return data_samples


class IoUMetric:
    """mmseg IoUMetric (intersect_and_union / compute_metrics for iou_metrics=['mIoU'])."""

    def __init__(self, ignore_index=255, iou_metrics=("mIoU",), nan_to_num=None, beta=1, collect_device="cpu",
                 output_dir=None, format_only=False, prefix=None, **kwargs):
        self.ignore_index = ignore_index
        self.metrics = list(iou_metrics)
        self.nan_to_num = nan_to_num
        self.output_dir = output_dir
        self.format_only = format_only
        self.results = []
        self.dataset_meta = None

    @staticmethod
    def intersect_and_union(pred_label, label, num_classes, ignore_index):
        mask = label != ignore_index
        pred_label = pred_label[mask]
        label = label[mask]
        intersect = pred_label[pred_label == label]
        area_intersect = torch.histc(intersect.float(), bins=num_classes, min=0, max=num_classes - 1).cpu()
        area_pred_label = torch.histc(pred_label.float(), bins=num_classes, min=0, max=num_classes - 1).cpu()
        area_label = torch.histc(label.float(), bins=num_classes, min=0, max=num_classes - 1).cpu()
        area_union = area_pred_label + area_label - area_intersect
        return area_intersect, area_union, area_pred_label, area_label

    def compute_metrics(self, results):
        import numpy as np
        results = tuple(zip(*results))
        assert len(results) == 4
        ti, tu, tp, tl = (sum(r) for r in results)
        all_acc = ti.sum() / tl.sum()
        ret = OrderedDict({"aAcc": all_acc})
        ret["IoU"] = ti / tu
        ret["Acc"] = ti / tl
        ret = {k: v.numpy() for k, v in ret.items()}
        summary = OrderedDict({k: np.round(np.nanmean(v) * 100, 2) for k, v in ret.items()})
        out = {}
        for k, v in summary.items():
            out[k if k == "aAcc" else "m" + k] = v
        return out


# ----------------------------------------------------------------------------- peft LoRA [3P 0.10.0]
class LoraConfig:
    def __init__(self, r=8, lora_alpha=8, target_modules=None, lora_dropout=0.0, bias="none", **kw):
        self.r, self.lora_alpha, self.target_modules, self.lora_dropout, self.bias = r, lora_alpha, target_modules, lora_dropout, bias


class LoraLinear(nn.Module):
    """y = base_layer(x) + lora_B(lora_A(dropout(x))) * (lora_alpha / r); A kaiming-uniform, B zeros."""

    def __init__(self, base_layer, r, lora_alpha, lora_dropout):
        super().__init__()
        self.base_layer = base_layer
        self.lora_dropout = nn.ModuleDict({"default": nn.Dropout(lora_dropout) if lora_dropout > 0 else nn.Identity()})
        self.lora_A = nn.ModuleDict({"default": nn.Linear(base_layer.in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({"default": nn.Linear(r, base_layer.out_features, bias=False)})
        self.scaling = {"default": lora_alpha / r}
        nn.init.kaiming_uniform_(self.lora_A["default"].weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B["default"].weight)

    def forward(self, x):
        result = self.base_layer(x)
        a, b, d = self.lora_A["default"], self.lora_B["default"], self.lora_dropout["default"]
        return result + b(a(d(x))) * self.scaling["default"]

    # peft 0.10.0 tuners_utils.BaseTunerLayer [3P]: `.weight` / `.bias` of a wrapped layer are the BASE layer's tensors.
    # eva_02.py:337-339 reads `self.q_proj.weight` and calls F.linear itself, so the LoRA adapters peft attaches to
    # q_proj / k_proj / v_proj never enter the EVA02 forward pass; only `attn.proj` (called as a module, :379) is adapted.
    @property
    def weight(self):
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias


class LoraModel(nn.Module):
    def __init__(self, model, config):
        super().__init__()
        self.model = model
        targets = list(config.target_modules)
        for name, mod in list(model.named_modules()):
            if isinstance(mod, nn.Linear) and any(name == t or name.endswith("." + t) for t in targets):
                parent_name, _, child = name.rpartition(".")
                parent = model.get_submodule(parent_name) if parent_name else model
                setattr(parent, child, LoraLinear(mod, config.r, config.lora_alpha, config.lora_dropout))

    def forward(self, *a, **k):
        return self.model(*a, **k)


class PeftModel(nn.Module):
    def __init__(self, model, config):
        super().__init__()
        self.base_model = LoraModel(model, config)
        self.peft_config = {"default": config}

    def forward(self, *a, **k):
        return self.base_model(*a, **k)


def get_peft_model(model, config):
    return PeftModel(model, config)


# ----------------------------------------------------------------------------- module plumbing
class _Inert(types.ModuleType):
    """Module whose unknown attributes resolve to inert callables/classes."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})
        setattr(self, name, obj)
        return obj


_STUB_PREFIXES = ("mmseg", "mmengine", "mmcv", "mmdet", "peft", "timm", "xformers", "matplotlib", "prettytable", "ftfy")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_PREFIXES:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Inert(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False


def install():
    """Idempotently install the stand-ins and the `rein` namespace packages."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    os.environ["XFORMERS_DISABLED"] = "1"  # dino_layers/attention.py:21 -> pure-torch attention
    sys.meta_path.insert(0, _StubFinder())

    def mod(name, **attrs):
        m = importlib.import_module(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    mod("mmseg.registry", MODELS=MODELS, METRICS=METRICS)
    mod("mmseg.models.builder", BACKBONES=MODELS, MODELS=MODELS)
    mod("mmseg.models.segmentors", EncoderDecoder=EncoderDecoder)
    mod("mmseg.models.decode_heads.decode_head", BaseDecodeHead=BaseDecodeHead)
    mod("mmseg.models.utils", resize=resize)
    mod("mmseg.models.losses", accuracy=accuracy)
    mod("mmseg.structures", SegDataSample=SegDataSample)
    mod("mmseg.utils", SampleList=list, add_prefix=add_prefix)
    mod("mmseg.evaluation.metrics.iou_metric", IoUMetric=IoUMetric)
    mod("mmengine.model", BaseModule=BaseModule, is_model_wrapper=lambda m: False)
    mod("mmengine.logging", MMLogger=_Logger, print_log=print_log)
    mod("mmcv.cnn", ConvModule=ConvModule, build_norm_layer=build_norm_layer)
    mod("peft", LoraConfig=LoraConfig, get_peft_model=get_peft_model)
    tl = mod("timm.models.layers", to_2tuple=lambda x: (x, x) if not isinstance(x, (tuple, list)) else tuple(x),
             trunc_normal_=nn.init.trunc_normal_)
    tl.drop_path = lambda x, p=0.0, training=False: x
    def _mea(q, k, v, attn_bias=None, p=0.0, scale=None):
        # xformers.ops.memory_efficient_attention on (B, N, H, D) tensors = softmax(q k^T / sqrt D) v [3P]
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), scale=scale)
        return o.transpose(1, 2)
    mod("xformers.ops", memory_efficient_attention=_mea)
    mod("matplotlib", use=lambda *a, **k: None)
    mod("matplotlib.pyplot")

    rein_root = REFERENCE_ROOT / "rein"
    for name, sub in [("rein", ""), ("rein.models", "models"), ("rein.models.backbones", "models/backbones"),
                      ("rein.models.heads", "models/heads"), ("rein.models.segmentors", "models/segmentors"),
                      ("rein.utils", "utils")]:
        m = types.ModuleType(name)
        m.__path__ = [str(rein_root / sub)]
        m.__package__ = name
        sys.modules[name] = m
    sys.modules["rein.utils"].subplotimg = lambda *a, **k: None
    from importlib import import_module as _imp
    w = _imp("rein.utils.wrappers")
    sys.modules["rein.utils"].resize = w.resize
    sys.modules["rein.utils"].crop = w.crop
    _installed = True


def load(*names):
    """Import reference modules by dotted name (relative to `rein.`) after installing the shim."""
    install()
    mods = [importlib.import_module("rein." + n) for n in names]
    return mods[0] if len(mods) == 1 else mods


def build_reference_segmentor(model_cfg: dict, backbone_ckpt_path: str):
    """MODELS.build of the reference's LoraBackboneEncoderDecoder from a config dict shaped like
    configs/_base_/models/lora_dinov2_linear.py (the `checkpoint` entry is replaced)."""
    load("models.backbones.dino_v2", "models.heads.linear_head", "models.segmentors.Lora_encoder_decoder")
    cfg = dict(model_cfg)
    cfg["checkpoint"] = backbone_ckpt_path
    cfg.pop("data_preprocessor", None)
    m = MODELS.build(cfg)
    m.eval()
    return m


def build_reference_ms_segmentor(model_cfg: dict, backbone_ckpt_path: str):
    """MODELS.build of the reference's MsVFMEncoderDecoder (LoRABackbone + LinearHead + VFMHead / MaskTransformerDecoder)
    from a config dict shaped like configs/_base_/models/lora_dinov2_ms_masked.py."""
    load("models.backbones.dino_v2", "models.backbones.lora_backbone", "models.heads.linear_head", "models.heads.Transformer",
         "models.heads.VFMHead", "models.segmentors.Ms_VFM_encoder_decoder")
    import copy
    cfg = copy.deepcopy(model_cfg)
    cfg["backbone"]["checkpoint"] = backbone_ckpt_path
    cfg["train_cfg"] = ConfigDict({k: (ConfigDict(v) if isinstance(v, dict) else v) for k, v in cfg.get("train_cfg", {}).items()})
    m = MODELS.build(cfg)
    m.eval()
    return m


def build_reference_eva_segmentor(model_cfg: dict, backbone_ckpt_path: str):
    """MODELS.build of the shim EncoderDecoder holding the reference's own LoRABackbone(EVA2) and LinearHead, from a config
    shaped like configs/_base_/models/lora_eva02_linear.py."""
    load("models.backbones.eva_02", "models.backbones.lora_backbone", "models.heads.linear_head")
    if "EncoderDecoder" not in MODELS.module_dict:
        MODELS.register_module(name="EncoderDecoder", module=EncoderDecoder)
    import copy
    cfg = copy.deepcopy(model_cfg)
    cfg["backbone"]["checkpoint"] = backbone_ckpt_path
    cfg.pop("data_preprocessor", None)
    m = MODELS.build(cfg)
    m.eval()
    return m


def build_reference_sam_segmentor(model_cfg: dict, backbone_ckpt_path: str):
    """MODELS.build of the shim EncoderDecoder holding the reference's own LoRABackbone(SAMViT) and LinearHead, from a
    config shaped like configs/_base_/models/lora_sam_linear.py."""
    load("models.backbones.sam_vit", "models.backbones.lora_backbone", "models.heads.linear_head")
    if "EncoderDecoder" not in MODELS.module_dict:
        MODELS.register_module(name="EncoderDecoder", module=EncoderDecoder)
    import copy
    cfg = copy.deepcopy(model_cfg)
    cfg["backbone"]["checkpoint"] = backbone_ckpt_path
    cfg.pop("data_preprocessor", None)
    m = MODELS.build(cfg)
    m.eval()
    return m
