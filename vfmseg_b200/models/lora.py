"""LoRA wrapping with peft 0.10.0's module tree and state-dict names, so checkpoints trained with
the reference (rein/models/segmentors/Lora_encoder_decoder.py:17-24, backbones/lora_backbone.py:16-23)
load unchanged:  <wrapper>.base_model.model.<path>.{base_layer.*, lora_A.default.weight, lora_B.default.weight}.

Inference semantics: W' = W + (lora_alpha / r) * B @ A, merged once when the engine packs weights.
peft is not installed in this image; this is a from-scratch container for the same parameters.
"""
from __future__ import annotations

import math
from typing import Sequence

import torch
import torch.nn as nn


class LoraConfig:
    def __init__(self, r=8, lora_alpha=8, target_modules: Sequence[str] = (), lora_dropout=0.0, bias="none", **_):
        self.r = r
        self.lora_alpha = lora_alpha
        self.target_modules = list(target_modules)
        self.lora_dropout = lora_dropout
        self.bias = bias


class LoraLinear(nn.Module):
    """Parameter container for a LoRA-adapted nn.Linear (A: kaiming-uniform(a=sqrt 5), B: zeros)."""

    def __init__(self, base_layer: nn.Linear, r: int, lora_alpha: float):
        super().__init__()
        self.base_layer = base_layer
        self.lora_A = nn.ModuleDict({"default": nn.Linear(base_layer.in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({"default": nn.Linear(r, base_layer.out_features, bias=False)})
        self.scaling = {"default": lora_alpha / r}
        nn.init.kaiming_uniform_(self.lora_A["default"].weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B["default"].weight)

    @property
    def in_features(self):
        return self.base_layer.in_features

    @property
    def out_features(self):
        return self.base_layer.out_features

    def merged_weight(self) -> torch.Tensor:
        return self.base_layer.weight + self.scaling["default"] * (self.lora_B["default"].weight @ self.lora_A["default"].weight)


class LoraModel(nn.Module):
    def __init__(self, model: nn.Module, config: LoraConfig):
        super().__init__()
        self.model = model
        n = 0
        for name, mod in list(model.named_modules()):
            if isinstance(mod, nn.Linear) and any(name == t or name.endswith("." + t) for t in config.target_modules):
                parent_name, _, child = name.rpartition(".")
                parent = model.get_submodule(parent_name) if parent_name else model
                setattr(parent, child, LoraLinear(mod, config.r, config.lora_alpha))
                n += 1
        if n == 0:
            raise ValueError(f"LoRA target_modules {config.target_modules} matched nothing")
        model._lora_scale = config.lora_alpha / config.r  # read by the engine when merging

    def forward(self, *a, **k):
        return self.model(*a, **k)


class PeftModel(nn.Module):
    def __init__(self, model: nn.Module, config: LoraConfig):
        super().__init__()
        self.base_model = LoraModel(model, config)
        self.peft_config = {"default": config}

    def forward(self, *a, **k):
        return self.base_model(*a, **k)


def get_peft_model(model: nn.Module, config: LoraConfig) -> PeftModel:
    return PeftModel(model, config)



def _register_lora_backbone():
    from ..registry import MODELS

    @MODELS.register_module()
    class LoRABackbone(nn.Module):
        """rein/models/backbones/lora_backbone.py:10-44: builds the inner backbone from the registry, wraps it with
        LoRA (state-dict prefix `model.base_model.model.`) and loads the frozen backbone checkpoint with the target
        modules renamed to `<t>.base_layer` (:27-35)."""

        def __init__(self, backbone, checkpoint=None, Lora_config=None, **kwargs):
            super().__init__()
            inner = MODELS.build(backbone)
            self.Lora_config = LoraConfig(r=Lora_config["r"], lora_alpha=Lora_config["lora_alpha"],
                                          target_modules=Lora_config["target_modules"],
                                          lora_dropout=Lora_config.get("lora_dropout", 0.0), bias="none")
            self.model = get_peft_model(inner, self.Lora_config)
            if checkpoint is not None:
                self.load_pretrained_backbone(checkpoint, Lora_config["target_modules"])

        def load_pretrained_backbone(self, checkpoint, target_modules):
            original = torch.load(checkpoint, map_location="cpu") if isinstance(checkpoint, str) else checkpoint
            new = {}
            for name, weight in original.items():
                for t in target_modules:
                    if t in name:
                        name = name.replace(t, t + ".base_layer")
                    new[name] = weight
            self.model.base_model.model.load_state_dict(new, strict=False)

        @property
        def inner(self):
            return self.model.base_model.model

        def forward(self, x):
            return self.model(x)

    return LoRABackbone


LoRABackbone = _register_lora_backbone()
