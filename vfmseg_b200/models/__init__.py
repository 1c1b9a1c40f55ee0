from .dino_v2 import DinoVisionTransformer
from .linear_head import LinearHead
from .lora import LoraConfig, LoraLinear, PeftModel, get_peft_model
from .segmentors import LoraBackboneEncoderDecoder, SegDataPreProcessor

__all__ = ["DinoVisionTransformer", "LinearHead", "LoraBackboneEncoderDecoder", "SegDataPreProcessor",
           "LoraConfig", "LoraLinear", "PeftModel", "get_peft_model"]
