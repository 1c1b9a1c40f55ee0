from .dino_v2 import DinoVisionTransformer
from .eva_02 import EVA2
from .linear_head import LinearHead
from .sam_vit import SAMViT
from .lora import LoRABackbone, LoraConfig, LoraLinear, PeftModel, get_peft_model
from .segmentors import EncoderDecoder, LoraBackboneEncoderDecoder, MsVFMEncoderDecoder, SegDataPreProcessor
from .vfm_head import MaskTransformerDecoder, TransformerDecoder, VFMHead

__all__ = ["DinoVisionTransformer", "LinearHead", "LoraBackboneEncoderDecoder", "SegDataPreProcessor", "MsVFMEncoderDecoder",
           "VFMHead", "TransformerDecoder", "MaskTransformerDecoder", "LoRABackbone", "EVA2", "SAMViT", "EncoderDecoder",
           "LoraConfig", "LoraLinear", "PeftModel", "get_peft_model"]
