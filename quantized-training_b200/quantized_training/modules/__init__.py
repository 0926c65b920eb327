from . import qat, quantizable
from .lora import LoraLinear, apply_lora

__all__ = ["qat", "quantizable", "LoraLinear", "apply_lora"]
