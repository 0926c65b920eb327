from .linear import Linear
from .lora import Linear as LoraLinear

__all__ = ["Linear", "LoraLinear"]
