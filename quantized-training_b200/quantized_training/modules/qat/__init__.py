from .conv import Conv1d, Conv2d, Conv3d
from .linear import Linear
from .lora import Linear as LoraLinear

__all__ = ["Conv1d", "Conv2d", "Conv3d", "Linear", "LoraLinear"]
