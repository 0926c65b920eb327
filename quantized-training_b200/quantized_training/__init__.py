"""quantized_training -- B200-native build of the fake-quant hot path.

Drop-in for the import surface the reference's drivers use
(``from quantized_training import add_qspec_args, quantize`` ...), with the numerics
executed by hand-written sm_100a kernels behind a C ABI (include/qt_b200.h).
CUDA only: there is no CPU or eager fallback.
"""
from . import _C
from . import decomposed  # registers torch.ops.quantized_ops.{vmap, quantize, dequantize} (CUDA)
from .fake_quantize import FusedAmaxObsFakeQuantize, get_quantization_map
from .host_io import HostPipeline, fake_quantize_host
from .qconfig import QConfig, get_qconfig
from .quantize import convert, get_quantized_model, prepare, propagate_config, quantize, replace_softmax
from .quantizer import QScheme, QuantizationSpec
from .training_args import add_qspec_args

# qscheme constants, as the reference exposes them at package level
per_tensor_symmetric = QScheme.PER_TENSOR_SYMMETRIC
per_channel_symmetric = QScheme.PER_CHANNEL_SYMMETRIC
microscaling = QScheme.MICROSCALING
group_wise_affine = QScheme.GROUP_WISE_AFFINE

__all__ = [
    "FusedAmaxObsFakeQuantize",
    "HostPipeline",
    "QConfig",
    "QScheme",
    "QuantizationSpec",
    "add_qspec_args",
    "fake_quantize_host",
    "convert",
    "get_qconfig",
    "get_quantization_map",
    "get_quantized_model",
    "prepare",
    "propagate_config",
    "quantize",
    "replace_softmax",
]
