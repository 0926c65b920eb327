"""What `quantize` swaps and what `prepare` hooks (reference: quantization_mappings.py:16-72).

* DEFAULT_QAT_MODULE_MAPPINGS   float module  -> QAT module (weight fake-quant inside)
* TRANSFORMER_MODULE_MAPPINGS   HF block      -> quantizable block (hookable matmul/mul/add/softmax)
* QCONFIG_PROPAGATE_MODULE_CLASS_LIST  op-group name (--quantize_forward / --quantize_backprop) -> classes
The Llama entry is enabled (commented out at the reference's HEAD).  ConvBn fusions, DistilBERT, GPT-2 and
Whisper entries belong to model families outside this build's scope.
"""
from typing import Any, Callable, Dict

import torch.nn as nn
from transformers.activations import GELUActivation
from transformers.models import bert, llama, mobilebert, roberta
from transformers.pytorch_utils import Conv1D

from .modules import lora as _lora
from .modules import qat as nnqat
from .modules import quantizable

DEFAULT_QAT_MODULE_MAPPINGS: Dict[Callable, Any] = {
    nn.Conv2d: nnqat.Conv2d,
    nn.Conv3d: nnqat.Conv3d,
    nn.Linear: nnqat.Linear,
    _lora.LoraLinear: nnqat.LoraLinear,
}
try:  # peft adapters, when the package exists
    from peft.tuners.lora import Linear as _PeftLoraLinear
    DEFAULT_QAT_MODULE_MAPPINGS[_PeftLoraLinear] = nnqat.LoraLinear
except Exception:  # pragma: no cover
    pass

_B, _R, _M, _L = (bert.modeling_bert, roberta.modeling_roberta, mobilebert.modeling_mobilebert,
                  llama.modeling_llama)
TRANSFORMER_MODULE_MAPPINGS: Dict[Callable, Any] = {
    _B.BertLayer: quantizable.BertLayer,
    _R.RobertaLayer: quantizable.BertLayer,
    _B.BertSelfAttention: quantizable.BertSelfAttention,
    _B.BertSelfOutput: quantizable.BertSelfOutput,
    _B.BertOutput: quantizable.BertOutput,
    _R.RobertaSelfAttention: quantizable.BertSelfAttention,
    _R.RobertaSelfOutput: quantizable.BertSelfOutput,
    _R.RobertaOutput: quantizable.BertOutput,
    _M.MobileBertLayer: quantizable.MobileBertLayer,
    _M.MobileBertSelfAttention: quantizable.MobileBertSelfAttention,
    _M.MobileBertSelfOutput: quantizable.MobileBertSelfOutput,
    _M.FFNOutput: quantizable.FFNOutput,
    _M.MobileBertOutput: quantizable.MobileBertOutput,
    _L.LlamaDecoderLayer: quantizable.LlamaDecoderLayer,
}

QCONFIG_PROPAGATE_MODULE_CLASS_LIST = {
    "activation": [nn.ReLU, nn.GELU, nn.Softmax, GELUActivation],
    "gemm": [nn.Conv1d, nn.Conv2d, nn.Conv3d, nn.Linear, Conv1D, quantizable.MatmulFunctional],
    "layernorm": [nn.LayerNorm, _L.LlamaRMSNorm, _M.NoNorm],
    "residual": [quantizable.AddFunctional],
    "scaling": [quantizable.MulFunctional],
}
