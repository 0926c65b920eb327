from .functional_modules import AddFunctional, MatmulFunctional, MulFunctional
from .modeling_bert import BertLayer, BertOutput, BertSelfAttention, BertSelfOutput
from .modeling_llama import LlamaAttention, LlamaDecoderLayer
from .modeling_mobilebert import (FFNOutput, MobileBertLayer, MobileBertOutput, MobileBertSelfAttention,
                                  MobileBertSelfOutput, OutputBottleneck)

__all__ = [
    "AddFunctional", "BertLayer", "BertOutput", "BertSelfAttention", "BertSelfOutput", "FFNOutput", "LlamaAttention",
    "LlamaDecoderLayer", "MatmulFunctional", "MobileBertLayer", "MobileBertOutput", "MobileBertSelfAttention",
    "MobileBertSelfOutput", "MulFunctional", "OutputBottleneck",
]
