from .quantizer import QScheme, QuantizationSpec, QuantizationSpecBase, get_quant_min_max

__all__ = ["QuantizationSpec", "QScheme", "QuantizationSpecBase", "get_quant_min_max"]
