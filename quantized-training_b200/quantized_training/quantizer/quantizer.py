"""qspec mini-language: ``dtype[,key=value...]``.

Host-side mirror of the reference's ``quantizer/quantizer.py`` (QScheme :18-22, key
abbreviations :24-33, value types :42-51, get_quant_min_max :53-94, QuantizationSpec
:96-146).  Same strings, same defaults, same errors; the numeric limits come from the
C ABI (``qt_format_min_max``) so Python and the kernels cannot drift apart.
"""
import re
from dataclasses import dataclass, fields
from enum import Enum
from typing import Any, Callable, List, Optional, Tuple, Union

from .. import _C

__all__ = ["QScheme", "QuantizationSpec", "QuantizationSpecBase", "get_quant_min_max"]


class QScheme(Enum):
    PER_TENSOR_SYMMETRIC = "per_tensor_symmetric"
    PER_CHANNEL_SYMMETRIC = "per_channel_symmetric"
    MICROSCALING = "microscaling"
    GROUP_WISE_AFFINE = "group_wise_affine"


class QuantizationSpecBase:
    """Stand-in for torch.ao.quantization.quantizer.QuantizationSpecBase (gone from torch >= 2.11)."""


def _int_or_tuple(text: str) -> Union[int, Tuple[int, ...]]:
    text = text.strip()
    if text[:1] == "(" and text[-1:] == ")":
        return tuple(int(tok) for tok in text[1:-1].split(","))
    return int(text)


# full key -> (abbreviation, converter)
_KEYS = {
    "quant_min": ("qmin", float),
    "quant_max": ("qmax", float),
    "qscheme": ("qs", QScheme),
    "amax_history_len": ("ahl", int),
    "ch_axis": ("ax", _int_or_tuple),
    "block_size": ("bs", _int_or_tuple),
    "scale_dtype": ("scale", str),
    "outlier_threshold": ("outlier", float),
}
ABBREV_MAP = {abbrev: full for full, (abbrev, _) in _KEYS.items()}
PARAMS_TYPE = {full: conv for full, (_, conv) in _KEYS.items()}
_TOP_LEVEL_COMMA = re.compile(r",(?![^()]*\))")  # commas inside (...) belong to a tuple value


def get_quant_min_max(dtype: str):
    """(qmin, qmax) of a dtype string; ValueError("Unsupported dtype: ...") otherwise."""
    lo, hi = _C.format_min_max(dtype)
    if re.fullmatch(r"u?int\d+", dtype, re.IGNORECASE):
        return int(lo), int(hi)
    if re.fullmatch(r"posit\d+_\d+", dtype, re.IGNORECASE) or re.fullmatch(r"nf\d+(?:_\d+)?", dtype, re.IGNORECASE):
        return (int(lo), int(hi)) if hi == int(hi) else (lo, hi)
    return lo, hi


@dataclass(eq=True)
class QuantizationSpec(QuantizationSpecBase):
    """How to quantize one tensor: dtype string plus optional dynamic-scaling parameters."""

    dtype: str
    observer_or_fake_quant_ctr: Optional[Callable[..., Any]] = None  # filled in __post_init__
    quant_min: Optional[float] = None
    quant_max: Optional[float] = None
    qscheme: Optional[QScheme] = None
    amax_history_len: Optional[int] = None
    ch_axis: Optional[Union[int, List[int]]] = None
    block_size: Optional[Union[int, List[int]]] = None
    scale_dtype: Optional[str] = None
    outlier_threshold: Optional[float] = None
    is_dynamic: bool = False

    @staticmethod
    def from_str(s):
        if isinstance(s, QuantizationSpec):  # the reference double-parses --error; be idempotent
            return s
        if not s:
            raise ValueError("String quantization_spec is None or empty")
        head, *rest = _TOP_LEVEL_COMMA.split(s)
        params = {"dtype": head}
        for item in rest:
            if "=" not in item:
                raise ValueError(f"Expected key=value format but got '{item}'")
            key, value = item.split("=")
            key = ABBREV_MAP.get(key, key)
            if key not in PARAMS_TYPE:
                raise ValueError(f"Unknown argument '{key}'. Valid keys: {', '.join(PARAMS_TYPE)}")
            params[key] = PARAMS_TYPE[key](value)
        scheme = params.get("qscheme")
        if scheme is not None:
            lo, hi = get_quant_min_max(head)
            params.setdefault("quant_min", float(lo))
            params.setdefault("quant_max", float(hi))
            if scheme in (QScheme.PER_TENSOR_SYMMETRIC, QScheme.PER_CHANNEL_SYMMETRIC):
                params.setdefault("amax_history_len", 16)
        return QuantizationSpec(**params)

    def __post_init__(self):
        if self.observer_or_fake_quant_ctr is None:
            from ..fake_quantize import FusedAmaxObsFakeQuantize
            self.observer_or_fake_quant_ctr = FusedAmaxObsFakeQuantize
        if self.qscheme is not None and self.quant_max is None:
            raise ValueError("quant_max is required for quantization.")
        if self.qscheme in (QScheme.MICROSCALING, QScheme.GROUP_WISE_AFFINE) and self.block_size is None:
            raise ValueError("block_size is required for microscaling.")

    def fake_quant_kwargs(self):
        """Constructor kwargs for FusedAmaxObsFakeQuantize (what the reference gets from dataclasses.asdict)."""
        skip = {"observer_or_fake_quant_ctr", "is_dynamic"}
        return {f.name: getattr(self, f.name) for f in fields(self) if f.name not in skip}
