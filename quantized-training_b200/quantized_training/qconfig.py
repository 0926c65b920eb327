"""QConfig: the three fake-quant *constructors* attached to every module
(mirror of the reference's qconfig.py:14-58)."""
from collections import namedtuple

from torch import nn

from .fake_quantize import FusedAmaxObsFakeQuantize
from .quantizer import QuantizationSpec

__all__ = ["QConfig", "get_qconfig"]


class QConfig(namedtuple("QConfig", ["activation", "weight", "error"])):
    """Constructors (not instances) for activation, weight and error (activation-gradient)
    fake-quantizers; ``nn.Identity`` where a tensor class is left unquantized."""

    def __new__(cls, activation, weight, error):
        return super().__new__(cls, activation, weight, error)


def _create_fake_quant(quantization_spec, record_histogram, force_scale_power_of_two):
    if quantization_spec is None:
        return nn.Identity
    # accepts a string or an already parsed QuantizationSpec (argparse parses --error eagerly)
    spec = QuantizationSpec.from_str(quantization_spec)
    return FusedAmaxObsFakeQuantize.with_args(
        **spec.fake_quant_kwargs(),
        record_histogram=record_histogram,
        force_scale_power_of_two=force_scale_power_of_two,
    )


def get_qconfig(activation, weight, error, record_histogram=False, force_scale_power_of_two=False):
    make = lambda spec: _create_fake_quant(spec, record_histogram, force_scale_power_of_two)  # noqa: E731
    return QConfig(activation=make(activation), weight=make(weight), error=make(error))
