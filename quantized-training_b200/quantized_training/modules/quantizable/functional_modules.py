"""Hookable stand-ins for the tensor ops between GEMMs.

`quantize()` can only attach fake-quantizers to *modules*, so attention blocks call these
instead of ``torch.matmul`` / ``*`` / ``+`` (reference: quantizable/functional_modules.py:8-27).
Op groups: MatmulFunctional -> "gemm", MulFunctional -> "scaling", AddFunctional -> "residual".
"""
from typing import Union

import torch
from torch import Tensor, nn

from ... import ops

__all__ = ["AddFunctional", "MulFunctional", "MatmulFunctional"]


class AddFunctional(nn.Module):
    def forward(self, x: Tensor, y: Union[Tensor, float]) -> Tensor:
        return torch.add(x, y)


class MulFunctional(nn.Module):
    def forward(self, x: Tensor, y: Union[Tensor, float]) -> Tensor:
        return torch.mul(x, y)


class MatmulFunctional(nn.Module):
    def forward(self, x: Tensor, y: Tensor) -> Tensor:
        return ops.matmul(x, y)
