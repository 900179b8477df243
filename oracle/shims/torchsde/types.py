"""Shim of torchsde.types: typing aliases only. TEST INFRASTRUCTURE ONLY."""
from typing import Sequence, Union, Optional, Any, Dict, Tuple, Callable  # noqa: F401

import torch

Tensor = torch.Tensor
Tensors = Sequence[Tensor]
TensorOrTensors = Union[Tensor, Tensors]
Scalar = Union[float, Tensor]
Vector = Union[Sequence[float], Tensor]
Module = torch.nn.Module
Modules = Sequence[Module]
ModuleOrModules = Union[Module, Modules]
Size = Sequence[int]
