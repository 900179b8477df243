"""Shim of torchsde._core.misc — only the helpers the reference's vendored sdeint.py touches."""
import warnings

import torch


def assert_no_grad(names, maybe_tensors):
    for name, maybe_tensor in zip(names, maybe_tensors):
        if torch.is_tensor(maybe_tensor) and maybe_tensor.requires_grad:
            raise ValueError(f"Argument {name} must not require gradient.")


def handle_unused_kwargs(unused_kwargs, msg=None):
    if len(unused_kwargs) > 0:
        if msg is not None:
            warnings.warn(f"{msg}: Unexpected arguments {unused_kwargs}")
        else:
            warnings.warn(f"Unexpected arguments {unused_kwargs}")


def is_strictly_increasing(ts):
    return all(x < y for x, y in zip(ts[:-1], ts[1:]))


def batch_mvp(m, v):
    return torch.bmm(m, v.unsqueeze(-1)).squeeze(dim=-1)


def vjp(outputs, inputs, **kwargs):
    if torch.is_tensor(inputs):
        inputs = [inputs]
    if torch.is_tensor(outputs):
        outputs = [outputs]
    outputs = [o for o in outputs if o.requires_grad]
    _vjp = torch.autograd.grad(outputs, inputs, **kwargs)
    return [torch.zeros_like(i) if v is None else v for v, i in zip(_vjp, inputs)]


def jvp(outputs, inputs, grad_inputs=None, **kwargs):
    raise NotImplementedError("shim: jvp is not on the reference's Euler path")
