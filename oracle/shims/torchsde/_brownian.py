"""Shim of torchsde._brownian: i.i.d. N(0, tb-ta) increments (what BrownianInterval yields for the solver's
non-overlapping, in-order queries) plus FixedIncrements for parity runs with caller-supplied dW."""
import torch

from .settings import LEVY_AREA_APPROXIMATIONS


class BaseBrownian:
    def __call__(self, ta, tb=None, return_U=False, return_A=False):
        raise NotImplementedError


class BrownianInterval(BaseBrownian):
    def __init__(self, t0=0., t1=1., size=None, dtype=None, device=None, entropy=None, dt=None, tol=0.,
                 pool_size=8, cache_size=45, halfway_tree=False,
                 levy_area_approximation=LEVY_AREA_APPROXIMATIONS.none, W=None, H=None):
        self.shape = tuple(size)
        self.dtype = dtype
        self.device = device
        self.levy_area_approximation = levy_area_approximation
        self._gen = torch.Generator(device='cpu')
        self._gen.manual_seed(0 if entropy is None else int(entropy))
        self.queries = []

    def __call__(self, ta, tb=None, return_U=False, return_A=False):
        self.queries.append((float(ta), float(tb)))
        h = tb - ta
        z = torch.randn(self.shape, dtype=self.dtype, generator=self._gen)
        return (z * torch.sqrt(torch.as_tensor(h, dtype=self.dtype))).to(self.device)


class FixedIncrements(BaseBrownian):
    """bm(ta, tb) returns dW[k] on the k-th query and logs (ta, tb)."""

    def __init__(self, dW):
        self.dW = dW
        self.shape = tuple(dW.shape[1:])
        self.dtype = dW.dtype
        self.device = dW.device
        self.levy_area_approximation = LEVY_AREA_APPROXIMATIONS.none
        self.queries = []

    def __call__(self, ta, tb=None, return_U=False, return_A=False):
        k = len(self.queries)
        self.queries.append((ta.clone() if torch.is_tensor(ta) else ta, tb.clone() if torch.is_tensor(tb) else tb))
        return self.dW[k]
