from .euler import Euler
from ...settings import METHODS, SDE_TYPES


def select(method, sde_type):
    if method == METHODS.euler and sde_type == SDE_TYPES.ito:
        return Euler
    raise NotImplementedError(f"shim: only method='euler' / Ito is on the reference path, got {method}/{sde_type}")
