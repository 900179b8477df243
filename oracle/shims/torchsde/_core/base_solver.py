"""Shim of torchsde._core.base_solver.BaseSDESolver (0.2.5), restated for the fixed-step path.

The integrate loop is the library twin of the reference's vendored copy (models/utils/sdeint.py:326-384).
"""
import abc

import torch

from . import adaptive_stepping  # noqa: F401  (re-exported: reference does `from ...base_solver import interp, adaptive_stepping`)
from . import interp
from ..settings import NOISE_TYPES  # noqa: F401


class BaseSDESolver(metaclass=abc.ABCMeta):
    strong_order = None
    weak_order = None
    sde_type = None
    noise_types = None
    levy_area_approximations = None

    def __init__(self, sde, bm, dt, adaptive, rtol, atol, dt_min, options, **kwargs):
        super(BaseSDESolver, self).__init__(**kwargs)
        if sde.sde_type != self.sde_type:
            raise ValueError(f"SDE is of type {sde.sde_type} but solver is for type {self.sde_type}")
        if sde.noise_type not in self.noise_types:
            raise ValueError(f"SDE has noise type {sde.noise_type} but solver only supports noise types "
                             f"{self.noise_types}")
        if bm.levy_area_approximation not in self.levy_area_approximations:
            raise ValueError(f"SDE solver requires one of {self.levy_area_approximations} set as the "
                             f"`levy_area_approximation` on the Brownian motion.")
        self.sde = sde
        self.bm = bm
        self.dt = dt
        self.adaptive = adaptive
        self.rtol = rtol
        self.atol = atol
        self.dt_min = dt_min
        self.options = options

    def __repr__(self):
        return f"{self.__class__.__name__} of strong order: {self.strong_order}, and weak order: {self.weak_order}"

    def init_extra_solver_state(self, t0, y0):
        return ()

    @abc.abstractmethod
    def step(self, t0, t1, y0, extra0):
        raise NotImplementedError

    def integrate(self, y0, ts, extra0):
        step_size = self.dt
        prev_t = curr_t = ts[0]
        prev_y = curr_y = y0
        curr_extra = extra0
        ys = [y0]
        for out_t in ts[1:]:
            while curr_t < out_t:
                next_t = min(curr_t + step_size, ts[-1])
                if self.adaptive:
                    raise NotImplementedError("shim: adaptive=False on the reference path")
                prev_t, prev_y = curr_t, curr_y
                curr_y, curr_extra = self.step(curr_t, next_t, curr_y, curr_extra)
                curr_t = next_t
            ys.append(interp.linear_interp(t0=prev_t, y0=prev_y, t1=curr_t, y1=curr_y, t=out_t))
        return torch.stack(ys, dim=0), curr_extra
