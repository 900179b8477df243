"""Shim of torchsde._core.base_sde (0.2.5): BaseSDE + the library's ForwardSDE glue for diagonal Euler."""
import abc

from torch import nn

from ..settings import NOISE_TYPES, SDE_TYPES


class BaseSDE(abc.ABC, nn.Module):
    def __init__(self, noise_type, sde_type):
        super(BaseSDE, self).__init__()
        if noise_type not in NOISE_TYPES:
            raise ValueError(f"Expected noise type in {NOISE_TYPES}, but found {noise_type}")
        if sde_type not in SDE_TYPES:
            raise ValueError(f"Expected sde type in {SDE_TYPES}, but found {sde_type}")
        self.noise_type = noise_type
        self.sde_type = sde_type


class ForwardSDE(BaseSDE):
    """Library wrapper used by torchsde.sdeint (decoder call site dec_hivt_nusargo_sde.py:88)."""

    def __init__(self, sde, fast_dg_ga_jvp_column_sum=False):
        super(ForwardSDE, self).__init__(sde_type=sde.sde_type, noise_type=sde.noise_type)
        self._base_sde = sde
        if hasattr(sde, 'f_and_g_prod'):
            self.f_and_g_prod = sde.f_and_g_prod
        elif hasattr(sde, 'f') and hasattr(sde, 'g_prod'):
            self.f_and_g_prod = self.f_and_g_prod_default1
        else:
            self.f_and_g_prod = self.f_and_g_prod_default2
        self.f = getattr(sde, 'f', None)
        self.g = getattr(sde, 'g', None)
        self.f_and_g = getattr(sde, 'f_and_g', self.f_and_g_default)
        self.g_prod = getattr(sde, 'g_prod', self.g_prod_default)
        if sde.noise_type != NOISE_TYPES.diagonal:
            raise NotImplementedError("shim: only diagonal noise is on the reference path")

    def f_and_g_default(self, t, y):
        return self.f(t, y), self.g(t, y)

    def prod(self, g, v):
        return g * v

    def g_prod_default(self, t, y, v):
        return self.prod(self.g(t, y), v)

    def f_and_g_prod_default1(self, t, y, v):
        return self.f(t, y), self.g_prod(t, y, v)

    def f_and_g_prod_default2(self, t, y, v):
        f, g = self.f_and_g(t, y)
        return f, self.prod(g, v)


class RenameMethodsSDE(BaseSDE):
    def __init__(self, sde, drift='f', diffusion='g', prior_drift='h', diffusion_prod='g_prod',
                 drift_and_diffusion='f_and_g', drift_and_diffusion_prod='f_and_g_prod'):
        super(RenameMethodsSDE, self).__init__(noise_type=sde.noise_type, sde_type=sde.sde_type)
        self._base_sde = sde
        for name, value in zip(('f', 'g', 'h', 'g_prod', 'f_and_g', 'f_and_g_prod'),
                               (drift, diffusion, prior_drift, diffusion_prod, drift_and_diffusion,
                                drift_and_diffusion_prod)):
            try:
                setattr(self, name, getattr(sde, value))
            except AttributeError:
                pass


class SDEIto(BaseSDE):
    def __init__(self, noise_type):
        super(SDEIto, self).__init__(noise_type=noise_type, sde_type=SDE_TYPES.ito)


class SDEStratonovich(BaseSDE):
    def __init__(self, noise_type):
        super(SDEStratonovich, self).__init__(noise_type=noise_type, sde_type=SDE_TYPES.stratonovich)


class SDELogqp(BaseSDE):
    def __init__(self, sde):
        raise NotImplementedError("shim: logqp=False everywhere in the reference (SURVEY §0.7)")
