"""Shim of torchsde.sdeint (0.2.5), restated. Mirrors the reference's vendored fork (models/utils/sdeint.py:110-197,
check_contract :827-995) minus the nus_mask plumbing; includes the side-effect-free f/g probe (:913-921 twin)."""
import torch

from ._brownian import BrownianInterval
from ._core import base_sde, methods, misc
from .settings import LEVY_AREA_APPROXIMATIONS, METHODS, NOISE_TYPES, SDE_TYPES


def check_contract(sde, y0, ts, bm, method, adaptive, options, names, logqp):
    if not hasattr(sde, "noise_type") or sde.noise_type not in NOISE_TYPES:
        raise ValueError("sde noise_type missing/invalid")
    if not hasattr(sde, "sde_type") or sde.sde_type not in SDE_TYPES:
        raise ValueError("sde sde_type missing/invalid")
    if not torch.is_tensor(y0) or y0.dim() != 2:
        raise ValueError("`y0` must be a 2-dimensional tensor of shape (batch, channels).")
    if logqp:
        raise NotImplementedError("shim: logqp=False on the reference path")
    if method is None:
        method = METHODS.srk if sde.sde_type == SDE_TYPES.ito else METHODS.midpoint
    if method not in METHODS:
        raise ValueError(f"Expected method in {METHODS}, but found {method}.")
    if not torch.is_tensor(ts):
        ts = torch.tensor(ts, dtype=y0.dtype, device=y0.device)
    if not misc.is_strictly_increasing(ts):
        raise ValueError("Evaluation times `ts` must be strictly increasing.")
    batch_sizes, state_sizes, noise_sizes = [y0.size(0)], [y0.size(1)], []
    if bm is not None:
        batch_sizes.append(bm.shape[0])
        noise_sizes.append(bm.shape[1])
    f_shape = tuple(sde.f(ts[0], y0).size())      # the library's wasted probe evaluation
    g_shape = tuple(sde.g(ts[0], y0).size())
    batch_sizes += [f_shape[0], g_shape[0]]
    state_sizes += [f_shape[1], g_shape[1]]
    noise_sizes.append(g_shape[1])
    if len(set(batch_sizes)) != 1 or len(set(state_sizes)) != 1 or len(set(noise_sizes)) != 1:
        raise ValueError("Batch/state/noise sizes not consistent.")
    sde = base_sde.ForwardSDE(sde)
    if bm is None:
        bm = BrownianInterval(t0=ts[0], t1=ts[-1], size=(batch_sizes[0], noise_sizes[0]), dtype=y0.dtype,
                              device=y0.device, levy_area_approximation=LEVY_AREA_APPROXIMATIONS.none)
    options = {} if options is None else options.copy()
    return sde, y0, ts, bm, method, options


def sdeint(sde, y0, ts, bm=None, method=None, dt=1e-3, adaptive=False, rtol=1e-5, atol=1e-4, dt_min=1e-5,
           options=None, names=None, logqp=False, extra=False, extra_solver_state=None, **unused_kwargs):
    misc.handle_unused_kwargs(unused_kwargs, msg="`sdeint`")
    sde, y0, ts, bm, method, options = check_contract(sde, y0, ts, bm, method, adaptive, options, names, logqp)
    misc.assert_no_grad(['ts', 'dt', 'rtol', 'atol', 'dt_min'], [ts, dt, rtol, atol, dt_min])
    solver_fn = methods.select(method=method, sde_type=sde.sde_type)
    solver = solver_fn(sde=sde, bm=bm, dt=dt, adaptive=adaptive, rtol=rtol, atol=atol, dt_min=dt_min, options=options)
    if extra_solver_state is None:
        extra_solver_state = solver.init_extra_solver_state(ts[0], y0)
    ys, extra_solver_state = solver.integrate(y0, ts, extra_solver_state)
    if extra:
        return ys, extra_solver_state
    return ys


def sdeint_adjoint(*a, **k):
    raise NotImplementedError("shim: adjoint=false in the reference config (yml:41)")
