"""Drop-in replacements for the reference's two solver call sites.

    sdeint(sde, y0, ts, ...)                 <- ``from torchsde import sdeint``            dec_hivt_nusargo_sde.py:11,88
    sdeint_dual(sde, y0, ts, nus_mask, ...)  <- ``from models.utils.sdeint import sdeint_dual``  enc…sep2.py:23,149,274
                                                (definition models/utils/sdeint.py:110-197)

Same names, positional/keyword arguments and error behaviour (ValueError for contract violations, sdeint.py:836-971);
configurations the reference never uses (adaptive, logqp, non-Euler, non-diagonal) raise NotImplementedError — there is
no silent fallback to another solver or to the CPU.
"""
import threading
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops
from .schedule import EulerSchedule, euler_schedule

_defaults = {'mode': 'tc_f16', 'seed': None, 'torch_seed': None, 'calls': 0}
_seed_lock = threading.Lock()
_MASK63 = 2**63 - 1


def set_default_mode(mode: str) -> None:
    """'exact' (fp32 FFMA validation kernel) or 'tc_f16' (tcgen05 tensor-core kernel, default)."""
    if mode not in _lib.MODES:
        raise ValueError(f"mode must be one of {sorted(_lib.MODES)}")
    _defaults['mode'] = mode


def get_default_mode() -> str:
    return _defaults['mode']


def manual_seed(seed: int) -> None:
    """Pin the base seed of the in-kernel Philox Brownian increments used when ``bm`` is None (every solver call draws an
    independent stream derived from (base seed, call index); the call index restarts at 0 here).

    Without this call the base seed follows torch: it is derived from ``torch.initial_seed()`` — so ``torch.manual_seed`` /
    ``pl.seed_everything`` make runs reproducible, and an unseeded process draws fresh noise every run like the reference's
    ``BrownianInterval(entropy=None)`` (models/utils/sdeint.py:983-984) — and, under ``torch.distributed``, from the rank, so
    data-parallel ranks do not replay each other's increments.  The op itself never consumes torch's global RNG."""
    with _seed_lock:
        _defaults['seed'], _defaults['torch_seed'], _defaults['calls'] = int(seed) & _MASK63, None, 0


def set_device_seed(word: Optional[torch.Tensor]) -> None:
    """Keep (part of) the Philox key in device memory: ``word`` = int64 CUDA tensor with one element, or None to go back to host
    seeds (for the word's device).  Every call on that device then draws its stream from (host seed of the call + word[0]).  This is
    what makes the path CUDA-graph friendly: capture a whole step once (the per-call host seeds become constants of the graph), then
    ``word += 1`` between replays gives every replay fresh Brownian increments.  Forward and backward of a step must see the same value:
    bump it after the backward."""
    from . import ops
    if word is None:
        ops.SEED_DEV.clear()
        return
    if not (torch.is_tensor(word) and word.is_cuda and word.dtype == torch.int64 and word.numel() == 1):
        raise ValueError("device seed: an int64 CUDA tensor with one element")
    ops.SEED_DEV[str(word.device)] = word


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & (2**64 - 1)
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
    return x ^ (x >> 31)


def _rank() -> int:
    import torch.distributed as dist
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def _next_call_seed() -> int:
    with _seed_lock:
        if _defaults['seed'] is None or _defaults['torch_seed'] is not None:
            ts = int(torch.initial_seed())                 # follows torch.manual_seed / seed_everything; re-derived when it changes
            if ts != _defaults['torch_seed']:
                _defaults['seed'] = _splitmix64(_splitmix64(ts & (2**64 - 1)) ^ (0xB200 + _rank())) & _MASK63
                _defaults['torch_seed'], _defaults['calls'] = ts, 0
        k = _defaults['calls']
        _defaults['calls'] = k + 1
        return (_defaults['seed'] + 0x9E3779B97F4A7C15 * (k + 1)) & _MASK63


# ---------------------------------------------------------------------------------------------------------------------
# sde object protocol: weights are read from the caller's modules on every call
# ---------------------------------------------------------------------------------------------------------------------
def _mlp_params(net: nn.Module, out_dim: int, what: str) -> List[torch.Tensor]:
    seq = getattr(net, 'net', None)
    if not isinstance(seq, nn.Sequential) or len(seq) != 5:
        raise NotImplementedError(f"{what}: expected `.net` = Sequential(Linear, Tanh, Linear, Tanh, Linear) "
                                  f"(dec_hivt_nusargo_sde.py:111-117 / enc…sep2.py:376-388 with sde_layers=2)")
    lin = [seq[0], seq[2], seq[4]]
    if not all(isinstance(m, nn.Linear) for m in lin) or not all(isinstance(seq[i], nn.Tanh) for i in (1, 3)):
        raise NotImplementedError(f"{what}: unsupported layer types in `.net`")
    want = [(64, 66), (64, 64), (out_dim, 64)]
    for m, shp in zip(lin, want):
        if tuple(m.weight.shape) != shp or m.bias is None:
            raise NotImplementedError(f"{what}: Linear{tuple(m.weight.shape)} unsupported, the fused op handles "
                                      f"embed_dim=64 nets {want} only")
    return [lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, lin[2].weight, lin[2].bias]


def _check_sde_types(sde):
    if not hasattr(sde, "noise_type"):
        raise ValueError("sde does not have the attribute noise_type.")              # sdeint.py:836-837
    if not hasattr(sde, "sde_type"):
        raise ValueError("sde does not have the attribute sde_type.")                # sdeint.py:842-843
    if sde.noise_type != 'diagonal' or sde.sde_type != 'ito':
        raise NotImplementedError(f"fused solve supports diagonal Ito SDEs only, got {sde.noise_type}/{sde.sde_type}")


def _common_checks(y0, adaptive, logqp, extra, names, options, extra_solver_state, unused):
    if unused:
        import warnings
        warnings.warn(f"`sdeint`: Unexpected arguments {unused}")                    # misc.handle_unused_kwargs
    if not torch.is_tensor(y0):
        raise ValueError("`y0` must be a torch.Tensor.")                             # sdeint.py:848-849
    if y0.dim() != 2:
        raise ValueError("`y0` must be a 2-dimensional tensor of shape (batch, channels).")   # :850-851
    if adaptive or logqp or extra or names or extra_solver_state is not None:
        raise NotImplementedError("fused solve: adaptive / logqp / extra / names / extra_solver_state are unsupported "
                                  "(the reference never enables them)")
    if not y0.is_cuda:
        raise RuntimeError("trajsde_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    if y0.shape[1] != 64:
        raise NotImplementedError("fused solve handles embed_dim == 64 only")


def _as_ts(ts, y0):
    if not torch.is_tensor(ts):
        if not isinstance(ts, (tuple, list)) or not all(isinstance(t, (float, int)) for t in ts):
            raise ValueError("Evaluation times `ts` must be a 1-D Tensor or list/tuple of floats.")   # sdeint.py:872-874
        ts = torch.tensor(ts, dtype=torch.float32)
    if ts.requires_grad:
        raise ValueError("Argument ts must not require gradient.")                   # misc.assert_no_grad
    return ts


def _materialise_bm(bm, sched: EulerSchedule, rows: int, device) -> torch.Tensor:
    """``bm`` protocol: a tensor dW[S,rows,64]; an object carrying ``.dW``; or a torchsde-style callable
    ``bm(t0, t1) -> [rows,64]`` which is queried once per schedule step like Euler.step does (sdeint.py:480)."""
    if torch.is_tensor(bm):
        dW = bm
    elif torch.is_tensor(getattr(bm, 'dW', None)):
        dW = bm.dW
    elif callable(bm):
        shp = tuple(getattr(bm, 'shape', (rows, 64)))
        if len(shp) != 2:
            raise ValueError("`bm` must be of shape (batch, noise_channels).")       # sdeint.py:886-887
        if shp != (rows, 64):
            raise ValueError("Batch sizes not consistent." if shp[0] != rows else "Noise sizes not consistent.")
        t0 = torch.from_numpy(sched.t0)
        t1 = t0 + torch.from_numpy(sched.h)
        dW = torch.stack([bm(t0[k], t1[k]) for k in range(sched.n_steps)])
    else:
        raise ValueError("`bm` must be a tensor dW[S,rows,64], expose `.dW`, or be callable bm(t0, t1)")
    if tuple(dW.shape) != (sched.n_steps, rows, 64):
        raise ValueError(f"Brownian increments must have shape ({sched.n_steps}, {rows}, 64), got {tuple(dW.shape)}")
    return dW.to(device=device, dtype=torch.float32)


def _solve(sde, params, y0, ts, dt, bm, nus_mask, mode, seed, row_offset, rows_major=False):
    sched = euler_schedule(ts, float(dt))
    dev = y0.device
    ds = ops.DeviceSchedule.get(sched, dev)
    dW = None if bm is None else _materialise_bm(bm, sched, y0.shape[0], dev)
    if seed is None:
        seed = _next_call_seed() if dW is None else 0
    need_grad = torch.is_grad_enabled() and (y0.requires_grad or any(p.requires_grad for p in params))
    mode_id = _lib.MODES[mode or _defaults['mode']]
    ys, g_last = ops.euler_call(y0, list(params), ds.step_tab, ds.out_begin, ds.out_w, sched.n_outputs, dW, nus_mask,
                                int(seed), int(row_offset), 0, mode_id, need_grad, bool(rows_major))
    for name in ('fnfe', 'gnfe'):                      # NFE counters the reference bumps per f/g call (dec…sde.py:177,193)
        if hasattr(sde, name):
            setattr(sde, name, getattr(sde, name) + sched.n_steps)
    return ys, g_last


def sdeint(sde, y0: torch.Tensor, ts, bm=None, method: Optional[str] = None, dt: float = 1e-3, adaptive: bool = False,
           rtol: float = 1e-5, atol: float = 1e-4, dt_min: float = 1e-5, options: Optional[Dict[str, Any]] = None,
           names: Optional[Dict[str, str]] = None, logqp: bool = False, extra: bool = False,
           extra_solver_state=None, *, mode: Optional[str] = None, seed: Optional[int] = None, row_offset: int = 0,
           rows_major: bool = False, **unused_kwargs) -> torch.Tensor:
    """torchsde.sdeint for the reference decoder (dec_hivt_nusargo_sde.py:88): returns ys[T, rows, 64], ys[0] == y0.

    Extensions (keyword-only): ``mode`` ('exact' | 'tc_f16'), ``seed`` (Philox seed when ``bm`` is None), ``row_offset``
    (global id of row 0, so scene-sharded ranks draw the noise of the unsharded batch), ``rows_major`` (store ys physically as
    [rows, T, 64]; the returned [T, rows, 64] view then makes the decoder's ``[1:].permute(1,0,2)`` unit-stride per row)."""
    _check_sde_types(sde)
    _common_checks(y0, adaptive, logqp, extra, names, options, extra_solver_state, unused_kwargs)
    if method != 'euler':
        raise NotImplementedError(f"fused solve implements method='euler' only (reference yml:76), got {method!r}")
    ts = _as_ts(ts, y0)
    if not (hasattr(sde, 'f_func') and hasattr(sde, 'g_func')):
        raise NotImplementedError("sde must expose `f_func` and `g_func` (decoder LSDEFunc, dec…sde.py:160-167)")
    params = _mlp_params(sde.f_func, 64, 'f_func') + _mlp_params(sde.g_func, 1, 'g_func')
    ys, _ = _solve(sde, params, y0, ts, dt, bm, None, mode, seed, row_offset, rows_major)
    return ys


def sdeint_dual(sde, y0: torch.Tensor, ts, nus_mask: torch.Tensor, bm=None, method: Optional[str] = None,
                dt: float = 1e-3, adaptive: bool = False, rtol: float = 1e-5, atol: float = 1e-4, dt_min: float = 1e-5,
                options: Optional[Dict[str, Any]] = None, names: Optional[Dict[str, str]] = None, logqp: bool = False,
                extra: bool = False, extra_solver_state=None, *, mode: Optional[str] = None, seed: Optional[int] = None,
                row_offset: int = 0, **unused_kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
    """The reference's vendored ``sdeint_dual`` (models/utils/sdeint.py:110-197): ``(ys[T,rows,64], g[rows,64])`` where
    ``g`` is the diffusion evaluated at the start of the last step, broadcast over the 64 channels as an expand view
    (reference: ``.repeat(1, 64)``, enc…sep2.py:480-481).  ``method`` is ignored like in the reference (:182)."""
    _check_sde_types(sde)
    _common_checks(y0, adaptive, logqp, extra, names, options, extra_solver_state, unused_kwargs)
    ts = _as_ts(ts, y0)
    if not (hasattr(sde, 'f_func') and hasattr(sde, 'g_nus') and hasattr(sde, 'g_argo')):
        raise NotImplementedError("sde must expose `f_func`, `g_nus`, `g_argo` (encoder LSDEFunc, enc…sep2.py:442-448)")
    if not torch.is_tensor(nus_mask) or nus_mask.dtype != torch.bool or tuple(nus_mask.shape) != (y0.shape[0],):
        raise ValueError("`nus_mask` must be a bool tensor of shape (batch,)")
    params = (_mlp_params(sde.f_func, 64, 'f_func') + _mlp_params(sde.g_nus, 1, 'g_nus') +
              _mlp_params(sde.g_argo, 1, 'g_argo'))
    ys, g_last = _solve(sde, params, y0, ts, dt, bm, nus_mask.to(y0.device), mode, seed, row_offset)
    return ys, g_last.unsqueeze(1).expand(-1, 64)
