"""Fused decoder heads: ``self.decoder(sol_y)`` and ``self.scale(sol_y)`` of the reference's ``SDEDecoder.forward``
(models/decoders/dec_hivt_nusargo_sde.py:50-61, 96, 98) as ONE launch that reads the solver's latents once.

    loc, scale_raw = decoder_heads(dec.decoder, dec.scale, sol_y)        # sol_y [rows, T, 64], any row / time strides
    scale = F.elu_(scale_raw, alpha=1.0) + 1.0 + min_scale               # :98-99 stays with the caller

Forward: tensor-core kernel (csrc/heads.cu).  Under autograd the call is differentiable: the backward (csrc/heads_bwd.cu) processes only
the (point, head) pairs that received a gradient — the reference's winner-takes-all L2 loss (losses/L2.py:12-20) reaches ~5 % of them —
and returns dL/dsol_y plus the gradients of all head parameters.  ``decoder_heads_from_solution`` takes the solver's full ``ys``
(slab 0 = y0 included) so that no slice-backward copy of the 3 GB gradient is needed.
"""
import ctypes as C
from typing import List, Optional, Tuple

import torch

from . import _lib
from .ops import LAUNCHES, _stream_ptr

_HEAD_SHAPES = [(64, 64), (64,), (64,), (64,), (2, 64), (2,)]


def head_params(head: torch.nn.Module) -> List[torch.Tensor]:
    """[w1, b1, ln_g, ln_b, w2, b2] of ``nn.Sequential(Linear(64,64), LayerNorm(64), ReLU, Linear(64,2))`` (dec…sde.py:50-54)."""
    lin1, ln, lin2 = head[0], head[1], head[3]
    ps = [lin1.weight, lin1.bias, ln.weight, ln.bias, lin2.weight, lin2.bias]
    if any(tuple(p.shape) != sh for p, sh in zip(ps, _HEAD_SHAPES)):
        raise NotImplementedError("fused decoder heads support Linear(64,64) / LayerNorm(64) / ReLU / Linear(64,2) only")
    return ps


def _heads_fwd_impl(x: torch.Tensor, params: List[torch.Tensor], n_heads: int, ln_eps: float,
                    cat4: Optional[float] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """x[rows, T, 64] (last dim unit-stride) -> (out0[rows, T, 2], out1[rows, T, 2]); out1 is empty when n_heads == 1.
    ``cat4`` = the decoder's ``min_scale``: out0 is the stage's result ``cat(loc, elu(scale) + 1 + min_scale)`` [rows, T, 4]
    (dec_hivt_nusargo_sde.py:98-100, TRAJSDE_HEADS_FLAG_CAT4) and out1 is empty."""
    if x.dim() != 3 or x.shape[2] != 64 or x.dtype != torch.float32 or (x.numel() > 0 and x.stride(2) != 1):
        raise ValueError("`sol_y` must be float32 of shape (rows, T, 64) with a unit-stride last dimension")
    if n_heads not in (1, 2) or len(params) != 6 * n_heads:
        raise ValueError("params: 6 tensors per head")
    rows, T = x.shape[0], x.shape[1]
    dev = x.device
    ps = [p.detach().contiguous() for p in params]
    a = _lib.HeadsArgs()
    a.struct_bytes = C.sizeof(_lib.HeadsArgs)
    if cat4 is not None and n_heads != 2:
        raise ValueError("the fused out['loc'] result needs both heads (uncertain=True)")
    a.mode, a.rows, a.dim, a.flags, a.n_t, a.n_heads = _lib.MODE_TC_F16, rows, 64, (_lib.HEADS_FLAG_CAT4 if cat4 is not None else 0), T, n_heads
    for h in range(n_heads):
        for name, t in zip(('w1', 'b1', 'ln_g', 'ln_b', 'w2', 'b2'), ps[6 * h:6 * h + 6]):
            setattr(a.head[h], name, t.data_ptr())
    a.ln_eps = ln_eps
    a.min_scale = 0.0 if cat4 is None else float(cat4)
    xd = x.detach()
    if rows > 0 and T > 0 and (xd.data_ptr() % 16 != 0 or xd.stride(0) % 4 != 0 or xd.stride(1) % 4 != 0):
        xd = xd.contiguous()
    a.x, a.x_row_stride, a.x_t_stride = xd.data_ptr(), max(xd.stride(0), 64), max(xd.stride(1), 64)
    if cat4 is not None:
        outs = [torch.empty((rows, T, 4), dtype=torch.float32, device=dev), torch.empty((0, T, 2), dtype=torch.float32, device=dev)]
        a.out[0] = outs[0].data_ptr()
    else:
        outs = [torch.empty((rows, T, 2), dtype=torch.float32, device=dev),
                torch.empty((rows, T, 2) if n_heads == 2 else (0, T, 2), dtype=torch.float32, device=dev)]
        for h in range(n_heads):
            a.out[h] = outs[h].data_ptr()
    L = _lib.lib()
    need = _lib.check(L.trajsde_heads_workspace_bytes(_lib.MODE_TC_F16), "trajsde_heads_workspace_bytes")
    ws = torch.empty((need,), dtype=torch.uint8, device=dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), need
    with torch.cuda.device(dev):
        _lib.check(L.trajsde_heads_fwd(C.byref(a), _stream_ptr(dev)), "trajsde_heads_fwd")
    if rows > 0 and T > 0:
        LAUNCHES['n'] += 2
    return outs[0], outs[1]


heads_fwd = torch.library.custom_op("trajsde::heads_fwd", _heads_fwd_impl, mutates_args=(), device_types="cuda")


@heads_fwd.register_fake
def _(x, params, n_heads, ln_eps, cat4=None):
    rows, T = x.shape[0], x.shape[1]
    if cat4 is not None:
        return x.new_empty((rows, T, 4)), x.new_empty((0, T, 2))
    return x.new_empty((rows, T, 2)), x.new_empty((rows, T, 2) if n_heads == 2 else (0, T, 2))


def _heads_bwd_impl(x: torch.Tensor, params: List[torch.Tensor], n_heads: int, ln_eps: float, grad0: Optional[torch.Tensor],
                    grad1: Optional[torch.Tensor], grad_x: torch.Tensor, row_flags: Optional[torch.Tensor] = None,
                    cat4: bool = False, grad_amax: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
    """Gradients of the 6 * n_heads head tensors; dL/dx is ACCUMULATED into the zero-filled ``grad_x`` (same shape as x, any row / t
    strides, unit channel stride) for the points that carry a non-zero dL/dout.  ``grad_h`` [rows, T, 2] or None per head.
    ``cat4``: ``grad0`` is dL/d out['loc'] [rows, T, 4] of the fused result (channels 2..3 pass through the ELU derivative)."""
    rows, T = x.shape[0], x.shape[1]
    dev = x.device
    ps = [p.detach().contiguous() for p in params]
    gps = [torch.empty_like(p) for p in ps]
    a = _lib.HeadsBwdArgs()
    a.struct_bytes = C.sizeof(_lib.HeadsBwdArgs)
    a.mode, a.rows, a.dim, a.flags, a.n_t, a.n_heads = _lib.MODE_TC_F16, rows, 64, (_lib.HEADS_FLAG_CAT4 if cat4 else 0), T, n_heads
    for h in range(n_heads):
        for name, t, g in zip(('w1', 'b1', 'ln_g', 'ln_b', 'w2', 'b2'), ps[6 * h:6 * h + 6], gps[6 * h:6 * h + 6]):
            setattr(a.head[h], name, t.data_ptr())
            setattr(a.grad_head[h], name, g.data_ptr())
    a.ln_eps = ln_eps
    xd = x.detach()
    if rows > 0 and T > 0 and (xd.data_ptr() % 16 != 0 or xd.stride(0) % 4 != 0 or xd.stride(1) % 4 != 0 or xd.stride(2) != 1):
        xd = xd.contiguous()
    a.x, a.x_row_stride, a.x_t_stride = xd.data_ptr(), max(xd.stride(0), 64), max(xd.stride(1), 64)
    keep = []
    for h, g in enumerate((grad0, grad1)[:n_heads]):
        if g is not None:
            g = g.contiguous()
            keep.append(g)
            a.grad_out[h] = g.data_ptr()
    if grad_x.stride(2) != 1 or grad_x.stride(0) % 4 or grad_x.stride(1) % 4 or (grad_x.numel() and grad_x.data_ptr() % 16):
        raise ValueError("grad_x: unit channel stride, 16-byte aligned, row / t strides multiples of 4 elements")
    a.grad_x, a.gx_row_stride, a.gx_t_stride = grad_x.data_ptr(), max(grad_x.stride(0), 64), max(grad_x.stride(1), 64)
    if row_flags is not None:      # grad_x may be uninitialised: the call flags the rows that carry a gradient and zero-fills only those
        a.row_flags = row_flags.data_ptr()
    if grad_amax is not None:      # float32 [1]: receives max |value written to grad_x| (the solver backward's loss-scale input)
        a.grad_amax = grad_amax.data_ptr()
    L = _lib.lib()
    need = _lib.check(L.trajsde_heads_bwd_workspace_bytes(_lib.MODE_TC_F16), "trajsde_heads_bwd_workspace_bytes")
    ws = torch.empty((need,), dtype=torch.uint8, device=dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), need
    with torch.cuda.device(dev):
        _lib.check(L.trajsde_heads_bwd(C.byref(a), _stream_ptr(dev)), "trajsde_heads_bwd")
    LAUNCHES['n'] += 4 if row_flags is not None else 2
    return gps


class _HeadsFn(torch.autograd.Function):
    """Both heads on ``sol_y`` [rows, T, 64] (a view is fine).  ``full``: the first argument is the solver's whole ``ys`` [T+1, rows, 64]
    and the heads read ``ys[1:].permute(1, 0, 2)`` — the gradient is then produced directly in ``ys``'s shape and layout (slab 0 zero),
    so autograd has no slice-backward copy to make."""

    @staticmethod
    def forward(ctx, x, full, n_heads, ln_eps, cat4, *params):
        sol = x[1:].permute(1, 0, 2) if full else x
        o0, o1 = _heads_fwd_impl(sol, list(params), n_heads, ln_eps, cat4)
        ctx.save_for_backward(x, *params)
        ctx.meta = (full, n_heads, ln_eps, cat4 is not None)
        ctx.set_materialize_grads(False)
        return o0, o1

    @staticmethod
    def backward(ctx, g0, g1):
        x, *params = ctx.saved_tensors
        full, n_heads, ln_eps, cat4 = ctx.meta
        if full:        # ys is dense in one of the two storage layouts: the gradient takes the same strides (slab 0 stays zero)
            gfull = torch.empty_strided(x.size(), x.stride(), dtype=x.dtype, device=x.device).zero_()
        else:
            gfull = torch.zeros(x.shape, dtype=x.dtype, device=x.device)
        sol = x[1:].permute(1, 0, 2) if full else x
        gsol = gfull[1:].permute(1, 0, 2) if full else gfull
        gps = _heads_bwd_impl(sol, list(params), n_heads, ln_eps, g0, g1 if n_heads == 2 and not cat4 else None, gsol, cat4=cat4)
        return (gfull, None, None, None, None) + tuple(gps)


class _SolveHeadsFn(torch.autograd.Function):
    """The decoder's solve and both heads as ONE autograd node (what "heads as an epilogue of the solver" means for training): the
    solution ``ys`` never becomes an autograd tensor, so its 3 GB gradient is an internal buffer — the heads backward writes only the
    rows that carry a gradient (a winner-takes-all loss reaches one mode in ten) and hands the solver backward the list of those rows;
    no zero-fill of the rest, no scan for it, no slice / permute copies."""

    @staticmethod
    def forward(ctx, y0, step_tab, out_begin, out_w, n_outputs, seed, row_offset, mode, n_heads, ln_eps, n_sde, cat4, *params):
        from . import ops
        sde_params, head_params_ = list(params[:n_sde]), list(params[n_sde:])
        ys, _, states = ops._euler_fwd_impl(y0, sde_params, step_tab, out_begin, out_w, n_outputs, None, None, seed, row_offset, 0, mode, True, True)
        o0, o1 = _heads_fwd_impl(ys[1:].permute(1, 0, 2), head_params_, n_heads, ln_eps, cat4)
        ctx.save_for_backward(ys, states, step_tab, out_begin, out_w, *params)
        ctx.meta = (n_outputs, seed, row_offset, mode, n_heads, ln_eps, n_sde, cat4 is not None)
        ctx.set_materialize_grads(False)
        return o0, o1

    @staticmethod
    def backward(ctx, g0, g1):
        from . import ops
        ys, states, step_tab, out_begin, out_w, *params = ctx.saved_tensors
        n_outputs, seed, row_offset, mode, n_heads, ln_eps, n_sde, cat4 = ctx.meta
        sde_params, head_params_ = list(params[:n_sde]), list(params[n_sde:])
        rows = ys.shape[1]
        sparse = ops.row_flags_supported(mode, False) and bool(ops.SKIP_ZERO_ROWS) and rows > 0
        if sparse:          # uninitialised gradient buffer in ys's layout: only flagged rows are ever written or read
            gys = torch.empty_strided(ys.size(), ys.stride(), dtype=ys.dtype, device=ys.device)
            gys[0].zero_()                                  # dL/dys[0] = 0 (the heads read ys[1:]); 256 B per row
            flags = torch.empty((rows,), dtype=torch.uint8, device=ys.device)
            amax = torch.empty((1,), dtype=torch.float32, device=ys.device)      # max |dL/dys|, written by the heads backward
        else:
            gys = torch.empty_strided(ys.size(), ys.stride(), dtype=ys.dtype, device=ys.device).zero_()
            flags = amax = None
        gps = _heads_bwd_impl(ys[1:].permute(1, 0, 2), head_params_, n_heads, ln_eps, g0, g1 if n_heads == 2 and not cat4 else None,
                              gys[1:].permute(1, 0, 2), flags, cat4=cat4, grad_amax=amax)
        grads = ops._euler_bwd_impl(gys, None, states, sde_params, step_tab, out_begin, out_w, n_outputs, None, None, seed, row_offset, 0,
                                    mode, flags, amax)
        return (grads[0],) + (None,) * 11 + tuple(grads[1:]) + tuple(gps)


def solve_and_heads(sde, loc_head: torch.nn.Module, scale_head: Optional[torch.nn.Module], y0: torch.Tensor, ts, dt: float, *,
                    mode: Optional[str] = None, seed: Optional[int] = None, row_offset: int = 0, bm=None,
                    cat_min_scale: Optional[float] = None):
    """``loc, scale_raw = heads(sdeint(sde, y0, ts, dt=dt, method='euler')[1:].permute(1, 0, 2))`` — dec_hivt_nusargo_sde.py:88, 95, 98 —
    as one differentiable node (in-kernel Philox noise).  With ``bm`` (caller-supplied increments: validation) or without autograd it
    is the plain composition of ``sdeint`` and ``decoder_heads_from_solution``.
    ``cat_min_scale`` (the decoder's ``min_scale``; needs ``scale_head``): returns ``(out4, None)`` with
    ``out4 = cat(loc, elu(scale_raw) + 1 + min_scale)`` [rows, T, 4] — lines :98-100 written by the heads kernel itself, their
    backward folded into the heads backward."""
    from . import ops, solver
    from .schedule import euler_schedule
    heads, hparams, eps = _head_args(loc_head, scale_head)
    sde_params = solver._mlp_params(sde.f_func, 64, 'f_func') + solver._mlp_params(sde.g_func, 1, 'g_func')
    need_grad = torch.is_grad_enabled() and (y0.requires_grad or any(p.requires_grad for p in sde_params + hparams))
    if cat_min_scale is not None and scale_head is None:
        raise ValueError("cat_min_scale needs the scale head (uncertain=True)")
    if bm is not None or not need_grad:
        ys = solver.sdeint(sde, y0, ts, bm=bm, dt=dt, method='euler', mode=mode, seed=seed, row_offset=row_offset, rows_major=True)
        return decoder_heads_from_solution(loc_head, scale_head, ys, cat_min_scale=cat_min_scale)
    solver._check_sde_types(sde)
    if not y0.is_cuda or y0.dim() != 2 or y0.shape[1] != 64:
        raise RuntimeError("trajsde_b200 runs on CUDA (sm_100a) only, y0 [rows, 64]; there is no CPU fallback")
    sched = euler_schedule(solver._as_ts(ts, y0), float(dt))
    ds = ops.DeviceSchedule.get(sched, y0.device)
    mode_id = _lib.MODES[mode or solver.get_default_mode()]
    seed = solver._next_call_seed() if seed is None else int(seed)
    o0, o1 = _SolveHeadsFn.apply(y0, ds.step_tab, ds.out_begin, ds.out_w, sched.n_outputs, seed, int(row_offset), mode_id, len(heads), eps,
                                 len(sde_params), None if cat_min_scale is None else float(cat_min_scale), *sde_params, *hparams)
    for name in ('fnfe', 'gnfe'):
        if hasattr(sde, name):
            setattr(sde, name, getattr(sde, name) + sched.n_steps)
    return o0, (o1 if scale_head is not None and cat_min_scale is None else None)


def _needs_grad(x: torch.Tensor, heads) -> bool:
    return torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for h in heads for p in h.parameters()))


def _head_args(loc_head, scale_head):
    heads = [loc_head] + ([scale_head] if scale_head is not None else [])
    eps = {float(h[1].eps) for h in heads}
    if len(eps) != 1:
        raise NotImplementedError("both heads must share the LayerNorm eps")
    return heads, [p for h in heads for p in head_params(h)], eps.pop()


def decoder_heads(loc_head: torch.nn.Module, scale_head: Optional[torch.nn.Module], sol_y: torch.Tensor,
                  cat_min_scale: Optional[float] = None):
    """(loc[rows,T,2], scale_raw[rows,T,2] or None) = (loc_head(sol_y), scale_head(sol_y)) in one fused launch; differentiable.
    ``cat_min_scale``: (``cat(loc, elu(scale_raw) + 1 + min_scale)`` [rows,T,4], None) instead — the stage's out['loc'] (dec…sde.py:98-100)."""
    if not sol_y.is_cuda:
        raise RuntimeError("trajsde_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    if cat_min_scale is not None and scale_head is None:
        raise ValueError("cat_min_scale needs the scale head (uncertain=True)")
    heads, params, eps = _head_args(loc_head, scale_head)
    cat4 = None if cat_min_scale is None else float(cat_min_scale)
    if _needs_grad(sol_y, heads):
        o0, o1 = _HeadsFn.apply(sol_y, False, len(heads), eps, cat4, *params)
    else:
        o0, o1 = _heads_fwd_impl(sol_y, params, len(heads), eps, cat4)
    return o0, (o1 if scale_head is not None and cat4 is None else None)


def decoder_heads_from_solution(loc_head: torch.nn.Module, scale_head: Optional[torch.nn.Module], ys: torch.Tensor,
                                cat_min_scale: Optional[float] = None):
    """The same on the solver's full output ``ys`` [T+1, rows, 64] (``sdeint``'s return value, either storage layout): the heads read
    ``ys[1:].permute(1, 0, 2)`` (dec…sde.py:88) and, in training, hand dL/dys back in ``ys``'s own layout — no slice-backward copy."""
    if not ys.is_cuda:
        raise RuntimeError("trajsde_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    if cat_min_scale is not None and scale_head is None:
        raise ValueError("cat_min_scale needs the scale head (uncertain=True)")
    heads, params, eps = _head_args(loc_head, scale_head)
    cat4 = None if cat_min_scale is None else float(cat_min_scale)
    if _needs_grad(ys, heads):
        o0, o1 = _HeadsFn.apply(ys, True, len(heads), eps, cat4, *params)
    else:
        o0, o1 = _heads_fwd_impl(ys[1:].permute(1, 0, 2), params, len(heads), eps, cat4)
    return o0, (o1 if scale_head is not None and cat4 is None else None)


class FusedHeadPair:
    """No-edit drop-in for ``SDEDecoder.forward``'s two head calls: ``install_heads`` binds ``loc_forward`` / ``scale_forward`` as the
    ``forward`` of the decoder's ``self.decoder`` / ``self.scale`` instances.  The first call on a given ``sol_y`` runs the fused
    launch for BOTH heads and keeps the other head's result for the call that follows (dec…sde.py:96 then :98); under autograd both
    results are outputs of one differentiable node.  Inputs the kernels do not serve (non-CUDA, not 3-D) go to ``nn.Sequential.forward``."""

    def __init__(self, loc_head, scale_head):
        self.loc_head, self.scale_head = loc_head, scale_head
        self._key, self._pending = None, None

    @staticmethod
    def _key_of(x):
        # inference tensors (torch.inference_mode(): Lightning's validate / test / predict loops) do not track a version counter
        version = None if x.is_inference() else x._version
        return (x.data_ptr(), version, tuple(x.shape), tuple(x.stride()), torch.is_grad_enabled())

    @staticmethod
    def _served(x):
        return x.is_cuda and x.dim() == 3 and x.shape[2] == 64 and x.dtype == torch.float32

    def _both(self, x):
        return decoder_heads(self.loc_head, self.scale_head, x)

    def loc_forward(self, x):
        if not self._served(x):
            return torch.nn.Sequential.forward(self.loc_head, x)
        if self._key == self._key_of(x) and self._pending is not None and self._pending[0] == 'loc':
            out, self._key, self._pending = self._pending[1], None, None
            return out
        loc, scale = self._both(x)
        self._key, self._pending = self._key_of(x), ('scale', scale)
        return loc

    def scale_forward(self, x):
        if not self._served(x):
            return torch.nn.Sequential.forward(self.scale_head, x)
        if self._key == self._key_of(x) and self._pending is not None and self._pending[0] == 'scale':
            out, self._key, self._pending = self._pending[1], None, None
            return out
        loc, scale = self._both(x)
        self._key, self._pending = self._key_of(x), ('loc', loc)
        return scale


def install_heads(decoder) -> dict:
    """Bind the fused pair on ``decoder.decoder`` / ``decoder.scale`` (instance attributes; classes untouched).  Returns what
    ``uninstall_heads`` needs.  No-op (empty dict) when the decoder has no ``scale`` head (``uncertain: false``) or other widths."""
    loc, scale = getattr(decoder, 'decoder', None), getattr(decoder, 'scale', None)
    if loc is None or scale is None:
        return {}
    try:
        head_params(loc), head_params(scale)
    except (NotImplementedError, IndexError, AttributeError, TypeError):
        return {}
    pair = FusedHeadPair(loc, scale)
    saved = {'loc': (loc, loc.__dict__.get('forward')), 'scale': (scale, scale.__dict__.get('forward')), 'pair': pair}
    loc.forward = pair.loc_forward
    scale.forward = pair.scale_forward
    return saved


def uninstall_heads(saved: Optional[dict]) -> None:
    for key in ('loc', 'scale'):
        if saved and key in saved:
            mod, orig = saved[key]
            if orig is None:
                mod.__dict__.pop('forward', None)
            else:
                mod.forward = orig
