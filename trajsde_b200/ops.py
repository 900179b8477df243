"""torch custom ops over the C ABI (include/trajsde_b200.h).  CUDA only, no CPU / eager fallback.

    trajsde::euler_fwd   fused Euler–Maruyama solve (all steps, one launch)        -> ys, g_last, states
    trajsde::euler_bwd   discretise-then-optimise backward (recompute from states)   -> grad_y0, grad_params...

The ops do not own parameters: they read the caller's ``nn.Linear`` tensors on every call (reference checkpoints load
unchanged) and return gradients for them through ``register_autograd``.
"""
import ctypes as C
import os as _os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .schedule import EulerSchedule

_N_MLP = 6  # w1 b1 w2 b2 w3 b3

# kernels launched by this library since import (bench.py reports the count inside its timed region)
LAUNCHES = {'n': 0}
_KERNELS_PER_FWD = {_lib.MODE_EXACT_F32: 1, _lib.MODE_TC_F16: 2}   # TC: weight-pack + persistent solve

# A/B switch (tests, bench): run the fp32 CUDA-core backward kernels even in 'tc_f16' mode (TRAJSDE_BWD_FLAG_EXACT_KERNELS)
BWD_EXACT_KERNELS = False
# A/B switch: run the encoder-recurrence backward as 2 launches per iteration instead of the single persistent launch
# (TRAJSDE_BWD_FLAG_PER_STEP_LAUNCHES; also the escape hatch when the status word reports TRAJSDE_STATUS_SWEEP_TIMEOUT)
ENC_BWD_PER_STEP = False
# TRAJSDE_BWD_FLAG_SKIP_ZERO_ROWS: the tensor-core backward of a single-diffusion solve (the decoder) first finds the rows whose incoming
# gradient is all zero (one scan of grad_ys on the device, no host sync) and sweeps only the others.  Exact (those rows contribute
# nothing), and worth ~8x under the reference's winner-takes-all L2 loss (losses/L2.py:17-20: 1 of 10 modes per actor gets a gradient);
# a sampled pre-scan recognises (nearly) dense cotangents and then skips the full scan.
SKIP_ZERO_ROWS = True

# per-device int32 status word the tensor-core backward kernels OR their TRAJSDE_STATUS_* bits into.  The word is never read on the
# hot path: after every tensor-core backward call a snapshot is copied to pinned host memory on the same stream (4 bytes, no sync)
# and the NEXT call into this module looks at the snapshot once its copy event has completed — so a clipped gradient surfaces within
# one optimizer step as a warning or an exception (policy below) without ever blocking the stream.
ADJOINT_RANGE_POLICIES = ('warn', 'raise', 'ignore')
_POLICY = {'adjoint_range': _os.environ.get('TRAJSDE_ADJOINT_RANGE', 'warn')}
if _POLICY['adjoint_range'] not in ADJOINT_RANGE_POLICIES:
    raise ValueError(f"TRAJSDE_ADJOINT_RANGE must be one of {ADJOINT_RANGE_POLICIES}")


class AdjointRangeError(FloatingPointError):
    """Raised (policy 'raise') when a tensor-core backward call reported TRAJSDE_STATUS_ADJOINT_RANGE."""


_ADJOINT_MSG = ("trajsde_b200: a tensor-core backward call reported TRAJSDE_STATUS_ADJOINT_RANGE — the loss-scaled adjoint outgrew the "
                "fp16 delta range, so the gradients of that step may be clipped.  Rerun the step with ops.BWD_EXACT_KERNELS = True or "
                "mode='exact' (policy: trajsde_b200.ops.set_adjoint_range_policy('warn' | 'raise' | 'ignore'), env TRAJSDE_ADJOINT_RANGE)")


class _StatusMonitor:
    def __init__(self, device: torch.device):
        self.word = torch.zeros((1,), dtype=torch.int32, device=device)
        self.host = torch.zeros((1,), dtype=torch.int32).pin_memory()
        self.event: Optional[torch.cuda.Event] = None
        self.seen = 0                                      # bits reported so far and not yet cleared by backward_status()

    def snapshot(self):
        """Enqueue a 4-byte D2H copy of the status word behind the backward kernels just launched (if none is pending)."""
        if _POLICY['adjoint_range'] == 'ignore' or self.event is not None or torch.cuda.is_current_stream_capturing():
            return      # inside a CUDA-graph capture the word still accumulates; poll it with backward_status() between replays
        self.host.copy_(self.word, non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record(torch.cuda.current_stream(self.word.device))

    def poll(self, block: bool = False):
        if self.event is None or torch.cuda.is_current_stream_capturing() or not (block or self.event.query()):
            return
        if block:
            self.event.synchronize()
        self.event = None
        v = int(self.host[0])
        new_bits = v & ~self.seen
        self.seen |= v
        if new_bits & _lib.STATUS_SWEEP_TIMEOUT:            # results of that call are invalid, whatever the policy
            self.word.zero_()
            self.seen = 0
            raise RuntimeError("trajsde_b200: trajsde_enc_bwd reported TRAJSDE_STATUS_SWEEP_TIMEOUT — the single-launch encoder backward was "
                               "not fully co-resident on the device (shared GPU?) and gave up after ~30 s; the gradients of that step are "
                               "invalid.  Set trajsde_b200.ops.ENC_BWD_PER_STEP = True and rerun the step.")
        if new_bits & _lib.STATUS_ADJOINT_RANGE:
            pol = _POLICY['adjoint_range']
            if pol == 'raise':
                self.word.zero_()
                self.seen = 0
                raise AdjointRangeError(_ADJOINT_MSG)
            if pol == 'warn':
                import warnings
                warnings.warn(_ADJOINT_MSG, RuntimeWarning, stacklevel=3)


_STATUS = {}


def _monitor(device) -> _StatusMonitor:
    m = _STATUS.get(str(device))
    if m is None:
        m = _StatusMonitor(torch.device(device))
        _STATUS[str(device)] = m
    return m


def _status_word(device) -> torch.Tensor:
    return _monitor(device).word


def set_adjoint_range_policy(policy: str) -> None:
    """What happens when a tensor-core backward reports a clipped adjoint: 'warn' (default, RuntimeWarning at the next call into the
    library), 'raise' (AdjointRangeError at the next call) or 'ignore' (poll ``backward_status`` yourself)."""
    if policy not in ADJOINT_RANGE_POLICIES:
        raise ValueError(f"policy must be one of {ADJOINT_RANGE_POLICIES}")
    _POLICY['adjoint_range'] = policy


def poll_status(device=None, block: bool = False) -> None:
    """Look at the pending status snapshots (non-blocking unless ``block``); called by every op of this module on entry, and by
    ``FlatGradBucket.all_reduce_mean`` / user code once per optimizer step with ``block=True`` for a same-step guarantee."""
    for key, m in list(_STATUS.items()):
        if device is None or key == str(torch.device(device)):
            m.poll(block)


def backward_status(device, clear: bool = True) -> int:
    """TRAJSDE_STATUS_* bits raised by tensor-core backward calls on ``device`` since the last clear (synchronises).
    Bit ``_lib.STATUS_ADJOINT_RANGE``: the loss-scaled adjoint outgrew the fp16 delta range — gradients may be clipped; rerun
    that step with ``ops.BWD_EXACT_KERNELS = True`` or ``mode='exact'``."""
    m = _monitor(torch.device(device))
    v = int(m.word.item())
    if clear:
        m.word.zero_()
        m.seen = 0
        m.event = None
    return v


# ---------------------------------------------------------------------------------------------------------------------
# device-resident schedule tables
# ---------------------------------------------------------------------------------------------------------------------
class DeviceSchedule:
    """TrajsdeSchedule tables uploaded once per (schedule, device)."""
    _cache = {}

    def __init__(self, sched: EulerSchedule, device: torch.device):
        self.n_steps, self.n_outputs = sched.n_steps, sched.n_outputs
        self.step_tab = torch.from_numpy(sched.step_tab()).to(device)
        self.out_begin = torch.from_numpy(sched.out_begin()).to(device)
        ow = sched.out_w() if sched.n_outputs else np.zeros((1, 2), np.float32)
        self.out_w = torch.from_numpy(ow).to(device)

    @classmethod
    def get(cls, sched: EulerSchedule, device: torch.device) -> "DeviceSchedule":
        # keyed by the schedule's CONTENT (step starts / sizes / output map), never by object identity
        key = (sched.t0.tobytes(), sched.h.tobytes(), sched.out_k.tobytes(), sched.w0.tobytes(), sched.w1.tobytes(), str(device))
        hit = cls._cache.get(key)
        if hit is None:
            if len(cls._cache) > 256:
                cls._cache.clear()
            hit = cls(sched, device)
            cls._cache[key] = hit
        return hit


# Device-resident noise seed (TrajsdeNoise.seed_dev): {device string: int64 tensor [1]}.  When set for a device every solver / encoder
# call on it keys its Philox stream by (host seed + that word): a captured CUDA graph then draws fresh noise on every replay once the
# word is bumped between replays (the host seed of each call is baked into the graph as an offset).  See solver.set_device_seed.
SEED_DEV = {}


def _apply_seed_dev(noise, dev) -> None:
    t = SEED_DEV.get(str(dev))
    if t is not None:
        noise.seed_dev = t.data_ptr()


def _capturing() -> bool:
    return torch.cuda.is_current_stream_capturing()


def row_flags_supported(mode: int, dual: bool) -> bool:
    """Can ``euler_bwd`` take the heads backward's row flags (and leave unflagged ``grad_ys`` rows unread)?"""
    return mode == _lib.MODE_TC_F16 and not dual and not BWD_EXACT_KERNELS


def _mlp_struct(ts: Sequence[torch.Tensor]) -> _lib.Mlp:
    m = _lib.Mlp()
    m.w1, m.b1, m.w2, m.b2, m.w3, m.b3 = (t.data_ptr() for t in ts)
    return m


def _check_params(params: Sequence[torch.Tensor], dual: bool, device):
    n = _N_MLP * (3 if dual else 2)
    if len(params) != n:
        raise ValueError(f"expected {n} parameter tensors, got {len(params)}")
    shapes = [(64, 66), (64,), (64, 64), (64,), (64, 64), (64,)] + [(64, 66), (64,), (64, 64), (64,), (1, 64), (1,)] * (2 if dual else 1)
    out = []
    for t, shp in zip(params, shapes):
        if tuple(t.shape) != shp:
            raise ValueError(f"parameter shape {tuple(t.shape)} != expected {shp} (only the reference's 66-64-64-64/1 nets are supported)")
        if t.dtype != torch.float32 or t.device != device:
            raise ValueError("parameters must be float32 on the same CUDA device as y0")
        out.append(t.detach().contiguous())
    return out


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


# ---------------------------------------------------------------------------------------------------------------------
# forward
# ---------------------------------------------------------------------------------------------------------------------
def _euler_fwd_impl(y0: torch.Tensor, params: List[torch.Tensor], step_tab: torch.Tensor, out_begin: torch.Tensor,
              out_w: torch.Tensor, n_outputs: int, dw: Optional[torch.Tensor], alt_mask: Optional[torch.Tensor],
              seed: int, row_offset: int, step_offset: int, mode: int, save_states: bool, rows_major: bool,
              ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """ys[T,rows,64] (T = n_outputs+1, ys[0] = y0), g_last[rows], states[S,rows,64] (empty unless save_states).
    ``rows_major``: allocate ys physically as [rows, T, 64] and return the [T, rows, 64] permuted view, so the reference's
    ``[1:].permute(1, 0, 2)`` (dec_hivt_nusargo_sde.py:88) hands unit-stride rows to the decoder heads."""
    if y0.dim() != 2 or y0.shape[1] != 64 or y0.dtype != torch.float32:
        raise ValueError("`y0` must be float32 of shape (rows, 64)")
    dev = y0.device
    _monitor(dev).poll()
    rows = y0.shape[0]
    S = step_tab.shape[0]
    dual = alt_mask is not None
    ps = _check_params(params, dual, dev)
    y0c = y0.detach()
    if y0c.stride(1) != 1 or y0c.stride(0) % 4 != 0 or y0c.data_ptr() % 16 != 0:
        y0c = y0c.contiguous()
    if rows_major:
        ys = torch.empty((rows, n_outputs + 1, 64), dtype=torch.float32, device=dev).permute(1, 0, 2)
    else:
        ys = torch.empty((n_outputs + 1, rows, 64), dtype=torch.float32, device=dev)
    g_last = torch.empty((rows,), dtype=torch.float32, device=dev)
    states = torch.empty((S if save_states else 0, rows, 64), dtype=torch.float32, device=dev)
    a = _lib.EulerFwdArgs()
    a.struct_bytes = C.sizeof(_lib.EulerFwdArgs)
    a.mode, a.rows, a.dim, a.flags = mode, rows, 64, 0
    a.sched.n_steps, a.sched.n_outputs = S, n_outputs
    a.sched.step_tab, a.sched.out_begin, a.sched.out_w = step_tab.data_ptr(), out_begin.data_ptr(), out_w.data_ptr()
    a.drift, a.diffusion = _mlp_struct(ps[0:6]), _mlp_struct(ps[6:12])
    mask_u8 = None
    if dual:
        if alt_mask.shape != (rows,) or alt_mask.dtype not in (torch.bool, torch.uint8):
            raise ValueError("`nus_mask` must be a bool tensor of shape (rows,)")
        mask_u8 = alt_mask.contiguous().view(torch.uint8)
        a.diffusion_alt = _mlp_struct(ps[12:18])
        a.alt_mask = mask_u8.data_ptr()
    if dw is not None:
        if tuple(dw.shape) != (S, rows, 64) or dw.dtype != torch.float32:
            raise ValueError(f"`dW` must be float32 of shape ({S}, {rows}, 64): one slab per schedule step")
        dw = dw.contiguous()
        a.noise.dw = dw.data_ptr()
    a.noise.seed, a.noise.row_offset, a.noise.step_offset = seed & (2**63 - 1), row_offset, step_offset
    _apply_seed_dev(a.noise, dev)
    a.y0, a.y0_row_stride = y0c.data_ptr(), y0c.stride(0)
    a.ys, a.ys_t_stride, a.ys_row_stride = ys.data_ptr(), ys.stride(0), ys.stride(1)
    a.g_last = g_last.data_ptr()
    a.states = states.data_ptr() if save_states and rows > 0 else None
    L = _lib.lib()
    need = _lib.check(L.trajsde_euler_fwd_workspace_bytes(mode, rows, S, int(dual)), "trajsde_euler_fwd_workspace_bytes")
    ws = torch.empty((max(need, 1),), dtype=torch.uint8, device=dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), need
    with torch.cuda.device(dev):
        _lib.check(L.trajsde_euler_fwd(C.byref(a), _stream_ptr(dev)), "trajsde_euler_fwd")
    if rows > 0:
        LAUNCHES['n'] += _KERNELS_PER_FWD.get(mode, 1)
    return ys, g_last, states


euler_fwd = torch.library.custom_op("trajsde::euler_fwd", _euler_fwd_impl, mutates_args=(), device_types="cuda")


@euler_fwd.register_fake
def _(y0, params, step_tab, out_begin, out_w, n_outputs, dw, alt_mask, seed, row_offset, step_offset, mode, save_states,
      rows_major):
    rows, S = y0.shape[0], step_tab.shape[0]
    ys = y0.new_empty((rows, n_outputs + 1, 64)).permute(1, 0, 2) if rows_major else y0.new_empty((n_outputs + 1, rows, 64))
    return (ys, y0.new_empty((rows,)),
            y0.new_empty((S if save_states else 0, rows, 64)))


# ---------------------------------------------------------------------------------------------------------------------
# backward
# ---------------------------------------------------------------------------------------------------------------------
def _euler_bwd_impl(grad_ys: Optional[torch.Tensor], grad_g: Optional[torch.Tensor], states: torch.Tensor,
              params: List[torch.Tensor], step_tab: torch.Tensor, out_begin: torch.Tensor, out_w: torch.Tensor,
              n_outputs: int, dw: Optional[torch.Tensor], alt_mask: Optional[torch.Tensor], seed: int, row_offset: int,
              step_offset: int, mode: int, row_flags: Optional[torch.Tensor] = None,
              grad_amax: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
    """[grad_y0] + gradients of every tensor in ``params`` (same order/shapes).  ``row_flags`` (uint8 [rows], from
    ``trajsde_heads_bwd``): rows flagged 0 have no incoming gradient and their ``grad_ys`` entries may be uninitialised — honoured by
    the tensor-core kernels of a single-diffusion solve only (the caller checks ``row_flags_supported``).  ``grad_amax`` (float32 [1],
    also from ``trajsde_heads_bwd``): max |grad_ys| of the flagged rows — with it the call scans no gradient at all before the sweep."""
    dev = states.device
    _monitor(dev).poll()
    S, rows = states.shape[0], states.shape[1]
    dual = alt_mask is not None
    ps = _check_params(params, dual, dev)
    grad_y0 = torch.empty((rows, 64), dtype=torch.float32, device=dev)
    gparams = [torch.empty_like(p) for p in ps]       # the fixed-order reduce writes every element
    a = _lib.EulerBwdArgs()
    a.struct_bytes = C.sizeof(_lib.EulerBwdArgs)
    skip_zero = bool(SKIP_ZERO_ROWS) and not dual and mode == _lib.MODE_TC_F16 and not BWD_EXACT_KERNELS and grad_ys is not None and rows >= 2048
    if row_flags is not None:
        if not row_flags_supported(mode, dual) or grad_ys is None:
            raise ValueError("row_flags need the tensor-core backward of a single-diffusion solve")
        skip_zero = True
        a.row_flags = row_flags.data_ptr()
        if grad_amax is not None:
            a.grad_amax = grad_amax.data_ptr()
    a.mode, a.rows, a.dim, a.flags = mode, rows, 64, (1 if BWD_EXACT_KERNELS else 0) | (2 if skip_zero else 0)
    a.sched.n_steps, a.sched.n_outputs = S, n_outputs
    a.sched.step_tab, a.sched.out_begin, a.sched.out_w = step_tab.data_ptr(), out_begin.data_ptr(), out_w.data_ptr()
    a.drift, a.diffusion = _mlp_struct(ps[0:6]), _mlp_struct(ps[6:12])
    a.grad_drift, a.grad_diffusion = _mlp_struct(gparams[0:6]), _mlp_struct(gparams[6:12])
    mask_u8 = None
    if dual:
        mask_u8 = alt_mask.contiguous().view(torch.uint8)
        a.diffusion_alt, a.grad_diffusion_alt = _mlp_struct(ps[12:18]), _mlp_struct(gparams[12:18])
        a.alt_mask = mask_u8.data_ptr()
    if dw is not None:
        dw = dw.contiguous()
        a.noise.dw = dw.data_ptr()
    a.noise.seed, a.noise.row_offset, a.noise.step_offset = seed & (2**63 - 1), row_offset, step_offset
    _apply_seed_dev(a.noise, dev)
    states = states.contiguous()
    a.states = states.data_ptr()
    if grad_ys is not None:
        if grad_ys.stride(2) != 1 or grad_ys.stride(0) % 4 or grad_ys.stride(1) % 4 or grad_ys.data_ptr() % 16:
            grad_ys = grad_ys.contiguous()
        a.grad_ys, a.grad_ys_t_stride, a.grad_ys_row_stride = grad_ys.data_ptr(), grad_ys.stride(0), grad_ys.stride(1)
    if grad_g is not None:
        grad_g = grad_g.contiguous()
        a.grad_g_last = grad_g.data_ptr()
    a.grad_y0 = grad_y0.data_ptr()
    a.status = _status_word(dev).data_ptr()
    L = _lib.lib()
    need = _lib.check(L.trajsde_euler_bwd_workspace_bytes(mode, rows, S, int(dual)), "trajsde_euler_bwd_workspace_bytes")
    ws = torch.empty((max(need, 1),), dtype=torch.uint8, device=dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), need
    with torch.cuda.device(dev):
        _lib.check(L.trajsde_euler_bwd(C.byref(a), _stream_ptr(dev)), "trajsde_euler_bwd")
        if rows > 0 and mode == _lib.MODE_TC_F16 and not BWD_EXACT_KERNELS:
            _monitor(dev).snapshot()
    if rows > 0:
        tc = mode == _lib.MODE_TC_F16 and not BWD_EXACT_KERNELS
        # TC: pack + absmax + fused dgrad/wgrad + reduce; exact: (dgrad + wgrad) per diffusion net + reduce
        sampled = tc and grad_ys is not None and rows * grad_ys.shape[0] >= (1 << 16)   # sampled absmax + its conditional full scan
        if skip_zero:
            # row activity (sampled + full, or flagged rows only) + compaction + pack + fused dgrad/wgrad + reduce
            LAUNCHES['n'] += (4 if grad_amax is not None and grad_g is None else 5) if row_flags is not None else 6
        else:
            LAUNCHES['n'] += ((5 if dual else 4) + int(sampled)) if tc else (5 if dual else 3)
    return [grad_y0] + gparams


euler_bwd = torch.library.custom_op("trajsde::euler_bwd", _euler_bwd_impl, mutates_args=(), device_types="cuda")


@euler_bwd.register_fake
def _(grad_ys, grad_g, states, params, step_tab, out_begin, out_w, n_outputs, dw, alt_mask, seed, row_offset,
      step_offset, mode, row_flags=None):
    return [states.new_empty((states.shape[1], 64))] + [torch.empty_like(p) for p in params]


def _setup_context(ctx, inputs, output):
    (y0, params, step_tab, out_begin, out_w, n_outputs, dw, alt_mask, seed, row_offset, step_offset, mode,
     save_states, _rows_major) = inputs
    _, _, states = output
    ctx.has_states = bool(save_states)
    ctx.save_for_backward(states, step_tab, out_begin, out_w, *params)
    ctx.dw, ctx.alt_mask = dw, alt_mask
    ctx.meta = (n_outputs, seed, row_offset, step_offset, mode)
    ctx.n_params = len(params)


def _backward(ctx, grad_ys, grad_g, grad_states):
    if not ctx.has_states:
        raise RuntimeError("trajsde::euler_fwd was run with save_states=False; backward needs the saved step states")
    states, step_tab, out_begin, out_w, *params = ctx.saved_tensors
    n_outputs, seed, row_offset, step_offset, mode = ctx.meta
    grads = euler_bwd(grad_ys, grad_g, states, list(params), step_tab, out_begin, out_w, n_outputs, ctx.dw, ctx.alt_mask,
                      seed, row_offset, step_offset, mode)
    return (grads[0], list(grads[1:]), None, None, None, None, None, None, None, None, None, None, None, None)


euler_fwd.register_autograd(_backward, setup_context=_setup_context)


# ---------------------------------------------------------------------------------------------------------------------
# fused encoder recurrence: one forward launch, one backward call (reverse sweep enqueued by the library)
# ---------------------------------------------------------------------------------------------------------------------
def _enc_fwd_impl(h0: torch.Tensor, aa_out: torch.Tensor, obs_mask: torch.Tensor, slot: torch.Tensor, params: List[torch.Tensor],
            gru_params: List[torch.Tensor], step_tab: torch.Tensor, dw: Optional[torch.Tensor],
            alt_mask: Optional[torch.Tensor], seed: int, row_offset: int, step_offset: int, save_y1: bool,
            ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """latent[S,rows,64] (post-GRU state of every iteration), g[S,rows] (pre-step diffusion of every iteration),
    y1[S,rows,64] (pre-GRU state of every iteration; empty unless ``save_y1`` — the backward needs it)."""
    dev = h0.device
    _monitor(dev).poll()
    rows, S = h0.shape[0], step_tab.shape[0]
    dual = alt_mask is not None
    ps = _check_params(params, dual, dev)
    gshapes = [(64, 128), (64,), (64, 64), (64,)] * 3
    if len(gru_params) != 12 or any(tuple(t.shape) != sh for t, sh in zip(gru_params, gshapes)):
        raise ValueError("gru_params must be the 12 GRU_Unit tensors (update/reset/new_state: Linear(128,64), Linear(64,64))")
    gs = [t.detach().contiguous() for t in gru_params]
    if aa_out.dim() != 3 or aa_out.shape[1] != rows or aa_out.shape[2] != 64 or aa_out.dtype != torch.float32:
        raise ValueError("`aa_out` must be float32 [n_slots, rows, 64]")
    n_slots = aa_out.shape[0]
    if tuple(obs_mask.shape) != (rows, n_slots) or obs_mask.dtype not in (torch.bool, torch.uint8):
        raise ValueError("`actors_mask` must be bool [rows, n_slots]")
    if slot.shape != (S,) or slot.dtype != torch.int32:
        raise ValueError("`slot` must be int32 [n_steps]")
    h0c = h0.detach()
    if h0c.stride(1) != 1 or h0c.stride(0) % 4 != 0 or h0c.data_ptr() % 16 != 0:
        h0c = h0c.contiguous()
    aa = aa_out.detach().contiguous()
    om = obs_mask.contiguous().view(torch.uint8)
    latent = torch.empty((S, rows, 64), dtype=torch.float32, device=dev)
    g_out = torch.empty((S, rows), dtype=torch.float32, device=dev)
    y1s = torch.empty((S if save_y1 else 0, rows, 64), dtype=torch.float32, device=dev)
    a = _lib.EncFwdArgs()
    a.struct_bytes = C.sizeof(_lib.EncFwdArgs)
    a.mode, a.rows, a.dim, a.flags = _lib.MODE_TC_F16, rows, 64, 0
    a.sched.n_steps, a.sched.n_outputs = S, 0
    a.sched.step_tab = step_tab.data_ptr()
    a.drift, a.diffusion = _mlp_struct(ps[0:6]), _mlp_struct(ps[6:12])
    mask_u8 = None
    if dual:
        mask_u8 = alt_mask.contiguous().view(torch.uint8)
        a.diffusion_alt = _mlp_struct(ps[12:18])
        a.alt_mask = mask_u8.data_ptr()
    for name, t in zip(('u1', 'ub1', 'u2', 'ub2', 'r1', 'rb1', 'r2', 'rb2', 'n1', 'nb1', 'n2', 'nb2'), gs):
        setattr(a.gru, name, t.data_ptr())
    if dw is not None:
        if tuple(dw.shape) != (S, rows, 64) or dw.dtype != torch.float32:
            raise ValueError(f"`dW` must be float32 of shape ({S}, {rows}, 64)")
        dw = dw.contiguous()
        a.noise.dw = dw.data_ptr()
    a.noise.seed, a.noise.row_offset, a.noise.step_offset = seed & (2**63 - 1), row_offset, step_offset
    _apply_seed_dev(a.noise, dev)
    a.h0, a.h0_row_stride = h0c.data_ptr(), h0c.stride(0)
    a.aa_out, a.slot = aa.data_ptr(), slot.data_ptr()
    a.obs_mask, a.obs_mask_row_stride = om.data_ptr(), om.stride(0)
    a.latent, a.g_out = latent.data_ptr(), g_out.data_ptr()
    a.y1_out = y1s.data_ptr() if save_y1 and rows > 0 else None
    L = _lib.lib()
    need = _lib.check(L.trajsde_enc_fwd_workspace_bytes(_lib.MODE_TC_F16, rows, S, int(dual)), "trajsde_enc_fwd_workspace_bytes")
    ws = torch.empty((max(need, 1),), dtype=torch.uint8, device=dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), need
    with torch.cuda.device(dev):
        _lib.check(L.trajsde_enc_fwd(C.byref(a), _stream_ptr(dev)), "trajsde_enc_fwd")
    if rows > 0:
        LAUNCHES['n'] += 2
    return latent, g_out, y1s


enc_fwd = torch.library.custom_op("trajsde::enc_fwd", _enc_fwd_impl, mutates_args=(), device_types="cuda")


@enc_fwd.register_fake
def _(h0, aa_out, obs_mask, slot, params, gru_params, step_tab, dw, alt_mask, seed, row_offset, step_offset, save_y1):
    rows, S = h0.shape[0], step_tab.shape[0]
    return h0.new_empty((S, rows, 64)), h0.new_empty((S, rows)), h0.new_empty((S if save_y1 else 0, rows, 64))


_GRU_NAMES = ('u1', 'ub1', 'u2', 'ub2', 'r1', 'rb1', 'r2', 'rb2', 'n1', 'nb1', 'n2', 'nb2')


def _enc_bwd_impl(grad_latent: Optional[torch.Tensor], grad_g: Optional[torch.Tensor], latent: torch.Tensor, y1s: torch.Tensor,
            h0: torch.Tensor, aa_out: torch.Tensor, obs_mask: torch.Tensor, slot: torch.Tensor, params: List[torch.Tensor],
            gru_params: List[torch.Tensor], step_tab: torch.Tensor, dw: Optional[torch.Tensor],
            alt_mask: Optional[torch.Tensor], seed: int, row_offset: int, step_offset: int) -> List[torch.Tensor]:
    """[grad_h0, grad_aa_out] + gradients of ``params`` + gradients of ``gru_params`` (same order / shapes)."""
    dev = latent.device
    _monitor(dev).poll()
    S, rows = latent.shape[0], latent.shape[1]
    dual = alt_mask is not None
    ps = _check_params(params, dual, dev)
    gs = [t.detach().contiguous() for t in gru_params]
    n_slots = aa_out.shape[0]
    grad_h0 = torch.empty((rows, 64), dtype=torch.float32, device=dev)
    grad_aa = torch.zeros((n_slots, rows, 64), dtype=torch.float32, device=dev)
    gparams = [torch.empty_like(p) for p in ps]       # the fixed-order reduces write every element
    ggru = [torch.empty_like(p) for p in gs]
    a = _lib.EncBwdArgs()
    a.struct_bytes = C.sizeof(_lib.EncBwdArgs)
    per_step = bool(ENC_BWD_PER_STEP) or S > 128 or rows > 40960      # mirrors enc_bwd.cu (launch accounting only)
    a.mode, a.rows, a.dim, a.flags = _lib.MODE_TC_F16, rows, 64, (4 if per_step else 0)
    a.sched.n_steps, a.sched.n_outputs = S, 0
    a.sched.step_tab = step_tab.data_ptr()
    a.drift, a.diffusion = _mlp_struct(ps[0:6]), _mlp_struct(ps[6:12])
    a.grad_drift, a.grad_diffusion = _mlp_struct(gparams[0:6]), _mlp_struct(gparams[6:12])
    mask_u8 = None
    if dual:
        mask_u8 = alt_mask.contiguous().view(torch.uint8)
        a.diffusion_alt, a.grad_diffusion_alt = _mlp_struct(ps[12:18]), _mlp_struct(gparams[12:18])
        a.alt_mask = mask_u8.data_ptr()
    for name, t, gt in zip(_GRU_NAMES, gs, ggru):
        setattr(a.gru, name, t.data_ptr())
        setattr(a.grad_gru, name, gt.data_ptr())
    if dw is not None:
        dw = dw.contiguous()
        a.noise.dw = dw.data_ptr()
    a.noise.seed, a.noise.row_offset, a.noise.step_offset = seed & (2**63 - 1), row_offset, step_offset
    _apply_seed_dev(a.noise, dev)
    h0c = h0.detach().contiguous()
    aa = aa_out.detach().contiguous()
    om = obs_mask.contiguous().view(torch.uint8)
    lat, y1c = latent.contiguous(), y1s.contiguous()
    a.h0, a.aa_out, a.n_slots, a.slot = h0c.data_ptr(), aa.data_ptr(), n_slots, slot.data_ptr()
    a.obs_mask, a.obs_mask_row_stride = om.data_ptr(), om.stride(0)
    a.latent, a.y1 = lat.data_ptr(), y1c.data_ptr()
    if grad_latent is not None:
        grad_latent = grad_latent.contiguous()
        a.grad_latent = grad_latent.data_ptr()
    if grad_g is not None:
        grad_g = grad_g.contiguous()
        a.grad_g = grad_g.data_ptr()
    a.grad_h0, a.grad_aa_out = grad_h0.data_ptr(), grad_aa.data_ptr()
    a.status = _status_word(dev).data_ptr()
    L = _lib.lib()
    need = _lib.check(L.trajsde_enc_bwd_workspace_bytes(_lib.MODE_TC_F16, rows, S, int(dual)), "trajsde_enc_bwd_workspace_bytes")
    ws = torch.empty((max(need, 1),), dtype=torch.uint8, device=dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), need
    with torch.cuda.device(dev):
        _lib.check(L.trajsde_enc_bwd(C.byref(a), _stream_ptr(dev)), "trajsde_enc_bwd")
        if rows > 0:
            _monitor(dev).snapshot()
    if rows > 0:
        # tables + absmax x2 + GRU pack + pack per net + the sweep (one persistent launch, or per iteration a GRU backward + a fused SDE
        # backward) + two reduces
        LAUNCHES['n'] += 3 + (2 if dual else 1) + (S * 2 if per_step else 1) + 2 + int(grad_latent is not None and rows * S >= (1 << 16))
    return [grad_h0, grad_aa] + gparams + ggru


enc_bwd = torch.library.custom_op("trajsde::enc_bwd", _enc_bwd_impl, mutates_args=(), device_types="cuda")


@enc_bwd.register_fake
def _(grad_latent, grad_g, latent, y1s, h0, aa_out, obs_mask, slot, params, gru_params, step_tab, dw, alt_mask, seed,
      row_offset, step_offset):
    return ([latent.new_empty((latent.shape[1], 64)), torch.empty_like(aa_out)] + [torch.empty_like(p) for p in params] +
            [torch.empty_like(p) for p in gru_params])


def _enc_setup_context(ctx, inputs, output):
    (h0, aa_out, obs_mask, slot, params, gru_params, step_tab, dw, alt_mask, seed, row_offset, step_offset, save_y1) = inputs
    latent, _, y1s = output
    ctx.has_y1 = bool(save_y1)
    ctx.save_for_backward(latent, y1s, h0, aa_out, obs_mask, slot, step_tab, *params, *gru_params)
    ctx.dw, ctx.alt_mask = dw, alt_mask
    ctx.meta = (seed, row_offset, step_offset, len(params))


def _enc_backward(ctx, grad_latent, grad_g, grad_y1):
    if not ctx.has_y1:
        raise RuntimeError("trajsde::enc_fwd was run with save_y1=False; backward needs the saved pre-GRU states")
    latent, y1s, h0, aa_out, obs_mask, slot, step_tab, *rest = ctx.saved_tensors
    seed, row_offset, step_offset, n_p = ctx.meta
    params, gru_params = list(rest[:n_p]), list(rest[n_p:])
    grads = enc_bwd(grad_latent, grad_g, latent, y1s, h0, aa_out, obs_mask, slot, params, gru_params, step_tab, ctx.dw,
                    ctx.alt_mask, seed, row_offset, step_offset)
    return (grads[0], grads[1], None, None, list(grads[2:2 + n_p]), list(grads[2 + n_p:]), None, None, None, None, None, None, None)


enc_fwd.register_autograd(_enc_backward, setup_context=_enc_setup_context)


# ---------------------------------------------------------------------------------------------------------------------
# stand-alone GRU_Unit jump (for the drop-in path that keeps the reference's encoder loop)
# ---------------------------------------------------------------------------------------------------------------------
_GRU_SHAPES = [(64, 128), (64,), (64, 64), (64,)] * 3


def _gru_args(h_cur, x, mask, gru_params):
    dev = h_cur.device
    rows = h_cur.shape[0]
    if h_cur.dim() != 2 or h_cur.shape[1] != 64 or tuple(x.shape) != (rows, 64) or h_cur.dtype != torch.float32 or x.dtype != torch.float32:
        raise ValueError("`h_cur` and `input_tensor` must be float32 of shape (rows, 64)")
    if tuple(mask.shape) != (rows,) or mask.dtype not in (torch.bool, torch.uint8):
        raise ValueError("`mask` must be a bool tensor of shape (rows,)")
    if len(gru_params) != 12 or any(tuple(t.shape) != sh for t, sh in zip(gru_params, _GRU_SHAPES)):
        raise ValueError("gru_params must be the 12 GRU_Unit tensors (update/reset/new_state: Linear(128,64), Linear(64,64))")
    gs = [t.detach().contiguous() for t in gru_params]
    keep = [h_cur.detach().contiguous(), x.detach().contiguous(), mask.contiguous().view(torch.uint8)] + gs
    a = _lib.GruArgs()
    a.struct_bytes = C.sizeof(_lib.GruArgs)
    a.mode, a.rows, a.dim, a.flags = _lib.MODE_TC_F16, rows, 64, 0
    for name, t in zip(_GRU_NAMES, gs):
        setattr(a.gru, name, t.data_ptr())
    a.h_cur, a.x, a.mask = keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr()
    L = _lib.lib()
    need = _lib.check(L.trajsde_gru_workspace_bytes(_lib.MODE_TC_F16, rows), "trajsde_gru_workspace_bytes")
    ws = torch.empty((max(need, 1),), dtype=torch.uint8, device=dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), need
    keep.append(ws)
    return a, keep, L


def _gru_fwd_impl(h_cur: torch.Tensor, x: torch.Tensor, mask: torch.Tensor, gru_params: List[torch.Tensor]) -> torch.Tensor:
    """GRU_Unit.forward (models/utils/ode_utils.py:136-152) as one tensor-core launch: h_next[rows,64]."""
    a, keep, L = _gru_args(h_cur, x, mask, gru_params)
    out = torch.empty_like(keep[0])
    a.h_next = out.data_ptr()
    with torch.cuda.device(h_cur.device):
        _lib.check(L.trajsde_gru_fwd(C.byref(a), _stream_ptr(h_cur.device)), "trajsde_gru_fwd")
    if h_cur.shape[0] > 0:
        LAUNCHES['n'] += 2
    return out


gru_fwd = torch.library.custom_op("trajsde::gru_fwd", _gru_fwd_impl, mutates_args=(), device_types="cuda")


@gru_fwd.register_fake
def _(h_cur, x, mask, gru_params):
    return torch.empty_like(h_cur)


def _gru_bwd_impl(grad_out: torch.Tensor, h_cur: torch.Tensor, x: torch.Tensor, mask: torch.Tensor,
            gru_params: List[torch.Tensor]) -> List[torch.Tensor]:
    """[grad_h_cur, grad_x] + gradients of the 12 GRU tensors."""
    a, keep, L = _gru_args(h_cur, x, mask, gru_params)
    go = grad_out.contiguous()
    gh, gx = torch.empty_like(keep[0]), torch.empty_like(keep[0])
    gg = [torch.empty_like(t) for t in keep[3:15]]
    a.grad_h_next, a.grad_h_cur, a.grad_x = go.data_ptr(), gh.data_ptr(), gx.data_ptr()
    for name, t in zip(_GRU_NAMES, gg):
        setattr(a.grad_gru, name, t.data_ptr())
    with torch.cuda.device(h_cur.device):
        _lib.check(L.trajsde_gru_bwd(C.byref(a), _stream_ptr(h_cur.device)), "trajsde_gru_bwd")
    if h_cur.shape[0] > 0:
        LAUNCHES['n'] += 4
    return [gh, gx] + gg


gru_bwd = torch.library.custom_op("trajsde::gru_bwd", _gru_bwd_impl, mutates_args=(), device_types="cuda")


@gru_bwd.register_fake
def _(grad_out, h_cur, x, mask, gru_params):
    return [torch.empty_like(h_cur), torch.empty_like(x)] + [torch.empty_like(p) for p in gru_params]


def _gru_setup_context(ctx, inputs, output):
    h_cur, x, mask, gru_params = inputs
    ctx.save_for_backward(h_cur, x, mask, *gru_params)


def _gru_backward(ctx, grad_out):
    h_cur, x, mask, *gru_params = ctx.saved_tensors
    grads = gru_bwd(grad_out, h_cur, x, mask, list(gru_params))
    return grads[0], grads[1], None, list(grads[2:])


gru_fwd.register_autograd(_gru_backward, setup_context=_gru_setup_context)


# ---------------------------------------------------------------------------------------------------------------------
# eager fast path
# ---------------------------------------------------------------------------------------------------------------------
# The torch.library operators above are the registered ops (dispatcher, fake tensors, torch.compile).  Eager callers — solver.py,
# encoder.py, i.e. the drop-in functions — go through thin autograd.Function wrappers over the SAME implementation functions
# (`_*_impl`, the plain Python callables the operators were registered from): that skips
# ~0.2 ms of dispatcher / pytree work per call, which is what the reference's 21-iteration encoder loop is made of.
# TRAJSDE_USE_DISPATCHER=1 routes everything through the registered ops instead.
USE_DISPATCHER = _os.environ.get('TRAJSDE_USE_DISPATCHER', '0') == '1'


def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


class _EulerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y0, step_tab, out_begin, out_w, n_outputs, dw, alt_mask, seed, row_offset, step_offset, mode, rows_major, *params):
        ys, g_last, states = _euler_fwd_impl(y0, list(params), step_tab, out_begin, out_w, n_outputs, dw, alt_mask, seed, row_offset,
                                             step_offset, mode, True, rows_major)
        ctx.save_for_backward(states, step_tab, out_begin, out_w, *params)
        ctx.dw, ctx.alt_mask = dw, alt_mask
        ctx.meta = (n_outputs, seed, row_offset, step_offset, mode)
        ctx.set_materialize_grads(False)               # an unused output arrives as None, not as a zero tensor
        return ys, g_last

    @staticmethod
    def backward(ctx, grad_ys, grad_g):
        states, step_tab, out_begin, out_w, *params = ctx.saved_tensors
        n_outputs, seed, row_offset, step_offset, mode = ctx.meta
        grads = _euler_bwd_impl(grad_ys, grad_g, states, list(params), step_tab, out_begin, out_w, n_outputs, ctx.dw, ctx.alt_mask,
                                seed, row_offset, step_offset, mode)
        return (grads[0],) + (None,) * 11 + tuple(grads[1:])


def euler_call(y0, params, step_tab, out_begin, out_w, n_outputs, dw, alt_mask, seed, row_offset, step_offset, mode, save_states,
               rows_major):
    """ys, g_last of the fused solve, differentiable w.r.t. y0 and params when ``save_states``."""
    if USE_DISPATCHER:
        ys, g_last, _ = euler_fwd(y0, params, step_tab, out_begin, out_w, n_outputs, dw, alt_mask, seed, row_offset, step_offset, mode,
                                  save_states, rows_major)
        return ys, g_last
    if save_states and _needs_grad(y0, *params):
        return _EulerFn.apply(y0, step_tab, out_begin, out_w, n_outputs, dw, alt_mask, seed, row_offset, step_offset, mode, rows_major,
                              *params)
    ys, g_last, _ = _euler_fwd_impl(y0, params, step_tab, out_begin, out_w, n_outputs, dw, alt_mask, seed, row_offset, step_offset, mode,
                                    False, rows_major)
    return ys, g_last


class _GruFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h_cur, x, mask, *gru_params):
        ctx.save_for_backward(h_cur, x, mask, *gru_params)
        return _gru_fwd_impl(h_cur, x, mask, list(gru_params))

    @staticmethod
    def backward(ctx, grad_out):
        h_cur, x, mask, *gru_params = ctx.saved_tensors
        grads = _gru_bwd_impl(grad_out, h_cur, x, mask, list(gru_params))
        return (grads[0], grads[1], None) + tuple(grads[2:])


def gru_call(h_cur, x, mask, gru_params):
    if USE_DISPATCHER:
        return gru_fwd(h_cur, x, mask, gru_params)
    if _needs_grad(h_cur, x, *gru_params):
        return _GruFn.apply(h_cur, x, mask, *gru_params)
    return _gru_fwd_impl(h_cur, x, mask, gru_params)


class _EncFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h0, aa_out, obs_mask, slot, step_tab, dw, alt_mask, seed, row_offset, step_offset, n_p, *all_params):
        params, gru_params = list(all_params[:n_p]), list(all_params[n_p:])
        latent, g, y1s = _enc_fwd_impl(h0, aa_out, obs_mask, slot, params, gru_params, step_tab, dw, alt_mask, seed, row_offset,
                                       step_offset, True)
        ctx.save_for_backward(latent, y1s, h0, aa_out, obs_mask, slot, step_tab, *all_params)
        ctx.dw, ctx.alt_mask = dw, alt_mask
        ctx.meta = (seed, row_offset, step_offset, n_p)
        ctx.set_materialize_grads(False)
        return latent, g

    @staticmethod
    def backward(ctx, grad_latent, grad_g):
        latent, y1s, h0, aa_out, obs_mask, slot, step_tab, *all_params = ctx.saved_tensors
        seed, row_offset, step_offset, n_p = ctx.meta
        grads = _enc_bwd_impl(grad_latent, grad_g, latent, y1s, h0, aa_out, obs_mask, slot, list(all_params[:n_p]), list(all_params[n_p:]),
                              step_tab, ctx.dw, ctx.alt_mask, seed, row_offset, step_offset)
        return (grads[0], grads[1]) + (None,) * 9 + tuple(grads[2:])


def enc_call(h0, aa_out, obs_mask, slot, params, gru_params, step_tab, dw, alt_mask, seed, row_offset, step_offset, need_grad):
    """latent, g of the fused encoder recurrence, differentiable when ``need_grad``."""
    if USE_DISPATCHER:
        latent, g, _ = enc_fwd(h0, aa_out, obs_mask, slot, params, gru_params, step_tab, dw, alt_mask, seed, row_offset, step_offset,
                               need_grad)
        return latent, g
    if need_grad:
        return _EncFn.apply(h0, aa_out, obs_mask, slot, step_tab, dw, alt_mask, seed, row_offset, step_offset, len(params),
                            *params, *gru_params)
    latent, g, _ = _enc_fwd_impl(h0, aa_out, obs_mask, slot, params, gru_params, step_tab, dw, alt_mask, seed, row_offset, step_offset,
                                 False)
    return latent, g


# ---------------------------------------------------------------------------------------------------------------------
# Brownian increments exactly as the kernels draw them (for replaying Philox runs through the oracle)
# ---------------------------------------------------------------------------------------------------------------------
def philox_dw(dsched: DeviceSchedule, rows: int, seed: int, device, row_offset: int = 0, step_offset: int = 0) -> torch.Tensor:
    out = torch.empty((dsched.n_steps, rows, 64), dtype=torch.float32, device=device)
    s = _lib.Schedule()
    s.n_steps, s.n_outputs = dsched.n_steps, dsched.n_outputs
    s.step_tab, s.out_begin, s.out_w = dsched.step_tab.data_ptr(), dsched.out_begin.data_ptr(), dsched.out_w.data_ptr()
    n = _lib.Noise()
    n.seed, n.row_offset, n.step_offset = seed & (2**63 - 1), row_offset, step_offset
    _apply_seed_dev(n, torch.device(device))
    with torch.cuda.device(device):
        _lib.check(_lib.lib().trajsde_philox_dw(C.byref(s), C.byref(n), rows, out.data_ptr(), _stream_ptr(device)),
                   "trajsde_philox_dw")
    LAUNCHES['n'] += 1
    return out
