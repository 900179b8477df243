"""Host-side mirror of the encoder's SDE recurrence (enc_hivt_nusargo_sde_sep2.py:128-196): 21 x [one-step sdeint_dual +
GRU jump], diffusion read-out and the eos gather.  The AA/AL graph attention around it stays on the reference path.

Two execution paths, both on our CUDA kernels:
  * fused  (default in TC mode): ONE persistent kernel for the whole recurrence (csrc/enc_tc.cu); differentiable — the
    backward is one library call (csrc/enc_bwd.cu: reverse sweep of GRU backward + one-step SDE backward kernels);
  * stepwise: 21 x fused one-step `sdeint_dual` launch + the GRU jump in plain torch — differentiable through the
    per-step ops, and the only path in 'exact' mode.
"""
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib, ops
from .schedule import encoder_schedule, encoder_time_pairs
from .solver import _mlp_params, _next_call_seed, get_default_mode, sdeint_dual

_GRU_KEYS = (('update_gate', 0), ('update_gate', 2), ('reset_gate', 0), ('reset_gate', 2), ('new_state_net', 0), ('new_state_net', 2))


def _gru_params(gru_unit):
    out = []
    for name, idx in _GRU_KEYS:
        lin = getattr(gru_unit, name)[idx]
        out += [lin.weight, lin.bias]
    return out


_slot_cache = {}


def _enc_tables(max_past_t, hist, dt, device):
    key = (float(max_past_t), int(hist), float(dt), str(device))
    hit = _slot_cache.get(key)
    if hit is None:
        sched = encoder_schedule(max_past_t, hist, dt)
        slots = np.array([t for _, _, t in encoder_time_pairs(max_past_t, hist)], dtype=np.int32)
        hit = (torch.from_numpy(sched.step_tab()).to(device), torch.from_numpy(slots).to(device))
        _slot_cache[key] = hit
    return hit


def gru_jump(gru_unit, h_cur: torch.Tensor, input_tensor: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """``GRU_Unit.forward(h_cur, input_tensor, mask)`` (models/utils/ode_utils.py:136-152) as one tensor-core launch with a fused
    backward; what ``install()`` binds ``encoder.gru_unit.forward`` to.  CUDA / 64-wide layers only — no fallback."""
    if not h_cur.is_cuda:
        raise RuntimeError("trajsde_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    return ops.gru_call(h_cur, input_tensor, mask, _gru_params(gru_unit))


def encoder_recurrence(sde, gru_unit, h0: torch.Tensor, aa_out: torch.Tensor, actors_mask: torch.Tensor,
                       nus_mask: torch.Tensor, *, dt: float = 0.1, max_past_t: float = 2.0, dW: Optional[torch.Tensor] = None,
                       seed: Optional[int] = None, mode: Optional[str] = None, fused: Optional[bool] = None,
                       row_offset: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns ``latent_ys[21, rows, 64]`` (post-GRU state of every loop iteration) and ``g[21, rows]`` (pre-step
    diffusion of every iteration).  ``dW[21, rows, 64]`` optionally supplies the Brownian increments per iteration."""
    hist = aa_out.shape[0]
    mode = mode or get_default_mode()
    params = (_mlp_params(sde.f_func, 64, 'f_func') + _mlp_params(sde.g_nus, 1, 'g_nus') + _mlp_params(sde.g_argo, 1, 'g_argo'))
    gparams = _gru_params(gru_unit)
    need_grad = torch.is_grad_enabled() and (h0.requires_grad or aa_out.requires_grad or
                                             any(p.requires_grad for p in params + gparams))
    if fused is None:
        fused = mode == 'tc_f16' and hist <= 32
    if fused:
        if mode != 'tc_f16':
            raise NotImplementedError("the fused encoder recurrence exists in 'tc_f16' mode only")
        if not h0.is_cuda:
            raise RuntimeError("trajsde_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        step_tab, slots = _enc_tables(max_past_t, hist, dt, h0.device)
        if seed is None:
            seed = _next_call_seed() if dW is None else 0
        return ops.enc_call(h0, aa_out, actors_mask, slots, list(params), list(gparams), step_tab, dW, nus_mask, int(seed),
                            int(row_offset), 0, bool(need_grad))
    h = h0
    latent, gs = [], []
    for idx, (prev_t, t_i, t) in enumerate(encoder_time_pairs(max_past_t, hist)):
        ts = torch.tensor([prev_t, t_i])                                   # enc…sep2.py:142
        bm = None if dW is None else dW[idx:idx + 1]
        ys, g = sdeint_dual(sde, h, ts, nus_mask, bm=bm, dt=dt, method='euler', mode=mode,
                            seed=None if seed is None else seed + idx, row_offset=row_offset)
        h = gru_unit(input_tensor=aa_out[t], h_cur=ys[-1], mask=actors_mask[:, t])   # :165-169
        latent.append(h)
        gs.append(g[:, 0])
    return torch.stack(latent), torch.stack(gs)


def eos_gather(latent_ys: torch.Tensor, bos_mask: torch.Tensor, ref_time: int = 20) -> torch.Tensor:
    """out[n] = latent_ys[ref_time - argmax(bos_mask[n]), n]   (enc…sep2.py:187-188)."""
    eos = ref_time - torch.argmax(bos_mask.float(), dim=1)
    return latent_ys[eos, torch.arange(latent_ys.size(1), device=latent_ys.device), :]


def encoder_recurrence_ood(sde, gru_unit, aa_out: torch.Tensor, actors_mask: torch.Tensor, nus_mask: torch.Tensor,
                           bos_mask: torch.Tensor, *, eval_iter: int = 10, dt: float = 0.1, max_past_t: float = 2.0,
                           seed: Optional[int] = None, mode: Optional[str] = None, ref_time: int = 20):
    """Monte-Carlo encoder of ``forward_ood`` (enc_hivt_nusargo_sde_sep2.py:252-313): ``eval_iter`` independent passes of the
    recurrence from a ZERO hidden state (:257), eos gather per pass (:309-310), then mean latent ``[N,64]`` and per-actor
    std ``outs.std(0).mean(-1)`` ``[N]`` (:311-313).  In TC mode all passes run as one fused-kernel launch."""
    rows = aa_out.shape[1]
    base = _next_call_seed() if seed is None else int(seed)
    mode = mode or get_default_mode()
    with torch.no_grad():
        if mode == 'tc_f16':
            # the eval_iter passes are independent rows: ONE fused launch over eval_iter x rows; pass j draws the Philox stream of
            # global rows [j*rows, (j+1)*rows) (the generator is keyed by global row), so passes get independent increments
            k = int(eval_iter)
            h0 = torch.zeros((k * rows, 64), dtype=torch.float32, device=aa_out.device)
            lat, _ = encoder_recurrence(sde, gru_unit, h0, aa_out.repeat(1, k, 1), actors_mask.repeat(k, 1), nus_mask.repeat(k),
                                        dt=dt, max_past_t=max_past_t, seed=base, mode=mode)
            outs = eos_gather(lat, bos_mask.repeat(k, 1), ref_time).view(k, rows, 64)
        else:
            h0 = torch.zeros((rows, 64), dtype=torch.float32, device=aa_out.device)
            outs = []
            for j in range(eval_iter):
                lat, _ = encoder_recurrence(sde, gru_unit, h0, aa_out, actors_mask, nus_mask, dt=dt, max_past_t=max_past_t,
                                            seed=(base + 0x51ED27 * (j + 1)) & (2**63 - 1), mode=mode)
                outs.append(eos_gather(lat, bos_mask, ref_time))
            outs = torch.stack(outs)
    return outs.mean(0), outs.std(0).mean(-1)
