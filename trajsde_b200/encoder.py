"""Host-side mirror of the encoder's SDE recurrence (enc_hivt_nusargo_sde_sep2.py:128-196): 21 x [one-step sdeint_dual +
GRU jump], diffusion read-out and the eos gather.  The AA/AL graph attention around it stays on the reference path."""
from typing import Optional, Tuple

import torch

from .schedule import encoder_time_pairs
from .solver import sdeint_dual


def encoder_recurrence(sde, gru_unit, h0: torch.Tensor, aa_out: torch.Tensor, actors_mask: torch.Tensor,
                       nus_mask: torch.Tensor, *, dt: float = 0.1, max_past_t: float = 2.0, dW: Optional[torch.Tensor] = None,
                       seed: Optional[int] = None, mode: Optional[str] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns ``latent_ys[21, rows, 64]`` (post-GRU state of every loop iteration) and ``g[21, rows]`` (pre-step
    diffusion of every iteration).  ``dW[21, rows, 64]`` optionally supplies the Brownian increments per iteration."""
    hist = aa_out.shape[0]
    h = h0
    latent, gs = [], []
    for idx, (prev_t, t_i, t) in enumerate(encoder_time_pairs(max_past_t, hist)):
        ts = torch.tensor([prev_t, t_i])                                   # enc…sep2.py:142
        bm = None if dW is None else dW[idx:idx + 1]
        ys, g = sdeint_dual(sde, h, ts, nus_mask, bm=bm, dt=dt, method='euler', mode=mode,
                            seed=None if seed is None else seed + idx)
        h = gru_unit(input_tensor=aa_out[t], h_cur=ys[-1], mask=actors_mask[:, t])   # :165-169
        latent.append(h)
        gs.append(g[:, 0])
    return torch.stack(latent), torch.stack(gs)


def eos_gather(latent_ys: torch.Tensor, bos_mask: torch.Tensor, ref_time: int = 20) -> torch.Tensor:
    """out[n] = latent_ys[ref_time - argmax(bos_mask[n]), n]   (enc…sep2.py:187-188)."""
    eos = ref_time - torch.argmax(bos_mask.float(), dim=1)
    return latent_ys[eos, torch.arange(latent_ys.size(1), device=latent_ys.device), :]
