"""CPU restatement (plain torch) of TrajSDE's Euler–Maruyama hot path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Parity: pinned against the reference's own files run through ``oracle/shims`` (tests/golden/*.npz, tests/test_oracle_golden.py).

Parameter containers are plain dicts of tensors keyed like the reference ``state_dict``:
  drift      ``{'0.weight':[64,66], '0.bias':[64], '2.weight':[64,64], '2.bias', '4.weight':[64,64], '4.bias'}``  (FFunc.net)
  diffusion  same keys, ``'4.weight':[1,64]``, ``'4.bias':[1]``                                                   (GFunc.net)
  gru        ``{'update_gate.0.weight':[64,128], 'update_gate.0.bias', 'update_gate.2.weight':[64,64], ... 'reset_gate.*',
               'new_state_net.*'}``                                                                              (GRU_Unit)
All functions are dtype-generic (fp32 = bit-faithful restatement, fp64 = gradient/accuracy reference) and differentiable.
"""
from typing import Dict, List, Optional, Tuple

import torch

P = Dict[str, torch.Tensor]


# ------------------------------------------------------------------------------------------------------------------
# Step schedule — reference: models/utils/sdeint.py:340-384 (BaseSDESolver_private.integrate; torchsde 0.2.5
# BaseSDESolver.integrate is identical), interpolation weights from torchsde._core.interp.linear_interp.
# ------------------------------------------------------------------------------------------------------------------
def euler_schedule_ref(ts: torch.Tensor, dt: float) -> Dict[str, torch.Tensor]:
    """Literal replay of the reference time loop with float32 0-d CPU tensors.

    Returns float32 tensors ``t0[S], h[S]`` (step start / size) and per output j=1..T-1 ``out_k[j-1]`` (index of the
    last step taken before the output), ``w0, w1`` such that ``ys[j] = w0*Y[k] + w1*Y[k+1]`` (Y[0]=y0).
    """
    ts = ts.detach().to('cpu')
    step_size = dt
    prev_t = curr_t = ts[0]
    t0s: List[torch.Tensor] = []
    hs: List[torch.Tensor] = []
    out_k, w0s, w1s = [], [], []
    for out_t in ts[1:]:
        while curr_t < out_t:                                  # sdeint.py:350
            next_t = min(curr_t + step_size, ts[-1])           # sdeint.py:351
            prev_t = curr_t                                    # sdeint.py:378
            t0s.append(curr_t)
            hs.append(next_t - curr_t)                         # Euler_private.step: dt = t1 - t0, sdeint.py:479
            curr_t = next_t                                    # sdeint.py:380
        if not t0s:
            raise ValueError("schedule: first output interval takes no step (ts not increasing?)")
        out_k.append(len(t0s) - 1)
        # linear_interp(t0=prev_t, y0=prev_y, t1=curr_t, y1=curr_y, t=out_t)          sdeint.py:382
        w0s.append((curr_t - out_t) / (curr_t - prev_t))
        w1s.append((out_t - prev_t) / (curr_t - prev_t))
    f32 = lambda xs: torch.stack([x.to(torch.float32) for x in xs])  # noqa: E731
    return {'t0': f32(t0s), 'h': f32(hs), 'out_k': torch.tensor(out_k, dtype=torch.int32),
            'w0': f32(w0s), 'w1': f32(w1s)}


# ------------------------------------------------------------------------------------------------------------------
# Drift / diffusion nets — reference: dec_hivt_nusargo_sde.py:107-127,141-158 ; enc_hivt_nusargo_sde_sep2.py:372-398,
# 412-440 (identical shapes for sde_layers=2).
# ------------------------------------------------------------------------------------------------------------------
def _time_cat(t, y: torch.Tensor) -> torch.Tensor:
    # FFunc.forward: _t = torch.ones(n,1) * float(t); _t = _t.to(y); cat((y, sin(_t), cos(_t)), -1)   dec…sde.py:124-126
    _t = torch.ones(y.size(0), 1) * float(t)
    _t = _t.to(y)
    return torch.cat((y, torch.sin(_t), torch.cos(_t)), dim=-1)


def _lin(p: P, i: int, x: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.linear(x, p[f'{i}.weight'].to(x.dtype), p[f'{i}.bias'].to(x.dtype))


def drift_ref(pf: P, t, y: torch.Tensor) -> torch.Tensor:
    """f(t, y) — FFunc.net: Linear(66,64) Tanh Linear(64,64) Tanh Linear(64,64)."""
    x = _time_cat(t, y)
    return _lin(pf, 4, torch.tanh(_lin(pf, 2, torch.tanh(_lin(pf, 0, x)))))


def diffusion_ref(pg: P, t, y: torch.Tensor) -> torch.Tensor:
    """g(t, y) — GFunc.net: Linear(66,64) Tanh Linear(64,64) Tanh Linear(64,1), then sigmoid; shape [rows,1]."""
    x = _time_cat(t, y)
    return torch.sigmoid(_lin(pg, 4, torch.tanh(_lin(pg, 2, torch.tanh(_lin(pg, 0, x))))))


def diffusion_dual_ref(pg_nus: P, pg_argo: P, t, y: torch.Tensor, nus_mask: torch.Tensor) -> torch.Tensor:
    """Encoder LSDEFunc.g (enc…sep2.py:470-482): g_nus on rows with nus_mask, g_argo on the rest; [rows,1].

    Restated with torch.where instead of masked gather/scatter (row-wise nets => same values) so it stays
    differentiable without index_put; the reference's [rows,64] ``repeat`` is applied by the caller.
    """
    g0 = diffusion_ref(pg_nus, t, y)
    g1 = diffusion_ref(pg_argo, t, y)
    return torch.where(nus_mask.unsqueeze(-1), g0, g1)


# ------------------------------------------------------------------------------------------------------------------
# Euler–Maruyama solve — reference: Euler_private.step sdeint.py:477-485 (library Euler.step :447-465 comment),
# prod_diagonal :544, integrate :340-384, linear_interp.
# ------------------------------------------------------------------------------------------------------------------
def euler_solve_ref(pf: P, pg: P, y0: torch.Tensor, ts: torch.Tensor, dt: float, dW: torch.Tensor,
                    nus_mask: Optional[torch.Tensor] = None, pg_argo: Optional[P] = None,
                    return_states: bool = False, probe: bool = False):
    """``ys[T,rows,64], g_last[rows,1]`` (+ ``states[S+1,rows,64]``) for caller-supplied increments ``dW[S,rows,64]``.

    ``dW[k]`` is consumed by schedule step k (one slab per Euler step, incl. the sliver step).
    With ``nus_mask``/``pg_argo`` this is the encoder's ``sdeint_dual`` (``pg`` is then g_nus); ``g_last`` is the
    diffusion evaluated at the start of the last step (sdeint.py:384 returns it misnamed ``g_prod``).
    Works in ``y0.dtype``; time arithmetic always in float32 like the reference.
    """
    sched = euler_schedule_ref(ts, dt)
    S = sched['t0'].numel()
    if probe:   # check_contract's shape probe: one wasted f and g evaluation per call (sdeint.py:913-921); timing fidelity only
        drift_ref(pf, ts[0], y0)
        diffusion_ref(pg, ts[0], y0) if nus_mask is None else diffusion_dual_ref(pg, pg_argo, ts[0], y0, nus_mask)
    assert dW.shape[0] == S, f"dW must have one slab per schedule step: {dW.shape[0]} vs {S}"
    dtype = y0.dtype
    Y = [y0]
    g = None
    y = y0
    for k in range(S):
        t0 = sched['t0'][k]
        h = sched['h'][k].to(dtype)
        f = drift_ref(pf, t0, y)
        if nus_mask is None:
            g = diffusion_ref(pg, t0, y)
        else:
            g = diffusion_dual_ref(pg, pg_argo, t0, y, nus_mask)
        g_full = g.repeat(1, y.size(1))                        # LSDEFunc.g: .repeat(1, embed_dim)   dec…sde.py:194
        y = y + f * h + g_full * dW[k].to(dtype)               # Euler step + prod_diagonal          sdeint.py:484,544
        Y.append(y)
    ys = [y0]
    for j in range(sched['out_k'].numel()):
        k = int(sched['out_k'][j])
        w0, w1 = sched['w0'][j].to(dtype), sched['w1'][j].to(dtype)
        ys.append(w0 * Y[k] + w1 * Y[k + 1])                   # linear_interp two-term form
    out = (torch.stack(ys, dim=0), g)
    if return_states:
        out = out + (torch.stack(Y, dim=0),)
    return out


# ------------------------------------------------------------------------------------------------------------------
# GRU jump — reference: models/utils/ode_utils.py:111-152 (GRU_Unit)
# ------------------------------------------------------------------------------------------------------------------
def gru_ref(pgru: P, h_cur: torch.Tensor, x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    def two(prefix, inp):
        w0, b0 = pgru[f'{prefix}.0.weight'].to(inp.dtype), pgru[f'{prefix}.0.bias'].to(inp.dtype)
        w2, b2 = pgru[f'{prefix}.2.weight'].to(inp.dtype), pgru[f'{prefix}.2.bias'].to(inp.dtype)
        return torch.nn.functional.linear(torch.tanh(torch.nn.functional.linear(inp, w0, b0)), w2, b2)

    y_concat = torch.cat([h_cur, x], -1)                                   # ode_utils.py:137
    update_gate = torch.sigmoid(two('update_gate', y_concat))              # :139
    reset_gate = torch.sigmoid(two('reset_gate', y_concat))                # :140
    combined = torch.cat([x, reset_gate * h_cur], dim=1)                   # :142
    new_state = two('new_state_net', combined)                             # :143
    h_next = (1 - update_gate) * new_state + update_gate * h_cur           # :145
    m = mask.unsqueeze(-1)
    return m * h_next + ~m * h_cur                                         # :150


# ------------------------------------------------------------------------------------------------------------------
# Encoder recurrence — reference: enc_hivt_nusargo_sde_sep2.py:128-196 (time grid, 21×[sdeint_dual + GRU], eos gather)
# ------------------------------------------------------------------------------------------------------------------
def encoder_time_pairs_ref(max_past_t: float = 2.0, historical_steps: int = 21) -> List[Tuple[torch.Tensor, torch.Tensor, int]]:
    """[(prev_t, t_i, data_slot)] for run_backwards=True, exactly as enc…sep2.py:128-135,175-179 builds them."""
    pts = -1 * torch.linspace(-max_past_t, 0, historical_steps)
    prev_t, t_i = pts[-1] - 0.01, pts[-1]
    pairs = []
    order = list(reversed(range(historical_steps)))
    for idx, t in enumerate(order):
        pairs.append((prev_t, t_i, t))
        if idx + 1 < historical_steps:
            prev_t, t_i = pts[t], pts[t - 1]
    return pairs


def encoder_recurrence_ref(pf: P, pg_nus: P, pg_argo: P, pgru: P, h0: torch.Tensor, aa_out: torch.Tensor,
                           actors_mask: torch.Tensor, nus_mask: torch.Tensor, dW: torch.Tensor, dt: float = 0.1,
                           max_past_t: float = 2.0, probe: bool = False):
    """21×(one-step ``sdeint_dual`` + ``GRU_Unit``).  ``h0[rows,64]``, ``aa_out[21,rows,64]``, ``actors_mask[rows,21]``,
    ``dW[21,rows,64]`` (slab idx = loop iteration).  Returns ``latent_ys[21,rows,64]`` (post-GRU, loop order) and
    ``g[21,rows,1]`` (pre-step diffusion of each iteration)."""
    hist = aa_out.shape[0]
    prev_hidden = h0
    latent, gs = [], []
    for idx, (prev_t, t_i, t) in enumerate(encoder_time_pairs_ref(max_past_t, hist)):
        time_points = torch.tensor([prev_t, t_i])                          # enc…sep2.py:142
        ys, g = euler_solve_ref(pf, pg_nus, prev_hidden, time_points, dt, dW[idx:idx + 1], nus_mask, pg_argo, probe=probe)
        yi_ode = ys[-1]                                                    # :165
        yi = gru_ref(pgru, yi_ode, aa_out[t].to(h0.dtype), actors_mask[:, t])  # :169
        prev_hidden = yi
        latent.append(yi)
        gs.append(g)
    return torch.stack(latent), torch.stack(gs)


def encoder_eos_gather_ref(latent_ys: torch.Tensor, bos_mask: torch.Tensor, ref_time: int = 20) -> torch.Tensor:
    """out[n] = latent_ys[ref_time - argmax(bos_mask[n]), n]   (enc…sep2.py:187-188)."""
    eos = ref_time - torch.argmax(bos_mask.float(), dim=1)
    return latent_ys[eos, torch.arange(latent_ys.size(1)), :]


# ------------------------------------------------------------------------------------------------------------------
# Consumers used for ADE/FDE agreement — decoder heads dec…sde.py:50-61,95-99 ; metrics/ade_t.py:44-66, fde_t.py:45-57
# ------------------------------------------------------------------------------------------------------------------
def decoder_loc_head_ref(phead: P, sol_y: torch.Tensor) -> torch.Tensor:
    """self.decoder: Linear(64,64) LayerNorm ReLU Linear(64,2) applied to sol_y[rows,T,64] -> [rows,T,2]."""
    x = torch.nn.functional.linear(sol_y, phead['0.weight'], phead['0.bias'])
    x = torch.nn.functional.layer_norm(x, (x.size(-1),), phead['1.weight'], phead['1.bias'])
    return torch.nn.functional.linear(torch.relu(x), phead['3.weight'], phead['3.bias'])


def decoder_scale_ref(phead: P, sol_y: torch.Tensor, min_scale: float) -> torch.Tensor:
    """self.scale head + dec…sde.py:98-99: elu(head(sol_y)) + 1 + min_scale (same layer stack as self.decoder)."""
    return torch.nn.functional.elu(decoder_loc_head_ref(phead, sol_y), alpha=1.0) + 1.0 + min_scale


def min_ade_fde_ref(loc: torch.Tensor, target: torch.Tensor, reg_mask: torch.Tensor) -> Tuple[float, float]:
    """loc[modes,N,T,2], target[N,T,2], reg_mask[N,T] -> (minADE, minFDE) with best mode by ADE ('nuScenes' branch,
    ade_t.py:55-57) and FDE at each agent's last valid slot."""
    l2 = torch.norm(loc - target.unsqueeze(0), p=2, dim=-1)
    keep = reg_mask.any(-1)
    l2, m = l2[:, keep], reg_mask[keep]
    l2 = l2 * m.unsqueeze(0)
    ade = l2.sum(-1) / m.sum(-1).unsqueeze(0)
    best = torch.argmin(ade, dim=0)
    n = torch.arange(m.size(0))
    last = m.size(1) - 1 - torch.argmax(m.flip(1).float(), dim=1)
    return float(ade[best, n].mean()), float(l2[best, n, last].mean())


def ade_t_ref(pred: torch.Tensor, target: torch.Tensor, reg_mask: torch.Tensor) -> float:
    """ADE_T.update/compute, 'nuScenes' branch (metrics/ade_t.py:44-66): pred[modes,N,T,2], target[N,T,2], reg_mask[N,T]; best mode
    by masked ADE, averaged over the actors that have any valid future slot."""
    l2 = torch.norm(pred - target.unsqueeze(0), p=2, dim=-1)                # :48
    keep = reg_mask.any(-1)                                                 # :49
    l2, m = l2[:, keep].clone(), reg_mask[keep]                             # :50
    l2[:, ~m] = 0                                                           # :51
    ade = l2.sum(-1) / m.sum(-1).unsqueeze(0)                               # :54
    best = torch.argmin(ade, dim=0)                                         # :57
    return float(ade[best, torch.arange(int(keep.sum()))].sum() / keep.sum())   # :68-73


def fde_t_ref(pred: torch.Tensor, target: torch.Tensor, reg_mask: torch.Tensor, source: torch.Tensor, end_idcs=(59, 29)) -> float:
    """FDE_T.update/compute (metrics/fde_t.py:45-57): displacement at the per-source end slot (actors sorted by source), best mode by
    that displacement, over the actors whose end slot is valid."""
    NA = pred.size(1)
    c0, c1 = int((source == 0).sum()), int((source == 1).sum())
    end = torch.repeat_interleave(torch.tensor(list(end_idcs)), torch.tensor([c0, c1]))   # :48-49
    n = torch.arange(NA)
    l2 = torch.norm(pred[:, n, end, :] - target[n, end].unsqueeze(0), p=2, dim=-1)          # :51
    ok = reg_mask[n, end]                                                                   # :52
    l2 = l2[:, ok]
    best = torch.argmin(l2, dim=0)
    return float(l2[best, torch.arange(int(ok.sum()))].sum() / ok.sum())


# ------------------------------------------------------------------------------------------------------------------
# the decoder stage around the solve and the two training losses
#   SDEDecoder.forward   models/decoders/dec_hivt_nusargo_sde.py:77-105  (aggr_embed :26-29,82 ; heads :50-61,95-99 ; pi :63-67,92-94)
#   L2                   losses/L2.py:10-27
#   DiffBCE              losses/diff_BCE.py:11-16
# ------------------------------------------------------------------------------------------------------------------
def _lin_ln_relu(p: P, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """nn.Sequential(Linear, LayerNorm, ReLU) with state_dict keys ``{prefix}.0.*`` / ``{prefix}.1.*``."""
    z = torch.nn.functional.linear(x, p[f'{prefix}.0.weight'].to(x.dtype), p[f'{prefix}.0.bias'].to(x.dtype))
    z = torch.nn.functional.layer_norm(z, (z.size(-1),), p[f'{prefix}.1.weight'].to(x.dtype), p[f'{prefix}.1.bias'].to(x.dtype))
    return torch.relu(z)


def aggr_embed_ref(p: P, local_embed: torch.Tensor, global_embed: torch.Tensor) -> torch.Tensor:
    """hidden_0[modes*N, 64] = aggr_embed(cat(global_embed, local_embed.expand(modes, N, 64)))   dec…sde.py:82-85.
    ``p``: the decoder's state_dict (keys ``aggr_embed.0.weight`` [64,128], ``aggr_embed.0.bias``, ``aggr_embed.1.weight/bias``)."""
    modes = global_embed.size(0)
    x = torch.cat((global_embed, local_embed.expand(modes, *local_embed.shape)), dim=-1)
    return _lin_ln_relu(p, 'aggr_embed', x).reshape(modes * local_embed.size(0), -1)


def decoder_forward_ref(p: P, local_embed: torch.Tensor, global_embed: torch.Tensor, padding_mask: torch.Tensor, dW: torch.Tensor,
                        ts: Optional[torch.Tensor] = None, dt: float = 0.1, min_scale: float = 0.001, future_steps: int = 60):
    """The whole ``SDEDecoder.forward`` (dec…sde.py:77-105) with caller-supplied increments: returns the reference's ``out`` dict
    (``loc`` [modes,N,F,4] = cat(loc, scale), ``pi`` [N,modes], ``reg_mask`` [N,F]) plus ``hidden_0`` and ``ys`` for step-wise checks.
    ``p`` = the decoder's state_dict; works in the dtype of the embeddings."""
    dtype = local_embed.dtype
    modes, N = global_embed.size(0), local_embed.size(0)
    ts = torch.linspace(0, 6, future_steps + 1) if ts is None else ts                      # :72
    hidden_0 = aggr_embed_ref(p, local_embed, global_embed)
    sub = lambda pre: {k[len(pre) + 1:]: v for k, v in p.items() if k.startswith(pre + '.')}  # noqa: E731
    ys, _ = euler_solve_ref(sub('lsde_func.f_func.net'), sub('lsde_func.g_func.net'), hidden_0, ts, dt, dW.to(dtype))
    sol_y = ys[1:].permute(1, 0, 2)                                                        # :88
    xpi = torch.cat((local_embed.expand(modes, *local_embed.shape), global_embed), dim=-1)   # :92-93
    pi = torch.nn.functional.linear(_lin_ln_relu(p, 'pi', xpi), p['pi.3.weight'].to(dtype), p['pi.3.bias'].to(dtype)).squeeze(-1).t()
    cast = lambda d: {k: v.to(dtype) for k, v in d.items()}                                # noqa: E731
    loc = decoder_loc_head_ref(cast(sub('decoder')), sol_y).view(modes, N, future_steps, 2)  # :95
    scale = torch.nn.functional.elu(decoder_loc_head_ref(cast(sub('scale')), sol_y), alpha=1.0).view(modes, -1, future_steps, 2) + 1.0
    scale = scale + min_scale                                                              # :98-99
    return {'loc': torch.cat((loc, scale), dim=-1), 'pi': pi, 'reg_mask': ~padding_mask[:, -future_steps:],   # :100,104
            'hidden_0': hidden_0, 'ys': ys}


def l2_loss_ref(loc4: torch.Tensor, target: torch.Tensor, reg_mask: torch.Tensor) -> torch.Tensor:
    """losses/L2.py:10-27: best mode per actor by mean masked displacement, then the mean displacement of that mode over the valid
    (actor, slot) pairs.  ``loc4`` [modes,N,F,4] (cat of loc and scale, :12), target [N,F,2], reg_mask [N,F]."""
    loc, _ = loc4.chunk(2, dim=-1)                                          # :12
    l2 = torch.norm(target.unsqueeze(0) - loc, p=2, dim=-1)                 # :15
    ade = l2.clone()
    ade[:, ~reg_mask] = 0                                                   # :17-18
    best = torch.argmin(ade.mean(-1), dim=0)                                # :19
    minl2 = l2[best, torch.arange(l2.size(1))]                              # :20
    if reg_mask.sum() > 0:
        return minl2[reg_mask].mean()                                       # :22-24
    return torch.zeros((), dtype=loc4.dtype)                                # :27 (returns the int 0)


def diff_bce_ref(diff_in: torch.Tensor, diff_out: torch.Tensor) -> torch.Tensor:
    """losses/diff_BCE.py:11-16 with the labels the encoder attaches (enc…sep2.py:194-195: in -> 0, out -> 1), mean reduction;
    nn.BCELoss clamps each log term at -100."""
    loss_in = -torch.clamp(torch.log1p(-diff_in), min=-100.0).mean()
    loss_out = -torch.clamp(torch.log(diff_out), min=-100.0).mean()
    return loss_in + loss_out
