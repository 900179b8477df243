"""Generate tests/golden/*.npz by running the reference's own files (oracle/ref_runner.py) in the dev container.

    python tests/golden/make_golden.py

Fixtures (all float32 unless noted, seeds fixed):
  decoder_solve.npz   reference SDEDecoder.lsde_func + torchsde.sdeint call (dec…sde.py:88): 48 rows, 61 steps, dW supplied;
                      weights (non-zero biases), y0, dW, ys, Brownian query times (ta,tb), loc/scale head parameters and outputs.
  encoder_loop.npz    reference sdeint_dual + GRU_Unit loop (enc…sep2.py:128-182): 40 rows, 21 steps, dual g, masks.
  decoder_stage.npz   the reference's whole `SDEDecoder.forward` (dec…sde.py:77-105: aggr_embed -> sdeint -> heads -> elu -> cat, pi) on
                      seeded embeddings with supplied dW, its `L2` / `DiffBCE` losses (losses/L2.py, losses/diff_BCE.py) with their
                      gradients through the reference solver, and `ADE_T` / `FDE_T` (metrics/ade_t.py, metrics/fde_t.py) of the result.
  encoder_stage.npz   the reference's whole `LocalEncoderSDESepPara2.forward` (enc…sep2.py:66-202: AA encoder -> 21 x [sdeint_dual + GRU jump] ->
                      gathers -> AL encoder) on a small synthetic graph batch, through the functional PyG stand-in of oracle/shims, eval
                      mode, supplied dW: what the SDE recurrence receives (aa_out, masks), what it hands on (latents at eos, agents'
                      diffusion), the stage outputs, and the gradients of a loss on them.
  schedule.npz        step schedules for the grids of SURVEY App. A (F = 10,20,30,50,60,100,200 and the encoder pairs),
                      from the literal torch replay + the (ta,tb) the reference solver actually queried.
Cannot run on the GPU box (/root/reference absent there) — the outputs are committed.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_runner as rr  # noqa: E402
from oracle import sde_oracle as so  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def flat(prefix, params):
    return {f'{prefix}/{k}': v.numpy() for k, v in params.items()}


def decoder_fixture():
    dec = rr.build_reference_decoder(seed=0, bias_std=0.1)
    g = torch.Generator().manual_seed(1234)
    rows = 48
    y0 = torch.relu(torch.randn(rows, 64, generator=g))
    sched = so.euler_schedule_ref(dec.ts_pred, dec.min_stepsize)
    S = sched['h'].numel()
    dW = torch.randn(S, rows, 64, generator=g) * torch.sqrt(sched['h']).view(S, 1, 1)
    ys, queries = rr.run_reference_decoder_solve(dec, y0, dW)
    loc, scale = rr.run_reference_decoder_heads(dec, ys)
    q = np.array([[float(a), float(b)] for a, b in queries], dtype=np.float32)
    d = dict(y0=y0.numpy(), dW=dW.numpy(), ys=ys.numpy(), ts=dec.ts_pred.numpy(), dt=np.float64(dec.min_stepsize),
             queries=q, loc=loc.numpy(), scale=scale.numpy(), fnfe=np.int64(dec.lsde_func.fnfe))
    d.update(flat('f', rr.net_params(dec.lsde_func.f_func.net)))
    d.update(flat('g', rr.net_params(dec.lsde_func.g_func.net)))
    d.update(flat('head', rr.net_params(dec.decoder)))
    d.update(flat('scale_head', rr.net_params(dec.scale)))
    d['min_scale'] = np.float64(dec.min_scale)
    np.savez_compressed(os.path.join(OUT, 'decoder_solve.npz'), **d)
    print('decoder_solve: ys', ys.shape, 'queries', len(queries), 'fnfe', dec.lsde_func.fnfe)


def encoder_fixture():
    lsde, gru = rr.build_reference_encoder_sde(seed=3, bias_std=0.1)
    g = torch.Generator().manual_seed(4321)
    rows, hist = 40, 21
    hidden = torch.randn(64, generator=g) * 0.02
    h0 = hidden.unsqueeze(0).repeat(rows, 1)
    aa_out = torch.randn(hist, rows, 64, generator=g)
    actors_mask = torch.rand(rows, hist, generator=g) > 0.25
    nus_mask = torch.rand(rows, generator=g) > 0.5
    pairs = so.encoder_time_pairs_ref(2.0, hist)
    hs = torch.stack([so.euler_schedule_ref(torch.tensor([a, b]), 0.1)['h'][0] for a, b, _ in pairs])
    dW = torch.randn(hist, rows, 64, generator=g) * torch.sqrt(hs).view(hist, 1, 1)
    latent, gs, queries = rr.run_reference_encoder_loop(lsde, gru, h0, aa_out, actors_mask, nus_mask, dW)
    q = np.array([[float(a), float(b)] for a, b in queries], dtype=np.float32)
    d = dict(h0=h0.numpy(), aa_out=aa_out.numpy(), actors_mask=actors_mask.numpy(), nus_mask=nus_mask.numpy(),
             dW=dW.numpy(), latent_ys=latent.numpy(), g=gs.numpy(), queries=q)
    d.update(flat('f', rr.net_params(lsde.f_func.net)))
    d.update(flat('g_nus', rr.net_params(lsde.g_nus.net)))
    d.update(flat('g_argo', rr.net_params(lsde.g_argo.net)))
    d.update(flat('gru', rr.net_params(gru)))
    np.savez_compressed(os.path.join(OUT, 'encoder_loop.npz'), **d)
    print('encoder_loop: latent', latent.shape, 'g', gs.shape, 'queries', len(queries))


def schedule_fixture():
    """Schedules cross-checked against the (ta,tb) sequence the reference solver queries its Brownian motion with."""
    m = rr.load_reference()
    dec = rr.build_reference_decoder(seed=0, bias_std=0.0)
    d = {}
    for F in (10, 20, 30, 50, 60, 100, 200):
        ts = torch.linspace(0, 0.1 * F, F + 1)
        sched = so.euler_schedule_ref(ts, 0.1)
        S = sched['h'].numel()
        bm = m['torchsde'].FixedIncrements(torch.zeros(S + 2, 2, 64))
        with torch.no_grad():
            m['dec'].sdeint(dec.lsde_func, torch.zeros(2, 64), ts, bm=bm, dt=0.1, dt_min=0.1, rtol=1e-3, atol=1e-3,
                            method='euler')
        q = np.array([[float(a), float(b)] for a, b in bm.queries], dtype=np.float32)
        assert q.shape[0] == S, (F, q.shape, S)
        for k in ('t0', 'h', 'out_k', 'w0', 'w1'):
            d[f'F{F}/{k}'] = sched[k].numpy()
        d[f'F{F}/queries'] = q
        d[f'F{F}/ts'] = ts.numpy()
    pairs = so.encoder_time_pairs_ref(2.0, 21)
    d['enc/pairs'] = np.array([[float(a), float(b)] for a, b, _ in pairs], dtype=np.float32)
    d['enc/slot'] = np.array([t for _, _, t in pairs], dtype=np.int32)
    d['enc/t0'] = np.stack([so.euler_schedule_ref(torch.tensor([a, b]), 0.1)['t0'].numpy()[0] for a, b, _ in pairs])
    d['enc/h'] = np.stack([so.euler_schedule_ref(torch.tensor([a, b]), 0.1)['h'].numpy()[0] for a, b, _ in pairs])
    np.savez_compressed(os.path.join(OUT, 'schedule.npz'), **d)
    print('schedule: grids', [k for k in d if k.endswith('/h')])


def stage_fixture():
    """SDEDecoder.forward + losses + metrics, run verbatim from the reference (training-mode graph: gradients included)."""
    from importlib.machinery import SourceFileLoader
    m = rr.load_reference()
    ref = rr.REFERENCE_ROOT
    L2 = SourceFileLoader('L2', os.path.join(ref, 'losses/L2.py')).load_module('L2').L2
    DiffBCE = SourceFileLoader('DiffBCE', os.path.join(ref, 'losses/diff_BCE.py')).load_module('DiffBCE').DiffBCE
    ADE_T = SourceFileLoader('ADE_T', os.path.join(ref, 'metrics/ade_t.py')).load_module('ADE_T').ADE_T
    FDE_T = SourceFileLoader('FDE_T', os.path.join(ref, 'metrics/fde_t.py')).load_module('FDE_T').FDE_T
    dec = rr.build_reference_decoder(seed=5, bias_std=0.1).train()
    g = torch.Generator().manual_seed(777)
    N, modes, F_, H = 12, 10, 60, 21
    local_embed = torch.randn(N, 64, generator=g).requires_grad_(True)
    global_embed = torch.randn(modes, N, 64, generator=g).requires_grad_(True)
    source = torch.tensor([0] * 5 + [1] * 7)                    # per-actor here: one agent per scene, nuScenes scenes first
    pad = torch.ones(N, H + F_, dtype=torch.bool)               # padding_mask: True = padded
    pad[source == 0, 0:H:5] = False                             # nuScenes-shaped: past slots {0,5,..,20}, future slots {4,9,..,59}
    pad[source == 0, H + 4::5] = False
    pad[source == 1, 1:H] = False                               # Argoverse-shaped: past 1..20, future 0..29
    pad[source == 1, H:H + 30] = False
    pad[3, H + 40:] = True                                      # ragged: one actor loses its tail, one has no future at all
    pad[9, H:] = True
    y = torch.randn(N, F_, 2, generator=g).cumsum(1) * 0.5
    data = {'padding_mask': pad, 'y': y}
    sched = so.euler_schedule_ref(dec.ts_pred, dec.min_stepsize)
    S = sched['h'].numel()
    dW = torch.randn(S, modes * N, 64, generator=g) * torch.sqrt(sched['h']).view(S, 1, 1)
    glob = type(dec).forward.__globals__
    orig = glob['sdeint']
    captured = {}

    def sdeint_with_dw(sde, y0, ts, **kw):
        captured['hidden_0'] = y0.detach().clone()
        ys = orig(sde, y0, ts, bm=m['torchsde'].FixedIncrements(dW), **kw)
        captured['ys'] = ys.detach().clone()
        return ys

    glob['sdeint'] = sdeint_with_dw
    try:
        out = dec(data, local_embed, global_embed)
    finally:
        glob['sdeint'] = orig
    diff_in = (torch.rand(4, 64, generator=g) * 0.8 + 0.1).requires_grad_(True)      # what the encoder hands DiffBCE: g in (0,1), repeated
    diff_out = (torch.rand(4, 64, generator=g) * 0.8 + 0.1).requires_grad_(True)
    out['diff_in'], out['diff_out'] = diff_in, diff_out
    out['label_in'], out['label_out'] = torch.full_like(diff_in, 0), torch.full_like(diff_out, 1)   # enc…sep2.py:194-195
    l2 = L2(reduction='mean')(data, out)
    bce = DiffBCE(reduction='mean')(data, out)
    (l2 + bce).backward()                                        # loss_weights [1, 1]  (yml:86)
    agent_index = torch.arange(N)
    ade, fde = ADE_T('nuScenes', [59, 29]), FDE_T('nuScenes', [59, 29])
    pred = out['loc'][:, agent_index, :, :2].detach()
    ade.update(pred, y[agent_index], out['reg_mask'][agent_index], source)
    fde.update(pred, y[agent_index], out['reg_mask'][agent_index], source)
    d = dict(local_embed=local_embed.detach().numpy(), global_embed=global_embed.detach().numpy(), padding_mask=pad.numpy(), y=y.numpy(),
             source=source.numpy(), dW=dW.numpy(), hidden_0=captured['hidden_0'].numpy(), ys=captured['ys'].numpy(),
             loc=out['loc'].detach().numpy(), pi=out['pi'].detach().numpy(), reg_mask=out['reg_mask'].numpy(),
             diff_in=diff_in.detach().numpy(), diff_out=diff_out.detach().numpy(), loss_l2=np.float64(l2.item()),
             loss_bce=np.float64(bce.item()), ade=np.float64(float(ade.compute())), fde=np.float64(float(fde.compute())),
             grad_local_embed=local_embed.grad.numpy(), grad_global_embed=global_embed.grad.numpy(),
             grad_diff_in=diff_in.grad.numpy(), grad_diff_out=diff_out.grad.numpy(), min_scale=np.float64(dec.min_scale))
    d.update({f'param/{k}': v.detach().numpy() for k, v in dec.state_dict().items()})
    d.update({f'grad/{k}': (p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)) for k, p in dec.named_parameters()})
    np.savez_compressed(os.path.join(OUT, 'decoder_stage.npz'), **d)
    print('decoder_stage: loc', tuple(out['loc'].shape), 'L2', l2.item(), 'BCE', bce.item(), 'ADE', float(ade.compute()), 'FDE', float(fde.compute()))


def synthetic_graph_batch(scenes=4, agents=6, lanes=9, seed=11):
    """A small TemporalData-shaped batch (models/utils/util.py:20-75) with the fields LocalEncoderSDESepPara2.forward reads."""
    from models.utils.util import TemporalData
    g = torch.Generator().manual_seed(seed)
    n = scenes * agents
    x = torch.randn(n, 21, 2, generator=g)
    positions = torch.randn(n, 81, 2, generator=g).cumsum(1)
    pad = torch.rand(n, 81, generator=g) < 0.2
    pad[:, 20] = False                                          # every actor is observed at the reference time
    bos = torch.zeros(n, 21, dtype=torch.bool)
    bos[torch.arange(n), torch.argmax((~pad[:, :21]).float(), 1)] = True
    ei = torch.tensor([[i, j] for s in range(scenes) for i in range(s * agents, (s + 1) * agents)
                       for j in range(s * agents, (s + 1) * agents) if i != j]).t()
    la = torch.stack((torch.randint(0, lanes, (5 * n,), generator=g), torch.arange(n).repeat(5)))
    data = TemporalData(x=x, positions=positions, edge_index=ei, y=torch.randn(n, 60, 2, generator=g), num_nodes=n, padding_mask=pad,
                        bos_mask=bos, rotate_angles=torch.rand(n, generator=g) * 6.28, lane_positions=torch.randn(lanes, 10, 2, generator=g),
                        lane_vectors=torch.randn(lanes, 2, generator=g), lane_paddings=torch.zeros(lanes, 10),
                        lane_actor_index=la, lane_actor_vectors=torch.randn(5 * n, 2, generator=g))
    data['agent_index'] = torch.arange(scenes) * agents
    data.batch = torch.arange(scenes).repeat_interleave(agents)
    data.source = torch.arange(scenes) % 2                      # nuScenes and Argoverse scenes alternate
    ang = data['rotate_angles']
    rot = torch.empty(n, 2, 2)                                  # model_base_mix_sde.py:76-83
    rot[:, 0, 0], rot[:, 0, 1], rot[:, 1, 0], rot[:, 1, 1] = torch.cos(ang), -torch.sin(ang), torch.sin(ang), torch.cos(ang)
    data['rotate_mat'] = rot
    return data


ENC_KW = dict(historical_steps=21, node_dim=2, edge_dim=2, embed_dim=64, num_heads=8, dropout=0.1, parallel=True, local_radius=50,
              sde_layers=2, ref_time=20, max_past_t=2, run_backwards=True, minimum_step=0.1, rtol=0.001, atol=0.001, method='euler')


def encoder_stage_fixture():
    from importlib.machinery import SourceFileLoader
    m = rr.load_reference()
    DiffBCE = SourceFileLoader('DiffBCE', os.path.join(rr.REFERENCE_ROOT, 'losses/diff_BCE.py')).load_module('DiffBCE').DiffBCE
    torch.manual_seed(21)
    enc = m['enc'].LocalEncoderSDESepPara2(**ENC_KW).eval()     # eval: no dropout in the HiVT blocks
    rr._perturb_biases(enc.gru_unit, 0.1, torch.Generator().manual_seed(22))
    rr._perturb_biases(enc.lsde_func, 0.1, torch.Generator().manual_seed(23))
    data = synthetic_graph_batch()
    n, n_fake = data.x.shape[0], data['agent_index'].numel()
    rows = n + n_fake
    pairs = so.encoder_time_pairs_ref(2.0, 21)
    hs = torch.stack([so.euler_schedule_ref(torch.tensor([a, b]), 0.1)['h'][0] for a, b, _ in pairs])
    dW = torch.randn(21, rows, 64, generator=torch.Generator().manual_seed(24)) * torch.sqrt(hs).view(21, 1, 1)
    glob = type(enc).forward.__globals__
    orig, calls, cap = glob['sdeint_dual'], {'n': 0}, {'x': [], 'mask': [], 'slot': []}

    def with_dw(sde, y0, ts, nus_mask, **kw):
        k = calls['n']
        calls['n'] += 1
        cap['nus_mask'] = nus_mask.clone()
        return orig(sde, y0, ts, nus_mask, bm=m['torchsde'].FixedIncrements(dW[k:k + 1]), **kw)

    def gru_pre(mod, args, kwargs):
        kwargs['input_tensor'].retain_grad()
        cap['x'].append(kwargs['input_tensor'])
        cap['mask'].append(kwargs['mask'].clone())

    def al_pre(mod, args, kwargs):
        kwargs['x'][1].retain_grad()
        cap['pre_al'] = kwargs['x'][1]

    h1 = enc.gru_unit.register_forward_pre_hook(gru_pre, with_kwargs=True)
    h2 = enc.al_encoder.register_forward_pre_hook(al_pre, with_kwargs=True)
    glob['sdeint_dual'] = with_dw
    torch.manual_seed(25)                                       # the perturbed target copies draw from the global RNG (enc…sep2.py:95)
    try:
        out, d_in, d_out, l_in, l_out = enc(data)
    finally:
        glob['sdeint_dual'] = orig
        h1.remove(); h2.remove()
    assert calls['n'] == 21
    loss = out.square().mean() + DiffBCE(reduction='mean')(data, {'diff_in': d_in, 'diff_out': d_out, 'label_in': l_in, 'label_out': l_out})
    loss.backward()
    slots = [t for _, _, t in pairs]                            # iteration k consumed aa_out[slots[k]]
    aa_out = torch.zeros(21, rows, 64)
    g_aa = torch.zeros(21, rows, 64)
    actors_mask = torch.zeros(rows, 21, dtype=torch.bool)
    for k, t in enumerate(slots):
        aa_out[t], actors_mask[:, t] = cap['x'][k].detach(), cap['mask'][k]
        g_aa[t] = cap['x'][k].grad if cap['x'][k].grad is not None else 0
    d = dict(aa_out=aa_out.numpy(), actors_mask=actors_mask.numpy(), nus_mask=cap['nus_mask'].numpy(), dW=dW.numpy(),
             bos_mask=data['bos_mask'].numpy(), agent_index=data['agent_index'].numpy(), pre_al=cap['pre_al'].detach().numpy(),
             out=out.detach().numpy(), diff_in=d_in.detach().numpy(), diff_out=d_out.detach().numpy(), label_in=l_in.numpy(),
             label_out=l_out.numpy(), loss=np.float64(loss.item()), grad_aa_out=g_aa.numpy(), grad_pre_al=cap['pre_al'].grad.numpy(),
             torch_seed_fake_agents=np.int64(25), init_seed=np.int64(21))
    keep = ('gru_unit.', 'lsde_func.f_func', 'lsde_func.g_nus', 'lsde_func.g_argo', 'hidden')
    d.update({f'param/{k}': v.detach().numpy() for k, v in enc.state_dict().items() if k.startswith(keep)})
    d.update({f'grad/{k}': p.grad.numpy() for k, p in enc.named_parameters() if k.startswith(keep) and p.grad is not None})
    np.savez_compressed(os.path.join(OUT, 'encoder_stage.npz'), **d)
    print('encoder_stage: out', tuple(out.shape), 'diff_in', tuple(d_in.shape), 'rows', rows, 'loss', loss.item())


if __name__ == '__main__':
    torch.set_num_threads(1)       # fixture bytes must not depend on the thread count of the generating host
    decoder_fixture()
    encoder_fixture()
    schedule_fixture()
    stage_fixture()
    encoder_stage_fixture()
