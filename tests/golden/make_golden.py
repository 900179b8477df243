"""Generate tests/golden/*.npz by running the reference's own files (oracle/ref_runner.py) in the dev container.

    python tests/golden/make_golden.py

Fixtures (all float32 unless noted, seeds fixed):
  decoder_solve.npz   reference SDEDecoder.lsde_func + torchsde.sdeint call (dec…sde.py:88): 48 rows, 61 steps, dW supplied;
                      weights (non-zero biases), y0, dW, ys, Brownian query times (ta,tb), loc/scale head parameters and outputs.
  encoder_loop.npz    reference sdeint_dual + GRU_Unit loop (enc…sep2.py:128-182): 40 rows, 21 steps, dual g, masks.
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


if __name__ == '__main__':
    torch.set_num_threads(1)       # fixture bytes must not depend on the thread count of the generating host
    decoder_fixture()
    encoder_fixture()
    schedule_fixture()
