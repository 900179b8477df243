"""BASELINE configs[1] at FULL size on one B200 (1024 scenes x 20 agents: decoder 204,800 rows x 61 steps, encoder 21,504 rows x 21
iterations, heads 12.3 M points): the oracle cannot run these sizes, so parity is checked through size-independent properties —
rows are independent (any row range solved alone, keyed by its global row offset, reproduces its part of the full solve), the
result is deterministic, `ys[0] == y0`, both storage layouts agree, weight gradients are additive over row ranges — plus
spot checks of scattered row blocks against the CPU oracle under the increments the kernels actually drew.

Tolerances as in the small-size tests (tc_f16: atol 2e-2 on latents up to ~16 — measured 7e-3; 3e-2 of the max-norm on gradients — measured <= 6e-3)."""
import pytest
import torch

import trajsde_b200 as tb
from helpers import DecoderSDE, EncoderSDE, init_like_reference, net_params
from oracle import sde_oracle as so
from trajsde_b200 import encoder as enc_mod
from trajsde_b200 import heads as hd
from trajsde_b200 import ops, synthetic as syn
from trajsde_b200.schedule import euler_schedule

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
SCENES, AGENTS = 1024, 20
TOL = dict(atol=2e-2, rtol=0)


@pytest.fixture(scope='module')
def batch():
    return syn.make_batch(SCENES, AGENTS, seed=77, mixed_sources=True)


def test_decoder_full_size_properties_and_oracle_spot_checks(batch):
    sde = init_like_reference(DecoderSDE(), seed=21).to(DEV)
    ts = torch.linspace(0, 6, 61)
    sched = euler_schedule(ts, 0.1)
    y0 = batch.dec_y0.to(DEV)
    M = y0.shape[0]
    assert M == 204800
    with torch.no_grad():
        ys = tb.sdeint(sde, y0, ts, dt=0.1, method='euler', mode='tc_f16', seed=99)
        assert ys.shape == (61, M, 64) and torch.equal(ys[0], y0)
        assert torch.isfinite(ys).all()
        assert torch.equal(ys, tb.sdeint(sde, y0, ts, dt=0.1, method='euler', mode='tc_f16', seed=99))          # deterministic
        ys_rm = tb.sdeint(sde, y0, ts, dt=0.1, method='euler', mode='tc_f16', seed=99, rows_major=True)         # same values, [rows,T,64] storage
        assert ys_rm.stride(1) == 61 * 64 and torch.equal(ys_rm, ys)
        del ys_rm
        # row independence: ranges that start and end off the 128-row tile grid, solved alone with their global offset
        for a, b in ((0, 1000), (77_777, 91_003), (204_800 - 4_321, 204_800)):
            part = tb.sdeint(sde, y0[a:b], ts, dt=0.1, method='euler', mode='tc_f16', seed=99, row_offset=a)
            assert torch.allclose(part, ys[:, a:b], atol=1e-4, rtol=0), (a, b)
        # oracle spot checks: 8 blocks of 8 rows, replaying the increments the kernel drew for exactly those rows
        dsched = ops.DeviceSchedule.get(sched, torch.device(DEV))
        pf, pg = net_params(sde.f_func), net_params(sde.g_func)
        for a in torch.linspace(0, M - 8, 8).long().tolist():
            dW = ops.philox_dw(dsched, 8, 99, torch.device(DEV), row_offset=a)
            ref, _ = so.euler_solve_ref(pf, pg, y0[a:a + 8].cpu(), ts, 0.1, dW.cpu())
            assert torch.allclose(ys[:, a:a + 8].cpu(), ref, **TOL), a


def test_decoder_full_size_gradients_are_additive_over_row_ranges(batch):
    """dL/dW of the full batch = sum over disjoint row ranges; dL/dy0 of a range = its slice (one launch of 204,800 rows x 61 steps
    against three launches over its parts; ranges cut off the tile grid)."""
    sde = init_like_reference(DecoderSDE(), seed=22).to(DEV)
    ts = torch.linspace(0, 6, 61)
    y0 = batch.dec_y0.to(DEV)
    M = y0.shape[0]
    cot = torch.randn(61, M, 64, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5)) * (1.0 / (M * 60))
    cot[0].zero_()

    def run(a, b):
        for p in sde.parameters():
            p.grad = None
        y = y0[a:b].detach().clone().requires_grad_(True)
        ys = tb.sdeint(sde, y, ts, dt=0.1, method='euler', mode='tc_f16', seed=123, row_offset=a)
        ys.backward(cot[:, a:b])
        return y.grad, [p.grad.clone() for p in sde.parameters()]

    gy_full, gw_full = run(0, M)
    cuts = [0, 50_001, 131_313, M]
    gw_sum = [torch.zeros_like(g) for g in gw_full]
    for a, b in zip(cuts[:-1], cuts[1:]):
        gy, gw = run(a, b)
        scale = gy_full.abs().max()
        assert (gy - gy_full[a:b]).abs().max() <= 3e-2 * scale
        for s, g in zip(gw_sum, gw):
            s += g
    for name, full, parts in zip([n for n, _ in sde.named_parameters()], gw_full, gw_sum):
        assert torch.isfinite(full).all()
        assert (full - parts).abs().max() <= 3e-2 * full.abs().max() + 1e-12, name
    assert ops.backward_status(torch.device(DEV)) == 0         # adjoint stayed inside the fp16 operand range
    # oracle spot checks of dL/dy0 at full size: fp64 autograd through the oracle solve of 8-row blocks under the increments the
    # kernel drew for exactly those rows (rows are independent, so dL/dy0 of a block needs only that block)
    from trajsde_b200.schedule import euler_schedule
    dsched = ops.DeviceSchedule.get(euler_schedule(ts, 0.1), torch.device(DEV))
    pf = {k: v.double() for k, v in net_params(sde.f_func).items()}
    pg = {k: v.double() for k, v in net_params(sde.g_func).items()}
    scale, worst = float(gy_full.abs().max()), 0.0
    for a in torch.linspace(0, M - 8, 8).long().tolist():
        dW = ops.philox_dw(dsched, 8, 123, torch.device(DEV), row_offset=a).cpu().double()
        y = y0[a:a + 8].cpu().double().requires_grad_(True)
        ref_ys, _ = so.euler_solve_ref(pf, pg, y, ts, 0.1, dW)
        (ref_ys * cot[:, a:a + 8].cpu().double()).sum().backward()
        worst = max(worst, float((gy_full[a:a + 8].cpu().double() - y.grad).abs().max()) / scale)
    print(f"full-size dL/dy0 vs fp64 oracle autograd: worst block error {worst:.2e} of the max-norm")
    assert worst < 2e-3                                     # measured 4.2e-4


def test_encoder_full_size_properties_and_oracle_spot_checks(batch):
    enc = init_like_reference(EncoderSDE(), seed=31).to(DEV)
    gru = syn.init_reference_style(syn.GRUUnit(), 3, bias_std=0.1).to(DEV)
    tr = {k: getattr(batch, k).to(DEV) for k in ('enc_h0', 'aa_out', 'actors_mask', 'nus_mask')}
    E = tr['enc_h0'].shape[0]
    assert E == 21504
    with torch.no_grad():
        lat, g = enc_mod.encoder_recurrence(enc, gru, tr['enc_h0'], tr['aa_out'], tr['actors_mask'], tr['nus_mask'], seed=7)
        assert lat.shape == (21, E, 64) and g.shape == (21, E) and torch.isfinite(lat).all()
        lat2, g2 = enc_mod.encoder_recurrence(enc, gru, tr['enc_h0'], tr['aa_out'], tr['actors_mask'], tr['nus_mask'], seed=7)
        assert torch.equal(lat, lat2) and torch.equal(g, g2)
        for a, b in ((0, 333), (10_001, 12_345), (E - 999, E)):
            pl, pg_ = enc_mod.encoder_recurrence(enc, gru, tr['enc_h0'][a:b], tr['aa_out'][:, a:b].contiguous(), tr['actors_mask'][a:b],
                                                 tr['nus_mask'][a:b], seed=7, row_offset=a)
            assert torch.allclose(pl, lat[:, a:b], atol=1e-4, rtol=0) and torch.allclose(pg_, g[:, a:b], atol=1e-5, rtol=0)
        # oracle spot checks under supplied increments (the recurrence compounds 21 steps + 21 GRU jumps)
        dW = torch.randn(21, E, 64, device=DEV, generator=torch.Generator(device=DEV).manual_seed(8)) * (0.1 ** 0.5)
        lat, g = enc_mod.encoder_recurrence(enc, gru, tr['enc_h0'], tr['aa_out'], tr['actors_mask'], tr['nus_mask'], dW=dW)
        pe = [net_params(enc.f_func), net_params(enc.g_nus), net_params(enc.g_argo)]
        pgru = {k: v.detach().cpu() for k, v in gru.state_dict().items()}
        for a in torch.linspace(0, E - 16, 6).long().tolist():
            sl = slice(a, a + 16)
            rl, rg = so.encoder_recurrence_ref(pe[0], pe[1], pe[2], pgru, batch.enc_h0[sl], batch.aa_out[:, sl], batch.actors_mask[sl],
                                               batch.nus_mask[sl], dW[:, sl].cpu())
            assert torch.allclose(lat[:, sl].cpu(), rl, **TOL), a
            assert torch.allclose(g[:, sl].cpu(), rg[..., 0], atol=3e-3, rtol=0), a


def test_heads_full_size_spot_checks_and_row_independence(batch):
    import torch.nn as nn
    mk = lambda sd: init_like_reference(nn.Sequential(nn.Linear(64, 64), nn.LayerNorm(64), nn.ReLU(inplace=True), nn.Linear(64, 2)), sd).to(DEV)  # noqa: E731
    loc_h, sc_h = mk(41), mk(42)
    M, T = 204800, 60
    store = torch.randn(M, T + 1, 64, device=DEV, generator=torch.Generator(device=DEV).manual_seed(9)) * 2.0
    sol_y = store[:, 1:]
    with torch.no_grad():
        loc, sc = hd.decoder_heads(loc_h, sc_h, sol_y)
        assert loc.shape == (M, T, 2) and torch.isfinite(loc).all() and torch.isfinite(sc).all()
        part_loc, part_sc = hd.decoder_heads(loc_h, sc_h, sol_y[100_003:100_900])
        assert torch.equal(part_loc, loc[100_003:100_900]) and torch.equal(part_sc, sc[100_003:100_900])
        pl = {k: v.detach().cpu() for k, v in loc_h.state_dict().items()}
        ps = {k: v.detach().cpu() for k, v in sc_h.state_dict().items()}
        for a in torch.linspace(0, M - 4, 10).long().tolist():
            x = sol_y[a:a + 4].cpu()
            assert torch.allclose(loc[a:a + 4].cpu(), so.decoder_loc_head_ref(pl, x), atol=2e-2, rtol=2e-2)
            assert torch.allclose(sc[a:a + 4].cpu(), so.decoder_loc_head_ref(ps, x), atol=2e-2, rtol=2e-2)
