"""Row (g) of the scope table: the reference-shaped stage ``forward`` with ``install()`` active, on the GPU, against the fixture the
REAL reference stage produced (tests/golden/decoder_stage.npz): outputs, the L2 / DiffBCE losses, ADE_T / FDE_T, and the gradients of a
training step (reference's own backward).  Also: the tc_f16 ADE/FDE pin, forward_ood against an oracle replay of the drawn increments,
the non-silent adjoint-range status, fused heads under torch.inference_mode(), the empty-shard encoder backward.

Tolerances (tc_f16 = fp16 operands + MUFU tanh over 61 steps; measured values are printed by the tests):
    latents          atol 2e-2                      (measured 7e-3)
    loc / scale      atol 9e-3 .. 1.1e-2            (measured 2.9e-3 fixture, 3.6e-3 at 204,800 rows)
    L2 loss, ADE_T, FDE_T   |delta| <= 5e-4 .. 1e-3 (metres; measured <= 3.4e-4)   north_star's ADE/FDE agreement
    gradients        3e-2 of the max-norm           (measured 1.8e-2 through the whole stage, 4e-4 for dL/dy0 of the solve alone)
"""
import warnings

import pytest
import torch

import ref_shaped
import trajsde_b200 as tb
from helpers import DecoderSDE, EncoderSDE, init_like_reference, net_params
from oracle import sde_oracle as so
from trajsde_b200 import encoder as enc_mod
from trajsde_b200 import ops, patch, synthetic as syn
from trajsde_b200.schedule import encoder_schedule

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _t(d, k, dtype=None):
    x = torch.from_numpy(d[k]).to(DEV)
    return x if dtype is None else x.to(dtype)


def _installed_stage(d, mode):
    sd = {k[len('param/'):]: torch.from_numpy(d[k]) for k in d if k.startswith('param/')}
    dec = ref_shaped.RefShapedDecoder().load_reference_state_dict(sd).to(DEV)
    old = tb.get_default_mode()
    tb.set_default_mode(mode)
    saved = patch.install(decoder=dec)
    g = type(dec).forward.__globals__
    installed = g['sdeint']
    dW = _t(d, 'dW')
    g['sdeint'] = lambda sde, y0, ts, **kw: installed(sde, y0, ts, bm=dW, **kw)        # the fixture's increments instead of Philox

    def undo():
        g['sdeint'] = installed
        patch.uninstall(saved)
        tb.set_default_mode(old)
    return dec, saved, undo


@pytest.mark.parametrize('ctx', ['no_grad', 'inference_mode'])
@pytest.mark.parametrize('mode', ['exact', 'tc_f16'])
def test_stage_forward_interop_vs_reference_fixture(mode, ctx, golden_stage):
    d = golden_stage
    dec, saved, undo = _installed_stage(d, mode)
    try:
        assert ('heads' in saved) == (mode == 'tc_f16')
        n0 = ops.LAUNCHES['n']
        with (torch.no_grad() if ctx == 'no_grad' else torch.inference_mode()):       # Lightning validates / tests under inference_mode
            out = dec({'padding_mask': _t(d, 'padding_mask')}, _t(d, 'local_embed'), _t(d, 'global_embed'))
        launched = ops.LAUNCHES['n'] - n0
    finally:
        undo()
    assert launched == (4 if mode == 'tc_f16' else 1)          # tc: pack + solve + (pack + fused heads); exact: the solve
    loc, ref = out['loc'].cpu(), torch.from_numpy(d['loc'])
    assert loc.shape == (10, 12, 60, 4)
    err = (loc - ref).abs().max().item()
    print(f"[{mode}/{ctx}] stage forward: loc|scale max-abs {err:.3e}")
    assert err < (2e-4 if mode == 'exact' else 9e-3)                # measured 1.5e-6 / 2.9e-3
    assert torch.allclose(out['pi'].cpu(), torch.from_numpy(d['pi']), atol=1e-5, rtol=1e-5)
    assert torch.equal(out['reg_mask'].cpu(), torch.from_numpy(d['reg_mask']))
    y, rm, src = torch.from_numpy(d['y']), torch.from_numpy(d['reg_mask']), torch.from_numpy(d['source'])
    dl2 = abs(float(so.l2_loss_ref(loc, y, rm)) - float(d['loss_l2']))
    dade = abs(so.ade_t_ref(loc[..., :2], y, rm) - float(d['ade']))
    dfde = abs(so.fde_t_ref(loc[..., :2], y, rm, src) - float(d['fde']))
    print(f"[{mode}/{ctx}] |L2 - ref| {dl2:.2e}  |ADE - ref| {dade:.2e}  |FDE - ref| {dfde:.2e}")
    assert max(dl2, dade, dfde) < (1e-5 if mode == 'exact' else 5e-4)   # measured 5e-7 / 1.4e-4 (north_star bound: 1e-3 m)


@pytest.mark.parametrize('mode', ['exact', 'tc_f16'])
def test_training_step_through_the_stage_vs_reference_gradients(mode, golden_stage):
    """forward (autograd on) -> L2 + DiffBCE -> backward, against the gradients the reference's own stage / losses / solver produced
    for the same inputs and increments."""
    d = golden_stage
    dec, saved, undo = _installed_stage(d, mode)
    try:
        le, ge = _t(d, 'local_embed').requires_grad_(True), _t(d, 'global_embed').requires_grad_(True)
        di, do = _t(d, 'diff_in').requires_grad_(True), _t(d, 'diff_out').requires_grad_(True)
        out = dec({'padding_mask': _t(d, 'padding_mask')}, le, ge)
        l2 = so.l2_loss_ref(out['loc'], _t(d, 'y'), out['reg_mask'])
        bce = so.diff_bce_ref(di, do)
        (l2 + bce).backward()
        ops.poll_status(DEV, block=True)
    finally:
        undo()
    assert abs(float(l2) - float(d['loss_l2'])) < (1e-5 if mode == 'exact' else 1e-3)
    assert abs(float(bce) - float(d['loss_bce'])) < 1e-5
    tol = 2e-4 if mode == 'exact' else 3e-2
    worst = 0.0
    pairs = [('local_embed', le.grad, d['grad_local_embed']), ('global_embed', ge.grad, d['grad_global_embed']),
             ('diff_in', di.grad, d['grad_diff_in']), ('diff_out', do.grad, d['grad_diff_out'])]
    pairs += [(k, p.grad, d['grad/' + k]) for k, p in dec.named_parameters() if k != 'hidden']
    for name, got, ref in pairs:
        ref = torch.from_numpy(ref)
        if float(ref.abs().max()) == 0.0:
            assert got is None or float(got.abs().max()) == 0.0, name
            continue
        e = float((got.cpu() - ref).abs().max() / ref.abs().max())
        worst = max(worst, e)
        assert e < tol, (name, e)
    print(f"[{mode}] training step through the stage: worst relative gradient error {worst:.2e}")


def test_ade_fde_pin_at_full_size_tc_f16():
    """north_star: 'ADE/FDE agreement on decoded trajectories'.  204,800 decoder rows (1024 scenes x 20 agents x 10 modes) solved in
    tc_f16 mode with in-kernel Philox, decoded by the fused heads; 8 blocks of 4 actors x 10 modes are replayed through the oracle
    (same increments, dumped with trajsde_philox_dw) and compared on minADE / minFDE against a synthetic ground truth."""
    import torch.nn as nn
    from trajsde_b200 import heads as hd
    from trajsde_b200.schedule import euler_schedule
    b = syn.make_batch(1024, 20, seed=77)
    sde = init_like_reference(DecoderSDE(), seed=21).to(DEV)
    mk = lambda s: init_like_reference(nn.Sequential(nn.Linear(64, 64), nn.LayerNorm(64), nn.ReLU(inplace=True), nn.Linear(64, 2)), s).to(DEV)  # noqa: E731
    loc_h, sc_h = mk(41), mk(42)
    ts = torch.linspace(0, 6, 61)
    y0 = b.dec_y0.to(DEV)
    M, N, modes = y0.shape[0], 20480, 10
    with torch.no_grad():
        ys = tb.sdeint(sde, y0, ts, dt=0.1, method='euler', mode='tc_f16', seed=99, rows_major=True)
        loc, _ = hd.decoder_heads(loc_h, sc_h, ys[1:].permute(1, 0, 2))
    loc = loc.view(modes, N, 60, 2)
    dsched = ops.DeviceSchedule.get(euler_schedule(ts, 0.1), torch.device(DEV))
    pf, pg = net_params(sde.f_func), net_params(sde.g_func)
    ph = {k: v.detach().cpu() for k, v in loc_h.state_dict().items()}
    gen = torch.Generator().manual_seed(3)
    worst = (0.0, 0.0, 0.0)
    for a in torch.linspace(0, N - 4, 8).long().tolist():
        rows = [m * N + a for m in range(modes)]
        dW = torch.cat([ops.philox_dw(dsched, 4, 99, torch.device(DEV), row_offset=r) for r in rows], dim=1).cpu()   # [61, 40, 64]
        y0_blk = torch.cat([b.dec_y0[r:r + 4] for r in rows])
        ref_ys, _ = so.euler_solve_ref(pf, pg, y0_blk, ts, 0.1, dW)
        ref_loc = so.decoder_loc_head_ref(ph, ref_ys[1:].permute(1, 0, 2)).view(modes, 4, 60, 2)
        got = loc[:, a:a + 4].cpu()
        target = ref_loc[3] + torch.randn(4, 60, 2, generator=gen) * 0.3               # ground truth near one mode, 30 valid slots
        rm = torch.zeros(4, 60, dtype=torch.bool); rm[:, :30] = True
        src = torch.ones(4, dtype=torch.long)
        e_loc = float((got - ref_loc).abs().max())
        e_ade = abs(so.ade_t_ref(got, target, rm) - so.ade_t_ref(ref_loc, target, rm))
        e_fde = abs(so.fde_t_ref(got, target, rm, src) - so.fde_t_ref(ref_loc, target, rm, src))
        worst = (max(worst[0], e_loc), max(worst[1], e_ade), max(worst[2], e_fde))
    print(f"full-size tc_f16 chain vs oracle: loc max-abs {worst[0]:.2e}, |dADE| {worst[1]:.2e}, |dFDE| {worst[2]:.2e}")
    assert worst[0] < 1.1e-2 and worst[1] < 2e-4 and worst[2] < 1e-3     # measured 3.6e-3, 5.9e-5, 3.4e-4


def test_forward_ood_vs_oracle_replay():
    """forward_ood (enc…sep2.py:252-313): 10 Monte-Carlo passes from a ZERO state (:257) -> mean latent / per-actor std (:311-313).
    The fused launch's increments are dumped with trajsde_philox_dw (pass j = global rows [j*rows, (j+1)*rows)) and the 10 passes are
    replayed through the oracle's recurrence."""
    sde = init_like_reference(EncoderSDE(), seed=1, bias_std=0.1).to(DEV)
    gru = syn.init_reference_style(syn.GRUUnit(), 2, bias_std=0.1).to(DEV)
    b = syn.make_batch(6, 10, seed=3, mixed_sources=True)
    n, k = 60, 10
    aa, am, nm, bos = b.aa_out[:, :n].contiguous(), b.actors_mask[:n], b.nus_mask[:n], b.bos_mask
    mean, std = enc_mod.encoder_recurrence_ood(sde, gru, aa.to(DEV), am.to(DEV), nm.to(DEV), bos.to(DEV), eval_iter=k, seed=5, mode='tc_f16')
    dsched = ops.DeviceSchedule.get(encoder_schedule(), torch.device(DEV))
    dW = ops.philox_dw(dsched, k * n, 5, torch.device(DEV)).cpu()                      # [21, 10*60, 64]
    pe = [net_params(sde.f_func), net_params(sde.g_nus), net_params(sde.g_argo)]
    pgru = {kk: v.detach().cpu() for kk, v in gru.state_dict().items()}
    outs = []
    for j in range(k):
        lat, _ = so.encoder_recurrence_ref(pe[0], pe[1], pe[2], pgru, torch.zeros(n, 64), aa, am, nm, dW[:, j * n:(j + 1) * n])
        outs.append(so.encoder_eos_gather_ref(lat, bos))
    outs = torch.stack(outs)
    ref_mean, ref_std = outs.mean(0), outs.std(0).mean(-1)
    e_m, e_s = float((mean.cpu() - ref_mean).abs().max()), float((std.cpu() - ref_std).abs().max())
    print(f"forward_ood vs oracle replay: mean max-abs {e_m:.2e}, std max-abs {e_s:.2e} (std range {float(ref_std.min()):.3f}..{float(ref_std.max()):.3f})")
    assert e_m < 2e-3 and e_s < 2e-4                                 # measured 5.4e-4, 4.2e-5
    # the stepwise exact path (10 separate passes, per-pass seeds) is a different Monte-Carlo sample of the same distribution
    mean_e, std_e = enc_mod.encoder_recurrence_ood(sde, gru, aa.to(DEV), am.to(DEV), nm.to(DEV), bos.to(DEV), eval_iter=k, seed=5, mode='exact')
    assert mean_e.shape == mean.shape and (std_e > 0).all()


def test_adjoint_range_status_is_not_silent():
    """A clipped adjoint surfaces by itself at the next call into the library: RuntimeWarning (default policy) or AdjointRangeError."""
    ts = torch.linspace(0, 6, 61)
    y0 = torch.relu(torch.randn(200, 64, generator=torch.Generator().manual_seed(1))).to(DEV)
    sde = init_like_reference(DecoderSDE(), seed=2, bias_std=0.2).to(DEV)
    with torch.no_grad():
        for p_ in sde.f_func.parameters():
            p_.mul_(40.0)

    def step():
        y = y0.clone().requires_grad_(True)
        tb.sdeint(sde, y, ts, dt=0.1, method='euler', mode='tc_f16', seed=5)[-1].sum().backward()

    try:
        ops.backward_status(DEV)                                  # clear
        ops.set_adjoint_range_policy('warn')
        step()
        torch.cuda.synchronize()
        with pytest.warns(RuntimeWarning, match='ADJOINT_RANGE'):
            tb.sdeint(sde, y0, ts, dt=0.1, method='euler', mode='tc_f16', seed=5)      # any next call reports it
        ops.backward_status(DEV)
        ops.set_adjoint_range_policy('raise')
        step()
        with pytest.raises(ops.AdjointRangeError):
            ops.poll_status(DEV, block=True)                      # once per optimizer step (FlatGradBucket.all_reduce_mean does this)
        ops.set_adjoint_range_policy('warn')
        good = init_like_reference(DecoderSDE(), seed=2, bias_std=0.2).to(DEV)
        with warnings.catch_warnings():
            warnings.simplefilter('error')
            y = y0.clone().requires_grad_(True)
            tb.sdeint(good, y, ts, dt=0.1, method='euler', mode='tc_f16', seed=5)[-1].sum().backward()
            ops.poll_status(DEV, block=True)                      # reference-style nets: nothing to report
    finally:
        ops.set_adjoint_range_policy('warn')
        ops.backward_status(DEV)


def test_empty_shard_encoder_backward_returns_zero_parameter_gradients():
    """rows == 0 (an empty data-parallel shard): trajsde_enc_bwd must still WRITE the parameter gradients (zeros) — they are
    accumulated into .grad and all-reduced across ranks."""
    sde = init_like_reference(EncoderSDE(), seed=1).to(DEV)
    gru = syn.init_reference_style(syn.GRUUnit(), 2).to(DEV)
    h0 = torch.zeros(0, 64, device=DEV, requires_grad=True)
    aa = torch.zeros(21, 0, 64, device=DEV)
    lat, g = enc_mod.encoder_recurrence(sde, gru, h0, aa, torch.ones(0, 21, dtype=torch.bool, device=DEV),
                                        torch.zeros(0, dtype=torch.bool, device=DEV), fused=True, seed=3)
    assert lat.shape == (21, 0, 64) and g.shape == (21, 0)
    junk = torch.full((1 << 20,), float('nan'), device=DEV)      # poison the allocator's free blocks
    del junk
    (lat.sum() + g.sum()).backward()
    for p_ in list(sde.parameters()) + list(gru.parameters()):
        assert p_.grad is not None and torch.equal(p_.grad, torch.zeros_like(p_.grad))
