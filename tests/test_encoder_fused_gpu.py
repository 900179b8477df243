"""Fused encoder recurrence (21 x [Euler step + GRU jump] in one kernel) against the reference loop fixture, the oracle
and the stepwise path."""
import pytest
import torch

import trajsde_b200 as tb
from conftest import sub
from helpers import EncoderSDE, init_like_reference, load_net, net_params
from oracle import sde_oracle as so
from trajsde_b200 import encoder as enc
from trajsde_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def gru_from(params):
    g = syn.GRUUnit()
    g.load_state_dict(params)
    return g.to(DEV)


def test_fused_encoder_vs_reference_fixture(golden_encoder):
    """Reference's own sdeint_dual + GRU_Unit loop (tests/golden/encoder_loop.npz): 40 rows, dual g, ragged masks."""
    e = golden_encoder
    sde = EncoderSDE()
    load_net(sde.f_func, sub(e, 'f')); load_net(sde.g_nus, sub(e, 'g_nus')); load_net(sde.g_argo, sub(e, 'g_argo'))
    sde = sde.to(DEV)
    gru = gru_from(sub(e, 'gru'))
    t = lambda k: torch.from_numpy(e[k]).to(DEV)  # noqa: E731
    with torch.no_grad():
        lat, g = enc.encoder_recurrence(sde, gru, t('h0'), t('aa_out'), t('actors_mask'), t('nus_mask'), dW=t('dW'), mode='tc_f16',
                                        fused=True)
    assert lat.shape == (21, 40, 64) and g.shape == (21, 40)
    ref_lat, ref_g = torch.from_numpy(e['latent_ys']), torch.from_numpy(e['g'])[:, :, 0]
    err, gerr = (lat.cpu() - ref_lat).abs().max().item(), (g.cpu() - ref_g).abs().max().item()
    print(f"fused encoder vs reference fixture: latent max-abs {err:.3e}, g max-abs {gerr:.3e}")
    assert err < 2e-2 and gerr < 3e-3


@pytest.mark.parametrize('rows,mixed', [(1, True), (127, True), (128, False), (300, True), (1000, True)])
def test_fused_encoder_vs_oracle_and_stepwise(rows, mixed):
    sde = init_like_reference(EncoderSDE(), seed=rows, bias_std=0.2).to(DEV)
    gru = syn.init_reference_style(syn.GRUUnit(), rows + 1, bias_std=0.2).to(DEV)
    g = torch.Generator().manual_seed(rows)
    h0 = (torch.randn(64, generator=g) * 0.02).repeat(rows, 1)
    aa = torch.randn(21, rows, 64, generator=g)
    am = torch.rand(rows, 21, generator=g) > 0.3
    nm = (torch.rand(rows, generator=g) > 0.5) if mixed else torch.zeros(rows, dtype=torch.bool)
    dW = torch.randn(21, rows, 64, generator=g) * 0.3
    ref_lat, ref_g = so.encoder_recurrence_ref(net_params(sde.f_func), net_params(sde.g_nus), net_params(sde.g_argo),
                                               {k: v.detach().cpu() for k, v in gru.state_dict().items()}, h0, aa, am, nm, dW)
    with torch.no_grad():
        lat, gg = enc.encoder_recurrence(sde, gru, h0.to(DEV), aa.to(DEV), am.to(DEV), nm.to(DEV), dW=dW.to(DEV), fused=True)
        lat_s, gg_s = enc.encoder_recurrence(sde, gru, h0.to(DEV), aa.to(DEV), am.to(DEV), nm.to(DEV), dW=dW.to(DEV), fused=False,
                                             mode='exact')
    assert (lat_s.cpu() - ref_lat).abs().max() < 5e-5                         # stepwise exact path = oracle
    assert (lat.cpu() - ref_lat).abs().max() < 2e-2
    assert (gg.cpu() - ref_g[:, :, 0]).abs().max() < 3e-3
    out = enc.eos_gather(lat, torch.nn.functional.one_hot(torch.randint(0, 21, (rows,), generator=g), 21).bool().to(DEV))
    assert out.shape == (rows, 64)


def test_fused_encoder_philox_deterministic_and_sharding_invariant():
    rows = 260
    sde = init_like_reference(EncoderSDE(), seed=1).to(DEV)
    gru = syn.init_reference_style(syn.GRUUnit(), 2).to(DEV)
    b = syn.make_batch(13, 19, seed=3, mixed_sources=True)
    h0, aa, am, nm = (x.to(DEV) for x in (b.enc_h0, b.aa_out, b.actors_mask, b.nus_mask))
    assert h0.shape[0] == rows
    with torch.no_grad():
        a1, g1 = enc.encoder_recurrence(sde, gru, h0, aa, am, nm, seed=11, fused=True)
        a2, g2 = enc.encoder_recurrence(sde, gru, h0, aa, am, nm, seed=11, fused=True)
        a3, _ = enc.encoder_recurrence(sde, gru, h0, aa, am, nm, seed=12, fused=True)
        p1, _ = enc.encoder_recurrence(sde, gru, h0[100:], aa[:, 100:].contiguous(), am[100:], nm[100:], seed=11, fused=True,
                                       row_offset=100)
    assert torch.equal(a1, a2) and torch.equal(g1, g2) and not torch.equal(a1, a3)
    assert torch.allclose(p1, a1[:, 100:], atol=1e-4, rtol=0)
    assert torch.isfinite(a1).all() and (g1 > 0).all() and (g1 < 1).all()


def test_fused_encoder_trains_and_rejects_exact_mode():
    sde = init_like_reference(EncoderSDE(), seed=1).to(DEV)
    gru = syn.GRUUnit().to(DEV)
    h0 = torch.zeros(4, 64, device=DEV)
    aa = torch.zeros(21, 4, 64, device=DEV)
    am = torch.ones(4, 21, dtype=torch.bool, device=DEV)
    nm = torch.zeros(4, dtype=torch.bool, device=DEV)
    lat, g = enc.encoder_recurrence(sde, gru, h0, aa, am, nm, fused=True, seed=3)      # grad enabled + parameters require grad
    assert lat.requires_grad and g.requires_grad
    (lat.sum() + g.sum()).backward()
    assert all(p_.grad is not None and torch.isfinite(p_.grad).all() for p_ in gru.parameters())
    assert sde.g_nus.net[0].weight.grad.abs().max() == 0                                 # no nuScenes row in this batch
    assert sde.g_argo.net[0].weight.grad.abs().max() > 0
    with torch.no_grad(), pytest.raises(NotImplementedError):
        enc.encoder_recurrence(sde, gru, h0, aa, am, nm, fused=True, mode='exact')


def test_forward_ood_monte_carlo_encoder():
    """forward_ood (enc…sep2.py:252-313): 10 passes from a zero state -> mean latent and per-actor std."""
    sde = init_like_reference(EncoderSDE(), seed=1).to(DEV)
    gru = syn.init_reference_style(syn.GRUUnit(), 2).to(DEV)
    b = syn.make_batch(6, 10, seed=3, mixed_sources=True)
    n = 60
    aa, am, nm, bos = b.aa_out[:, :n].contiguous().to(DEV), b.actors_mask[:n].to(DEV), b.nus_mask[:n].to(DEV), b.bos_mask.to(DEV)
    mean, std = enc.encoder_recurrence_ood(sde, gru, aa, am, nm, bos, eval_iter=10, seed=5)
    mean2, std2 = enc.encoder_recurrence_ood(sde, gru, aa, am, nm, bos, eval_iter=10, seed=5)
    assert mean.shape == (n, 64) and std.shape == (n,)
    assert torch.equal(mean, mean2) and torch.equal(std, std2)
    assert (std > 0).all() and torch.isfinite(mean).all()


def test_rows_major_output_layout_matches_and_is_unit_stride():
    from helpers import DecoderSDE
    sde = init_like_reference(DecoderSDE(), seed=4).to(DEV)
    ts = torch.linspace(0, 6, 61)
    y0 = torch.relu(torch.randn(300, 64, generator=torch.Generator().manual_seed(4))).to(DEV)
    for mode in ('exact', 'tc_f16'):
        a = tb.sdeint(sde, y0, ts, dt=0.1, method='euler', mode=mode, seed=9)
        b_ = tb.sdeint(sde, y0, ts, dt=0.1, method='euler', mode=mode, seed=9, rows_major=True)
        assert b_.shape == a.shape and torch.equal(a, b_)
        sol_y = b_[1:].permute(1, 0, 2)                     # what SDEDecoder.forward hands to its heads (dec…sde.py:88)
        assert sol_y.stride() == (61 * 64, 64, 1)


@pytest.mark.parametrize('rows,mixed', [(90, True), (200, False)])
def test_fused_encoder_gradients_vs_fp64_autograd(rows, mixed):
    """trajsde_enc_bwd (reverse sweep: fp32 GRU backward + tensor-core one-step SDE backward per diffusion net) against fp64
    autograd through the oracle's recurrence: gradients of h0, aa_out, every SDE weight (f, g_nus, g_argo) and every GRU
    weight, with cotangents on the latents AND on the per-iteration diffusion g (DiffBCE, losses/diff_BCE.py:11-16)."""
    sde = init_like_reference(EncoderSDE(), seed=rows, bias_std=0.2).to(DEV)
    gru = syn.init_reference_style(syn.GRUUnit(), rows + 1, bias_std=0.2).to(DEV)
    g = torch.Generator().manual_seed(rows)
    h0 = torch.randn(rows, 64, generator=g) * 0.3
    aa = torch.randn(21, rows, 64, generator=g)
    am = torch.rand(rows, 21, generator=g) > 0.3
    nm = (torch.rand(rows, generator=g) > 0.5) if mixed else torch.ones(rows, dtype=torch.bool)
    dW = torch.randn(21, rows, 64, generator=g) * 0.3
    cot = torch.randn(21, rows, 64, generator=g)
    cot_g = torch.randn(21, rows, generator=g)

    nets = [net_params(sde.f_func), net_params(sde.g_nus), net_params(sde.g_argo), {k: v.detach().cpu() for k, v in gru.state_dict().items()}]
    P = [{k: v.double().clone().requires_grad_(True) for k, v in n.items()} for n in nets]
    h0d, aad = h0.double().requires_grad_(True), aa.double().requires_grad_(True)
    lat_r, g_r = so.encoder_recurrence_ref(P[0], P[1], P[2], P[3], h0d, aad, am, nm, dW.double())
    loss = (lat_r * cot.double()).sum() + (g_r[:, :, 0] * cot_g.double()).sum()
    leaves = [h0d, aad] + [t for n in P for t in n.values()]
    ref = torch.autograd.grad(loss, leaves, allow_unused=True)
    names = ['h0', 'aa_out'] + [f'net{i}.{k}' for i, n in enumerate(P) for k in n]

    h = h0.to(DEV).requires_grad_(True)
    a = aa.to(DEV).requires_grad_(True)
    lat, gg = enc.encoder_recurrence(sde, gru, h, a, am.to(DEV), nm.to(DEV), dW=dW.to(DEV), mode='tc_f16', fused=True)
    ((lat * cot.to(DEV)).sum() + (gg * cot_g.to(DEV)).sum()).backward()
    got = [h.grad, a.grad] + [p_.grad for net in (sde.f_func, sde.g_nus, sde.g_argo) for _, p_ in net.net.named_parameters()] + \
          [gru.get_parameter(k).grad for k in nets[3]]
    assert len(got) == len(ref)
    for n, x, r in zip(names, got, ref):
        if r is None or float(r.abs().max()) == 0.0:        # g_argo unused when every row is nuScenes
            assert x is None or float(x.abs().max()) == 0.0, n
            continue
        e = float((x.double().cpu() - r).abs().max() / r.abs().max())
        print(f"rows={rows} {n}: rel err {e:.2e}")
        assert e < 3e-2, (n, e)


@pytest.mark.parametrize('rows,mixed,philox', [(90, True, False), (700, True, True), (6500, True, False), (8000, False, True), (21504, True, True)])
def test_single_launch_sweep_matches_per_step_launches(rows, mixed, philox):
    """trajsde_enc_bwd's default form (ONE persistent launch, CTAs in GRU / SDE-step roles handing tiles over through progress counters,
    enc_bwd_sweep.cu) against the per-iteration form (TRAJSDE_BWD_FLAG_PER_STEP_LAUNCHES): the per-row results (dL/dh0, dL/daa_out) are
    the same arithmetic on the same tiles -> bit-identical; the weight gradients differ only in how the per-CTA partial sums are grouped.
    Row counts cover one tile, fewer tiles than CTAs per role, and several tiles per CTA (the pipelined case); run twice for the
    stale-counter / reuse-of-workspace case."""
    from trajsde_b200 import ops
    sde = init_like_reference(EncoderSDE(), seed=rows, bias_std=0.2).to(DEV)
    gru = syn.init_reference_style(syn.GRUUnit(), rows + 1, bias_std=0.2).to(DEV)
    g = torch.Generator().manual_seed(rows)
    h0 = (torch.randn(rows, 64, generator=g) * 0.3).to(DEV)
    aa = torch.randn(21, rows, 64, generator=g).to(DEV)
    am = (torch.rand(rows, 21, generator=g) > 0.3).to(DEV)
    nm = ((torch.rand(rows, generator=g) > 0.5) if mixed else torch.ones(rows, dtype=torch.bool)).to(DEV)
    dW = None if philox else (torch.randn(21, rows, 64, generator=g) * 0.3).to(DEV)
    cot = torch.randn(21, rows, 64, generator=g).to(DEV)
    cot_g = torch.randn(21, rows, generator=g).to(DEV)
    params = [p_ for net in (sde.f_func, sde.g_nus, sde.g_argo) for _, p_ in net.net.named_parameters()] + list(gru.parameters())

    def run(per_step):
        ops.ENC_BWD_PER_STEP = per_step
        try:
            h, a = h0.clone().requires_grad_(True), aa.clone().requires_grad_(True)
            for p_ in params:
                p_.grad = None
            lat, gg = enc.encoder_recurrence(sde, gru, h, a, am, nm, dW=dW, mode='tc_f16', fused=True, seed=5)
            ((lat * cot).sum() + (gg * cot_g).sum()).backward()
            torch.cuda.synchronize()
            return [h.grad, a.grad] + [None if p_.grad is None else p_.grad.clone() for p_ in params]
        finally:
            ops.ENC_BWD_PER_STEP = False

    ref = run(True)
    for rep in range(2):
        got = run(False)
        assert ops.backward_status(DEV) == 0
        assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]), rep
        for i, (x, r) in enumerate(zip(got[2:], ref[2:])):
            if r is None or float(r.abs().max()) == 0.0:
                assert x is None or float(x.abs().max()) == 0.0, i
                continue
            e = float((x - r).abs().max() / r.abs().max())
            assert e < 2e-5, (i, e)


@pytest.mark.parametrize('rows', [200, 9000])
def test_single_diffusion_recurrence_equals_dual_with_one_source(rows):
    """The C ABI also serves a recurrence with ONE diffusion net (alt_mask == NULL: two roles / one SDE pass in the single-launch backward,
    the <*, false> forward variants).  With every row on the first net the dual call must give the same latents, g and gradients of h0,
    aa_out, f, g and the GRU — in both forms of the backward sweep."""
    from trajsde_b200 import ops
    from trajsde_b200.encoder import _enc_tables, _gru_params
    from trajsde_b200.solver import _mlp_params
    sde = init_like_reference(EncoderSDE(), seed=rows, bias_std=0.2).to(DEV)
    gru = syn.init_reference_style(syn.GRUUnit(), rows + 1, bias_std=0.2).to(DEV)
    g = torch.Generator().manual_seed(rows)
    h0 = (torch.randn(rows, 64, generator=g) * 0.3).to(DEV)
    aa = torch.randn(21, rows, 64, generator=g).to(DEV)
    am = (torch.rand(rows, 21, generator=g) > 0.3).to(DEV)
    cot = torch.randn(21, rows, 64, generator=g).to(DEV)
    cot_g = torch.randn(21, rows, generator=g).to(DEV)
    step_tab, slots = _enc_tables(2.0, 21, 0.1, torch.device(DEV))
    p_f, p_g, p_a = _mlp_params(sde.f_func, 64, 'f_func'), _mlp_params(sde.g_nus, 1, 'g_nus'), _mlp_params(sde.g_argo, 1, 'g_argo')
    gp = _gru_params(gru)
    all_params = list(p_f) + list(p_g) + list(p_a) + list(gp)

    def run(dual, per_step):
        ops.ENC_BWD_PER_STEP = per_step
        try:
            for p_ in all_params:
                p_.grad = None
            h, a = h0.clone().requires_grad_(True), aa.clone().requires_grad_(True)
            params = list(p_f) + list(p_g) + (list(p_a) if dual else [])
            nm = torch.ones(rows, dtype=torch.bool, device=DEV) if dual else None
            lat, gg = ops.enc_call(h, a, am, slots, params, list(gp), step_tab, None, nm, 77, 0, 0, True)
            ((lat * cot).sum() + (gg * cot_g).sum()).backward()
            torch.cuda.synchronize()
            return lat.detach(), gg.detach(), [h.grad, a.grad] + [p_.grad.clone() for p_ in list(p_f) + list(p_g) + list(gp)]
        finally:
            ops.ENC_BWD_PER_STEP = False

    lat_d, g_d, gr_d = run(True, False)
    for per_step in (False, True):
        lat_s, g_s, gr_s = run(False, per_step)
        assert torch.equal(lat_s, lat_d) and torch.equal(g_s, g_d)
        assert ops.backward_status(DEV) == 0
        for i, (x, r) in enumerate(zip(gr_s, gr_d)):
            e = float((x - r).abs().max() / r.abs().max())
            assert e < 2e-5, (per_step, i, e)


@pytest.mark.parametrize('rows', [1, 130, 700])
def test_gru_jump_forward_and_gradients(rows):
    """Stand-alone fused GRU_Unit jump (trajsde_gru_fwd / trajsde_gru_bwd) against the oracle's gru_ref and fp64 autograd."""
    gru = syn.init_reference_style(syn.GRUUnit(), rows, bias_std=0.2).to(DEV)
    g = torch.Generator().manual_seed(rows)
    h, x = torch.randn(rows, 64, generator=g), torch.randn(rows, 64, generator=g)
    m = torch.rand(rows, generator=g) > 0.3
    cot = torch.randn(rows, 64, generator=g)
    P = {k: v.detach().cpu().double().requires_grad_(True) for k, v in gru.state_dict().items()}
    hd, xd = h.double().requires_grad_(True), x.double().requires_grad_(True)
    ref = so.gru_ref(P, hd, xd, m)
    gref = torch.autograd.grad((ref * cot.double()).sum(), [hd, xd] + list(P.values()))

    hg, xg = h.to(DEV).requires_grad_(True), x.to(DEV).requires_grad_(True)
    out = enc.gru_jump(gru, hg, xg, m.to(DEV))
    assert (out.detach().cpu().double() - ref.detach()).abs().max() < 1e-2
    assert torch.equal(out.detach()[~m.to(DEV)], hg.detach()[~m.to(DEV)])            # unobserved rows pass through untouched
    (out * cot.to(DEV)).sum().backward()
    got = [hg.grad, xg.grad] + [gru.get_parameter(k).grad for k in P]
    for n, a_, r in zip(['h_cur', 'x'] + list(P), got, gref):
        e = float((a_.double().cpu() - r).abs().max() / (r.abs().max() + 1e-30))
        assert e < 3e-2, (n, e)


def test_stepwise_encoder_with_installed_gru_matches_fused_recurrence():
    """The drop-in path (reference loop: sdeint_dual + gru_unit per iteration, both rebound by install()) and the fused recurrence
    kernel compute the same latents from the same Brownian increments."""
    from trajsde_b200 import patch
    rows = 150
    sde = init_like_reference(EncoderSDE(), seed=3, bias_std=0.2).to(DEV)
    gru = syn.init_reference_style(syn.GRUUnit(), 4, bias_std=0.2).to(DEV)
    g = torch.Generator().manual_seed(9)
    h0 = (torch.randn(rows, 64, generator=g) * 0.1).to(DEV)
    aa = torch.randn(21, rows, 64, generator=g).to(DEV)
    am = (torch.rand(rows, 21, generator=g) > 0.3).to(DEV)
    nm = (torch.rand(rows, generator=g) > 0.5).to(DEV)
    dW = (torch.randn(21, rows, 64, generator=g) * 0.3).to(DEV)
    glob = {'sdeint_dual': None}
    exec("class Stage:\n    def forward(self):\n        return sdeint_dual\n", glob)
    enc_stage = glob['Stage']()
    enc_stage.gru_unit = gru
    saved = patch.install(encoder=enc_stage)
    try:
        with torch.no_grad():
            lat_s, _ = enc.encoder_recurrence(sde, gru, h0, aa, am, nm, dW=dW, fused=False)       # loop: sdeint_dual op + patched GRU
            lat_f, _ = enc.encoder_recurrence(sde, gru, h0, aa, am, nm, dW=dW, fused=True)
    finally:
        patch.uninstall(saved)
    assert (lat_s - lat_f).abs().max() < 2e-2
