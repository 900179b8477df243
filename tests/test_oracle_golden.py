"""Pin oracle/sde_oracle.py against outputs of the reference's own code (tests/golden, made by make_golden.py) and, when
/root/reference is present (dev container), against fresh runs of the reference with other seeds."""
import numpy as np
import pytest
import torch

from conftest import sub
from oracle import ref_runner as rr
from oracle import sde_oracle as so


def test_decoder_solve_bit_exact_vs_golden(golden_decoder):
    d = golden_decoder
    ys, g = so.euler_solve_ref(sub(d, 'f'), sub(d, 'g'), torch.from_numpy(d['y0']), torch.from_numpy(d['ts']),
                               float(d['dt']), torch.from_numpy(d['dW']))
    assert ys.shape == (61, 48, 64)
    assert torch.equal(ys, torch.from_numpy(d['ys']))          # bit-for-bit
    assert int(d['fnfe']) == 62                                 # 61 steps + the contract probe
    assert torch.equal(ys[0], torch.from_numpy(d['y0']))


def test_decoder_heads_vs_golden(golden_decoder):
    d = golden_decoder
    loc = so.decoder_loc_head_ref(sub(d, 'head'), torch.from_numpy(d['ys'])[1:].permute(1, 0, 2))
    assert torch.allclose(loc, torch.from_numpy(d['loc']), atol=1e-5, rtol=1e-5)
    scale = so.decoder_scale_ref(sub(d, 'scale_head'), torch.from_numpy(d['ys'])[1:].permute(1, 0, 2), float(d['min_scale']))
    assert torch.allclose(scale, torch.from_numpy(d['scale']), atol=1e-5, rtol=1e-5)


def test_encoder_loop_vs_golden(golden_encoder):
    e = golden_encoder
    lat, gs = so.encoder_recurrence_ref(sub(e, 'f'), sub(e, 'g_nus'), sub(e, 'g_argo'), sub(e, 'gru'),
                                        torch.from_numpy(e['h0']), torch.from_numpy(e['aa_out']),
                                        torch.from_numpy(e['actors_mask']), torch.from_numpy(e['nus_mask']),
                                        torch.from_numpy(e['dW']))
    # the reference gathers rows per source before its GEMMs (enc…sep2.py:478-479) => different GEMM blocking, few ulp
    assert (lat - torch.from_numpy(e['latent_ys'])).abs().max() < 2e-6
    assert (gs.repeat(1, 1, 64) - torch.from_numpy(e['g'])).abs().max() < 5e-7


def test_fp64_oracle_close_to_fp32(golden_decoder):
    d = golden_decoder
    pf = {k: v.double() for k, v in sub(d, 'f').items()}
    pg = {k: v.double() for k, v in sub(d, 'g').items()}
    ys64, _ = so.euler_solve_ref(pf, pg, torch.from_numpy(d['y0']).double(), torch.from_numpy(d['ts']), float(d['dt']),
                                 torch.from_numpy(d['dW']).double())
    assert (ys64.float() - torch.from_numpy(d['ys'])).abs().max() < 2e-4


@pytest.mark.skipif(not rr.reference_available(), reason="reference tree absent (GPU box)")
@pytest.mark.parametrize('seed', [7, 11])
def test_oracle_vs_live_reference_decoder(seed):
    dec = rr.build_reference_decoder(seed=seed, bias_std=0.2)
    g = torch.Generator().manual_seed(seed)
    y0 = torch.relu(torch.randn(33, 64, generator=g))
    sched = so.euler_schedule_ref(dec.ts_pred, dec.min_stepsize)
    dW = torch.randn(61, 33, 64, generator=g) * torch.sqrt(sched['h']).view(-1, 1, 1)
    ys_ref, queries = rr.run_reference_decoder_solve(dec, y0, dW)
    ys, _ = so.euler_solve_ref(rr.net_params(dec.lsde_func.f_func.net), rr.net_params(dec.lsde_func.g_func.net), y0,
                               dec.ts_pred, dec.min_stepsize, dW)
    assert torch.equal(ys, ys_ref)
    assert len(queries) == 61


@pytest.mark.skipif(not rr.reference_available(), reason="reference tree absent (GPU box)")
def test_oracle_vs_live_reference_encoder():
    lsde, gru = rr.build_reference_encoder_sde(seed=5, bias_std=0.2)
    g = torch.Generator().manual_seed(5)
    rows = 19
    h0 = (torch.randn(64, generator=g) * 0.02).repeat(rows, 1)
    aa = torch.randn(21, rows, 64, generator=g)
    am = torch.rand(rows, 21, generator=g) > 0.3
    nm = torch.rand(rows, generator=g) > 0.5
    dW = torch.randn(21, rows, 64, generator=g) * 0.3
    lat_ref, g_ref, _ = rr.run_reference_encoder_loop(lsde, gru, h0, aa, am, nm, dW)
    lat, gs = so.encoder_recurrence_ref(rr.net_params(lsde.f_func.net), rr.net_params(lsde.g_nus.net),
                                        rr.net_params(lsde.g_argo.net), rr.net_params(gru), h0, aa, am, nm, dW)
    assert (lat - lat_ref).abs().max() < 2e-6
    assert (gs.repeat(1, 1, 64) - g_ref).abs().max() < 5e-7


# ---- decoder stage, losses, metrics (SURVEY §8(f)-4, row g): oracle vs the reference's own SDEDecoder.forward / L2 / DiffBCE / ADE_T / FDE_T ----
def _stage_inputs(d, dtype=torch.float32):
    p = {k[len('param/'):]: torch.from_numpy(d[k]).to(dtype) for k in d if k.startswith('param/')}
    le = torch.from_numpy(d['local_embed']).to(dtype)
    ge = torch.from_numpy(d['global_embed']).to(dtype)
    return p, le, ge, torch.from_numpy(d['padding_mask']), torch.from_numpy(d['dW'])


def test_decoder_stage_forward_vs_golden(golden_stage):
    d = golden_stage
    p, le, ge, pad, dW = _stage_inputs(d)
    out = so.decoder_forward_ref(p, le, ge, pad, dW, min_scale=float(d['min_scale']))
    assert torch.allclose(out['hidden_0'], torch.from_numpy(d['hidden_0']), atol=1e-6, rtol=1e-6)      # aggr_embed
    assert torch.allclose(out['ys'], torch.from_numpy(d['ys']), atol=2e-5, rtol=1e-5)                  # solve from the oracle's own hidden_0
    assert torch.allclose(out['loc'], torch.from_numpy(d['loc']), atol=2e-5, rtol=1e-5)                # [10,12,60,4] = cat(loc, scale)
    assert torch.allclose(out['pi'], torch.from_numpy(d['pi']), atol=1e-5, rtol=1e-5)
    assert torch.equal(out['reg_mask'], torch.from_numpy(d['reg_mask']))


def test_losses_and_metrics_vs_golden(golden_stage):
    d = golden_stage
    loc, y, rm = torch.from_numpy(d['loc']), torch.from_numpy(d['y']), torch.from_numpy(d['reg_mask'])
    assert abs(float(so.l2_loss_ref(loc, y, rm)) - float(d['loss_l2'])) < 1e-6
    assert abs(float(so.diff_bce_ref(torch.from_numpy(d['diff_in']), torch.from_numpy(d['diff_out']))) - float(d['loss_bce'])) < 1e-6
    assert abs(so.ade_t_ref(loc[..., :2], y, rm) - float(d['ade'])) < 1e-5
    assert abs(so.fde_t_ref(loc[..., :2], y, rm, torch.from_numpy(d['source'])) - float(d['fde'])) < 1e-5
    ade2, _ = so.min_ade_fde_ref(loc[..., :2], y, rm)          # the older helper agrees with the ADE_T restatement
    assert abs(ade2 - float(d['ade'])) < 1e-5


def test_training_gradients_through_the_stage_vs_golden(golden_stage):
    """fp64 autograd of the oracle chain (aggr_embed -> solve -> heads -> L2, + DiffBCE) reproduces the gradients the reference's
    own forward/backward produced in fp32 — this is the oracle the GPU training-step tests lean on."""
    d = golden_stage
    p, le, ge, pad, dW = _stage_inputs(d, torch.float64)
    for v in p.values():
        v.requires_grad_(True)
    le.requires_grad_(True); ge.requires_grad_(True)
    di = torch.from_numpy(d['diff_in']).double().requires_grad_(True)
    do = torch.from_numpy(d['diff_out']).double().requires_grad_(True)
    out = so.decoder_forward_ref(p, le, ge, pad, dW.double(), min_scale=float(d['min_scale']))
    loss = so.l2_loss_ref(out['loc'], torch.from_numpy(d['y']).double(), out['reg_mask']) + so.diff_bce_ref(di, do)
    loss.backward()

    def close(a, b, name):
        b = torch.from_numpy(b).double()
        assert (a - b).abs().max() <= 2e-4 * b.abs().max() + 1e-9, name

    close(le.grad, d['grad_local_embed'], 'local_embed')
    close(ge.grad, d['grad_global_embed'], 'global_embed')
    close(di.grad, d['grad_diff_in'], 'diff_in')
    close(do.grad, d['grad_diff_out'], 'diff_out')
    for k in ('aggr_embed.0.weight', 'aggr_embed.1.bias', 'lsde_func.f_func.net.0.weight', 'lsde_func.f_func.net.4.bias',
              'lsde_func.g_func.net.0.weight', 'lsde_func.g_func.net.4.weight', 'decoder.0.weight', 'decoder.1.weight', 'decoder.3.weight',
              'decoder.3.bias'):
        close(p[k].grad, d['grad/' + k], k)
    # the scale head and pi do not reach L2 (loc only, losses/L2.py:12,15): their reference gradients are exactly zero
    assert not d['grad/scale.0.weight'].any() and not d['grad/pi.0.weight'].any()


# ---- encoder stage (enc…sep2.py:66-202 run verbatim through the functional PyG stand-in): what the SDE recurrence receives and hands on ----
def _enc_stage_params(d, dtype=torch.float32):
    p = {k[len('param/'):]: torch.from_numpy(d[k]).to(dtype) for k in d if k.startswith('param/')}
    sub_ = lambda pre: {k[len(pre) + 1:]: v for k, v in p.items() if k.startswith(pre + '.')}  # noqa: E731
    return p, sub_('lsde_func.f_func.net'), sub_('lsde_func.g_nus.net'), sub_('lsde_func.g_argo.net'), sub_('gru_unit')


def encoder_stage_ref(d, dtype=torch.float32, leaves=False):
    """Oracle restatement of the SDE part of LocalEncoderSDESepPara2.forward on the fixture's captured inputs: recurrence, eos gather
    (:184-188), agents' diffusion read-out (:171, :190-194)."""
    p, pf, pgn, pga, pgru = _enc_stage_params(d, dtype)
    aa = torch.from_numpy(d['aa_out']).to(dtype)
    if leaves:
        for t in [aa] + list(p.values()):
            t.requires_grad_(True)
    rows, n = aa.shape[1], d['bos_mask'].shape[0]
    h0 = p['hidden'].unsqueeze(0).repeat(rows, 1)
    lat, g = so.encoder_recurrence_ref(pf, pgn, pga, pgru, h0, aa, torch.from_numpy(d['actors_mask']), torch.from_numpy(d['nus_mask']),
                                       torch.from_numpy(d['dW']).to(dtype))
    bos, ai = torch.from_numpy(d['bos_mask']), torch.from_numpy(d['agent_index'])
    eos = 20 - torch.argmax(bos.float(), dim=1)
    pre_al = lat[:, :n][eos, torch.arange(n)]
    new_ai = torch.cat((ai, torch.arange(n, rows)))
    diff = g[eos[ai].repeat(2), new_ai, 0]
    return pre_al, diff, aa, p


def test_encoder_stage_recurrence_vs_golden(golden_encoder_stage):
    d = golden_encoder_stage
    pre_al, diff, _, _ = encoder_stage_ref(d)
    assert torch.allclose(pre_al, torch.from_numpy(d['pre_al']), atol=3e-6, rtol=1e-5)
    ref_diff = torch.cat((torch.from_numpy(d['diff_in']), torch.from_numpy(d['diff_out'])))
    assert torch.allclose(diff.unsqueeze(-1).expand(-1, 64), ref_diff, atol=1e-6, rtol=1e-5)
    assert float(d['label_in'].max()) == 0 and float(d['label_out'].min()) == 1


def test_encoder_stage_gradients_vs_golden(golden_encoder_stage):
    """fp64 autograd of the oracle recurrence under the cotangents the reference's own backward delivered to it (dL/d pre_al from the
    AL encoder + out.square().mean(), DiffBCE on the agents' diffusion) reproduces the reference's gradients of aa_out, hidden, GRU, SDE."""
    d = golden_encoder_stage
    pre_al, diff, aa, p = encoder_stage_ref(d, torch.float64, leaves=True)
    d_in, d_out = torch.chunk(diff, 2, 0)
    loss = (pre_al * torch.from_numpy(d['grad_pre_al']).double()).sum() + so.diff_bce_ref(d_in.unsqueeze(-1).expand(-1, 64), d_out.unsqueeze(-1).expand(-1, 64))
    loss.backward()
    ga = torch.from_numpy(d['grad_aa_out']).double()
    assert (aa.grad - ga).abs().max() <= 2e-4 * ga.abs().max()
    for k in d:
        if k.startswith('grad/'):
            r = torch.from_numpy(d[k]).double()
            assert (p[k[5:]].grad - r).abs().max() <= 3e-4 * r.abs().max() + 1e-10, k
