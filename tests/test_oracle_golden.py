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
