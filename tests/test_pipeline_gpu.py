"""Host-fed execution (trajsde_b200.pipeline): slicing the batch for copy/compute overlap must not change any result."""
import pytest
import torch

import trajsde_b200 as tb
from trajsde_b200 import encoder as enc
from trajsde_b200 import synthetic as syn
from trajsde_b200.pipeline import HostFedSdePath

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


def test_run_batch_equals_direct_calls():
    enc_sde = syn.init_reference_style(syn.EncoderSDEFunc(), 1).to(DEV)
    dec_sde = syn.init_reference_style(syn.DecoderSDEFunc(), 2).to(DEV)
    gru = syn.init_reference_style(syn.GRUUnit(), 3).to(DEV)
    hb = syn.make_batch(9, 13, seed=4, mixed_sources=True, pin=True)            # 1170 decoder rows: ragged slices and tiles
    ts = torch.linspace(0, 6, 61)
    out_enc = torch.empty((hb.enc_rows, 64)).pin_memory()
    out_dec = torch.empty((hb.dec_rows, 64)).pin_memory()
    pipe = HostFedSdePath(enc_sde, gru, dec_sde, DEV, ts)
    for chunks in (1, 3, 4):
        out_enc.zero_(); out_dec.zero_()
        pipe.run_batch(hb, out_enc, out_dec, seed=50, dec_chunks=chunks)
        with torch.no_grad():
            lat, _ = enc.encoder_recurrence(enc_sde, gru, hb.enc_h0.to(DEV), hb.aa_out.to(DEV), hb.actors_mask.to(DEV),
                                            hb.nus_mask.to(DEV), seed=50)
            ys = tb.sdeint(dec_sde, hb.dec_y0.to(DEV), ts, dt=0.1, method='euler', seed=51)
        assert torch.equal(out_enc, lat[-1].cpu())
        # slices start on other tile boundaries than the unsliced solve: identical Philox increments, same arithmetic per row
        assert torch.allclose(out_dec, ys[-1].cpu(), atol=1e-4, rtol=0), float((out_dec - ys[-1].cpu()).abs().max())


def test_run_batch_returns_the_decoder_outputs_through_the_fused_heads():
    """e2e contract: what comes back to the host is the decoder's out['loc'] = cat(loc, elu(scale) + 1 + min_scale) [M, 60, 4]
    (dec…sde.py:95-100), computed on the device by the fused heads; aa_out may arrive as fp16."""
    import torch.nn as nn
    enc_sde = syn.init_reference_style(syn.EncoderSDEFunc(), 1).to(DEV)
    dec_sde = syn.init_reference_style(syn.DecoderSDEFunc(), 2).to(DEV)
    gru = syn.init_reference_style(syn.GRUUnit(), 3).to(DEV)
    mk = lambda sd: syn.init_reference_style(nn.Sequential(nn.Linear(64, 64), nn.LayerNorm(64), nn.ReLU(inplace=True), nn.Linear(64, 2)), sd).to(DEV)  # noqa: E731
    loc_h, sc_h = mk(7), mk(8)
    hb = syn.make_batch(9, 13, seed=4, mixed_sources=True, pin=True)
    aa_half = hb.aa_out.half().pin_memory()
    ts = torch.linspace(0, 6, 61)
    out_enc = torch.empty((hb.enc_rows, 64)).pin_memory()
    out_dec = torch.empty((hb.dec_rows, 60, 4)).pin_memory()
    pipe = HostFedSdePath(enc_sde, gru, dec_sde, DEV, ts)
    pipe.run_batch(hb, out_enc, out_dec, seed=50, dec_chunks=3, heads=(loc_h, sc_h), min_scale=0.001, aa_out_half=aa_half)
    with torch.no_grad():
        ys = tb.sdeint(dec_sde, hb.dec_y0.to(DEV), ts, dt=0.1, method='euler', seed=51)
        sol = ys[1:].permute(1, 0, 2)
        want = torch.cat((loc_h(sol), torch.nn.functional.elu(sc_h(sol)) + 1.0 + 0.001), dim=-1)
        lat, _ = enc.encoder_recurrence(enc_sde, gru, hb.enc_h0.to(DEV), aa_half.to(DEV).float(), hb.actors_mask.to(DEV), hb.nus_mask.to(DEV), seed=50)
    assert torch.allclose(out_dec, want.cpu(), atol=2e-2, rtol=2e-2)
    assert torch.equal(out_enc, lat[-1].cpu())
