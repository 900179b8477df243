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
