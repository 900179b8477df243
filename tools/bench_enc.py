"""Times the fused recurrent encoder forward (enc_fwd_tc_kernel) at configs[1] size: python tools/bench_enc.py [scenes] [iters]"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
from trajsde_b200 import synthetic as syn, encoder as enc_mod
dev = torch.device('cuda:0')
scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
enc_sde = syn.init_reference_style(syn.EncoderSDEFunc(), 1).to(dev); gru = syn.init_reference_style(syn.GRUUnit(), 3).to(dev)
for mixed in (False, True):
    b = syn.make_batch(scenes, 20, seed=5, mixed_sources=mixed)
    tr = {k: getattr(b, k).to(dev) for k in ('enc_h0', 'aa_out', 'actors_mask', 'nus_mask')}
    rows = tr['enc_h0'].shape[0]
    dW = torch.randn(21, rows, 64, device=dev) * 0.3
    for name, kw in (('dw', dict(dW=dW)), ('philox', dict(seed=300))):
        with torch.no_grad():
            for _ in range(3): enc_mod.encoder_recurrence(enc_sde, gru, tr['enc_h0'], tr['aa_out'], tr['actors_mask'], tr['nus_mask'], **kw)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters): enc_mod.encoder_recurrence(enc_sde, gru, tr['enc_h0'], tr['aa_out'], tr['actors_mask'], tr['nus_mask'], **kw)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"enc mixed={mixed} {name}: rows={rows} {ms:.3f} ms  {rows*21/ms*1e-6:.3f} G row-steps/s")
