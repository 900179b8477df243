"""Encoder fwd+bwd time vs scenes (tiles per SM): separates the per-launch fixed cost of the 42-launch backward chain from tile work."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
from trajsde_b200 import synthetic as syn, encoder as enc_mod
dev = torch.device('cuda:0')
enc_sde = syn.init_reference_style(syn.EncoderSDEFunc(), 1).to(dev); gru = syn.init_reference_style(syn.GRUUnit(), 3).to(dev)
for scenes in (6, 128, 400, 800, 1024, 1400, 2048):
    b = syn.make_batch(scenes, 20, seed=5, mixed_sources=True)
    tr = {k: getattr(b, k).to(dev) for k in ('enc_h0', 'aa_out', 'actors_mask', 'nus_mask')}
    rows = tr['enc_h0'].shape[0]
    go_l = torch.full((21, rows, 64), 1e-6, device=dev); go_g = torch.full((21, rows), 1e-6, device=dev)
    def step(i):
        aa = tr['aa_out'].detach().requires_grad_(True)
        lat, g = enc_mod.encoder_recurrence(enc_sde, gru, tr['enc_h0'], aa, tr['actors_mask'], tr['nus_mask'], seed=300 + i)
        torch.autograd.backward([lat, g], [go_l, go_g])
    def fwd(i):
        with torch.no_grad():
            enc_mod.encoder_recurrence(enc_sde, gru, tr['enc_h0'], tr['aa_out'], tr['actors_mask'], tr['nus_mask'], seed=300 + i)
    res = []
    for fn in (fwd, step):
        for i in range(3): fn(i)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10): fn(i)
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 10)
    tiles = (rows + 127) // 128
    print(f"scenes {scenes:5d} rows {rows:6d} tiles {tiles:4d} ({tiles / 148:.2f}/SM): fwd {res[0]:.3f} ms  fwd+bwd {res[1]:.3f} ms  bwd/launch {(res[1] - res[0]) / 42 * 1e3:.1f} us")
