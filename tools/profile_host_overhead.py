"""cProfile of the host side of the installed drop-in encoder loop (per-call Python overhead of the custom ops)."""
import sys, os, cProfile, pstats
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
from trajsde_b200 import synthetic as syn, encoder as enc_mod, patch
dev = torch.device('cuda:0')
enc_sde = syn.init_reference_style(syn.EncoderSDEFunc(), 1).to(dev); gru = syn.init_reference_style(syn.GRUUnit(), 3).to(dev)
b = syn.make_batch(128, 20, seed=5, mixed_sources=True)
tr = {k: getattr(b, k).to(dev) for k in ('enc_h0', 'aa_out', 'actors_mask', 'nus_mask')}
glob = {'sdeint_dual': None}
exec("class Stage:\n    def forward(self):\n        return sdeint_dual\n", glob)
stage = glob['Stage'](); stage.gru_unit = gru
patch.install(encoder=stage)
def step(i):
    aa = tr['aa_out'].detach().requires_grad_(True)
    lat, g = enc_mod.encoder_recurrence(enc_sde, gru, tr['enc_h0'], aa, tr['actors_mask'], tr['nus_mask'], seed=300 + i, fused=False)
    torch.autograd.backward([lat, g], [torch.full_like(lat, 1e-6), torch.full_like(g, 1e-6)])
for i in range(3): step(i)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for i in range(5): step(i)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
