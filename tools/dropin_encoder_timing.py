"""Encoder training step (fwd+bwd) through the three execution paths: reference-style loop with torch GRU, the same loop after
install() (fused sdeint_dual + fused GRU jump per iteration), and the fused recurrence (one forward launch, one backward call)."""
import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
from trajsde_b200 import synthetic as syn, encoder as enc_mod, patch
dev = torch.device('cuda:0')
scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 128
enc_sde = syn.init_reference_style(syn.EncoderSDEFunc(), 1).to(dev); gru = syn.init_reference_style(syn.GRUUnit(), 3).to(dev)
b = syn.make_batch(scenes, 20, seed=5, mixed_sources=True)
tr = {k: getattr(b, k).to(dev) for k in ('enc_h0', 'aa_out', 'actors_mask', 'nus_mask')}
glob = {'sdeint_dual': None}
exec("class Stage:\n    def forward(self):\n        return sdeint_dual\n", glob)
stage = glob['Stage'](); stage.gru_unit = gru


def step(i, fused):
    for p_ in list(enc_sde.parameters()) + list(gru.parameters()):
        p_.grad = None
    aa = tr['aa_out'].detach().requires_grad_(True)
    lat, g = enc_mod.encoder_recurrence(enc_sde, gru, tr['enc_h0'], aa, tr['actors_mask'], tr['nus_mask'], seed=300 + i, fused=fused)
    torch.autograd.backward([lat, g], [torch.full_like(lat, 1e-6), torch.full_like(g, 1e-6)])


def timeit(label, fused):
    for i in range(2):
        step(i, fused)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(5):
        step(i, fused)
    torch.cuda.synchronize(); print(f"{label:58s} {scenes} scenes: {(time.perf_counter() - t0) / 5 * 1e3:7.2f} ms fwd+bwd")


timeit("loop: fused sdeint_dual op + torch GRU_Unit", False)
saved = patch.install(encoder=stage)
timeit("loop after install(): fused sdeint_dual + fused GRU jump", False)
patch.uninstall(saved)
timeit("fused recurrence (enc_fwd_tc_kernel / trajsde_enc_bwd)", True)
