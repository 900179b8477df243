"""Manual gradient check of the fused encoder backward against fp64 autograd of the oracle (dev script; not collected by pytest)."""
import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import torch, time
from helpers import EncoderSDE, init_like_reference, net_params
from oracle import sde_oracle as so
from trajsde_b200 import encoder as enc, synthetic as syn
DEV='cuda:0'
rows, mixed = 90, True
sde = init_like_reference(EncoderSDE(), seed=rows, bias_std=0.2).to(DEV)
gru = syn.init_reference_style(syn.GRUUnit(), rows + 1, bias_std=0.2).to(DEV)
g = torch.Generator().manual_seed(rows)
h0 = torch.randn(rows, 64, generator=g) * 0.3
aa = torch.randn(21, rows, 64, generator=g)
am = torch.rand(rows, 21, generator=g) > 0.3
nm = (torch.rand(rows, generator=g) > 0.5)
dW = torch.randn(21, rows, 64, generator=g) * 0.3
cot = torch.randn(21, rows, 64, generator=g); cot_g = torch.randn(21, rows, generator=g)
nets = [net_params(sde.f_func), net_params(sde.g_nus), net_params(sde.g_argo), {k: v.detach().cpu() for k, v in gru.state_dict().items()}]
P = [{k: v.double().clone().requires_grad_(True) for k, v in n.items()} for n in nets]
h0d, aad = h0.double().requires_grad_(True), aa.double().requires_grad_(True)
lat_r, g_r = so.encoder_recurrence_ref(P[0], P[1], P[2], P[3], h0d, aad, am, nm, dW.double())
loss = (lat_r * cot.double()).sum() + (g_r[:, :, 0] * cot_g.double()).sum()
leaves = [h0d, aad] + [t for n in P for t in n.values()]
ref = torch.autograd.grad(loss, leaves, allow_unused=True)
names = ['h0', 'aa_out'] + [f'net{i}.{k}' for i, n in enumerate(P) for k in n]
for fused in (True, False):
    for p_ in list(sde.parameters())+list(gru.parameters()): p_.grad=None
    h = h0.to(DEV).requires_grad_(True); a = aa.to(DEV).requires_grad_(True)
    lat, gg = enc.encoder_recurrence(sde, gru, h, a, am.to(DEV), nm.to(DEV), dW=dW.to(DEV), mode='tc_f16', fused=fused)
    ((lat * cot.to(DEV)).sum() + (gg * cot_g.to(DEV)).sum()).backward(); torch.cuda.synchronize()
    got = [h.grad, a.grad] + [p_.grad for net in (sde.f_func, sde.g_nus, sde.g_argo) for _, p_ in net.net.named_parameters()] + [gru.get_parameter(k).grad for k in nets[3]]
    print('--- fused' if fused else '--- stepwise')
    for n, x, r in zip(names, got, ref):
        if r is None: continue
        e = float((x.double().cpu() - r).abs().max() / (r.abs().max()+1e-30))
        print(f"  {n:32s} rel err {e:.2e}   |ref| {float(r.abs().max()):.3e}")
# timing
for scenes in (128, 1024):
    b = syn.make_batch(scenes, 20, seed=5, mixed_sources=True)
    tr = {k: getattr(b, k).to(DEV) for k in ('enc_h0', 'aa_out', 'actors_mask', 'nus_mask')}
    def step(i):
        for p_ in list(sde.parameters())+list(gru.parameters()): p_.grad=None
        aa_ = tr['aa_out'].detach().requires_grad_(True)
        lat, gg = enc.encoder_recurrence(sde, gru, tr['enc_h0'], aa_, tr['actors_mask'], tr['nus_mask'], seed=300+i, fused=True)
        (lat.square().mean() + gg.mean()).backward()
    for i in range(2): step(i)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for i in range(5): step(i)
    torch.cuda.synchronize(); print(f"enc fwd+bwd fused {scenes} scenes: {(time.perf_counter()-t0)/5*1e3:.3f} ms")
