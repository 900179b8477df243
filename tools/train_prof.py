import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R)
import torch, time
import trajsde_b200 as tb
from trajsde_b200 import synthetic as syn, encoder as enc_mod
dev=torch.device('cuda:0')
scenes=int(sys.argv[1]) if len(sys.argv)>1 else 128
iters=int(sys.argv[2]) if len(sys.argv)>2 else 1
enc_sde=syn.init_reference_style(syn.EncoderSDEFunc(),1).to(dev); dec_sde=syn.init_reference_style(syn.DecoderSDEFunc(),2).to(dev); gru=syn.init_reference_style(syn.GRUUnit(),3).to(dev)
b=syn.make_batch(scenes,20,seed=5,mixed_sources=True)
tr={k:getattr(b,k).to(dev) for k in ('enc_h0','aa_out','actors_mask','nus_mask','dec_y0')}
ts=torch.linspace(0,6,61)
def step(i, what):
    for p_ in list(enc_sde.parameters())+list(dec_sde.parameters())+list(gru.parameters()): p_.grad=None
    if what=='enc':
        aa=tr['aa_out'].detach().requires_grad_(True)
        lat,g=enc_mod.encoder_recurrence(enc_sde,gru,tr['enc_h0'],aa,tr['actors_mask'],tr['nus_mask'],seed=300+i)
        go_l=torch.full_like(lat,1e-6); go_g=torch.full_like(g,1e-6)
        torch.autograd.backward([lat,g],[go_l,go_g])
    else:
        y0=tr['dec_y0'].detach().requires_grad_(True)
        ys=tb.sdeint(dec_sde,y0,ts,dt=0.1,method='euler',seed=400+i)
        ys.backward(torch.full_like(ys,1e-6))
for what in ('enc','dec'):
    for i in range(2): step(i,what)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for i in range(iters): step(i,what)
    torch.cuda.synchronize(); print(what, scenes, 'scenes fwd+bwd:', (time.perf_counter()-t0)/iters*1e3,'ms')
