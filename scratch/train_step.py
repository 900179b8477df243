import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R)
import torch, time
import trajsde_b200 as tb
from trajsde_b200 import synthetic as syn, encoder as enc_mod
from trajsde_b200.dist import FlatGradBucket
dev=torch.device('cuda:0')
scenes=int(sys.argv[1]) if len(sys.argv)>1 else 128
enc_sde=syn.init_reference_style(syn.EncoderSDEFunc(),1).to(dev); dec_sde=syn.init_reference_style(syn.DecoderSDEFunc(),2).to(dev); gru=syn.init_reference_style(syn.GRUUnit(),3).to(dev)
b=syn.make_batch(scenes,20,seed=5,mixed_sources=True)
tr={k:getattr(b,k).to(dev) for k in ('enc_h0','aa_out','actors_mask','nus_mask','dec_y0')}
ts=torch.linspace(0,6,61)
params=list(enc_sde.parameters())+list(dec_sde.parameters())+list(gru.parameters())
bucket=FlatGradBucket(params)
def step(i, what='both'):
    bucket.zero_()
    loss=0
    if what in ('both','enc'):
        lat,g=enc_mod.encoder_recurrence(enc_sde,gru,tr['enc_h0'],tr['aa_out'],tr['actors_mask'],tr['nus_mask'],seed=300+i,fused=False)
        loss=loss+lat.square().mean()+g.mean()
    if what in ('both','dec'):
        y0=tr['dec_y0'].detach().requires_grad_(True)
        ys=tb.sdeint(dec_sde,y0,ts,dt=0.1,method='euler',seed=400+i)
        loss=loss+ys[1:].square().mean()
    loss.backward()
for what in ('enc','dec','both'):
    for i in range(2): step(i,what)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for i in range(3): step(i,what)
    torch.cuda.synchronize(); print(what, scenes, 'scenes:', (time.perf_counter()-t0)/3*1e3,'ms')
