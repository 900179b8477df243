import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R)
import torch
from trajsde_b200 import synthetic as syn, encoder as enc_mod
dev=torch.device('cuda:0')
scenes=int(sys.argv[1]) if len(sys.argv)>1 else 1024
enc_sde=syn.init_reference_style(syn.EncoderSDEFunc(),1).to(dev); gru=syn.init_reference_style(syn.GRUUnit(),3).to(dev)
b=syn.make_batch(scenes,20,seed=5,mixed_sources=True)
tr={k:getattr(b,k).to(dev) for k in ('enc_h0','aa_out','actors_mask','nus_mask')}
aa=tr['aa_out'].detach().requires_grad_(True)
lat,g=enc_mod.encoder_recurrence(enc_sde,gru,tr['enc_h0'],aa,tr['actors_mask'],tr['nus_mask'],seed=300)
torch.autograd.backward([lat,g],[torch.full_like(lat,1e-6),torch.full_like(g,1e-6)])
torch.cuda.synchronize(); print('ok')
