import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import torch
import trajsde_b200 as tb
from helpers import DecoderSDE, init_like_reference
DEV='cuda:0'
rows=int(sys.argv[1]) if len(sys.argv)>1 else 204800
sde = init_like_reference(DecoderSDE(), seed=1).to(DEV)
ts=torch.linspace(0,6,61); y0=torch.relu(torch.randn(rows,64,device=DEV))
dW=torch.randn(61,rows,64,device=DEV)*0.3
for bm in (dW, None):
    y=y0.clone().requires_grad_(True)
    ys=tb.sdeint(sde,y,ts,bm=bm,dt=0.1,method='euler',mode='tc_f16',seed=3)
    ys.backward(torch.ones_like(ys)*1e-6); torch.cuda.synchronize()
print('ok')
