import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
import trajsde_b200 as tb
from helpers import EncoderSDE, load_net, init_like_reference, net_params
from oracle import sde_oracle as so
DEV='cuda:0'
sde = init_like_reference(EncoderSDE(), seed=3, bias_std=0.3).to(DEV)
rows=40
g=torch.Generator().manual_seed(0)
y0=torch.randn(rows,64,generator=g)
ts=torch.tensor([0.0,0.1])
dW=torch.randn(1,rows,64,generator=g)*0.3
for name,mask in [('all_nus',torch.ones(rows,dtype=torch.bool)),('all_argo',torch.zeros(rows,dtype=torch.bool)),('mixed',torch.rand(rows,generator=g)>0.5)]:
    ref_ys,ref_g=so.euler_solve_ref(net_params(sde.f_func),net_params(sde.g_nus),y0,ts,0.1,dW,mask,net_params(sde.g_argo))
    for mode in ('exact','tc_f16'):
        ys,gg=tb.sdeint_dual(sde,y0.to(DEV),ts,mask.to(DEV),bm=dW.to(DEV),dt=0.1,method='euler',mode=mode)
        eg=(gg[:,0].cpu()-ref_g[:,0]).abs()
        ey=(ys.cpu()-ref_ys).abs().max().item()
        print(name,mode,'g err max',eg.max().item(),'nus rows err',eg[mask].max().item() if mask.any() else None,'argo rows err',eg[~mask].max().item() if (~mask).any() else None,'ys err',ey)
# f check: zero diffusion influence -> compare f via ys with dW=0
dW0=torch.zeros_like(dW)
mask=torch.rand(rows,generator=g)>0.5
ref_ys,_=so.euler_solve_ref(net_params(sde.f_func),net_params(sde.g_nus),y0,ts,0.1,dW0,mask,net_params(sde.g_argo))
ys,_=tb.sdeint_dual(sde,y0.to(DEV),ts,mask.to(DEV),bm=dW0.to(DEV),dt=0.1,method='euler',mode='tc_f16')
print('f-only ys err (x10 = f err)', (ys.cpu()-ref_ys).abs().max().item())
