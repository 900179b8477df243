import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
import trajsde_b200 as tb
from trajsde_b200 import ops, _lib
from helpers import DecoderSDE, init_like_reference
DEV='cuda:0'
ts=torch.linspace(0,6,61)
y0=torch.relu(torch.randn(200,64,generator=torch.Generator().manual_seed(1))).to(DEV)
for blow in (1.0, 2.0, 3.0, 4.0, 6.0, 10.0, 40.0):
    sde=init_like_reference(DecoderSDE(),seed=2,bias_std=0.2).to(DEV)
    with torch.no_grad():
        for p_ in sde.f_func.parameters(): p_.mul_(blow)
    ops.backward_status(DEV)
    y=y0.clone().requires_grad_(True)
    ys=tb.sdeint(sde,y,ts,dt=0.1,method='euler',mode='tc_f16',seed=5)
    ys[-1].sum().backward()
    st=ops.backward_status(DEV)
    g_tc=y.grad.clone()
    ops.BWD_EXACT_KERNELS=True
    y2=y0.clone().requires_grad_(True)
    ys2=tb.sdeint(sde,y2,ts,dt=0.1,method='euler',mode='tc_f16',seed=5)
    ys2[-1].sum().backward(); ops.BWD_EXACT_KERNELS=False
    rel=float((g_tc-y2.grad).abs().max()/(y2.grad.abs().max()+1e-30))
    print(f"blow {blow}: status {st} |ys|max {float(ys.abs().max()):.3e} |grad_y0|max exact {float(y2.grad.abs().max()):.3e} tc-vs-exact rel {rel:.2e}")
