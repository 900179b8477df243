import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import torch, time
import trajsde_b200 as tb
from trajsde_b200 import ops
from helpers import DecoderSDE, init_like_reference, make_dw
from trajsde_b200.schedule import euler_schedule
DEV='cuda:0'
def run_case(F, rows, use_dw, scale):
    sde = init_like_reference(DecoderSDE(), seed=F, bias_std=0.2).to(DEV)
    ts = torch.linspace(0, 0.1*F, F+1); sched = euler_schedule(ts, 0.1)
    g = torch.Generator().manual_seed(F)
    y0 = torch.relu(torch.randn(rows,64,generator=g)).to(DEV)
    dW = make_dw(sched.h, rows, seed=F+1).to(DEV) if use_dw else None
    cot = (torch.randn(F+1,rows,64,generator=g)*scale).to(DEV)
    def run(exact):
        ops.BWD_EXACT_KERNELS = exact
        for p_ in sde.parameters(): p_.grad=None
        y=y0.clone().requires_grad_(True)
        ys=tb.sdeint(sde,y,ts,bm=dW,dt=0.1,method='euler',mode='tc_f16',seed=77)
        (ys*cot).sum().backward(); torch.cuda.synchronize()
        ops.BWD_EXACT_KERNELS=False
        return [y.grad.clone()]+[p_.grad.clone() if p_.grad is not None else torch.zeros_like(p_) for p_ in sde.parameters()]
    ref, got = run(True), run(False)
    names=['y0']+[n for n,_ in sde.named_parameters()]
    print(f"--- F={F} rows={rows} dw={use_dw} scale={scale}")
    for n,a,b in zip(names,got,ref):
        e=float((a-b).abs().max()/(b.abs().max()+1e-30))
        print(f"  {n:28s} rel err {e:.2e}  |ref|max {float(b.abs().max()):.3e} |got|max {float(a.abs().max()):.3e}")
for case in [(10,128,True,1.0),(60,300,True,1e-6),(20,129,False,1e-3),(100,64,True,1.0)]:
    run_case(*case)
# timing
for rows in (25600, 204800):
    sde = init_like_reference(DecoderSDE(), seed=1).to(DEV)
    ts=torch.linspace(0,6,61); y0=torch.relu(torch.randn(rows,64,device=DEV))
    for exact in (False,) if rows>30000 else (False,True):
        ops.BWD_EXACT_KERNELS=exact
        def step():
            y=y0.clone().requires_grad_(True)
            ys=tb.sdeint(sde,y,ts,dt=0.1,method='euler',mode='tc_f16',seed=3)
            e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
            g=torch.ones_like(ys)*1e-6
            e0.record(); ys.backward(g); e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1)
        step(); t=[step() for _ in range(3)]
        print(f"rows={rows} exact_bwd={exact}: backward {min(t):.3f} ms")
    ops.BWD_EXACT_KERNELS=False
