import sys, os, torch
sys.path.insert(0, os.getcwd())
import trajsde_b200 as tb
from trajsde_b200 import synthetic as syn, ops
dev='cuda:0'
sde=syn.init_reference_style(syn.DecoderSDEFunc(),2).to(dev)
rows=204800; ts=torch.linspace(0,6,61)
y0=torch.relu(torch.randn(rows,64,device=dev))
cot=torch.randn(rows,61,64,device=dev).permute(1,0,2)*1e-6
act=(torch.rand(rows,device=dev)<0.1)
cot_sparse=(cot*act.view(1,-1,1))
def run(c,skip,n=5):
    ops.SKIP_ZERO_ROWS=skip
    def step():
        y=y0.clone().requires_grad_(True)
        ys=tb.sdeint(sde,y,ts,dt=0.1,method='euler',seed=1,rows_major=True)
        ys.backward(c)
    for _ in range(2): step()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
for name,c in (('dense',cot),('sparse10',cot_sparse)):
    for skip in (False,True):
        print(name,'skip' if skip else 'full', f'{run(c,skip):.3f} ms fwd+bwd')
