import sys, os, time
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,R)
import torch
import trajsde_b200 as tb
from trajsde_b200 import synthetic as syn
from trajsde_b200.schedule import euler_schedule
dev='cuda:0'
rows=int(sys.argv[1]) if len(sys.argv)>1 else 204800
iters=int(sys.argv[2]) if len(sys.argv)>2 else 5
modes=sys.argv[3].split(",") if len(sys.argv)>3 else ["dw","philox"]
rm=len(sys.argv)>4 and sys.argv[4]=="rows_major"
sde=syn.init_reference_style(syn.DecoderSDEFunc(),2).to(dev)
ts=torch.linspace(0,6,61); sched=euler_schedule(ts,0.1)
y0=torch.relu(torch.randn(rows,64,device=dev))
dW=torch.randn(61,rows,64,device=dev)*0.3
for name in modes:
    bm=dW if name=='dw' else None
    with torch.no_grad():
        for _ in range(2): tb.sdeint(sde,y0,ts,bm=bm,dt=0.1,method='euler',mode='tc_f16',seed=1,rows_major=rm)
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): tb.sdeint(sde,y0,ts,bm=bm,dt=0.1,method='euler',mode='tc_f16',seed=1,rows_major=rm)
        e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/iters
    print(f"{name}: rows={rows} {ms:.3f} ms  {rows*61/ms*1e-6:.3f} G row-steps/s  clk/row-step/SM@1.9GHz={148*1.9e9/(rows*61/(ms*1e-3)):.1f}")
