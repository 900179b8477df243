import sys, os, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import trajsde_b200 as tb
from helpers import DecoderSDE, init_like_reference, make_dw, net_params
from oracle import sde_oracle as so
from trajsde_b200.schedule import euler_schedule
DEV='cuda:0'
for F in (10, 30, 60, 100, 200):
    sde = init_like_reference(DecoderSDE(), seed=F).to(DEV)
    ts = torch.linspace(0, 0.1 * F, F + 1)
    sched = euler_schedule(ts, 0.1)
    y0 = torch.relu(torch.randn(70, 64, generator=torch.Generator().manual_seed(F)))
    dW = make_dw(sched.h, 70, seed=F) * 0.5
    ref, _ = so.euler_solve_ref(net_params(sde.f_func), net_params(sde.g_func), y0, ts, 0.1, dW)
    ys = tb.sdeint(sde, y0.to(DEV), ts, bm=dW.to(DEV), dt=0.1, method='euler', mode='tc_f16').cpu()
    print(F, 'max-abs', float((ys-ref).abs().max()), 'latent max', float(ref.abs().max()), 'rel-to-max', float((ys-ref).abs().max()/ref.abs().max()))
