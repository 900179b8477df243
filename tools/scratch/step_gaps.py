import sys, os, torch, time
sys.path.insert(0, os.getcwd())
import trajsde_b200 as tb
from trajsde_b200 import synthetic as syn, encoder as enc_mod
from trajsde_b200.schedule import encoder_schedule, euler_schedule
dev=torch.device('cuda:0')
enc_sde=syn.init_reference_style(syn.EncoderSDEFunc(),1).to(dev); dec_sde=syn.init_reference_style(syn.DecoderSDEFunc(),2).to(dev); gru=syn.init_reference_style(syn.GRUUnit(),3).to(dev)
host=syn.make_batch(1024,20,seed=1000)
res={k:getattr(host,k).to(dev) for k in ('enc_h0','aa_out','actors_mask','nus_mask','dec_y0')}
E,M=host.enc_rows,host.dec_rows
ts=torch.linspace(0,6,61); sd,se=euler_schedule(ts,0.1),encoder_schedule()
gen=torch.Generator(device=dev).manual_seed(1)
dW_d=torch.randn(61,M,64,device=dev,generator=gen)*0.3; dW_e=torch.randn(21,E,64,device=dev,generator=gen)*0.3
ev=lambda: torch.cuda.Event(enable_timing=True)
def step(fixed, evs=None):
    with torch.no_grad():
        a=ev(); a.record()
        lat,g=enc_mod.encoder_recurrence(enc_sde,gru,res['enc_h0'],res['aa_out'],res['actors_mask'],res['nus_mask'],dW=dW_e if fixed else None,seed=0)
        b=ev(); b.record()
        ys=tb.sdeint(dec_sde,res['dec_y0'],ts,bm=dW_d if fixed else None,dt=0.1,method='euler',seed=100)
        c=ev(); c.record()
    if evs is not None: evs.append((a,b,c))
for fixed in (True, False):
    for _ in range(3): step(fixed)
    torch.cuda.synchronize(); evs=[]; t0=time.perf_counter(); e0=ev(); e0.record()
    for _ in range(10): step(fixed, evs)
    e1=ev(); e1.record(); torch.cuda.synchronize(); host_ms=(time.perf_counter()-t0)*100
    enc=sum(a.elapsed_time(b) for a,b,c in evs)/10; dec=sum(b.elapsed_time(c) for a,b,c in evs)/10
    print('fixed' if fixed else 'philox', f'total/step {e0.elapsed_time(e1)/10:.3f} ms  enc {enc:.3f} dec {dec:.3f}  host wall/step {host_ms:.3f}')
    t0=time.perf_counter()
    for _ in range(20): step(fixed)
    print('  host enqueue time/step (no sync)', (time.perf_counter()-t0)/20*1e3, 'ms'); torch.cuda.synchronize()
