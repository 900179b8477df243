"""Where the fixed per-launch cost of euler_bwd_tc_kernel goes (one tile, F Euler steps), from the -DTRAJSDE_BWD_TIMELINE build:
    bash bench_micro/build_timeline_lib.sh && TRAJSDE_LIB_PATH=bench_micro/libtrajsde_b200_tl.so python bench_micro/bwd_fixed_cost.py
"""
import ctypes as C, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import torch
import trajsde_b200 as tb
from trajsde_b200 import _lib
from helpers import DecoderSDE, init_like_reference
DEV = 'cuda:0'
sde = init_like_reference(DecoderSDE(), seed=1).to(DEV)
names = ['tile tail (+ per-step loop tail)', 'SS: transposes', 'wait bar_wg', 'SS: y, arrive, adjoint, df', 'wait P1', 'e1 + prefetch', 'wait P2', 'e2',
         'wait D1', 'e3', 'wait D2', 'e4', 'wait D3', 'PROLOGUE', 'FLUSH']
L = _lib.lib()
for F in (1, 4):
    ts = torch.linspace(0, 0.1 * F, F + 1)
    for rep in range(2):
        buf = (C.c_longlong * 16)()
        y = torch.relu(torch.randn(128, 64, device=DEV)).requires_grad_(True)
        ys = tb.sdeint(sde, y, ts, dt=0.1, method='euler', mode='tc_f16', seed=3)
        torch.cuda.synchronize(); L.trajsde_debug_bwd_timeline(buf)
        ys.backward(torch.full_like(ys, 1e-6)); torch.cuda.synchronize(); L.trajsde_debug_bwd_timeline(buf)
    print(f"--- one tile, {F} step(s): clocks of thread 0 (second call)")
    for n, v in zip(names, buf):
        print(f"  {n:40s} {v:9d}")
    print(f"  {'total':40s} {sum(buf[:15]):9d}")
