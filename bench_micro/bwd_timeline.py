"""Per-phase clock budget of euler_bwd_tc_kernel (thread 0 of CTA 0) from a library built with -DTRAJSDE_BWD_TIMELINE.

    bash bench_micro/build_timeline_lib.sh && TRAJSDE_LIB_PATH=bench_micro/libtrajsde_b200_tl.so python bench_micro/bwd_timeline.py
"""
import ctypes as C, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import torch
import trajsde_b200 as tb
from trajsde_b200 import _lib
from helpers import DecoderSDE, init_like_reference
DEV = 'cuda:0'
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 204800
sde = init_like_reference(DecoderSDE(), seed=1).to(DEV)
ts = torch.linspace(0, 6, 61); y0 = torch.relu(torch.randn(rows, 64, device=DEV))
dW = torch.randn(61, rows, 64, device=DEV) * 0.3
names = ['loop / e5 tail / tile tail', 'SS: transposes', 'wait bar_wg', 'SS: y, arrive, adjoint, q, E, df', 'wait P1', 'e1 + prefetch issue', 'wait P2', 'e2',
         'wait D1', 'e3', 'wait D2', 'e4', 'wait D3']
L = _lib.lib()
for label, bm in (('supplied dW', dW), ('philox', None)):
    buf = (C.c_longlong * 16)()
    y = y0.clone().requires_grad_(True)
    ys = tb.sdeint(sde, y, ts, bm=bm, dt=0.1, method='euler', mode='tc_f16', seed=3)
    g = torch.ones_like(ys) * 1e-6
    torch.cuda.synchronize(); L.trajsde_debug_bwd_timeline(buf)
    ys.backward(g); torch.cuda.synchronize(); L.trajsde_debug_bwd_timeline(buf)
    tiles = -(-rows // 128); per_cta = tiles // 148 + (1 if tiles % 148 else 0)
    steps = per_cta * 61
    print(f"--- {label}: rows {rows}, CTA 0 processed {per_cta} tiles x 61 steps; clocks per tile-step")
    tot = 0
    for n, v in zip(names, buf):
        print(f"  {n:40s} {v / steps:9.1f}"); tot += v / steps
    print(f"  {'total':40s} {tot:9.1f}")
