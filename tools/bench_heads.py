"""Fused decoder heads vs the reference nn.Sequential heads on the solver output of configs[1]: python tools/bench_heads.py [rows] [iters]"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import torch, torch.nn as nn
from trajsde_b200 import heads as hd
dev = 'cuda:0'
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 204800
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
T = 60
mk = lambda: nn.Sequential(nn.Linear(64, 64), nn.LayerNorm(64), nn.ReLU(inplace=True), nn.Linear(64, 2)).to(dev)
torch.manual_seed(0)
loc_h, sc_h = mk(), mk()
def timeit(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for layout in ('rows_major', 'time_major'):
    if layout == 'rows_major':
        store = torch.randn(rows, T + 1, 64, device=dev); sol_y = store[:, 1:]
    else:
        store = torch.randn(T + 1, rows, 64, device=dev); sol_y = store[1:].permute(1, 0, 2)
    with torch.no_grad():
        ms_f = timeit(lambda: hd.decoder_heads(loc_h, sc_h, sol_y))
        ms_e = timeit(lambda: (loc_h(sol_y), sc_h(sol_y)))
        a, b = hd.decoder_heads(loc_h, sc_h, sol_y)
        err = max((a - loc_h(sol_y)).abs().max().item(), (b - sc_h(sol_y)).abs().max().item())
    gb = rows * T * (256 + 16) / 1e9
    print(f"heads {layout}: rows={rows} fused {ms_f:.3f} ms ({gb / ms_f * 1e3:.0f} GB/s algorithmic)  eager torch {ms_e:.3f} ms  x{ms_e / ms_f:.1f}  max|diff| {err:.2e}")
    del store, sol_y
