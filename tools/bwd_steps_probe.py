"""Decoder backward on ONE tile for 1..16 Euler steps: fixed per-launch cost vs per-step cost of euler_bwd_tc_kernel (run under ncu launch list)."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
import trajsde_b200 as tb
from trajsde_b200 import synthetic as syn
dev = 'cuda:0'
sde = syn.init_reference_style(syn.DecoderSDEFunc(), 2).to(dev)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for F in (1, 2, 4, 8, 16, 1, 2, 4, 8, 16):
    ts = torch.linspace(0, 0.1 * F, F + 1)
    y0 = torch.relu(torch.randn(rows, 64, device=dev)).requires_grad_(True)
    ys = tb.sdeint(sde, y0, ts, dt=0.1, method='euler', mode='tc_f16', seed=1)
    ys.backward(torch.full_like(ys, 1e-6))
torch.cuda.synchronize()
