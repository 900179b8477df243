"""Per-step wait clocks of the decoder forward (thread 0 of CTA 0, slot 0): needs a -DTRAJSDE_FWD_TIMELINE build
    EXTRA=-DTRAJSDE_FWD_TIMELINE OUT=bench_micro/libtrajsde_b200_var.so bash bench_micro/build_variant_lib.sh
    TRAJSDE_LIB_PATH=bench_micro/libtrajsde_b200_var.so python bench_micro/fwd_timeline.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import trajsde_b200 as tb  # noqa: E402
from trajsde_b200 import _lib, synthetic as syn  # noqa: E402

dev = 'cuda:0'
rows = 204800
sde = syn.init_reference_style(syn.DecoderSDEFunc(), 2).to(dev)
ts = torch.linspace(0, 6, 61)
y0 = torch.relu(torch.randn(rows, 64, device=dev))
dW = torch.randn(61, rows, 64, device=dev) * 0.3
L = _lib.lib()
buf = (C.c_longlong * 8)()
seg = (C.c_longlong * 16)()
SEG = ['loop tail + ring wait', 'wait P1', 'epilogue 1 (+ draw)', 'wait P2f', 'epilogue 2a (+ draw)', 'wait P2g', 'epilogue 2b', 'X free + dW landed (or draw)',
       'wait P3', 'epilogue 3', 'fences + arrives']
for name, bm in (('dw', dW), ('philox', None)):
    with torch.no_grad():
        for _ in range(2):
            tb.sdeint(sde, y0, ts, bm=bm, dt=0.1, method='euler', mode='tc_f16', seed=1)
        torch.cuda.synchronize()
        L.trajsde_debug_fwd_timeline(buf)
        L.trajsde_debug_fwd_segments(seg)
        n = 4
        for _ in range(n):
            tb.sdeint(sde, y0, ts, bm=bm, dt=0.1, method='euler', mode='tc_f16', seed=1)
        torch.cuda.synchronize()
        L.trajsde_debug_fwd_timeline(buf)
    L.trajsde_debug_fwd_segments(seg)
    v = [buf[i] / n for i in range(8)]
    steps = max(v[4], 1)
    print(f"{name}: kernel {v[0] / 1e3:.0f} kclk, slot-0 steps {steps:.0f} -> {v[0] / steps:.0f} clk per slot-step; per step: wait X free {v[1] / steps:.0f}, "
          f"wait dW landed {v[2] / steps:.0f}, wait P3 {v[3] / steps:.0f}")
    for i, name_ in enumerate(SEG):
        print(f"    {name_:32s} {seg[i] / n / steps:8.0f} clk")

