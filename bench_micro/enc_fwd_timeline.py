"""Share of an encoder-recurrence iteration the epilogue warps spend waiting for the tensor core (thread 0 of CTA 0): needs a
-DTRAJSDE_ENC_TIMELINE build (EXTRA=-DTRAJSDE_ENC_TIMELINE OUT=bench_micro/libtrajsde_b200_var.so bash bench_micro/build_variant_lib.sh;
TRAJSDE_LIB_PATH=bench_micro/libtrajsde_b200_var.so python bench_micro/enc_fwd_timeline.py)"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from trajsde_b200 import _lib, encoder as enc_mod, synthetic as syn  # noqa: E402

dev = torch.device('cuda:0')
enc_sde = syn.init_reference_style(syn.EncoderSDEFunc(), 1).to(dev)
gru = syn.init_reference_style(syn.GRUUnit(), 3).to(dev)
L = _lib.lib()
buf = (C.c_longlong * 4)()
seg = (C.c_longlong * 20)()
PH = ['P1', 'P2f', 'P2g', 'P3', 'G1', 'G2', 'G3', 'G4']
for scenes in (128, 1024):
    b = syn.make_batch(scenes, 20, seed=5, mixed_sources=True)
    tr = {k: getattr(b, k).to(dev) for k in ('enc_h0', 'aa_out', 'actors_mask', 'nus_mask')}
    dW = torch.randn(21, tr['enc_h0'].shape[0], 64, device=dev) * 0.3
    for name, kw in (('dw', dict(dW=dW)), ('philox', dict(seed=300))):
        with torch.no_grad():
            for _ in range(2):
                enc_mod.encoder_recurrence(enc_sde, gru, tr['enc_h0'], tr['aa_out'], tr['actors_mask'], tr['nus_mask'], **kw)
            torch.cuda.synchronize()
            L.trajsde_debug_enc_timeline(buf)
            L.trajsde_debug_enc_segments(seg)
            n = 4
            for _ in range(n):
                enc_mod.encoder_recurrence(enc_sde, gru, tr['enc_h0'], tr['aa_out'], tr['actors_mask'], tr['nus_mask'], **kw)
            torch.cuda.synchronize()
            L.trajsde_debug_enc_timeline(buf)
        L.trajsde_debug_enc_segments(seg)
        its = max(buf[2] / n, 1)
        print(f"scenes={scenes} {name}: kernel {buf[0] / n / 1e3:.0f} kclk, {its:.0f} iterations on CTA 0 -> {buf[0] / n / its:.0f} clk per iteration, "
              f"{buf[1] / n / its:.0f} of them waiting for the tensor core (8 hand-shakes)")
        print('    ' + ' | '.join(f"{PH[k]}: {seg[2 * k] / n / its:.0f} + wait {seg[2 * k + 1] / n / its:.0f}" for k in range(8)) + f" | tail {seg[16] / n / its:.0f}")

