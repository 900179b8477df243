"""A/B of trajsde_enc_bwd: single persistent launch (default) vs two launches per iteration (ops.ENC_BWD_PER_STEP).
    python bench_micro/enc_bwd_ab.py        # encoder fwd+bwd at the cfg2 / cfg3 encoder sizes, CUDA events, 20 reps"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from trajsde_b200 import encoder as enc, ops, synthetic as syn  # noqa: E402

DEV = torch.device('cuda:0')


def run(rows, per_step, philox=True, reps=20):
    torch.manual_seed(0)
    sde = syn.init_reference_style(syn.EncoderSDEFunc(), 1, bias_std=0.1).to(DEV)
    gru = syn.init_reference_style(syn.GRUUnit(), 2, bias_std=0.1).to(DEV)
    h0 = (torch.randn(rows, 64) * 0.3).to(DEV).requires_grad_(True)
    aa = torch.randn(21, rows, 64).to(DEV).requires_grad_(True)
    am = (torch.rand(rows, 21) > 0.3).to(DEV)
    nm = (torch.rand(rows) > 0.5).to(DEV)
    dW = None if philox else (torch.randn(21, rows, 64) * 0.3).to(DEV)
    cot = torch.randn(21, rows, 64).to(DEV)
    ops.ENC_BWD_PER_STEP = per_step
    fwd, tot = [], []
    for i in range(reps + 3):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        lat, gg = enc.encoder_recurrence(sde, gru, h0, aa, am, nm, dW=dW, mode='tc_f16', fused=True, seed=5)
        e1.record()
        ((lat * cot).sum() + gg.sum()).backward()
        e2.record()
        torch.cuda.synchronize()
        if i >= 3:
            fwd.append(e0.elapsed_time(e1))
            tot.append(e0.elapsed_time(e2))
    ops.ENC_BWD_PER_STEP = False
    fwd.sort(), tot.sort()
    return fwd[len(fwd) // 2], tot[len(tot) // 2]


def timeline(rows):
    """needs a -DTRAJSDE_SWEEP_TIMELINE build (TRAJSDE_LIB_PATH=bench_micro/libtrajsde_b200_var.so)"""
    import ctypes as C
    from trajsde_b200 import _lib
    L = _lib.lib()
    buf = (C.c_longlong * 640)()
    run(rows, False, reps=2)
    L.trajsde_debug_sweep_timeline(buf)
    run(rows, False, reps=1)        # 4 calls
    L.trajsde_debug_sweep_timeline(buf)
    v = [buf[i] / 4 for i in range(640)]
    n = sum(1 for i in range(160) if v[4 * i] > 0)
    tiles = (rows + 127) // 128
    Gs, Gg = min(tiles, 148 // 4), min(tiles, 148 - 2 * (148 // 4))
    assert n == Gg + 2 * Gs, (n, Gg, Gs)
    for role in range(3):
        cs = range(0, Gg) if role == 0 else range(Gg + (role - 1) * Gs, Gg + role * Gs)
        G = len(cs)
        tot = sum(v[4 * c] for c in cs) / G
        wt = sum(v[4 * c + 1] for c in cs) / G
        pb = sum(v[4 * c + 2] for c in cs) / G
        nb = sum(v[4 * c + 3] for c in cs) / G
        print(f"rows={rows} role {role}: kernel {tot / 1e3:.1f} kclk  in wait {wt / 1e3:.1f} kclk  in publish {pb / 1e3:.1f} kclk  blocked waits {nb:.1f}  (per CTA, thread 0, mean of {G} CTAs)")


def gru_timeline(rows):
    """needs a -DTRAJSDE_GRU_TIMELINE build: clocks per GRU tile of CTA 0 of the single-launch sweep, by segment"""
    import ctypes as C
    from trajsde_b200 import _lib
    L = _lib.lib()
    buf = (C.c_longlong * 24)()
    run(rows, False, reps=2)
    L.trajsde_debug_gru_segments(buf)
    run(rows, False, reps=1)        # 4 calls
    L.trajsde_debug_gru_segments(buf)
    tiles = (rows + 127) // 128
    Gg = min(tiles, 148 - 2 * (148 // 4))
    n_tiles = 4 * 21 * (tiles // Gg + (1 if 0 < tiles % Gg else 0))        # CTA 0 gets the larger share
    names = ['loop / flush tail', 'tile start (role waits, loads, transposes)'] + [f'{ph}: {what}' for ph in ('F1', 'F2', 'F3', 'F4', 'B1', 'B2', 'B3', 'B4')
                                                                                    for what in ('work before wait', 'wait MMA')]
    names += ['B4 epilogue + stores + publish', 'wait weight-gradient MMAs', 'dU1|dR1 flush']
    tot = 0
    for i, nm in enumerate(names):
        print(f"  {nm:44s} {buf[i] / n_tiles:8.0f} clk per tile")
        tot += buf[i] / n_tiles
    print(f"  total {tot:.0f} clk per GRU tile (rows={rows}, {n_tiles // 4} tiles per call on CTA 0)")


if __name__ == '__main__':
    if '--gru-timeline' in sys.argv:
        for rows in (2688, 21504):
            gru_timeline(rows)
        sys.exit(0)
    if '--timeline' in sys.argv:
        for rows in (2688, 21504, 86016):
            timeline(rows)
        sys.exit(0)
    for rows in (2688, 21504, 86016):
        for per_step in (True, False):
            f, t = run(rows, per_step)
            print(f"rows={rows:7d} {'per-step launches' if per_step else 'single launch    '}: fwd {f:.3f} ms  fwd+bwd {t:.3f} ms  (bwd side {t - f:.3f} ms)  "
                  f"status={ops.backward_status(DEV)}")
