#!/usr/bin/env python
"""BASELINE configs[4]: solver-step sweep (10-200 steps) x agents-per-scene (8-128) throughput / roofline-fraction map on one B200.
The scene count is FIXED (--scenes, default 256), so the agents axis scales the work: decoder rows = 10 modes x A x scenes (20 k .. 328 k),
encoder rows = (A + 1) x scenes.  Per cell: the decoder solve with caller-supplied dW and with in-kernel Philox noise; per agents value
(the encoder always takes 21 steps): the fused encoder recurrence.  Prints one JSON object.

    python bench_sweep.py [--out profiles/r2_sweep.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=None)
    ap.add_argument('--iters', type=int, default=3)
    ap.add_argument('--scenes', type=int, default=256)
    args = ap.parse_args()
    import trajsde_b200 as tb
    from trajsde_b200 import encoder as enc_mod, synthetic as syn
    from trajsde_b200.schedule import euler_schedule
    dev = torch.device('cuda:0')
    hbm = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.isfile(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
    sde = syn.init_reference_style(syn.DecoderSDEFunc(), 2).to(dev)
    res = []
    for F in (10, 20, 50, 100, 200):
        ts = torch.linspace(0, 0.1 * F, F + 1)
        sched = euler_schedule(ts, 0.1)
        S, T = sched.n_steps, sched.n_outputs + 1
        for A in (8, 16, 32, 64, 128):
            scenes = args.scenes
            rows = scenes * A * 10
            y0 = torch.relu(torch.randn(rows, 64, device=dev))
            dW = torch.randn(S, rows, 64, device=dev) * 0.3
            row = {"future_steps": F, "euler_steps": S, "agents_per_scene": A, "scenes": scenes, "rows": rows}
            for name, bm in (("fixed_dw", dW), ("philox", None)):
                with torch.no_grad():
                    for _ in range(2):
                        tb.sdeint(sde, y0, ts, bm=bm, dt=0.1, method='euler', seed=1)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(args.iters):
                        tb.sdeint(sde, y0, ts, bm=bm, dt=0.1, method='euler', seed=1)
                    e1.record()
                    torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.iters
                nbytes = rows * 256 * (1 + T + (S if bm is not None else 0))
                row[name] = {"ms": ms, "agent_steps_per_s": rows * S / (ms * 1e-3), "hbm_frac": nbytes / (ms * 1e-3) / 1e9 / hbm}
            res.append(row)
            del dW, y0
            torch.cuda.empty_cache()
    enc_sde = syn.init_reference_style(syn.EncoderSDEFunc(), 1).to(dev)
    gru = syn.init_reference_style(syn.GRUUnit(), 3).to(dev)
    enc = []
    for A in (8, 16, 32, 64, 128):
        b = syn.make_batch(args.scenes, A, seed=A, mixed_sources=True)
        t = {k: getattr(b, k).to(dev) for k in ('enc_h0', 'aa_out', 'actors_mask', 'nus_mask')}
        with torch.no_grad():
            for _ in range(2):
                enc_mod.encoder_recurrence(enc_sde, gru, t['enc_h0'], t['aa_out'], t['actors_mask'], t['nus_mask'], seed=1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                enc_mod.encoder_recurrence(enc_sde, gru, t['enc_h0'], t['aa_out'], t['actors_mask'], t['nus_mask'], seed=1)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        enc.append({"agents_per_scene": A, "scenes": args.scenes, "rows": b.enc_rows, "ms": ms, "agent_steps_per_s": b.enc_rows * 21 / (ms * 1e-3),
                    "hbm_frac": b.enc_rows * 21 * 517 / (ms * 1e-3) / 1e9 / hbm})
    out = {"what": "decoder solve sweep (steps x agents) + fused encoder recurrence sweep (agents), tc_f16 kernels, CUDA-event time of the call, "
                   f"{args.scenes} scenes", "hbm_peak_gbs": hbm, "rows": res, "encoder": enc}
    s = json.dumps(out, indent=1)
    if args.out:
        open(args.out, 'w').write(s)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
