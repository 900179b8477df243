#!/usr/bin/env python
"""BASELINE configs[4]: solver-step sweep (10-200 steps) x agents-per-scene (8-128) throughput / roofline-fraction map of the
fused decoder solve on one B200.  Rows are kept near 2e5 (scenes = 2e5 / (10 modes x agents)).  Prints one JSON object.

    python bench_sweep.py [--out profiles/r1_sweep.json]
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
    args = ap.parse_args()
    import trajsde_b200 as tb
    from trajsde_b200 import synthetic as syn
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
            scenes = max(1, round(200_000 / (10 * A)))
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
    out = {"what": "decoder solve sweep, tc_f16 kernel, CUDA-event time of the sdeint call", "hbm_peak_gbs": hbm, "rows": res}
    s = json.dumps(out, indent=1)
    if args.out:
        open(args.out, 'w').write(s)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
