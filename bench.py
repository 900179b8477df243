#!/usr/bin/env python
"""bench.py — agent-SDE-steps/s of the fused SDE encoder+decoder forward on synthetic Argoverse-shaped batches.

    python bench.py [--gpus N --steps K --warmup W]            our arm (one rank per GPU under torch.distributed.run for N>1)
    python bench.py --impl reference [--steps K --warmup W]    the reference's CPU torchsde path (oracle port) on host cores

A "step" is one forward pass of the SDE hot path over one synthetic batch: the encoder recurrence (21 x [one-step
sdeint_dual + GRU jump], rows = agents + one perturbed target copy per scene) followed by the decoder solve (61 Euler
steps, rows = 10 modes x agents).  Workload at every N: BASELINE.json configs[1] per GPU — 1024 scenes x 20 agents —
scene-sharded with no forward collective (weak scaling; at N=8 this is configs[3]'s 8192-scene batch).
metric = agent-SDE-steps/s = (enc_rows*21 + dec_rows*61) / step time, whole job.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "agent_sde_steps_per_s"
UNIT = "agent-SDE-steps/s"
ENC_STEPS, DEC_STEPS = 21, 61


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), float(d.get('bf16_tflops', 1590.5)), 'measured'
    return 6650.0, 1590.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks/throttle sampler running during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '20',
                                       '-i', str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
            time.sleep(0.3)                       # let the first samples land before the timed region starts
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().strip().splitlines():
            c = [x.strip() for x in line.split(',')]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), c[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU torchsde path restated (oracle port), timed on the host cores
# ------------------------------------------------------------------------------------------------------------------------------
def cpu_reference_pass(scenes, agents, seed=0):
    """One forward of the SDE path the way the reference runs it on CPU: internally generated Brownian increments, the
    contract probe evaluation per sdeint call, GRU jump between encoder steps.  Returns (seconds, agent-SDE-steps)."""
    from oracle import sde_oracle as so
    from trajsde_b200 import synthetic as syn
    b = syn.make_batch(scenes, agents, seed=seed)
    enc, dec, gru = _cpu_modules()
    pe = {k: {n: v.detach() for n, v in getattr(enc, k).net.state_dict().items()} for k in ('f_func', 'g_nus', 'g_argo')}
    pd = {k: {n: v.detach() for n, v in getattr(dec, k).net.state_dict().items()} for k in ('f_func', 'g_func')}
    pg = {n: v.detach() for n, v in gru.state_dict().items()}
    ts = torch.linspace(0, 6, 61)
    t0 = time.perf_counter()
    with torch.no_grad():
        dW_e = torch.randn(ENC_STEPS, b.enc_rows, 64) * (0.1 ** 0.5)
        so.encoder_recurrence_ref(pe['f_func'], pe['g_nus'], pe['g_argo'], pg, b.enc_h0, b.aa_out, b.actors_mask, b.nus_mask,
                                  dW_e, probe=True)
        sched = so.euler_schedule_ref(ts, 0.1)
        dW_d = torch.randn(DEC_STEPS, b.dec_rows, 64) * torch.sqrt(sched['h']).view(-1, 1, 1)
        so.euler_solve_ref(pd['f_func'], pd['g_func'], b.dec_y0, ts, 0.1, dW_d, probe=True)
    dt = time.perf_counter() - t0
    return dt, b.enc_rows * ENC_STEPS + b.dec_rows * DEC_STEPS


_CPU_MODS = None


def _cpu_modules():
    global _CPU_MODS
    if _CPU_MODS is None:
        from trajsde_b200 import synthetic as syn
        _CPU_MODS = (syn.init_reference_style(syn.EncoderSDEFunc(), 1), syn.init_reference_style(syn.DecoderSDEFunc(), 2),
                     syn.init_reference_style(syn.GRUUnit(), 3))
    return _CPU_MODS


def cpu_baseline(budget_s=12.0, scenes=32, agents=20):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cpu_reference_pass(scenes, agents)             # warm-up
    best, n, t_start = float('inf'), 0, time.perf_counter()
    work = 0
    while n < 3 or (time.perf_counter() - t_start < budget_s and n < 50):
        dt, work = cpu_reference_pass(scenes, agents)
        best = min(best, dt)
        n += 1
    return {"value": work / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"BASELINE configs[0]: {scenes} scenes x {agents} agents (enc {scenes * (agents + 1)} rows x 21 steps + GRU, "
                      f"dec {scenes * agents * 10} rows x 61 steps), best of {n} passes, torch {torch.__version__} CPU, "
                      f"oracle port of the reference torchsde path incl. its contract-probe f/g evaluation and randn increments",
            "scenes_per_s": scenes / best}


def run_reference_arm(args, out=sys.stdout):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    scenes, agents = 32, 20
    for _ in range(max(args.warmup, 1)):
        cpu_reference_pass(scenes, agents)
    t0 = time.perf_counter()
    work = 0
    for _ in range(args.steps):
        _, w = cpu_reference_pass(scenes, agents)
        work += w
    dt = time.perf_counter() - t0
    v = work / dt
    sample = (f"each step = one forward of BASELINE configs[0] ({scenes} scenes x {agents} agents) — a bounded sample of the "
              f"configs[1] workload (same per-row work, 1/32 of the rows)")
    print(file=out, flush=True, *[json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2-shaped SDE encoder+decoder forward, CPU sample of 32 scenes x 20 agents per step"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "scenes_per_s": scenes * args.steps / dt,
    })])


# ------------------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------------------
def _claim_stdout():
    """The contract is ONE JSON line on stdout: route everything else that might print there (NCCL's version banner, library
    chatter) to stderr at file-descriptor level and return a handle on the real stdout for the final line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--scenes', type=int, default=1024, help='scenes per GPU')
    ap.add_argument('--agents', type=int, default=20)
    ap.add_argument('--mode', default='tc_f16', choices=['tc_f16', 'exact'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--train-scenes', type=int, default=128, help='scenes per GPU of the fwd+bwd training step (reference batch 128, yml:106)')
    ap.add_argument('--no-train', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='skip the CUDA-graph replay of the training step')
    ap.add_argument('--no-graph-ddp', action='store_true', help='N > 1: do not capture the training step (with its NCCL all-reduces) into a CUDA graph')
    ap.add_argument('--strong-scenes', type=int, default=8192, help='total scenes of the strong-scaling training entry (0 = skip)')
    ap.add_argument('--no-heads', action='store_true')
    ap.add_argument('--e2e-chunks', type=int, default=8, help='decoder row slices per e2e step (H2D / kernels / D2H overlap)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    out = _claim_stdout()
    if args.impl == 'reference':
        return run_reference_arm(args, out)

    import torch.distributed as dist
    import trajsde_b200 as tb
    from trajsde_b200 import encoder as enc_mod
    from trajsde_b200 import ops, synthetic as syn
    from trajsde_b200.schedule import encoder_schedule, euler_schedule

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    from trajsde_b200.dist import bind_host_to_gpu
    cores_bound = bind_host_to_gpu(local) if world > 1 else None      # host-fed path: pinned buffers next to this rank's GPU
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    mode = args.mode
    enc_sde = syn.init_reference_style(syn.EncoderSDEFunc(), 1).to(dev)
    dec_sde = syn.init_reference_style(syn.DecoderSDEFunc(), 2).to(dev)
    gru = syn.init_reference_style(syn.GRUUnit(), 3).to(dev)
    host = syn.make_batch(args.scenes, args.agents, seed=1000 + rank, pin=True)
    E, M = host.enc_rows, host.dec_rows
    work = E * ENC_STEPS + M * DEC_STEPS
    ts_dec = torch.linspace(0, 6, 61)
    sched_d, sched_e = euler_schedule(ts_dec, 0.1), encoder_schedule()

    # resident inputs for `value`
    res = {k: getattr(host, k).to(dev) for k in ('enc_h0', 'aa_out', 'actors_mask', 'nus_mask', 'dec_y0')}
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    dW_d = torch.randn(DEC_STEPS, M, 64, device=dev, generator=gen) * torch.sqrt(torch.from_numpy(sched_d.h)).to(dev).view(-1, 1, 1)
    dW_e = torch.randn(ENC_STEPS, E, 64, device=dev, generator=gen) * torch.sqrt(torch.from_numpy(sched_e.h)).to(dev).view(-1, 1, 1)

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    dec_events = []

    def step(inp, fixed_dw, seed=0, record=False):
        with torch.no_grad():
            lat, g = enc_mod.encoder_recurrence(enc_sde, gru, inp['enc_h0'], inp['aa_out'], inp['actors_mask'], inp['nus_mask'],
                                                dW=dW_e if fixed_dw else None, seed=seed, mode=mode)
            if record:
                e0, e1 = ev(), ev()
                e0.record()
            ys = tb.sdeint(dec_sde, inp['dec_y0'], ts_dec, bm=dW_d if fixed_dw else None, dt=0.1, dt_min=0.1, rtol=1e-3,
                           atol=1e-3, method='euler', mode=mode, seed=seed + 100, row_offset=rank * M)
            if record:
                e1.record()
                dec_events.append((e0, e1))
        return lat, ys

    def timed(fn, k):
        barrier()
        e0, e1 = ev(), ev()
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- value: fixed dW (configs[1]), inputs resident in HBM ------------------------------------------------------------------
    for i in range(args.warmup):
        step(res, True)
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = ops.LAUNCHES['n']
    ms_fixed = timed(lambda i: step(res, True, record=True), args.steps)
    launches = ops.LAUNCHES['n'] - n0
    if sampler is not None:                        # keep the GPU under the same load until enough clock samples exist
        t_end = time.perf_counter() + 1.0
        while time.perf_counter() < t_end:
            step(res, True)
        torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    dec_ms = sorted(a.elapsed_time(b) for a, b in dec_events)
    dec_ms_avg = sum(dec_ms) / len(dec_ms)
    dec_events.clear()

    # ---- Philox variant (bm=None, what the reference does in production), resident inputs -----------------------------------------
    for i in range(2):
        step(res, False, seed=i)
    ms_philox = timed(lambda i: step(res, False, seed=10 + i, record=True), args.steps)
    dec_ms_philox = sum(a.elapsed_time(b) for a, b in dec_events) / len(dec_events)
    dec_events.clear()

    # ---- e2e: public API with HOST buffers: every step copies all inputs from pinned host memory (the AA-encoder output as fp16: the
    # kernel rounds it to fp16 MMA operands anyway), runs bm=None (in-kernel Philox, like the reference's default BrownianInterval), the
    # fused heads on the device, and copies the decoder's RESULT out['loc'] = cat(loc, scale) [10 N, 60, 4] (dec…sde.py:95-100) plus the
    # encoder's final latents back; the batch arrives as `e2e_chunks` decoder slices whose copies overlap the kernels
    # (trajsde_b200.pipeline.HostFedSdePath) ---------------------------------------------------------------------------------------
    import torch.nn as nn
    from trajsde_b200.pipeline import HostFedSdePath
    mk_head = lambda sd: syn.init_reference_style(nn.Sequential(nn.Linear(64, 64), nn.LayerNorm(64), nn.ReLU(inplace=True),  # noqa: E731
                                                                nn.Linear(64, 2)), sd).to(dev)
    loc_h, sc_h = mk_head(7), mk_head(8)
    n_chunks = args.e2e_chunks
    aa_half = host.aa_out.half().pin_memory()
    out_enc = torch.empty((E, 64), dtype=torch.float32).pin_memory()
    out_dec = torch.empty((M, sched_d.n_outputs, 4), dtype=torch.float32).pin_memory()
    h2d = sum(getattr(host, k).numel() * getattr(host, k).element_size() for k in ('enc_h0', 'actors_mask', 'nus_mask', 'dec_y0')) + aa_half.numel() * 2
    d2h = (out_enc.numel() + out_dec.numel()) * 4
    pipe = HostFedSdePath(enc_sde, gru, dec_sde, dev, ts_dec, mode=mode)
    e2e_work = work

    def e2e_step(i):
        pipe.run_batch(host, out_enc, out_dec, seed=20 + 10 * i, dec_chunks=n_chunks, enc_row_offset=rank * E, dec_row_offset=rank * M,
                       heads=(loc_h, sc_h), aa_out_half=aa_half)

    for i in range(2):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps)
    # pinned-copy microbenchmark on the same box: what the host link gives this rank while every rank copies at once (the e2e floor)
    link = {}
    if mode == 'tc_f16':
        src = torch.empty((256 << 20,), dtype=torch.uint8).pin_memory()
        dst_d = torch.empty_like(src, device=dev)
        dst_h = torch.empty_like(src).pin_memory()
        s2 = torch.cuda.Stream(dev)
        def h2d_only(i):
            dst_d.copy_(src, non_blocking=True)
        def both(i):
            dst_d.copy_(src, non_blocking=True)
            with torch.cuda.stream(s2):
                dst_h.copy_(dst_d, non_blocking=True)
        h2d_only(0); both(0); torch.cuda.synchronize()
        t_h = timed(h2d_only, 4) / 4
        def both_timed(i):
            both(i)
            torch.cuda.current_stream().wait_stream(s2)
        t_b = timed(both_timed, 4) / 4
        link = {"h2d_gbs_per_rank": src.numel() / (t_h * 1e-3) / 1e9, "h2d_plus_d2h_gbs_per_rank_each_way": src.numel() / (t_b * 1e-3) / 1e9,
                "e2e_copy_floor_ms": max(h2d, d2h) / (src.numel() / (t_b * 1e-3)) * 1e3,
                "note": "256 MiB pinned copies timed on every rank at once; floor = max(H2D, D2H bytes) at the duplex rate"}
        del src, dst_d, dst_h
        torch.cuda.empty_cache()

    # ---- decoder heads on the solver output: fused launch vs the reference nn.Sequential heads; forward (inference) and
    # forward + backward under a winner-takes-all cotangent (one mode in ten, 30 valid slots: what losses/L2.py hands back) ---------------
    heads = None
    if not args.no_heads:
        from trajsde_b200 import heads as hd
        with torch.no_grad():
            ys_rm = tb.sdeint(dec_sde, res['dec_y0'], ts_dec, bm=dW_d, dt=0.1, method='euler', mode=mode, rows_major=True)
            sol_y = ys_rm[1:].permute(1, 0, 2)                                   # dec…sde.py:88: unit-stride rows here
            for _ in range(2):
                hd.decoder_heads(loc_h, sc_h, sol_y)
            ms_heads = timed(lambda i: hd.decoder_heads(loc_h, sc_h, sol_y), args.steps) / args.steps
            loc_f, sc_f = hd.decoder_heads(loc_h, sc_h, sol_y)
            loc_h(sol_y[:128])
            ms_heads_eager = timed(lambda i: (loc_h(sol_y), sc_h(sol_y)), 2) / 2
            err = max(float((loc_f - loc_h(sol_y)).abs().max()), float((sc_f - sc_h(sol_y)).abs().max()))
        cot = torch.zeros((M, sched_d.n_outputs, 2), device=dev)
        cot[::10, :30] = 1e-6
        ysg = ys_rm.detach().requires_grad_(True)

        def heads_fb(i):
            ysg.grad = None
            loc_, _ = hd.decoder_heads_from_solution(loc_h, sc_h, ysg)
            loc_.backward(cot)

        for _ in range(2):
            heads_fb(0)
        ms_heads_fb = timed(heads_fb, args.steps) / args.steps
        hbytes = M * sched_d.n_outputs * (256 + 16)
        heads = {"kernel": "heads_fwd_kernel (self.decoder + self.scale of SDEDecoder.forward, one launch)", "ms": ms_heads,
                 "points_per_s": world * M * sched_d.n_outputs / (ms_heads * 1e-3), "reference_torch_ms": ms_heads_eager,
                 "roofline_frac_hbm": hbytes / (ms_heads * 1e-3) / 1e9 / peaks()[0], "algorithmic_bytes": hbytes,
                 "max_abs_diff_vs_torch_fp32": err, "layout": "rows_major solver output",
                 "fwd_bwd_ms_sparse_cotangent": ms_heads_fb,
                 "fwd_bwd_note": "heads_fwd_kernel + heads_bwd_kernel (fp32, active points only) + the zero-fill of dL/dys; cotangent on 1 row in 10, 30 slots"}
        for p_ in list(loc_h.parameters()) + list(sc_h.parameters()):
            p_.grad = None
        del ys_rm, sol_y, loc_f, sc_f, cot, ysg
        torch.cuda.empty_cache()

    # ---- training step (BASELINE configs[2]/[3]): the SDE path and its direct consumers, for real: fused encoder recurrence -> eos
    # gather -> fused aggr_embed -> decoder solve -> fused heads -> out['loc'] -> L2 (winner-takes-all) + DiffBCE on the agents'
    # diffusion -> backward through all of it (heads_bwd, zero-row-skipping solver backward, aggr_embed_bwd, encoder backward) ->
    # all-reduce of the flat gradient buckets (decoder side launched on a side stream as soon as its gradients exist, under the encoder
    # backward) -> AdamW.  The HiVT stages that feed it (AA encoder -> aa_out, global interactor -> global_embed) are out of scope:
    # their outputs are inputs here and receive gradients. ------------------------------------------------------------------------------
    train = None
    if not args.no_train:
        from trajsde_b200 import stage as stg
        from trajsde_b200.dist import FlatGradBucket
        from trajsde_b200.stages import FusedDecoderMixin

        class DecStage(FusedDecoderMixin, syn.DecoderStage):
            pass

        dstage = syn.init_reference_style(DecStage(), 11).to(dev)
        hidden = torch.nn.Parameter(torch.randn(64, generator=torch.Generator().manual_seed(5)).to(dev) * 0.02)     # enc…sep2.py:61-62; same on every rank
        dec_params = list(dstage.parameters())
        enc_params = list(enc_sde.parameters()) + list(gru.parameters()) + [hidden]

        def measure_train(scenes_per_gpu, global_scenes=None):
            tbatch = syn.make_train_batch(scenes_per_gpu, args.agents, seed=2000 + rank, device=dev)
            b = tbatch.base
            Et, Mt, Nt = b.enc_rows, b.dec_rows, scenes_per_gpu * args.agents
            bk_dec, bk_enc = FlatGradBucket(dec_params, pack=True), FlatGradBucket(enc_params, pack=True)
            opt = torch.optim.AdamW(dec_params + enc_params, lr=1e-3, weight_decay=7e-4, fused=True)   # yml:2-3; fused: one launch per dtype/device group
            # instead of ~150 foreach / per-tensor launches (the eager step is launch-bound at 128 scenes)
            eos = 20 - torch.argmax(b.bos_mask.float(), dim=1)                                # enc…sep2.py:187
            ar = torch.arange(Nt, device=dev)
            new_agent_index = torch.cat((tbatch.agent_index, torch.arange(Nt, Et, device=dev)))   # :101
            agent_eos = eos[tbatch.agent_index].repeat(2)                                     # :190
            data = {'padding_mask': tbatch.padding_mask}
            pending = {}

            def train_step(i):
                bk_dec.zero_(); bk_enc.zero_()
                aa = b.aa_out.detach().requires_grad_(True)                    # the AA encoder upstream needs dL/daa_out
                ge = tbatch.global_embed.detach().requires_grad_(True)         # the global interactor upstream needs dL/dglobal_embed
                h0 = hidden.unsqueeze(0).repeat(Et, 1)                         # :78
                lat, g = enc_mod.encoder_recurrence(enc_sde, gru, h0, aa, b.actors_mask, b.nus_mask, seed=300 + i, mode=mode, row_offset=rank * Et)
                local_embed = lat[eos, ar]                                     # :184-188 (the AL encoder after it is out of scope)
                if world > 1:       # decoder-side gradients are complete when dL/dlocal_embed arrives: reduce them under the encoder backward
                    local_embed.register_hook(lambda gr: pending.__setitem__('dec', bk_dec.all_reduce_mean_async()))
                dstage.solver_kwargs = {'seed': 400 + i, 'row_offset': rank * Mt, 'mode': mode}
                out = dstage(data, local_embed, ge)
                diff_in, diff_out = torch.chunk(g[agent_eos, new_agent_index], 2, 0)          # :171, :190-194
                loss = stg.l2_loss(out['loc'], tbatch.y, out['reg_mask']) + stg.diff_bce_loss(diff_in, diff_out)   # loss_weights [1, 1]
                loss.backward()
                if world > 1:
                    pe = bk_enc.all_reduce_mean_async()
                    pending.pop('dec').wait()
                    pe.wait()
                else:
                    bk_enc.all_reduce_mean()                                   # single rank: only polls the backward status words
                opt.step()
                return loss

            for i in range(3):
                train_step(i)
            k_train = max(3, min(args.steps, 10))
            n0 = ops.LAUNCHES['n']
            ms = timed(train_step, k_train) / k_train
            nl = (ops.LAUNCHES['n'] - n0) // k_train
            loss_v = float(train_step(99).detach())
            in_sync = None
            if world > 1:       # every rank started from the same weights: with the bucket all-reduces in place they stay bit-identical
                chk = torch.stack([p_.detach().double().sum() for p_ in dec_params + enc_params]).sum().reshape(1)
                hi, lo = chk.clone(), chk.clone()
                dist.all_reduce(hi, op=dist.ReduceOp.MAX)
                dist.all_reduce(lo, op=dist.ReduceOp.MIN)
                in_sync = bool((hi == lo).item())
            twork = Et * ENC_STEPS + Mt * DEC_STEPS
            gs = global_scenes or world * scenes_per_gpu
            out = {"scenes_per_gpu": scenes_per_gpu, "global_scenes": gs, "ms_per_step": ms, "steps": k_train,
                   "agent_steps_per_s_fwd_bwd": world * twork / (ms * 1e-3), "scenes_per_s_fwd_bwd": world * scenes_per_gpu / (ms * 1e-3),
                   "allreduce_floats": (bk_dec.numel + bk_enc.numel) if world > 1 else 0, "gpu_launches_per_step": nl, "loss": loss_v,
                   "params_in_sync_across_ranks": in_sync}
            if not args.no_graph and (world == 1 or not args.no_graph_ddp):
                # the same step captured ONCE into a CUDA graph (~35 launches of this library, ~250 small aten launches and the Python between them become one replay): the Philox key
                # lives in device memory (trajsde_b200.set_device_seed) and is bumped inside the graph, so every replay draws fresh noise
                word = torch.zeros(1, dtype=torch.int64, device=dev)
                tb.set_device_seed(word)
                opt_eager, opt = opt, torch.optim.AdamW(dec_params + enc_params, lr=1e-3, weight_decay=7e-4, fused=True, capturable=True)

                def graph_body():
                    l_ = train_step(0)
                    word.add_(1)
                    return l_

                try:
                    side = torch.cuda.Stream(dev)
                    side.wait_stream(torch.cuda.current_stream(dev))
                    with torch.cuda.stream(side):
                        for _ in range(3):
                            graph_body()
                    torch.cuda.current_stream(dev).wait_stream(side)
                    graph = torch.cuda.CUDAGraph()
                    # thread_local: NCCL's watchdog thread may query its events while this thread captures (N > 1: the two bucket
                    # all-reduces and their side-stream fork/join are captured with the step)
                    with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                        loss_static = graph_body()
                    for _ in range(2):
                        graph.replay()
                    ms_g = timed(lambda i: graph.replay(), k_train) / k_train
                    out.update(ms_per_step_cuda_graph=ms_g, scenes_per_s_fwd_bwd_cuda_graph=world * scenes_per_gpu / (ms_g * 1e-3),
                               loss_cuda_graph=float(loss_static.detach()))
                    del graph, loss_static
                except RuntimeError as e:                      # the eager numbers above stand; say why the replay is missing
                    out.update(ms_per_step_cuda_graph=None, cuda_graph_error=str(e).splitlines()[0][:200])
                tb.set_device_seed(None)
                opt = opt_eager
            for p_ in dec_params + enc_params:
                p_.grad = None
            del bk_dec, bk_enc, opt, tbatch, b
            torch.cuda.empty_cache()
            return out

        train = {"note": "real step of the SDE path and its direct consumers: fused encoder recurrence (enc_fwd_tc_kernel / trajsde_enc_bwd), eos gather, "
                         "fused aggr_embed, decoder solve (euler_fwd_tc_kernel / euler_bwd_tc_kernel with zero-row skipping), fused heads fwd + bwd, "
                         "L2 (winner-takes-all) + DiffBCE kernels, in-kernel Philox noise, mixed nuScenes/Argoverse scenes, NCCL all-reduce of two flat "
                         "gradient buckets launched asynchronously on a side stream (the decoder's under the encoder backward), AdamW",
                 "cfg2_reference_batch": measure_train(args.train_scenes),          # BASELINE configs[2]: yml:106 batch 128
                 "cfg3_scene_sharded": measure_train(args.scenes)}                  # BASELINE configs[3] weak: 1024 scenes per GPU (8192 at N=8)
        if args.strong_scenes and args.strong_scenes % world == 0:                    # BASELINE configs[3] strong: 8192 scenes in total at every N
            train["cfg3_strong_8192"] = measure_train(args.strong_scenes // world, global_scenes=args.strong_scenes)

    # parity is NOT checked here: the oracle is test infrastructure (tests/test_full_size_gpu.py checks this very workload against it);
    # in this file only the cpu_baseline / reference-arm legs execute oracle code, as the thing being timed on the host cores
    parity = {"checked_by": "tests/test_full_size_gpu.py, tests/test_euler_fwd_gpu.py (pytest -m gpu)", "mode": mode}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_gbs, bf16_tf, peak_src = peaks()
    traffic = None
    # ncu-measured DRAM bytes per launch of the dominant kernel: the latest (by round tag) profiles/r*_traffic.json captured for this workload (written by
    # tools/gpu_visit.sh from an `ncu --set full` capture of the same source tree; its "commit" field says which)
    traffic_src = None
    try:
        import glob
        for fp in sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r*_traffic.json')), reverse=True):
            t = json.load(open(fp)).get('euler_fwd_tc_kernel<1,0>')
            if t and t['rows'] == M and t['steps'] == DEC_STEPS:
                traffic, traffic_src = t['dram_bytes'], os.path.basename(fp)
                break
    except (OSError, KeyError, ValueError):
        pass
    T = sched_d.n_outputs + 1
    dec_bytes_fixed = M * 256 * (1 + T + DEC_STEPS)             # y0 in + ys (incl. ys[0]) out + dW in
    dec_bytes_philox = M * 256 * (1 + T)
    ach = dec_bytes_fixed / (dec_ms_avg * 1e-3) / 1e9
    ach_p = dec_bytes_philox / (dec_ms_philox * 1e-3) / 1e9
    flops = 41856.0 * M * DEC_STEPS
    res_json = {
        "metric": METRIC, "value": world * work / (ms_fixed / args.steps * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_fixed / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate+state" if mode == 'tc_f16' else "f32",
        "data": "synthetic",
        "config": {"workload": f"BASELINE configs[1] per GPU: sdesepenc+sdedec forward, {args.scenes} scenes x {args.agents} agents "
                               f"(encoder {E} rows x 21 steps + GRU jump, decoder {M} rows x 61 steps -> 60 outputs), caller-supplied dW",
                   "kernel_mode": mode, "scenes_per_gpu": args.scenes, "parallelism": f"scene-sharded dp{world}, no forward collective",
                   "cache": "inputs larger than L2 (dW 3.2 GB + ys 3.2 GB per step vs 126 MB L2)",
                   "e2e_note": "e2e uses bm=None (in-kernel Philox, like the reference's BrownianInterval default); every step copies all "
                               "inputs from pinned host memory (decoder y0 first, in row slices solved as they land; encoder inputs behind "
                               "them, the AA-encoder output as fp16), runs the fused heads on the device and copies the decoder's RESULT "
                               "out['loc'] = cat(loc, scale) [10 N, 60, 4] and the encoder's final latents back (HostFedSdePath.run_batch)"},
        "scenes_per_s": world * args.scenes / (ms_fixed / args.steps * 1e-3),
        "philox": {"value": world * work / (ms_philox / args.steps * 1e-3), "ms_per_step": ms_philox / args.steps,
                   "decoder_ms": dec_ms_philox, "decoder_agent_steps_per_s": M * DEC_STEPS / (dec_ms_philox * 1e-3),
                   "roofline_frac_hbm": ach_p / hbm_gbs},
        "decoder": {"ms": dec_ms_avg, "agent_steps_per_s": M * DEC_STEPS / (dec_ms_avg * 1e-3), "rows": M, "steps": DEC_STEPS},
        "encoder": {"ms": ms_fixed / args.steps - dec_ms_avg, "rows": E, "steps": ENC_STEPS,
                    "agent_steps_per_s": E * ENC_STEPS / ((ms_fixed / args.steps - dec_ms_avg) * 1e-3),
                    "note": "one fused kernel: 21 x [Euler step of the dual-diffusion SDE + GRU jump] (enc_fwd_tc_kernel)"},
        "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_gbs, "unit": "GB/s", "frac": ach / hbm_gbs, "traffic": traffic,
                     "traffic_source": traffic_src,
                     "kernel": "euler_fwd_tc_kernel (decoder solve)", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": dec_bytes_fixed,
                     "tensor_frac_of_bf16_peak": flops / (dec_ms_avg * 1e-3) / 1e12 / bf16_tf,
                     "sfu_note": "informational third ceiling: 257 MUFU ops per agent-step at the measured 16/clk/SM"},
        "e2e": {"value": world * e2e_work / (ms_e2e / args.steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps, "micro_batches": n_chunks,
                "host_cores_bound_per_rank": cores_bound, "result": "decoder out['loc'] [M,60,4] fp32 + encoder final latents [N',64]",
                "host_link": link},
        "heads": heads,
        "train": train,
        "gpu_launches": launches,
        "clocks": clocks,
        "parity": parity,
    }
    if world == 1 and not args.no_cpu_baseline:
        res_json["cpu_baseline"] = cpu_baseline()
    print(json.dumps(res_json), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
