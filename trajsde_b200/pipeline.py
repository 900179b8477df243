"""Host-fed execution of the SDE path: micro-batches that live in pinned host memory are copied in, solved and copied out
on three CUDA streams so that PCIe traffic overlaps the fused kernels (rows of different scenes are independent, so a batch
may be processed as any number of scene chunks — results are identical row for row)."""
from typing import List, Optional, Sequence

import torch

from . import encoder as enc_mod
from .solver import sdeint
from .synthetic import SdeBatch

_KEYS = ('enc_h0', 'aa_out', 'actors_mask', 'nus_mask', 'dec_y0')


class HostFedSdePath:
    """Forward pass (encoder recurrence + decoder solve) over a list of pinned host micro-batches.

    `run()` enqueues, for every chunk c: H2D of its inputs (copy-in stream) -> fused encoder + decoder kernels (caller's
    stream) -> D2H of the final encoder latents and final decoder latents (copy-out stream), and returns after the last
    D2H has landed, i.e. when the host can read the results."""

    def __init__(self, enc_sde, gru, dec_sde, device, ts_dec: torch.Tensor, dt: float = 0.1, mode: Optional[str] = None):
        self.enc_sde, self.gru, self.dec_sde = enc_sde, gru, dec_sde
        self.device, self.ts_dec, self.dt, self.mode = torch.device(device), ts_dec, dt, mode
        self.copy_in = torch.cuda.Stream(self.device)
        self.copy_out = torch.cuda.Stream(self.device)
        self._dev: List[dict] = []
        self._ready: List[torch.cuda.Event] = []
        self._done: List[torch.cuda.Event] = []
        self._drained: List[torch.cuda.Event] = []

    def _ensure_buffers(self, chunks: Sequence[SdeBatch]):
        if len(self._dev) == len(chunks) and all(self._dev[i]['dec_y0'].shape == chunks[i].dec_y0.shape for i in range(len(chunks))):
            return
        self._dev = [{k: torch.empty_like(getattr(c, k), device=self.device) for k in _KEYS} for c in chunks]
        self._ready = [torch.cuda.Event() for _ in chunks]
        self._done = [torch.cuda.Event() for _ in chunks]
        self._drained = [torch.cuda.Event() for _ in chunks]
        for e in self._done + self._drained:
            e.record(torch.cuda.current_stream(self.device))

    def run(self, chunks: Sequence[SdeBatch], out_enc: Sequence[torch.Tensor], out_dec: Sequence[torch.Tensor], seed: int = 0,
            enc_row_offsets: Optional[Sequence[int]] = None, dec_row_offsets: Optional[Sequence[int]] = None) -> None:
        self._ensure_buffers(chunks)
        cur = torch.cuda.current_stream(self.device)
        keep = []
        with torch.no_grad():
            for c, hb in enumerate(chunks):
                with torch.cuda.stream(self.copy_in):
                    self.copy_in.wait_event(self._done[c])           # previous use of these device buffers has finished
                    for k in _KEYS:
                        self._dev[c][k].copy_(getattr(hb, k), non_blocking=True)
                    self._ready[c].record(self.copy_in)
                cur.wait_event(self._ready[c])
                d = self._dev[c]
                lat, _ = enc_mod.encoder_recurrence(self.enc_sde, self.gru, d['enc_h0'], d['aa_out'], d['actors_mask'], d['nus_mask'],
                                                    dt=self.dt, seed=seed + 2 * c, mode=self.mode,
                                                    row_offset=0 if enc_row_offsets is None else enc_row_offsets[c])
                ys = sdeint(self.dec_sde, d['dec_y0'], self.ts_dec, dt=self.dt, dt_min=self.dt, rtol=1e-3, atol=1e-3, method='euler',
                            mode=self.mode, seed=seed + 2 * c + 1, row_offset=0 if dec_row_offsets is None else dec_row_offsets[c])
                self._done[c].record(cur)
                keep.append((lat, ys))
                with torch.cuda.stream(self.copy_out):
                    self.copy_out.wait_event(self._done[c])
                    out_enc[c].copy_(lat[-1], non_blocking=True)
                    out_dec[c].copy_(ys[-1], non_blocking=True)
        self.copy_out.synchronize()                                   # the host reads the results now
        del keep

    def run_batch(self, hb: SdeBatch, out_enc: torch.Tensor, out_dec: torch.Tensor, seed: int = 0, dec_chunks: int = 4,
                  enc_row_offset: int = 0, dec_row_offset: int = 0, heads=None, min_scale: float = 0.001,
                  aa_out_half: Optional[torch.Tensor] = None) -> None:
        """One pinned host batch, copies ordered by what the kernels can start on first: the decoder's `dec_y0` goes over
        in `dec_chunks` row slices, each solved as soon as it lands (Philox streams are keyed by global row, so the slices
        reproduce the unsliced solve exactly); the encoder's inputs (two thirds of the bytes) stream in behind them while
        those solves run, and the latency-bound recurrence kernel then runs once over all rows.

        What comes back to the host is the stage's RESULT, not an internal state: with ``heads=(loc_head, scale_head)`` every
        decoder slice goes through the fused heads on the device and ``out_dec`` [M, 60, 4] receives ``cat(loc, elu(scale) + 1 +
        min_scale)`` — the decoder's ``out['loc']`` (dec…sde.py:95-100), 16 B per (row, t) instead of the 256 B latent; without heads
        the final decoder latents [M, 64] (round-1 behaviour).  ``out_enc`` [N', 64] receives the encoder's final latents.
        ``aa_out_half``: the AA-encoder output as fp16 on the host (the kernel rounds it to fp16 MMA operands anyway): one third fewer
        H2D bytes; converted to fp32 on the device."""
        dev, cur = self.device, torch.cuda.current_stream(self.device)
        M = hb.dec_y0.shape[0]
        if not self._dev or self._dev[0]['dec_y0'].shape != hb.dec_y0.shape or self._dev[0]['aa_out'].shape != hb.aa_out.shape:
            self._dev = [{k: torch.empty_like(getattr(hb, k), device=dev) for k in _KEYS}]
            self._dev[0]['aa_half'] = torch.empty(hb.aa_out.shape, dtype=torch.float16, device=dev)
            self._done = [torch.cuda.Event()]
            self._done[0].record(cur)
        d = self._dev[0]
        bounds = [M * c // dec_chunks for c in range(dec_chunks + 1)]
        ready = [torch.cuda.Event() for _ in range(dec_chunks + 1)]
        with torch.cuda.stream(self.copy_in):
            self.copy_in.wait_event(self._done[0])                   # previous call has finished with the device buffers
            for c in range(dec_chunks):
                d['dec_y0'][bounds[c]:bounds[c + 1]].copy_(hb.dec_y0[bounds[c]:bounds[c + 1]], non_blocking=True)
                ready[c].record(self.copy_in)
            for k in ('enc_h0', 'actors_mask', 'nus_mask'):
                d[k].copy_(getattr(hb, k), non_blocking=True)
            if aa_out_half is not None:
                d['aa_half'].copy_(aa_out_half, non_blocking=True)
            else:
                d['aa_out'].copy_(hb.aa_out, non_blocking=True)
            ready[dec_chunks].record(self.copy_in)
        keep = []
        with torch.no_grad():
            for c in range(dec_chunks):
                cur.wait_event(ready[c])
                lo, hi = bounds[c], bounds[c + 1]
                ys = sdeint(self.dec_sde, d['dec_y0'][lo:hi], self.ts_dec, dt=self.dt, dt_min=self.dt, rtol=1e-3, atol=1e-3,
                            method='euler', mode=self.mode, seed=seed + 1, row_offset=dec_row_offset + lo, rows_major=heads is not None)
                if heads is not None:
                    from .heads import decoder_heads_from_solution
                    res, _ = decoder_heads_from_solution(heads[0], heads[1], ys, cat_min_scale=min_scale)   # dec…sde.py:95-100 in one launch
                else:
                    res = ys[-1]
                ev = torch.cuda.Event()
                ev.record(cur)
                keep.append((ys, res))
                with torch.cuda.stream(self.copy_out):
                    self.copy_out.wait_event(ev)
                    out_dec[lo:hi].copy_(res, non_blocking=True)
            cur.wait_event(ready[dec_chunks])
            aa = d['aa_half'].float() if aa_out_half is not None else d['aa_out']
            lat, _ = enc_mod.encoder_recurrence(self.enc_sde, self.gru, d['enc_h0'], aa, d['actors_mask'], d['nus_mask'],
                                                dt=self.dt, seed=seed, mode=self.mode, row_offset=enc_row_offset)
            self._done[0].record(cur)
            keep.append(lat)
            with torch.cuda.stream(self.copy_out):
                self.copy_out.wait_event(self._done[0])
                out_enc.copy_(lat[-1], non_blocking=True)
        self.copy_out.synchronize()                                   # the host reads the results now
        del keep
