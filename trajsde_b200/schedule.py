"""Host-side step schedule of the fixed-step Euler solve.

The reference advances time with float32 0-d CPU tensors: ``next_t = min(curr_t + dt, ts[-1])`` inside
``while curr_t < out_t`` and linearly interpolates every requested output (models/utils/sdeint.py:340-384; torchsde 0.2.5
``BaseSDESolver.integrate`` is identical).  For ``linspace(0, 6, 61)`` / ``dt=0.1`` that is 61 steps, not 60, with a
3.3e-6 sliver step at the end and interpolation weights != (0, 1).  The schedule depends only on
``(ts, dt)``, so it is computed here once — in numpy float32, which rounds exactly like the reference's float32 tensors —
and handed to the kernels as small device tables.
"""
from dataclasses import dataclass
from functools import lru_cache
from typing import Tuple

import numpy as np
import torch


@dataclass(frozen=True)
class EulerSchedule:
    t0: np.ndarray       # [S] float32 step start times
    h: np.ndarray        # [S] float32 step sizes (t1 - t0)
    sin_t0: np.ndarray   # [S] float32
    cos_t0: np.ndarray   # [S] float32
    out_k: np.ndarray    # [T-1] int32: output j+1 interpolates between Y[out_k[j]] and Y[out_k[j]+1]
    w0: np.ndarray       # [T-1] float32
    w1: np.ndarray       # [T-1] float32

    @property
    def n_steps(self) -> int:
        return int(self.h.shape[0])

    @property
    def n_outputs(self) -> int:
        return int(self.out_k.shape[0])

    def step_tab(self) -> np.ndarray:
        """[S,4] float32 rows (t0, h, sin t0, cos t0): TrajsdeSchedule.step_tab."""
        return np.ascontiguousarray(np.stack([self.t0, self.h, self.sin_t0, self.cos_t0], axis=1), dtype=np.float32)

    def out_begin(self) -> np.ndarray:
        """[S+1] int32 CSR offsets of the outputs grouped by the step that completes them."""
        counts = np.bincount(self.out_k, minlength=self.n_steps).astype(np.int64)
        return np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)

    def out_w(self) -> np.ndarray:
        return np.ascontiguousarray(np.stack([self.w0, self.w1], axis=1), dtype=np.float32)


def _sincos_f32(t0: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    # The reference evaluates torch.sin/cos on a float32 tensor holding float(t) (dec_hivt_nusargo_sde.py:124-126).
    t = torch.from_numpy(np.ascontiguousarray(t0, dtype=np.float32))
    return torch.sin(t).numpy().copy(), torch.cos(t).numpy().copy()


def euler_schedule(ts, dt: float) -> EulerSchedule:
    """Replay of the reference time loop in float32.  ``ts``: 1-D sequence/array/tensor of output times (>= 2 entries,
    strictly increasing); ``dt``: python float step size (rounded to float32 when added, like torch does)."""
    if torch.is_tensor(ts):
        ts = ts.detach().to('cpu', torch.float32).numpy()
    ts = np.ascontiguousarray(ts, dtype=np.float32)
    if ts.ndim != 1 or ts.shape[0] < 2:
        raise ValueError("`ts` must be 1-D with at least two entries")
    if not np.all(ts[1:] > ts[:-1]):
        raise ValueError("Evaluation times `ts` must be strictly increasing.")  # sdeint.py:876-877
    dt32 = np.float32(dt)
    if not (dt32 > 0):
        raise ValueError("`dt` must be positive")
    return _euler_schedule_cached(ts.tobytes(), float(dt32))


@lru_cache(maxsize=64)
def _euler_schedule_cached(ts_bytes: bytes, dt: float) -> EulerSchedule:
    ts = np.frombuffer(ts_bytes, dtype=np.float32)
    dt32 = np.float32(dt)
    t_end = ts[-1]
    curr = prev = ts[0]
    t0s, hs, out_k, w0s, w1s = [], [], [], [], []
    for out_t in ts[1:]:
        while curr < out_t:
            nxt = np.float32(curr + dt32)
            if t_end < nxt:
                nxt = t_end
            prev = curr
            t0s.append(curr)
            hs.append(np.float32(nxt - curr))
            curr = nxt
            if len(t0s) > 10_000_000:
                raise ValueError("schedule: more than 1e7 steps (dt too small for ts?)")
        if not t0s:
            raise ValueError("schedule: first output interval takes no step")
        out_k.append(len(t0s) - 1)
        den = np.float32(curr - prev)
        w0s.append(np.float32(np.float32(curr - out_t) / den))
        w1s.append(np.float32(np.float32(out_t - prev) / den))
    t0 = np.asarray(t0s, dtype=np.float32)
    sin_t0, cos_t0 = _sincos_f32(t0)
    return EulerSchedule(t0=t0, h=np.asarray(hs, dtype=np.float32), sin_t0=sin_t0, cos_t0=cos_t0,
                         out_k=np.asarray(out_k, dtype=np.int32), w0=np.asarray(w0s, dtype=np.float32),
                         w1=np.asarray(w1s, dtype=np.float32))


def encoder_time_pairs(max_past_t: float = 2.0, historical_steps: int = 21):
    """(prev_t, t_i, data_slot) per iteration of the encoder loop for run_backwards=True, float32, exactly as
    enc_hivt_nusargo_sde_sep2.py:128-135,175-179 builds them: SDE time runs 0 -> max_past_t while data slots run
    historical_steps-1 -> 0; the first pair is (pts[-1] - 0.01, pts[-1])."""
    pts = (-1 * torch.linspace(-max_past_t, 0, historical_steps)).numpy()
    prev_t, t_i = np.float32(pts[-1] - np.float32(0.01)), pts[-1]
    pairs = []
    order = list(reversed(range(historical_steps)))
    for idx, t in enumerate(order):
        pairs.append((np.float32(prev_t), np.float32(t_i), t))
        if idx + 1 < historical_steps:
            prev_t, t_i = pts[t], pts[t - 1]
    return pairs


def encoder_schedule(max_past_t: float = 2.0, historical_steps: int = 21, dt: float = 0.1) -> EulerSchedule:
    """The 21 one-step ``sdeint_dual`` calls of the encoder concatenated into one schedule (step idx = loop iteration).
    Raises if any call would take more than one Euler step (never the case for the reference grid)."""
    t0s, hs = [], []
    for prev_t, t_i, _ in encoder_time_pairs(max_past_t, historical_steps):
        s = euler_schedule(np.array([prev_t, t_i], dtype=np.float32), dt)
        if s.n_steps != 1 or s.w0[0] != 0.0 or s.w1[0] != 1.0:
            raise NotImplementedError("encoder grid with more than one Euler step per observation is not supported")
        t0s.append(s.t0[0])
        hs.append(s.h[0])
    t0 = np.asarray(t0s, dtype=np.float32)
    sin_t0, cos_t0 = _sincos_f32(t0)
    n = len(t0s)
    return EulerSchedule(t0=t0, h=np.asarray(hs, dtype=np.float32), sin_t0=sin_t0, cos_t0=cos_t0,
                         out_k=np.arange(n, dtype=np.int32), w0=np.zeros(n, dtype=np.float32),
                         w1=np.ones(n, dtype=np.float32))
