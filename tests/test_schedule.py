"""Step-schedule known answers — product schedule vs oracle replay vs reference Brownian queries."""
import struct

import numpy as np
import pytest
import torch

from oracle import sde_oracle as so
from trajsde_b200.schedule import encoder_schedule, encoder_time_pairs, euler_schedule


def f32hex(x):
    return struct.pack('>f', float(x)).hex()


def test_decoder_schedule_known_answers():
    s = euler_schedule(torch.linspace(0, 6, 61), 0.1)
    assert s.n_steps == 61 and s.n_outputs == 60
    assert f32hex(s.t0[0]) == '00000000' and f32hex(s.h[0]) == '3dcccccd'
    assert f32hex(s.t0[6]) == '3f19999a' and f32hex(s.h[6]) == '3dccccd0'
    assert f32hex(s.t0[7]) == '3f333334'
    assert f32hex(s.t0[23]) == '40133333' and f32hex(s.h[23]) == '3dccccc0'
    assert f32hex(s.t0[24]) == '40199999'
    assert f32hex(s.t0[59]) == '40bcccc6'
    assert f32hex(s.t0[60]) == '40bffff9' and f32hex(s.h[60]) == '36600000'
    # every interval takes 1 step except output 24 (2 steps): k_j = j-1 (j<=23), j (j>=24)
    j = np.arange(1, 61)
    assert np.array_equal(s.out_k, np.where(j <= 23, j - 1, j))
    assert all((s.w0[i], s.w1[i]) == (0.0, 1.0) for i in range(6))
    assert f32hex(s.w0[6]) == '351ffffe' and f32hex(s.w1[6]) == '3f7ffff6'
    assert f32hex(s.w0[19]) == '3620000a' and f32hex(s.w1[19]) == '3f7fffd8'
    assert f32hex(s.w0[58]) == '3f7ffdd0' and f32hex(s.w1[58]) == '380c0009'
    assert (s.w0[59], s.w1[59]) == (0.0, 1.0)
    assert np.all(s.w0 + s.w1 == 1.0)


@pytest.mark.parametrize('F,S', [(10, 10), (20, 20), (30, 31), (50, 51), (60, 61), (100, 100), (200, 200)])
def test_product_schedule_matches_oracle_and_reference_queries(F, S, golden_schedule):
    ts = torch.linspace(0, 0.1 * F, F + 1)
    s = euler_schedule(ts, 0.1)
    o = so.euler_schedule_ref(ts, 0.1)
    assert s.n_steps == S
    for k in ('t0', 'h', 'out_k', 'w0', 'w1'):
        assert np.array_equal(getattr(s, k), o[k].numpy()), k
        assert np.array_equal(getattr(s, k), golden_schedule[f'F{F}/{k}']), k
    # (ta, tb) the reference solver queried its Brownian motion with (fixture made by tests/golden/make_golden.py)
    q = golden_schedule[f'F{F}/queries']
    assert np.array_equal(q[:, 0], s.t0)
    assert np.array_equal((q[:, 1] - q[:, 0]).astype(np.float32), s.h)
    ob = s.out_begin()
    assert ob[0] == 0 and ob[-1] == F and np.all(np.diff(ob) >= 0)
    assert np.array_equal(np.repeat(np.arange(S), np.diff(ob)), s.out_k)


def test_zero_step_interval_exists_for_F100():
    s = euler_schedule(torch.linspace(0, 10, 101), 0.1)
    counts = np.diff(s.out_begin())
    assert counts.max() == 2 and counts.min() == 0      # one step completes two outputs, one interval takes two steps


def test_encoder_schedule(golden_schedule):
    e = encoder_schedule()
    assert e.n_steps == 21
    assert np.array_equal(e.t0, golden_schedule['enc/t0']) and np.array_equal(e.h, golden_schedule['enc/h'])
    assert abs(float(e.t0[0]) + 0.00999999978) < 1e-9 and abs(float(e.h[0]) - 0.00999999978) < 1e-9
    assert [t for _, _, t in encoder_time_pairs()] == list(range(20, -1, -1))
    assert np.array_equal(np.array([[a, b] for a, b, _ in encoder_time_pairs()], np.float32), golden_schedule['enc/pairs'])
    o = so.encoder_time_pairs_ref()
    assert all(float(a) == float(x) and float(b) == float(y) for (a, b, _), (x, y, _) in zip(encoder_time_pairs(), o))


def test_schedule_errors():
    with pytest.raises(ValueError):
        euler_schedule(torch.tensor([0.0, 1.0, 1.0]), 0.1)
    with pytest.raises(ValueError):
        euler_schedule(torch.tensor([0.0]), 0.1)
    with pytest.raises(ValueError):
        euler_schedule(torch.tensor([0.0, 1.0]), 0.0)
