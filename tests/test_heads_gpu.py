"""Fused decoder heads (trajsde_heads_fwd) against the oracle's restatement of the reference heads
(dec_hivt_nusargo_sde.py:50-61, 96-99) and the committed reference fixture.

Tolerance (tc_f16: fp16 operands for the two 64x64 layers, fp32 LayerNorm / ReLU / 64->2 projections): atol 2e-2 + rtol 2e-2 on
head outputs of magnitude O(1); measured ~3e-3."""
import pytest
import torch
import torch.nn as nn

import trajsde_b200 as tb
from conftest import sub
from helpers import DecoderSDE, init_like_reference
from oracle import sde_oracle as so
from trajsde_b200 import heads as hd

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = dict(atol=2e-2, rtol=2e-2)


def make_head(seed):
    h = nn.Sequential(nn.Linear(64, 64), nn.LayerNorm(64), nn.ReLU(inplace=True), nn.Linear(64, 2))
    init_like_reference(h, seed)
    g = torch.Generator().manual_seed(seed + 100)
    with torch.no_grad():
        h[1].weight.copy_(1.0 + 0.2 * torch.randn(64, generator=g))
        h[1].bias.copy_(0.1 * torch.randn(64, generator=g))
    return h


def params_of(head):
    return {k: v.detach().cpu() for k, v in head.state_dict().items()}


def test_heads_vs_reference_fixture(golden_decoder):
    d = golden_decoder
    loc_h, sc_h = make_head(0), make_head(1)
    loc_h.load_state_dict(sub(d, 'head'))
    sc_h.load_state_dict(sub(d, 'scale_head'))
    loc_h, sc_h = loc_h.to(DEV), sc_h.to(DEV)
    ys = torch.from_numpy(d['ys']).to(DEV)
    sol_y = ys[1:].permute(1, 0, 2)                           # the reference's view (dec…sde.py:88): time-major storage
    with torch.no_grad():
        loc, scale_raw = hd.decoder_heads(loc_h, sc_h, sol_y)
    scale = torch.nn.functional.elu(scale_raw, alpha=1.0) + 1.0 + float(d['min_scale'])
    assert torch.allclose(loc.cpu(), torch.from_numpy(d['loc']), **TOL)
    assert torch.allclose(scale.cpu(), torch.from_numpy(d['scale']), **TOL)


@pytest.mark.parametrize('rows,T,layout', [(1, 1, 'rows'), (127, 3, 'rows'), (300, 60, 'rows'), (300, 60, 'time'), (2049, 7, 'time'),
                                           (0, 60, 'rows')])
def test_heads_vs_oracle_shapes_and_layouts(rows, T, layout):
    loc_h, sc_h = make_head(3).to(DEV), make_head(4).to(DEV)
    g = torch.Generator().manual_seed(rows * 131 + T)
    x = torch.randn(rows, T, 64, generator=g) * 3.0
    if layout == 'rows':                                      # rows_major solver output: [rows, T+1, 64] storage, t = 0 dropped
        store = torch.zeros(rows, T + 1, 64)
        store[:, 1:] = x
        xd = store.to(DEV)[:, 1:]
    else:                                                     # time-major storage [T+1, rows, 64] -> [1:].permute(1, 0, 2)
        store = torch.zeros(T + 1, rows, 64)
        store[1:] = x.permute(1, 0, 2)
        xd = store.to(DEV)[1:].permute(1, 0, 2)
    with torch.no_grad():
        loc, scale_raw = hd.decoder_heads(loc_h, sc_h, xd)
    assert loc.shape == (rows, T, 2) and scale_raw.shape == (rows, T, 2)
    ref_loc = so.decoder_loc_head_ref(params_of(loc_h), x)
    ref_sc = so.decoder_loc_head_ref(params_of(sc_h), x)
    assert torch.allclose(loc.cpu(), ref_loc, **TOL)
    assert torch.allclose(scale_raw.cpu(), ref_sc, **TOL)


def test_single_head_and_errors():
    loc_h = make_head(5).to(DEV)
    x = torch.randn(200, 5, 64, generator=torch.Generator().manual_seed(5)).to(DEV)
    with torch.no_grad():
        loc, none = hd.decoder_heads(loc_h, None, x)
    assert none is None
    assert torch.allclose(loc.cpu(), so.decoder_loc_head_ref(params_of(loc_h), x.cpu()), **TOL)
    with pytest.raises(NotImplementedError):                  # autograd requested: no silent fallback
        hd.decoder_heads(loc_h, None, x.clone().requires_grad_(True))
    with pytest.raises(RuntimeError):
        hd.decoder_heads(loc_h.cpu(), None, x.cpu())


def test_install_heads_drop_in_pair():
    """The reference forward calls self.decoder(sol_y) then self.scale(sol_y): one fused launch serves both; training goes to
    the original nn.Sequential."""
    class Dec(nn.Module):
        def __init__(self):
            super().__init__()
            self.decoder, self.scale = make_head(7), make_head(8)

    dec = Dec().to(DEV)
    x = torch.randn(130, 60, 64, generator=torch.Generator().manual_seed(9)).to(DEV)
    with torch.no_grad():
        want_loc, want_sc = dec.decoder(x), dec.scale(x)
    saved = hd.install_heads(dec)
    from trajsde_b200 import ops
    n0 = ops.LAUNCHES['n']
    with torch.no_grad():
        loc = dec.decoder(x)
        sc = dec.scale(x)
    assert ops.LAUNCHES['n'] - n0 == 2                         # pack + one fused kernel for both heads
    assert torch.allclose(loc, want_loc, **TOL) and torch.allclose(sc, want_sc, **TOL)
    xg = x.clone().requires_grad_(True)                        # autograd -> reference PyTorch heads
    out = dec.decoder(xg)
    out.sum().backward()
    assert xg.grad is not None and torch.allclose(out, want_loc, atol=1e-5, rtol=1e-5)
    hd.uninstall_heads(saved)
    assert 'forward' not in dec.decoder.__dict__
    with torch.no_grad():
        assert torch.equal(dec.decoder(x), want_loc)


def test_heads_on_solver_output_rows_major():
    """End of the decoder as the reference runs it: rows_major solve -> [1:].permute(1,0,2) view -> fused heads."""
    sde = init_like_reference(DecoderSDE(), seed=2).to(DEV)
    loc_h, sc_h = make_head(11).to(DEV), make_head(12).to(DEV)
    y0 = torch.relu(torch.randn(260, 64, generator=torch.Generator().manual_seed(3))).to(DEV)
    ts = torch.linspace(0, 6, 61)
    with torch.no_grad():
        ys = tb.sdeint(sde, y0, ts, dt=0.1, method='euler', mode='tc_f16', seed=5, rows_major=True)
        sol_y = ys[1:].permute(1, 0, 2)
        loc, sc = hd.decoder_heads(loc_h, sc_h, sol_y)
        assert torch.allclose(loc, loc_h(sol_y), **TOL) and torch.allclose(sc, sc_h(sol_y), **TOL)
