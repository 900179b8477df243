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
    xg = x.clone().requires_grad_(True)                       # autograd requested: the fused backward serves it (single head too)
    out, _ = hd.decoder_heads(loc_h, None, xg)
    out.sum().backward()
    assert xg.grad is not None and torch.isfinite(xg.grad).all() and loc_h[0].weight.grad is not None
    with pytest.raises(RuntimeError):
        hd.decoder_heads(loc_h.cpu(), None, x.cpu())


def test_install_heads_drop_in_pair():
    """The reference forward calls self.decoder(sol_y) then self.scale(sol_y): one fused launch serves both, with and without autograd."""
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
    xg = x.clone().requires_grad_(True)                        # autograd: one differentiable fused node serves both calls
    n0 = ops.LAUNCHES['n']
    out, out_sc = dec.decoder(xg), dec.scale(xg)
    assert ops.LAUNCHES['n'] - n0 == 2 and torch.allclose(out, want_loc, **TOL) and torch.allclose(out_sc, want_sc, **TOL)
    (out.sum() + 0.5 * out_sc.sum()).backward()
    ref = x.clone().requires_grad_(True)
    (torch.nn.Sequential.forward(dec.decoder, ref).sum() + 0.5 * torch.nn.Sequential.forward(dec.scale, ref).sum()).backward()
    assert (xg.grad - ref.grad).abs().max() <= 1e-4 * ref.grad.abs().max()
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


def _fp64_head_grads(heads, x, cots):
    """fp64 autograd through the oracle's restatement of the heads: dL/dx and dL/dparam for L = sum_h <out_h, cot_h>."""
    xd = x.double().clone().requires_grad_(True)
    P = [{k: v.double().clone().requires_grad_(True) for k, v in params_of(h).items()} for h in heads]
    loss = sum((so.decoder_loc_head_ref(p, xd) * c.double()).sum() for p, c in zip(P, cots) if c is not None)
    leaves = [xd] + [t for p in P for t in p.values()]
    g = torch.autograd.grad(loss, leaves, allow_unused=True)
    return g[0], [dict(zip(p.keys(), g[1 + 6 * i:7 + 6 * i])) for i, p in enumerate(P)]


@pytest.mark.parametrize('rows,T,frac,which', [(50, 7, 1.0, 'both'), (300, 60, 0.1, 'both'), (300, 60, 0.1, 'loc'), (2000, 60, 0.03, 'both'),
                                               (130, 5, 0.0, 'both')])
def test_heads_backward_vs_fp64_autograd(rows, T, frac, which):
    """trajsde_heads_bwd (fp32, active points only) against fp64 autograd of the oracle heads: dL/dsol_y and every head parameter,
    for dense, sparse (winner-takes-all shaped), single-head and empty cotangents.  The forward values that autograd differentiates
    are the tensor-core ones; the backward recomputes the activations in fp32 from the same inputs."""
    loc_h, sc_h = make_head(21).to(DEV), make_head(22).to(DEV)
    g = torch.Generator().manual_seed(rows + T)
    x = torch.randn(rows, T, 64, generator=g) * 2.0
    act = torch.rand(rows, T, generator=g) < frac
    c0 = torch.randn(rows, T, 2, generator=g) * act.unsqueeze(-1)
    c1 = torch.randn(rows, T, 2, generator=g) * (torch.rand(rows, T, generator=g) < frac).unsqueeze(-1) if which == 'both' else None
    gx_ref, gp_ref = _fp64_head_grads([loc_h, sc_h], x, [c0, c1])
    xg = x.to(DEV).requires_grad_(True)
    o0, o1 = hd.decoder_heads(loc_h, sc_h, xg)
    loss = (o0 * c0.to(DEV)).sum() + ((o1 * c1.to(DEV)).sum() if c1 is not None else 0.0)
    loss.backward()
    if frac == 0.0:
        assert float(xg.grad.abs().max()) == 0.0 and all(float(p_.grad.abs().max()) == 0.0 for p_ in loc_h.parameters())
        return
    assert float((xg.grad.double().cpu() - gx_ref).abs().max()) <= 1e-4 * float(gx_ref.abs().max())
    for h, head in enumerate((loc_h, sc_h)):
        for k, p_ in head.state_dict(keep_vars=True).items():
            r = gp_ref[h][k]
            if r is None or float(r.abs().max()) == 0.0:
                assert p_.grad is None or float(p_.grad.abs().max()) == 0.0, (h, k)
                continue
            e = float((p_.grad.double().cpu() - r).abs().max() / r.abs().max())
            assert e < 1e-4, (h, k, e)


@pytest.mark.parametrize('rows,T,frac', [(50, 7, 1.0), (300, 60, 0.1), (2000, 60, 0.03), (129, 1, 1.0)])
def test_fused_result_cat4_forward_and_backward(rows, T, frac):
    """TRAJSDE_HEADS_FLAG_CAT4: the heads kernel writes the stage's result out['loc'] = cat(loc, elu(scale) + 1 + min_scale) [rows, T, 4]
    (dec…sde.py:98-100) and the heads backward takes dL/d out['loc'] (scale channels through the ELU derivative).  Forward: the loc
    channels are bit-identical to the two-output launch, the scale channels equal torch's elu / add / add on its raw scale.  Backward:
    fp64 autograd through the oracle heads + elu + cat (the kernel recomputes the raw scale in fp32, so its ELU derivative is the exact
    one, not the derivative at the tensor-core forward value)."""
    import torch.nn.functional as F
    loc_h, sc_h = make_head(31).to(DEV), make_head(32).to(DEV)
    with torch.no_grad():
        sc_h[3].bias.sub_(0.5)                                  # both ELU branches populated
    min_scale = 1e-3
    g = torch.Generator().manual_seed(rows * 7 + T)
    x = torch.randn(rows, T, 64, generator=g) * 2.0
    cot = torch.randn(rows, T, 4, generator=g) * (torch.rand(rows, T, 1, generator=g) < frac)
    cot[..., 2:] *= (torch.rand(rows, T, 1, generator=g) < 0.5)               # points active in loc only / in both

    xd = x.double().clone().requires_grad_(True)
    P = [{k: v.double().clone().requires_grad_(True) for k, v in params_of(h).items()} for h in (loc_h, sc_h)]
    out_ref = torch.cat((so.decoder_loc_head_ref(P[0], xd), F.elu(so.decoder_loc_head_ref(P[1], xd)) + 1.0 + min_scale), dim=-1)
    leaves = [xd] + [t for p in P for t in p.values()]
    g_ref = torch.autograd.grad((out_ref * cot.double()).sum(), leaves)

    xg = x.to(DEV).requires_grad_(True)
    out, none = hd.decoder_heads(loc_h, sc_h, xg, cat_min_scale=min_scale)
    assert none is None and out.shape == (rows, T, 4)
    (out * cot.to(DEV)).sum().backward()
    with torch.no_grad():
        loc2, raw2 = hd.decoder_heads(loc_h, sc_h, xg.detach())
        o_i, _ = hd.decoder_heads(loc_h, sc_h, xg.detach(), cat_min_scale=min_scale)       # inference path (no autograd node)
    assert torch.equal(out.detach()[..., :2], loc2) and torch.equal(o_i, out.detach())
    assert torch.allclose(out.detach()[..., 2:], F.elu(raw2) + 1.0 + min_scale, atol=2e-6, rtol=2e-6)
    assert torch.allclose(out.detach().cpu().double(), out_ref.detach(), **TOL)
    sc = out.detach()[..., 2:]
    assert float((sc < 1.0 + min_scale).float().mean()) > 0.05 and float((sc > 1.0 + min_scale).float().mean()) > 0.05
    got = [xg.grad] + [p_.grad for h in (loc_h, sc_h) for p_ in h.state_dict(keep_vars=True).values()]
    for i, (a_, r_) in enumerate(zip(got, g_ref)):
        e = float((a_.double().cpu() - r_).abs().max() / r_.abs().max())
        assert e < 1e-4, (i, e)


@pytest.mark.parametrize('rows_major', [True, False])
def test_heads_from_solution_gradient_reaches_the_solver_in_place(rows_major):
    """decoder_heads_from_solution: heads on the solver's full ys; dL/dys comes back in ys's own layout (slab 0 zero) and the chain
    solver -> heads trains end to end (gradients of the SDE nets and of y0 are finite and non-zero)."""
    sde = init_like_reference(DecoderSDE(), seed=2).to(DEV)
    loc_h, sc_h = make_head(11).to(DEV), make_head(12).to(DEV)
    y0 = torch.relu(torch.randn(2100, 64, generator=torch.Generator().manual_seed(3))).to(DEV).requires_grad_(True)
    ts = torch.linspace(0, 6, 61)
    ys = tb.sdeint(sde, y0, ts, dt=0.1, method='euler', mode='tc_f16', seed=5, rows_major=rows_major)
    loc, sc = hd.decoder_heads_from_solution(loc_h, sc_h, ys)
    with torch.no_grad():
        want, _ = hd.decoder_heads(loc_h, sc_h, ys.detach()[1:].permute(1, 0, 2))
    assert torch.equal(loc.detach(), want)
    best = torch.zeros(2100, dtype=torch.bool, device=DEV)
    best[::10] = True                                          # one "mode" in ten receives a gradient
    (loc[best, :30].square().sum()).backward()
    assert float(y0.grad[best].abs().max()) > 0 and float(y0.grad[~best].abs().max()) == 0.0
    assert all(p_.grad is not None and torch.isfinite(p_.grad).all() for p_ in sde.parameters())
    assert sc_h[0].weight.grad is None or float(sc_h[0].weight.grad.abs().max()) == 0.0     # no gradient reached the scale head


def test_solve_and_heads_single_node_equals_the_composition():
    """solve_and_heads (one autograd node: the solution's gradient is an internal, only partly written buffer; the heads backward tells
    the solver backward which rows carry a gradient) against sdeint -> decoder_heads_from_solution on the same Philox seed."""
    from trajsde_b200 import ops
    sde = init_like_reference(DecoderSDE(), seed=2, bias_std=0.1).to(DEV)
    loc_h, sc_h = make_head(11).to(DEV), make_head(12).to(DEV)
    y_init = torch.relu(torch.randn(4100, 64, generator=torch.Generator().manual_seed(3))).to(DEV)
    ts = torch.linspace(0, 6, 61)
    cot = torch.zeros(4100, 60, 2, device=DEV)
    cot[3::10, :30] = torch.randn(410, 30, 2, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1)) * 1e-3

    def grads(fn):
        for p_ in list(sde.parameters()) + list(loc_h.parameters()) + list(sc_h.parameters()):
            p_.grad = None
        y0 = y_init.clone().requires_grad_(True)
        loc, sc = fn(y0)
        (loc * cot).sum().backward()
        return loc.detach(), y0.grad, [p_.grad.clone() for p_ in list(sde.parameters()) + list(loc_h.parameters())]

    n0 = ops.LAUNCHES['n']
    loc_a, gy_a, gp_a = grads(lambda y0: hd.solve_and_heads(sde, loc_h, sc_h, y0, ts, 0.1, seed=77, row_offset=5))
    assert ops.LAUNCHES['n'] - n0 == 4 + 4 + 4         # fwd: pack + solve + pack + heads; bwd: heads (flags, zero rows, main, reduce) + solver (compaction, pack, sweep, reduce: flags and max|grad| come from the heads backward)
    loc_b, gy_b, gp_b = grads(lambda y0: hd.decoder_heads_from_solution(
        loc_h, sc_h, tb.sdeint(sde, y0, ts, dt=0.1, method='euler', seed=77, row_offset=5, rows_major=True)))
    assert torch.equal(loc_a, loc_b)
    assert float((gy_a - gy_b).abs().max()) <= 3e-3 * float(gy_b.abs().max())
    assert float(gy_a[0::10].abs().max()) == 0.0 and float(gy_a[3::10].abs().max()) > 0
    for a_, b_ in zip(gp_a, gp_b):
        assert float((a_ - b_).abs().max()) <= 3e-3 * float(b_.abs().max()) + 1e-12
    assert sc_h[0].weight.grad is None or float(sc_h[0].weight.grad.abs().max()) == 0.0
    assert ops.backward_status(torch.device(DEV)) == 0
