"""GPU parity of the fused backward (discretise-then-optimise) against fp64 autograd through the oracle — the gradients
torch.autograd would produce through the reference solver (`adjoint: false`, yml:41): y0, every weight incl. the time
columns W1[:,64:66], every bias."""
import pytest
import torch

import trajsde_b200 as tb
from helpers import DecoderSDE, EncoderSDE, init_like_reference, make_dw, net_params
from oracle import sde_oracle as so
from trajsde_b200 import ops
from trajsde_b200.schedule import euler_schedule

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def oracle_grads(nets, y0, ts, dW, cot_ys, cot_g=None, nus_mask=None):
    """fp64 autograd through oracle.euler_solve_ref. nets: list of param dicts [f, g] or [f, g_nus, g_argo]."""
    P = [{k: v.double().clone().requires_grad_(True) for k, v in n.items()} for n in nets]
    y = y0.double().clone().requires_grad_(True)
    if nus_mask is None:
        ys, g = so.euler_solve_ref(P[0], P[1], y, ts, 0.1, dW.double())
    else:
        ys, g = so.euler_solve_ref(P[0], P[1], y, ts, 0.1, dW.double(), nus_mask, P[2])
    loss = (ys * cot_ys.double()).sum()
    if cot_g is not None:
        loss = loss + (g[:, 0] * cot_g.double()).sum()
    leaves = [y] + [t for n in P for t in n.values()]
    grads = torch.autograd.grad(loss, leaves)
    names = ['y0'] + [f'net{i}.{k}' for i, n in enumerate(P) for k in n]
    return dict(zip(names, grads))


def rel_err(a, b):
    return float((a.double().cpu() - b).abs().max() / (b.abs().max() + 1e-12))


@pytest.mark.parametrize('mode,tol', [('exact', 2e-4), ('tc_f16', 3e-2)])
@pytest.mark.parametrize('F,rows', [(10, 37), (60, 70)])
def test_decoder_gradients_vs_fp64_autograd(mode, tol, F, rows):
    sde = init_like_reference(DecoderSDE(), seed=F + rows, bias_std=0.2).to(DEV)
    ts = torch.linspace(0, 0.1 * F, F + 1)
    sched = euler_schedule(ts, 0.1)
    g = torch.Generator().manual_seed(F)
    y0 = torch.relu(torch.randn(rows, 64, generator=g))
    dW = make_dw(sched.h, rows, seed=F + 1)
    cot = torch.randn(F + 1, rows, 64, generator=g)
    ref = oracle_grads([net_params(sde.f_func), net_params(sde.g_func)], y0, ts, dW, cot)

    y = y0.to(DEV).requires_grad_(True)
    ys = tb.sdeint(sde, y, ts, bm=dW.to(DEV), dt=0.1, method='euler', mode=mode)
    (ys * cot.to(DEV)).sum().backward()
    assert rel_err(y.grad, ref['y0']) < tol, ('y0', rel_err(y.grad, ref['y0']))
    for i, net in enumerate((sde.f_func, sde.g_func)):
        for k, prm in net.net.named_parameters():
            e = rel_err(prm.grad, ref[f'net{i}.{k}'])
            print(f"[{mode}] F={F} net{i}.{k}: rel err {e:.2e}")
            assert e < tol, (i, k, e)


@pytest.mark.parametrize('mode,tol', [('exact', 2e-4), ('tc_f16', 3e-2)])
def test_sdeint_dual_gradients_incl_g_output(mode, tol):
    """Encoder call site: gradients flow through ys AND through the returned diffusion g (DiffBCE consumes it,
    losses/diff_BCE.py:11-16); rows are routed to g_nus / g_argo by nus_mask."""
    rows = 90
    sde = init_like_reference(EncoderSDE(), seed=5, bias_std=0.2).to(DEV)
    ts = torch.tensor([0.3, 0.4])
    sched = euler_schedule(ts, 0.1)
    g = torch.Generator().manual_seed(5)
    y0 = torch.randn(rows, 64, generator=g) * 0.5
    mask = torch.rand(rows, generator=g) > 0.4
    dW = make_dw(sched.h, rows, seed=6)
    cot = torch.randn(2, rows, 64, generator=g)
    cot_g = torch.randn(rows, generator=g)
    nets = [net_params(sde.f_func), net_params(sde.g_nus), net_params(sde.g_argo)]
    ref = oracle_grads(nets, y0, ts, dW, cot, cot_g, mask)

    y = y0.to(DEV).requires_grad_(True)
    ys, gg = tb.sdeint_dual(sde, y, ts, mask.to(DEV), bm=dW.to(DEV), dt=0.1, method='euler', mode=mode)
    ((ys * cot.to(DEV)).sum() + (gg[:, 0] * cot_g.to(DEV)).sum()).backward()
    assert rel_err(y.grad, ref['y0']) < tol
    for i, net in enumerate((sde.f_func, sde.g_nus, sde.g_argo)):
        for k, prm in net.net.named_parameters():
            e = rel_err(prm.grad, ref[f'net{i}.{k}'])
            assert e < tol, (i, k, e)


def test_philox_backward_equals_supplied_dw_backward():
    """The backward regenerates the forward's Philox increments: identical gradients to the supplied-dW run."""
    sde = init_like_reference(DecoderSDE(), seed=8).to(DEV)
    ts = torch.linspace(0, 2, 21)
    sched = euler_schedule(ts, 0.1)
    y0 = torch.relu(torch.randn(50, 64, generator=torch.Generator().manual_seed(8))).to(DEV)

    def run(bm, seed):
        for p_ in sde.parameters():
            p_.grad = None
        y = y0.clone().requires_grad_(True)
        ys = tb.sdeint(sde, y, ts, bm=bm, dt=0.1, method='euler', mode='exact', seed=seed)
        ys.square().mean().backward()
        return [y.grad.clone()] + [p_.grad.clone() for p_ in sde.parameters()]

    a = run(None, 31)
    dW = ops.philox_dw(ops.DeviceSchedule.get(sched, torch.device(DEV)), 50, 31, torch.device(DEV))
    b = run(dW, None)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_backward_is_deterministic_and_grad_free_inference():
    sde = init_like_reference(DecoderSDE(), seed=3).to(DEV)
    ts = torch.linspace(0, 6, 61)
    y0 = torch.relu(torch.randn(300, 64, generator=torch.Generator().manual_seed(3))).to(DEV)

    def grads():
        for p_ in sde.parameters():
            p_.grad = None
        ys = tb.sdeint(sde, y0, ts, dt=0.1, method='euler', mode='tc_f16', seed=5)
        ys[1:].mean().backward()
        return [p_.grad.clone() for p_ in sde.parameters()]

    a, b = grads(), grads()
    assert all(torch.equal(x, y) for x, y in zip(a, b))          # fixed-order reduction: bit-reproducible
    with torch.no_grad():
        ys = tb.sdeint(sde, y0, ts, dt=0.1, method='euler', mode='tc_f16', seed=5)
    assert not ys.requires_grad


@pytest.mark.parametrize('F,rows,use_dw', [(60, 300, True), (20, 129, False), (100, 64, True)])
def test_tc_backward_matches_exact_backward(F, rows, use_dw):
    """Fused tensor-core backward (euler_bwd_tc.cu: fp16 operands, loss-scaled adjoint, MN-major wgrad MMAs) against the fp32
    CUDA-core backward on the SAME saved states: isolates the backward arithmetic.  Covers a ragged last tile, in-kernel Philox
    regeneration, tiny incoming gradients (loss scale), and the F=100 schedule with a zero-step and a two-step interval."""
    sde = init_like_reference(DecoderSDE(), seed=F, bias_std=0.2).to(DEV)
    ts = torch.linspace(0, 0.1 * F, F + 1)
    sched = euler_schedule(ts, 0.1)
    g = torch.Generator().manual_seed(F)
    y0 = torch.relu(torch.randn(rows, 64, generator=g)).to(DEV)
    dW = make_dw(sched.h, rows, seed=F + 1).to(DEV) if use_dw else None
    cot = (torch.randn(F + 1, rows, 64, generator=g) * 1e-6).to(DEV)      # mean-reduced losses give gradients this small

    def run(exact):
        ops.BWD_EXACT_KERNELS = exact
        try:
            for p_ in sde.parameters():
                p_.grad = None
            y = y0.clone().requires_grad_(True)
            ys = tb.sdeint(sde, y, ts, bm=dW, dt=0.1, method='euler', mode='tc_f16', seed=77)
            (ys * cot).sum().backward()
            return [y.grad.clone()] + [p_.grad.clone() for p_ in sde.parameters()]
        finally:
            ops.BWD_EXACT_KERNELS = False

    ref, got = run(True), run(False)
    names = ['y0'] + [n for n, _ in sde.named_parameters()]
    for n, a, b in zip(names, got, ref):
        e = float((a - b).abs().max() / (b.abs().max() + 1e-30))
        print(f"F={F} {n}: rel err {e:.2e}")
        assert e < 1e-2, (n, e)


def test_adjoint_range_status_bit():
    """The TC backward carries the adjoint with a power-of-two loss scale; when the adjoint outgrows the fp16 delta range the kernel
    raises TRAJSDE_STATUS_ADJOINT_RANGE instead of clipping silently.  Reference-style weights never get there; a drift net blown
    up 40x (Jacobian norm >> 1 over 61 steps) does."""
    from trajsde_b200 import _lib
    ts = torch.linspace(0, 6, 61)
    y0 = torch.relu(torch.randn(200, 64, generator=torch.Generator().manual_seed(1))).to(DEV)
    for blow_up, expect in ((1.0, 0), (40.0, _lib.STATUS_ADJOINT_RANGE)):
        sde = init_like_reference(DecoderSDE(), seed=2, bias_std=0.2).to(DEV)
        with torch.no_grad():
            for p_ in sde.f_func.parameters():
                p_.mul_(blow_up)
        ops.backward_status(DEV)                                   # clear
        y = y0.clone().requires_grad_(True)
        ys = tb.sdeint(sde, y, ts, dt=0.1, method='euler', mode='tc_f16', seed=5)
        ys[-1].sum().backward()
        assert ops.backward_status(DEV) & _lib.STATUS_ADJOINT_RANGE == expect


def test_registered_torch_ops_equal_the_eager_fast_path():
    """solver.py / encoder.py call thin autograd.Function wrappers; the registered torch.library ops (trajsde::euler_fwd/bwd,
    gru_fwd/bwd) wrap the same implementations and must give identical outputs and gradients through the dispatcher."""
    from trajsde_b200 import synthetic as syn
    sde = init_like_reference(DecoderSDE(), seed=4).to(DEV)
    ts = torch.linspace(0, 2, 21)
    y0 = torch.relu(torch.randn(100, 64, generator=torch.Generator().manual_seed(4))).to(DEV)

    def run(dispatcher):
        ops.USE_DISPATCHER = dispatcher
        try:
            for p_ in sde.parameters():
                p_.grad = None
            y = y0.clone().requires_grad_(True)
            ys = tb.sdeint(sde, y, ts, dt=0.1, method='euler', mode='tc_f16', seed=5)
            ys[1:].square().sum().backward()
            return [ys.detach().clone(), y.grad.clone()] + [p_.grad.clone() for p_ in sde.parameters()]
        finally:
            ops.USE_DISPATCHER = False

    for a, b in zip(run(False), run(True)):
        assert torch.equal(a, b)
    assert torch.ops.trajsde.euler_fwd is not None and torch.ops.trajsde.gru_fwd is not None and torch.ops.trajsde.enc_fwd is not None
    gru = syn.init_reference_style(syn.GRUUnit(), 1).to(DEV)
    h = torch.randn(70, 64, device=DEV)
    m = torch.rand(70, device=DEV) > 0.5
    gp = [p_ for p_ in gru.parameters()]
    assert torch.equal(ops.gru_call(h, h * 0.5, m, gp), torch.ops.trajsde.gru_fwd(h, h * 0.5, m, gp))


def test_loss_scale_sampled_absmax_falls_back_to_the_full_scan():
    """The loss scale of the tensor-core backward comes from a SAMPLED scan of the incoming gradient (every 8th block of 32 rows) for
    large inputs; when every sampled row carries a zero gradient the full scan must run, else a 1e-7-sized gradient would be carried
    unscaled and flushed to zero by the fp16 delta operands.  Gradient lives only in unsampled rows here; exact mode is the reference."""
    rows, F = 70_000, 2
    sde = init_like_reference(DecoderSDE(), seed=5, bias_std=0.2).to(DEV)
    ts = torch.linspace(0, 0.1 * F, F + 1)
    g = torch.Generator(device=DEV).manual_seed(11)
    y0 = torch.relu(torch.randn(rows, 64, device=DEV, generator=g))
    cot = torch.randn(F + 1, rows, 64, device=DEV, generator=g) * 1e-7
    sampled_rows = ((torch.arange(rows, device=DEV) >> 5) % 8) == 0
    cot[:, sampled_rows] = 0.0
    out = {}
    for mode in ('exact', 'tc_f16'):
        for p in sde.parameters():
            p.grad = None
        y = y0.clone().requires_grad_(True)
        ys = tb.sdeint(sde, y, ts, dt=0.1, method='euler', mode=mode, seed=3)
        ys.backward(cot)
        out[mode] = (y.grad.clone(), [p.grad.clone() for p in sde.parameters()])
    gy_e, gw_e = out['exact']
    gy_t, gw_t = out['tc_f16']
    assert gy_e.abs().max() > 0
    assert (gy_t - gy_e).abs().max() <= 3e-2 * gy_e.abs().max()
    for a, b in zip(gw_t, gw_e):
        assert (a - b).abs().max() <= 3e-2 * b.abs().max() + 1e-14


@pytest.mark.parametrize('mode', ['exact', 'tc_f16'])
def test_rows_major_storage_gives_the_same_gradients(mode):
    """install() binds the decoder's sdeint with rows-major storage; the reference consumes `ys[1:].permute(1,0,2)` (dec…sde.py:88).
    Same loss through both storage layouts -> same gradients (the backward reads grad_ys through its strides)."""
    sde = init_like_reference(DecoderSDE(), seed=6, bias_std=0.2).to(DEV)
    ts = torch.linspace(0, 2, 21)
    g = torch.Generator().manual_seed(2)
    y0 = torch.relu(torch.randn(150, 64, generator=g)).to(DEV)
    w = torch.randn(150, 20, 64, generator=g).to(DEV)
    out = []
    for rm in (False, True):
        for p in sde.parameters():
            p.grad = None
        y = y0.clone().requires_grad_(True)
        sol_y = tb.sdeint(sde, y, ts, dt=0.1, method='euler', mode=mode, seed=12, rows_major=rm)[1:].permute(1, 0, 2)
        (sol_y * w).sum().backward()
        out.append([y.grad.clone()] + [p.grad.clone() for p in sde.parameters()])
    for a, b in zip(*out):
        assert torch.allclose(a, b, atol=1e-6 * float(b.abs().max()) + 1e-12, rtol=1e-5)


@pytest.mark.parametrize('rows,frac,rows_major', [(3000, 0.1, True), (5000, 0.0, False), (2500, 1.0, True), (70_000, 0.1, True)])
def test_zero_row_skipping_is_exact(rows, frac, rows_major):
    """TRAJSDE_BWD_FLAG_SKIP_ZERO_ROWS: rows with an all-zero incoming gradient (90 % of the decoder rows under the reference's
    winner-takes-all L2 loss) are left out of the reverse sweep.  Same gradients as the full sweep for sparse, empty and dense cotangents
    and both output layouts — up to the fp16 rounding of the delta operands: the two paths derive the power-of-two loss scale from a
    full and from a sampled scan of max|grad|, which may differ by a binade, and partition the partial sums differently; in-kernel Philox
    noise is regenerated for the compacted rows by their ORIGINAL row ids."""
    sde = init_like_reference(DecoderSDE(), seed=rows, bias_std=0.2).to(DEV)
    ts = torch.linspace(0, 6, 61)
    g = torch.Generator().manual_seed(rows)
    y0 = torch.relu(torch.randn(rows, 64, generator=g)).to(DEV)
    active = torch.rand(rows, generator=g) < frac
    cot = torch.randn(61, rows, 64, generator=g) * 1e-4
    cot[0].zero_()
    cot[5:, ~active] = 0                                         # inactive rows: nothing at all
    cot[1:5, ~active] = 0
    cot[0, active] = torch.randn(int(active.sum()), 64, generator=g) * 1e-4 if frac > 0 else 0     # ys[0] = y0 cotangent on active rows
    if rows == 3000:
        cot[:, 7] = 0
        cot[0, 7, 3] = 1e-4                                       # a row that is active through its slab-0 gradient only
    cot = cot.to(DEV)

    def run(skip):
        ops.SKIP_ZERO_ROWS = skip
        for p_ in sde.parameters():
            p_.grad = None
        y = y0.clone().requires_grad_(True)
        ys = tb.sdeint(sde, y, ts, dt=0.1, method='euler', mode='tc_f16', seed=31, row_offset=12345, rows_major=rows_major)
        ys.backward(cot)
        return y.grad.clone(), [p_.grad.clone() for p_ in sde.parameters()]

    try:
        gy_a, gw_a = run(False)
        gy_b, gw_b = run(True)
    finally:
        ops.SKIP_ZERO_ROWS = True
    scale = float(gy_a.abs().max())
    if frac == 0.0:
        assert scale == 0.0 and float(gy_b.abs().max()) == 0.0
        assert all(float(w.abs().max()) == 0.0 for w in gw_a + gw_b)
        return
    assert float((gy_a - gy_b).abs().max()) <= 3e-3 * scale
    assert torch.equal(gy_b[~active.to(DEV)] if rows != 3000 else gy_b[(~active).to(DEV) & (torch.arange(rows, device=DEV) != 7)],
                       torch.zeros_like(gy_b[~active.to(DEV)] if rows != 3000 else gy_b[(~active).to(DEV) & (torch.arange(rows, device=DEV) != 7)]))
    for name, a_, b_ in zip([n for n, _ in sde.named_parameters()], gw_a, gw_b):
        assert float((a_ - b_).abs().max()) <= 3e-3 * float(a_.abs().max()) + 1e-12, name
    assert ops.backward_status(torch.device(DEV)) == 0
