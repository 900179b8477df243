"""the fused aggr_embed prologue and the L2 / DiffBCE losses against the oracle restatements (oracle/sde_oracle.py) and
the fixture the REAL reference stage + losses produced (tests/golden/decoder_stage.npz), forward values and gradients (fp64 autograd)."""
import pytest
import torch
import torch.nn as nn

from oracle import sde_oracle as so
from trajsde_b200 import stage

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _t(d, k):
    return torch.from_numpy(d[k]).to(DEV)


def test_aggr_embed_vs_reference_fixture_and_fp64_gradients(golden_stage):
    d = golden_stage
    mod = nn.Sequential(nn.Linear(128, 64), nn.LayerNorm(64), nn.ReLU(inplace=True))
    mod.load_state_dict({k[len('param/aggr_embed.'):]: torch.from_numpy(v) for k, v in d.items() if k.startswith('param/aggr_embed.')})
    mod = mod.to(DEV)
    le, ge = _t(d, 'local_embed').requires_grad_(True), _t(d, 'global_embed').requires_grad_(True)
    h0 = stage.aggr_embed(mod, le, ge)
    assert h0.shape == (120, 64)
    assert torch.allclose(h0.detach().cpu(), torch.from_numpy(d['hidden_0']), atol=2e-6, rtol=1e-5)       # the reference's own hidden_0
    cot = torch.randn(120, 64, generator=torch.Generator().manual_seed(1))
    (h0 * cot.to(DEV)).sum().backward()
    p = {k[len('param/'):]: torch.from_numpy(v).double().requires_grad_(True) for k, v in d.items() if k.startswith('param/aggr_embed.')}
    led, ged = torch.from_numpy(d['local_embed']).double().requires_grad_(True), torch.from_numpy(d['global_embed']).double().requires_grad_(True)
    (so.aggr_embed_ref(p, led, ged) * cot.double()).sum().backward()
    pairs = [(le.grad, led.grad), (ge.grad, ged.grad), (mod[0].weight.grad, p['aggr_embed.0.weight'].grad), (mod[0].bias.grad, p['aggr_embed.0.bias'].grad),
             (mod[1].weight.grad, p['aggr_embed.1.weight'].grad), (mod[1].bias.grad, p['aggr_embed.1.bias'].grad)]
    for got, ref in pairs:
        assert float((got.double().cpu() - ref).abs().max()) <= 1e-4 * float(ref.abs().max())


@pytest.mark.parametrize('modes,n', [(10, 1), (10, 777), (6, 20480), (10, 0)])
def test_aggr_embed_shapes_vs_torch(modes, n):
    mod = nn.Sequential(nn.Linear(128, 64), nn.LayerNorm(64), nn.ReLU(inplace=True)).to(DEV)
    with torch.no_grad():
        mod[1].weight.add_(0.1 * torch.randn(64, device=DEV)); mod[1].bias.add_(0.1 * torch.randn(64, device=DEV))
    le, ge = torch.randn(n, 64, device=DEV), torch.randn(modes, n, 64, device=DEV)
    with torch.no_grad():
        got = stage.aggr_embed(mod, le, ge)
        want = mod(torch.cat((ge, le.expand(modes, *le.shape)), dim=-1)).reshape(modes * n, 64)
    assert got.shape == want.shape and torch.allclose(got, want, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize('modes,n', [(10, 1), (10, 777), (6, 20480), (10, 0)])
def test_pi_head_vs_torch_module(modes, n):
    """trajsde_pi_head_fwd (dec…sde.py:63-67, 92-94: the 4-layer pi head on cat(local.expand, global), squeezed and transposed) against
    the nn.Sequential it replaces; then its autograd path (cold: torch recompute) against autograd through the module."""
    mod = nn.Sequential(nn.Linear(128, 64), nn.LayerNorm(64), nn.ReLU(inplace=True), nn.Linear(64, 1)).to(DEV)
    with torch.no_grad():
        mod[1].weight.add_(0.1 * torch.randn(64, device=DEV)); mod[1].bias.add_(0.1 * torch.randn(64, device=DEV))
        mod[3].bias.add_(0.3)
    le, ge = torch.randn(n, 64, device=DEV), torch.randn(modes, n, 64, device=DEV)

    def ref(le_, ge_):
        return mod(torch.cat((le_.expand(modes, *le_.shape), ge_), dim=-1)).squeeze(-1).t()

    with torch.no_grad():
        got = stage.pi_head(mod, le, ge)
        want = ref(le, ge)
    assert got.shape == want.shape == (n, modes) and torch.allclose(got, want, atol=1e-5, rtol=1e-5)
    if n == 0:
        return
    cot = torch.randn(n, modes, device=DEV)
    res = []
    for fn in (lambda a, b: stage.pi_head(mod, a, b), ref):
        for p_ in mod.parameters():
            p_.grad = None
        a, b = le.clone().requires_grad_(True), ge.clone().requires_grad_(True)
        (fn(a, b) * cot).sum().backward()
        res.append([a.grad, b.grad] + [p_.grad.clone() for p_ in mod.parameters()])
    for x, r in zip(*res):
        assert torch.allclose(x, r, atol=1e-5 * float(r.abs().max()) + 1e-7, rtol=1e-4)
    # unused pi (the reference's training configuration): backward returns without touching the head
    for p_ in mod.parameters():
        p_.grad = None
    a = le.clone().requires_grad_(True)
    out = stage.pi_head(mod, a, ge)
    (a.sum() + 0.0 * out.detach().sum()).backward()
    assert all(p_.grad is None for p_ in mod.parameters())


def test_l2_and_bce_vs_reference_fixture(golden_stage):
    d = golden_stage
    loc = _t(d, 'loc').requires_grad_(True)
    loss, best = stage.l2_loss(loc, _t(d, 'y'), _t(d, 'reg_mask'), return_best=True)
    assert abs(float(loss) - float(d['loss_l2'])) < 1e-6
    di, do = _t(d, 'diff_in').requires_grad_(True), _t(d, 'diff_out').requires_grad_(True)
    bce = stage.diff_bce_loss(di, do)
    assert abs(float(bce) - float(d['loss_bce'])) < 1e-6
    (loss + bce).backward()
    assert torch.allclose(di.grad.cpu(), torch.from_numpy(d['grad_diff_in']), atol=1e-9, rtol=1e-5)      # the reference's own gradients
    assert torch.allclose(do.grad.cpu(), torch.from_numpy(d['grad_diff_out']), atol=1e-9, rtol=1e-5)
    locd = torch.from_numpy(d['loc']).double().requires_grad_(True)
    so.l2_loss_ref(locd, torch.from_numpy(d['y']).double(), torch.from_numpy(d['reg_mask'])).backward()
    assert float((loc.grad.double().cpu() - locd.grad).abs().max()) <= 1e-6 * float(locd.grad.abs().max())
    l2 = torch.norm(torch.from_numpy(d['y']).unsqueeze(0) - torch.from_numpy(d['loc'])[..., :2], dim=-1) * torch.from_numpy(d['reg_mask'])
    assert torch.equal(best.cpu().long(), torch.argmin(l2.mean(-1), dim=0))


@pytest.mark.parametrize('modes,n,T,c', [(10, 2048, 60, 4), (10, 333, 60, 2), (3, 1, 7, 4), (10, 50, 60, 4)])
def test_l2_loss_vs_oracle_random(modes, n, T, c):
    g = torch.Generator().manual_seed(n + T)
    loc = torch.randn(modes, n, T, c, generator=g) * 3
    y = torch.randn(n, T, 2, generator=g) * 3
    rm = torch.rand(n, T, generator=g) > (0.5 if n != 50 else 2.0)          # n == 50: nothing valid -> loss 0, gradient 0
    locg = loc.to(DEV).requires_grad_(True)
    loss = stage.l2_loss(locg, y.to(DEV), rm.to(DEV))
    (loss * 3.0).backward()
    locd = loc.double().requires_grad_(True)
    ref = so.l2_loss_ref(locd if c == 4 else torch.cat((locd, locd), -1), y.double(), rm)
    assert abs(float(loss) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref)))
    if rm.any():
        (ref * 3.0).backward()
        assert float((locg.grad.double().cpu() - locd.grad).abs().max()) <= 1e-5 * float(locd.grad.abs().max())
    else:
        assert float(locg.grad.abs().max()) == 0.0


def test_diff_bce_clamp_and_shapes():
    di = torch.tensor([0.0, 0.3, 1.0, 0.999999], device=DEV).requires_grad_(True)      # log(1 - 1) -> clamped at -100
    do = torch.tensor([[1e-30, 0.5], [0.0, 1.0]], device=DEV).requires_grad_(True)
    loss = stage.diff_bce_loss(di, do)
    ref = torch.nn.BCELoss()(di.detach(), torch.zeros_like(di)) + torch.nn.BCELoss()(do.detach(), torch.ones_like(do))
    assert abs(float(loss) - float(ref)) < 1e-4 * float(ref)
    loss.backward()
    assert di.grad.shape == di.shape and do.grad.shape == do.shape and torch.isfinite(di.grad).all() and torch.isfinite(do.grad).all()
