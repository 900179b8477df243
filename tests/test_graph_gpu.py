"""CUDA-graph friendliness of the path: the Philox key may live in device memory (TrajsdeNoise.seed_dev / trajsde_b200.set_device_seed), so
a step captured once draws fresh Brownian increments on every replay; every library call only enqueues work (no host sync, no
allocation outside torch's graph pool), so a whole training step — fused encoder, fused decoder stage, losses, backward — captures."""
import pytest
import torch

import trajsde_b200 as tb
from helpers import DecoderSDE, init_like_reference
from trajsde_b200 import ops, stage, synthetic as syn
from trajsde_b200 import encoder as enc_mod
from trajsde_b200.stages import FusedDecoderMixin

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


@pytest.fixture(autouse=True)
def _reset_device_seed():
    yield
    tb.set_device_seed(None)


def test_device_seed_adds_to_the_host_seed_and_graph_replays_draw_fresh_noise():
    sde = init_like_reference(DecoderSDE(), seed=3).to(DEV)
    ts = torch.linspace(0, 6, 61)
    y0 = torch.relu(torch.randn(300, 64, generator=torch.Generator().manual_seed(3))).to(DEV)
    with torch.no_grad():
        ref5 = tb.sdeint(sde, y0, ts, dt=0.1, method='euler', seed=5)
        ref8 = tb.sdeint(sde, y0, ts, dt=0.1, method='euler', seed=8)
        word = torch.zeros(1, dtype=torch.int64, device=DEV)
        tb.set_device_seed(word)
        assert torch.equal(tb.sdeint(sde, y0, ts, dt=0.1, method='euler', seed=5), ref5)
        word.fill_(3)
        assert torch.equal(tb.sdeint(sde, y0, ts, dt=0.1, method='euler', seed=5), ref8)
        # capture once, replay with a bumped word
        word.zero_()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            tb.sdeint(sde, y0, ts, dt=0.1, method='euler', seed=5)           # warm-up on the side stream (schedule tables, attributes)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            ys = tb.sdeint(sde, y0, ts, dt=0.1, method='euler', seed=5)
        g.replay()
        assert torch.equal(ys, ref5)
        word.add_(3)
        g.replay()
        assert torch.equal(ys, ref8)
    with pytest.raises(ValueError):
        tb.set_device_seed(torch.zeros(2, dtype=torch.int64, device=DEV))


def test_training_step_captures_into_one_cuda_graph():
    """fused encoder -> eos gather -> fused decoder stage -> L2 + DiffBCE -> backward, captured once; a replay with the same device seed
    reproduces the eager step's loss and gradients, a replay with another seed word differs."""
    class DecStage(FusedDecoderMixin, syn.DecoderStage):
        pass

    scenes, agents = 16, 8
    tbatch = syn.make_train_batch(scenes, agents, seed=5, device=DEV)
    b = tbatch.base
    enc_sde = syn.init_reference_style(syn.EncoderSDEFunc(), 1, bias_std=0.05).to(DEV)
    gru = syn.init_reference_style(syn.GRUUnit(), 3, bias_std=0.05).to(DEV)
    dstage = syn.init_reference_style(DecStage(), 11, bias_std=0.05).to(DEV)
    hidden = torch.nn.Parameter(torch.randn(64, device=DEV) * 0.02)
    params = list(enc_sde.parameters()) + list(gru.parameters()) + list(dstage.parameters()) + [hidden]
    E, N = b.enc_rows, scenes * agents
    eos = 20 - torch.argmax(b.bos_mask.float(), dim=1)
    ar = torch.arange(N, device=DEV)
    new_ai = torch.cat((tbatch.agent_index, torch.arange(N, E, device=DEV)))
    agent_eos = eos[tbatch.agent_index].repeat(2)
    word = torch.zeros(1, dtype=torch.int64, device=DEV)
    tb.set_device_seed(word)

    def step():
        for p_ in params:
            p_.grad = None
        lat, g = enc_mod.encoder_recurrence(enc_sde, gru, hidden.unsqueeze(0).repeat(E, 1), b.aa_out, b.actors_mask, b.nus_mask, seed=300)
        dstage.solver_kwargs = {'seed': 400}
        out = dstage({'padding_mask': tbatch.padding_mask}, lat[eos, ar], tbatch.global_embed)
        d_in, d_out = torch.chunk(g[agent_eos, new_ai], 2, 0)
        loss = stage.l2_loss(out['loc'], tbatch.y, out['reg_mask']) + stage.diff_bce_loss(d_in, d_out)
        loss.backward()
        return loss

    loss_eager = step().detach().clone()
    grads_eager = [None if p_.grad is None else p_.grad.clone() for p_ in params]     # pi / scale heads: no gradient under L2 + DiffBCE
    assert sum(ge is not None for ge in grads_eager) > 40
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(s)
    for p_ in params:
        p_.grad = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loss_static = step()
    g.replay()
    torch.cuda.synchronize()
    assert abs(float(loss_static) - float(loss_eager)) < 1e-6 * max(1.0, abs(float(loss_eager)))
    for p_, ge in zip(params, grads_eager):
        if ge is None:
            continue
        assert p_.grad is not None and float((p_.grad - ge).abs().max()) <= 1e-5 * float(ge.abs().max()) + 1e-12
    word.add_(1)
    g.replay()
    torch.cuda.synchronize()
    assert abs(float(loss_static) - float(loss_eager)) > 0          # fresh increments -> another sample of the loss
    assert ops.backward_status(DEV) == 0
