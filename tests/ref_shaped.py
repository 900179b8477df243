"""Reference-SHAPED stage modules for boxes where /root/reference is absent (the GPU box).  TEST SCAFFOLDING, own code.

``RefShapedDecoder`` has the parameter names (state_dict keys) and the ``forward`` call sequence of the reference's ``SDEDecoder``
(models/decoders/dec_hivt_nusargo_sde.py:15-105): aggr_embed -> module-global ``sdeint`` -> [1:].permute(1,0,2) -> heads -> elu_ -> cat,
so ``trajsde_b200.install()`` meets exactly what it meets on the real stage: a module global named ``sdeint`` and head instances
called ``decoder`` / ``scale``.  ``RefShapedEncoder`` keeps the reference's loop (enc…sep2.py:128-196) around a module-global
``sdeint_dual`` and a ``gru_unit`` instance, with PyG-free stand-ins where the reference runs its graph attention (AAEncoder /
ALEncoder are out of scope): per-slot MLPs with the same input/output shapes.
The fixtures under tests/golden/decoder_stage.npz were produced by the REAL reference class; loading its state_dict here and
reproducing its outputs is what pins this stand-in.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from helpers import DecoderSDE, EncoderSDE
from trajsde_b200.synthetic import GRUUnit


def _unbound(*a, **k):
    raise RuntimeError("torchsde is not installed here: trajsde_b200.install() must rebind this module global first")


sdeint = _unbound            # the reference binds these with `from torchsde import sdeint` / `from models.utils.sdeint import sdeint_dual`
sdeint_dual = _unbound


def _head(out_dim, in_dim=64):
    return nn.Sequential(nn.Linear(in_dim, 64), nn.LayerNorm(64), nn.ReLU(inplace=True), nn.Linear(64, out_dim))


class RefShapedDecoder(nn.Module):
    def __init__(self, num_modes=10, future_steps=60, max_fut_t=6, min_stepsize=0.1, min_scale=0.001, rtol=0.001, atol=0.001,
                 method='euler', uncertain=True):
        super().__init__()
        self.num_modes, self.future_steps, self.min_stepsize, self.min_scale = num_modes, future_steps, min_stepsize, min_scale
        self.rtol, self.atol, self.method, self.uncertain, self.hidden_size = rtol, atol, method, uncertain, 64
        self.aggr_embed = nn.Sequential(nn.Linear(128, 64), nn.LayerNorm(64), nn.ReLU(inplace=True))
        self.lsde_func = DecoderSDE()
        self.decoder = _head(2)
        if uncertain:
            self.scale = _head(2)
        self.pi = _head(1, 128)
        self.hidden = nn.Parameter(torch.zeros(64))
        self.ts_pred = torch.linspace(0, max_fut_t, future_steps + 1)

    def load_reference_state_dict(self, sd):
        """state_dict of the real SDEDecoder; its inert ``lsde_func.h_func`` scalars (HFunc, never evaluated) have no twin here."""
        own = self.state_dict()
        missing = [k for k in own if k not in sd]
        assert not missing, missing
        self.load_state_dict({k: v for k, v in sd.items() if k in own})
        return self

    def forward(self, data, local_embed, global_embed):
        expanded = local_embed.expand(self.num_modes, *local_embed.shape)
        loc_emb = self.aggr_embed(torch.cat((global_embed, expanded), dim=-1))
        num_actors = loc_emb.shape[1]
        hidden_0 = loc_emb.view(self.num_modes * num_actors, self.hidden_size)
        sol_y = sdeint(self.lsde_func, hidden_0, self.ts_pred, dt=self.min_stepsize, dt_min=self.min_stepsize, rtol=self.rtol, atol=self.atol,
                       method=self.method)[1:].permute(1, 0, 2)
        pi = self.pi(torch.cat((expanded, global_embed), dim=-1)).squeeze(-1).t()
        loc = self.decoder(sol_y).view(self.num_modes, num_actors, self.future_steps, 2)
        out = {'pi': pi}
        if self.uncertain:
            scale = F.elu_(self.scale(sol_y), alpha=1.0).view(self.num_modes, -1, self.future_steps, 2) + 1.0
            out['loc'] = torch.cat((loc, scale + self.min_scale), dim=-1)
        else:
            out['loc'] = loc
        out['reg_mask'] = ~data['padding_mask'][:, -self.future_steps:]
        return out


class _SlotMLP(nn.Module):
    """Stand-in with AAEncoder's interface shape: x[(21*N'), 2] -> [(21*N'), 64] (no neighbours: graph attention is out of scope)."""

    def __init__(self):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(2, 64), nn.ReLU(), nn.Linear(64, 64))

    def forward(self, x):
        return self.net(x)


class RefShapedEncoder(nn.Module):
    """The reference's loop (enc…sep2.py:128-196) over ``sdeint_dual`` + ``gru_unit``; ``data`` is a dict with ``x`` [N,21,2],
    ``padding_mask`` [N,81], ``bos_mask`` [N,21], ``agent_index`` [B], ``batch`` [N], ``source`` [B]."""

    def __init__(self, historical_steps=21, max_past_t=2, ref_time=20, minimum_step=0.1, rtol=0.001, atol=0.001, method='euler'):
        super().__init__()
        self.historical_steps, self.max_past_t, self.ref_time, self.minimum_step = historical_steps, max_past_t, ref_time, minimum_step
        self.rtol, self.atol, self.method, self.run_backwards, self.embed_dim = rtol, atol, method, True, 64
        self.aa_encoder = _SlotMLP()
        self.al_encoder = nn.Linear(64, 64)
        self.gru_unit = GRUUnit()
        self.lsde_func = EncoderSDE()
        self.real_label, self.fake_label = 0, 1
        self.hidden = nn.Parameter(torch.randn(64) * 0.02)

    # ---- the parts the reference computes with PyG (enc…sep2.py:66-127), PyG-free ----------------------------------------------------
    def prepare(self, data, noise=None):
        nus_mask = torch.isin(data['batch'], torch.where(data['source'] == 0)[0])
        actor_num = data['x'].shape[0]
        ai = data['agent_index']
        x_agent = data['x'][ai]
        noise = 2 * torch.randn_like(x_agent) if noise is None else noise
        x_actors = torch.cat((data['x'], x_agent + noise), dim=0)
        actors_pad = torch.cat((data['padding_mask'], data['padding_mask'][ai]), dim=0)
        actors_mask = ~actors_pad[:, :self.ref_time + 1]
        new_agent_index = torch.cat((ai, torch.arange(actor_num, actor_num + ai.size(0), device=ai.device)))
        nus_mask = torch.cat((nus_mask, data['source'] == 0), dim=0)
        aa_out = self.aa_encoder(x_actors.transpose(0, 1).reshape(-1, 2)).view(self.historical_steps, x_actors.shape[0], -1)
        return aa_out, actors_mask, nus_mask, new_agent_index

    def forward(self, data, noise=None):
        aa_out, actors_mask, nus_mask, new_agent_index = self.prepare(data, noise)
        n_rows = aa_out.shape[1]
        prev_hidden = self.hidden.unsqueeze(0).repeat(n_rows, 1)
        past_time_steps = -1 * torch.linspace(-self.max_past_t, 0, self.historical_steps)
        prev_t, t_i = past_time_steps[-1] - 0.01, past_time_steps[-1]
        latent_ys, diffusions = [], []
        for idx, t in enumerate(reversed(range(self.historical_steps))):
            time_points = torch.tensor([prev_t, t_i])
            pred_y, diff_noise = sdeint_dual(self.lsde_func, prev_hidden, time_points, nus_mask, dt=self.minimum_step, rtol=self.rtol,
                                             atol=self.atol, method=self.method)
            ode_sol = pred_y.permute(1, 2, 0)
            if torch.mean(ode_sol[:, :, 0] - prev_hidden) >= 0.001:          # the reference's per-iteration host sync (:160-163)
                raise RuntimeError("first point of the ODE is not equal to initial value")
            yi = self.gru_unit(input_tensor=aa_out[t], h_cur=ode_sol[:, :, -1], mask=actors_mask[:, t]).squeeze(0)
            diffusions.append(diff_noise[new_agent_index])
            prev_hidden = yi
            if idx + 1 < self.historical_steps:
                prev_t, t_i = past_time_steps[t], past_time_steps[t - 1]
            latent_ys.append(yi)
        latent_ys = torch.stack(latent_ys)[:, :-len(data['agent_index'])]
        diffusions = torch.stack(diffusions)
        eos_idcs = self.ref_time - torch.argmax(data['bos_mask'].float(), dim=1)
        out = latent_ys[eos_idcs, torch.arange(latent_ys.size(1)), :]
        agent_eos = eos_idcs[data['agent_index']]
        diff_out = diffusions[agent_eos.repeat(2), torch.arange(diffusions.size(1))]
        d_in, d_out = torch.chunk(diff_out, 2, 0)
        return self.al_encoder(out), d_in, d_out, torch.full_like(d_in, self.real_label), torch.full_like(d_out, self.fake_label)
