"""Synthetic Argoverse/nuScenes-shaped inputs of the SDE hot path: shapes, masks and value statistics of
what the reference encoder/decoder feed their solver calls.  No real data exists in this environment."""
import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

HIST, FUT, MODES, DIM = 21, 60, 10, 64


def _mlp(out_dim: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(DIM + 2, DIM), nn.Tanh(), nn.Linear(DIM, DIM), nn.Tanh(), nn.Linear(DIM, out_dim))


class _Net(nn.Module):
    def __init__(self, out_dim: int):
        super().__init__()
        self.net = _mlp(out_dim)


class DecoderSDEFunc(nn.Module):
    """Same attribute / state_dict layout as the decoder's LSDEFunc (dec_hivt_nusargo_sde.py:160-167)."""
    noise_type, sde_type = 'diagonal', 'ito'

    def __init__(self):
        super().__init__()
        self.f_func, self.g_func = _Net(DIM), _Net(1)
        self.fnfe = self.gnfe = self.hnfe = 0


class EncoderSDEFunc(nn.Module):
    """Same layout as the encoder's dual-diffusion LSDEFunc (enc_hivt_nusargo_sde_sep2.py:442-448)."""
    noise_type, sde_type = 'diagonal', 'ito'

    def __init__(self):
        super().__init__()
        self.f_func, self.g_nus, self.g_argo = _Net(DIM), _Net(1), _Net(1)
        self.fnfe = self.gnfe = self.hnfe = 0


class GRUUnit(nn.Module):
    """Same parameter layout as GRU_Unit (models/utils/ode_utils.py:111-134)."""

    def __init__(self, latent_dim: int = DIM, input_dim: int = DIM, n_units: int = DIM):
        super().__init__()
        self.update_gate = nn.Sequential(nn.Linear(latent_dim + input_dim, n_units), nn.Tanh(), nn.Linear(n_units, latent_dim), nn.Sigmoid())
        self.reset_gate = nn.Sequential(nn.Linear(latent_dim + input_dim, n_units), nn.Tanh(), nn.Linear(n_units, latent_dim), nn.Sigmoid())
        self.new_state_net = nn.Sequential(nn.Linear(latent_dim + input_dim, n_units), nn.Tanh(), nn.Linear(n_units, latent_dim))

    def forward(self, h_cur, input_tensor, mask):
        y_concat = torch.cat([h_cur, input_tensor], -1)
        u = self.update_gate(y_concat)
        r = self.reset_gate(y_concat)
        n = self.new_state_net(torch.cat([input_tensor, r * h_cur], dim=1))
        h_next = (1 - u) * n + u * h_cur
        m = mask.unsqueeze(-1)
        return m * h_next + ~m * h_cur


def init_reference_style(module: nn.Module, seed: int, bias_std: float = 0.0) -> nn.Module:
    """xavier_uniform_ weights, zero biases (models/utils/util.py:94-98); optional N(0,bias_std) biases for parity runs."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, nn.Linear):
                fan_out, fan_in = m.weight.shape
                a = math.sqrt(6.0 / (fan_in + fan_out))
                m.weight.copy_((torch.rand(m.weight.shape, generator=g) * 2 - 1) * a)
                m.bias.zero_()
                if bias_std > 0:
                    m.bias.add_(torch.randn(m.bias.shape, generator=g) * bias_std)
    return module


@dataclass
class SdeBatch:
    """Host-side (pinned when possible) inputs of one forward pass of the SDE path for `scenes` scenes."""
    scenes: int
    agents: int
    enc_h0: torch.Tensor        # [N', 64]  hidden.repeat(N',1), hidden ~ N(0, 0.02^2)        enc…sep2.py:61-62,78
    aa_out: torch.Tensor        # [21, N', 64] ~ N(0,1): output of the (out-of-scope) AA encoder   :107-121
    actors_mask: torch.Tensor   # [N', 21] bool: observed slots                                    :100
    nus_mask: torch.Tensor      # [N'] bool: row belongs to a nuScenes scene                       :73-74,103
    bos_mask: torch.Tensor      # [N, 21] bool: first observed slot                                :187
    dec_y0: torch.Tensor        # [10*N, 64] relu(N(0,1)): post aggr_embed statistics          dec…sde.py:26-29,82

    @property
    def enc_rows(self) -> int:
        return self.enc_h0.shape[0]

    @property
    def dec_rows(self) -> int:
        return self.dec_y0.shape[0]


def make_batch(scenes: int, agents_per_scene: int = 20, seed: int = 0, mixed_sources: bool = False,
               pin: bool = False) -> SdeBatch:
    g = torch.Generator().manual_seed(seed)
    n = scenes * agents_per_scene
    n_enc = n + scenes                                     # + one perturbed copy of every target agent (enc…sep2.py:94-103)
    hidden = torch.randn(DIM, generator=g) * 0.02
    enc_h0 = hidden.unsqueeze(0).repeat(n_enc, 1)
    aa_out = torch.randn(HIST, n_enc, DIM, generator=g)
    source = (torch.rand(scenes, generator=g) < 0.5).long() if mixed_sources else torch.ones(scenes, dtype=torch.long)
    batch = torch.arange(scenes).repeat_interleave(agents_per_scene)
    nus_mask = torch.cat([(source == 0)[batch], source == 0])
    # Argoverse-shaped past: slots 1..20 valid; nuScenes-shaped: slots {0,5,10,15,20} (nuScenes_Argoverse.py:92-103)
    argo = torch.zeros(HIST, dtype=torch.bool); argo[1:] = True
    nusc = torch.zeros(HIST, dtype=torch.bool); nusc[0::5] = True
    pattern = torch.where(nus_mask.unsqueeze(1), nusc.unsqueeze(0), argo.unsqueeze(0))
    actors_mask = pattern & (torch.rand(n_enc, HIST, generator=g) > 0.1)       # 10% extra random padding
    first = torch.argmax(actors_mask[:n].float(), dim=1)
    bos_mask = torch.zeros(n, HIST, dtype=torch.bool)
    bos_mask[torch.arange(n), first] = True
    dec_y0 = torch.relu(torch.randn(MODES * n, DIM, generator=g))
    out = SdeBatch(scenes, agents_per_scene, enc_h0, aa_out, actors_mask, nus_mask, bos_mask, dec_y0)
    if pin and torch.cuda.is_available():
        for k in ('enc_h0', 'aa_out', 'actors_mask', 'nus_mask', 'bos_mask', 'dec_y0'):
            setattr(out, k, getattr(out, k).pin_memory())
    return out


def _head(out_dim: int, in_dim: int = DIM) -> nn.Sequential:
    return nn.Sequential(nn.Linear(in_dim, DIM), nn.LayerNorm(DIM), nn.ReLU(inplace=True), nn.Linear(DIM, out_dim))


class DecoderStage(nn.Module):
    """Parameter containers of the decoder stage around the solve, with the reference's names and shapes (dec_hivt_nusargo_sde.py:26-29,
    46-67): ``aggr_embed``, ``lsde_func``, ``decoder``, ``scale``, ``pi``.  Forward = trajsde_b200.stages.FusedDecoderMixin."""

    def __init__(self, num_modes: int = MODES, future_steps: int = FUT, min_scale: float = 0.001):
        super().__init__()
        self.num_modes, self.future_steps, self.min_scale, self.uncertain = num_modes, future_steps, min_scale, True
        self.min_stepsize, self.rtol, self.atol, self.method = 0.1, 0.001, 0.001, 'euler'
        self.aggr_embed = nn.Sequential(nn.Linear(2 * DIM, DIM), nn.LayerNorm(DIM), nn.ReLU(inplace=True))
        self.lsde_func = DecoderSDEFunc()
        self.decoder, self.scale, self.pi = _head(2), _head(2), _head(1, 2 * DIM)
        self.ts_pred = torch.linspace(0, 6, future_steps + 1)


@dataclass
class TrainBatch:
    """Device-resident inputs of one training step of the SDE path and its direct consumers (the HiVT stages that produce them are out
    of scope: ``aa_out`` is the AA encoder's output, ``global_embed`` the global interactor's)."""
    base: SdeBatch
    global_embed: torch.Tensor     # [10, N, 64]
    y: torch.Tensor                # [N, 60, 2] ground-truth displacements
    padding_mask: torch.Tensor     # [N, 81] bool
    agent_index: torch.Tensor      # [scenes] long: the target agent of every scene


def make_train_batch(scenes: int, agents_per_scene: int = 20, seed: int = 0, device='cpu') -> TrainBatch:
    """Mixed nuScenes / Argoverse-shaped scenes (BASELINE configs[2]/[3]): past / future validity patterns of
    dataset/nuScenes_Argoerse/nuScenes_Argoverse.py:92-103 plus 10 % random padding."""
    b = make_batch(scenes, agents_per_scene, seed=seed, mixed_sources=True)
    g = torch.Generator().manual_seed(seed + 17)
    n = scenes * agents_per_scene
    nus = b.nus_mask[:n]
    fut_argo = torch.zeros(FUT, dtype=torch.bool); fut_argo[:30] = True
    fut_nusc = torch.zeros(FUT, dtype=torch.bool); fut_nusc[4::5] = True
    fut_valid = torch.where(nus.unsqueeze(1), fut_nusc.unsqueeze(0), fut_argo.unsqueeze(0)) & (torch.rand(n, FUT, generator=g) > 0.1)
    padding_mask = torch.cat((~b.actors_mask[:n], ~fut_valid), dim=1)
    tb_ = TrainBatch(b, torch.randn(MODES, n, DIM, generator=g), torch.randn(n, FUT, 2, generator=g).cumsum(1) * 0.5, padding_mask,
                     torch.arange(scenes) * agents_per_scene)
    if str(device) != 'cpu':
        for k in ('enc_h0', 'aa_out', 'actors_mask', 'nus_mask', 'bos_mask', 'dec_y0'):
            setattr(b, k, getattr(b, k).to(device))
        tb_.global_embed, tb_.y, tb_.padding_mask, tb_.agent_index = (t.to(device) for t in (tb_.global_embed, tb_.y, tb_.padding_mask, tb_.agent_index))
    return tb_
