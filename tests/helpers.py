"""Test helpers: build reference-shaped nn.Modules (own code, reference layout/state_dict keys) and synthetic inputs."""
import math

import torch
import torch.nn as nn


def _mlp(out_dim):
    return nn.Sequential(nn.Linear(66, 64), nn.Tanh(), nn.Linear(64, 64), nn.Tanh(), nn.Linear(64, out_dim))


class Net(nn.Module):
    def __init__(self, out_dim):
        super().__init__()
        self.net = _mlp(out_dim)


class DecoderSDE(nn.Module):
    """Attribute layout of the decoder's LSDEFunc (dec_hivt_nusargo_sde.py:160-167): f_func.net / g_func.net."""
    noise_type, sde_type = 'diagonal', 'ito'

    def __init__(self):
        super().__init__()
        self.f_func, self.g_func = Net(64), Net(1)
        self.fnfe = self.gnfe = 0


class EncoderSDE(nn.Module):
    """Attribute layout of the encoder's LSDEFunc (enc…sep2.py:442-448): f_func / g_nus / g_argo."""
    noise_type, sde_type = 'diagonal', 'ito'

    def __init__(self):
        super().__init__()
        self.f_func, self.g_nus, self.g_argo = Net(64), Net(1), Net(1)
        self.fnfe = self.gnfe = 0


def init_like_reference(module, seed, bias_std=0.1):
    """xavier_uniform weights / zero biases (models/utils/util.py:94-98) + N(0,bias_std) biases so bias paths are live."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, nn.Linear):
                fan_out, fan_in = m.weight.shape
                a = math.sqrt(6.0 / (fan_in + fan_out))
                m.weight.copy_((torch.rand(m.weight.shape, generator=g) * 2 - 1) * a)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * bias_std)
    return module


def load_net(net, params):
    with torch.no_grad():
        for k, v in params.items():
            i, kind = k.split('.')
            getattr(net.net[int(i)], kind).copy_(v)
    return net


def net_params(net):
    return {k: v.detach().cpu().clone() for k, v in net.net.state_dict().items()}


def make_dw(sched_h, rows, seed):
    g = torch.Generator().manual_seed(seed)
    S = len(sched_h)
    return torch.randn(S, rows, 64, generator=g) * torch.sqrt(torch.as_tensor(sched_h)).view(S, 1, 1)
