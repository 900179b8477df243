"""Drive the reference's OWN hot-path files (verbatim from /root/reference) behind oracle/shims.  TEST INFRASTRUCTURE ONLY.

Dev-container only: /root/reference does not exist on the GPU box, so nothing marked ``gpu``, ``smoke()`` or ``bench.py``
may call into this module.  It is used by tests/golden/make_golden.py (fixture generation) and by CPU tests that pin
``oracle/sde_oracle.py`` against the reference when the tree is present.

What runs verbatim from the reference:
  * ``FFunc/GFunc/HFunc/LSDEFunc`` and ``SDEDecoder`` (models/decoders/dec_hivt_nusargo_sde.py)
  * ``FFunc/GFunc/LSDEFunc`` (dual g) of models/encoders/enc_hivt_nusargo_sde_sep2.py
  * ``sdeint_dual``, ``check_contract``, ``Euler_private``, ``BaseSDESolver_private``, ``ForwardSDE_private``
    (models/utils/sdeint.py) and ``GRU_Unit`` (models/utils/ode_utils.py)
What is restated in the shim (not in the reference tree): torchsde 0.2.5 ``sdeint``/``Euler``/``BaseSDESolver``/
``linear_interp``/Brownian classes (oracle/shims/torchsde).
"""
import os
import sys
from importlib.machinery import SourceFileLoader
from typing import Dict

import torch

REFERENCE_ROOT = os.environ.get('TRAJSDE_REFERENCE_ROOT', '/root/reference')
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shims')

DEC_KW = dict(local_channels=64, global_channels=64, future_steps=60, num_modes=10, max_fut_t=6, ode_func_layers=3,
              uncertain=True, min_scale=0.001, rtol=0.001, atol=0.001, min_stepsize=0.1, method='euler')


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'models', 'utils', 'sdeint.py'))


_mods = {}


def load_reference():
    """Import the reference modules the way model_base_mix_sde.py:38-45 does (SourceFileLoader)."""
    if _mods:
        return _mods
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (REFERENCE_ROOT, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    _mods['dec'] = SourceFileLoader('SDEDecoder', os.path.join(
        REFERENCE_ROOT, 'models/decoders/dec_hivt_nusargo_sde.py')).load_module('SDEDecoder')
    _mods['enc'] = SourceFileLoader('LocalEncoderSDESepPara2', os.path.join(
        REFERENCE_ROOT, 'models/encoders/enc_hivt_nusargo_sde_sep2.py')).load_module('LocalEncoderSDESepPara2')
    import models.utils.ode_utils as ode_utils
    import models.utils.sdeint as ref_sdeint
    import models.utils.util as ref_util
    import torchsde
    _mods.update(ode_utils=ode_utils, sdeint=ref_sdeint, util=ref_util, torchsde=torchsde)
    return _mods


def _perturb_biases(module: torch.nn.Module, std: float, gen: torch.Generator):
    """init_weights zeroes every bias (util.py:94-98); parity fixtures add N(0,std) so bias paths are exercised."""
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, torch.nn.Linear) and m.bias is not None:
                m.bias.add_(torch.randn(m.bias.shape, generator=gen) * std)


def build_reference_decoder(seed: int = 0, bias_std: float = 0.1):
    m = load_reference()
    torch.manual_seed(seed)
    dec = m['dec'].SDEDecoder(**DEC_KW)
    _perturb_biases(dec, bias_std, torch.Generator().manual_seed(seed + 1))
    return dec.eval()


def build_reference_encoder_sde(seed: int = 0, bias_std: float = 0.1):
    """(lsde_func, gru_unit) built exactly like LocalEncoderSDESepPara2.__init__ :49-57,64 does (AA/AL encoders skipped:
    out of scope)."""
    m = load_reference()
    enc = m['enc']
    torch.manual_seed(seed)
    gru = m['ode_utils'].GRU_Unit(64, 64, n_units=64)
    lsde = enc.LSDEFunc(f=enc.FFunc(64, num_layers=2), g_nus=enc.GFunc(64, num_layers=2, sigma=0.5),
                        g_Argo2=enc.GFunc(64, num_layers=2, sigma=0.5), h=enc.HFunc(theta=1.0, mu=0.0), embed_dim=64)
    lsde.noise_type, lsde.sde_type = 'diagonal', 'ito'
    holder = torch.nn.ModuleDict({'gru_unit': gru, 'lsde_func': lsde})
    holder.apply(m['util'].init_weights)                       # enc…sep2.py:64 overrides GRU's N(0,0.1) init
    _perturb_biases(holder, bias_std, torch.Generator().manual_seed(seed + 1))
    return lsde.eval(), gru.eval()


def net_params(net: torch.nn.Module) -> Dict[str, torch.Tensor]:
    return {k: v.detach().clone() for k, v in net.state_dict().items()}


def run_reference_decoder_solve(dec, hidden_0: torch.Tensor, dW: torch.Tensor):
    """The call at dec_hivt_nusargo_sde.py:88, with a FixedIncrements Brownian so dW is caller-supplied.
    Returns (ys[61,rows,64], list of (ta,tb) Brownian queries)."""
    m = load_reference()
    bm = m['torchsde'].FixedIncrements(dW)
    with torch.no_grad():
        ys = m['dec'].sdeint(dec.lsde_func, hidden_0, dec.ts_pred, bm=bm, dt=dec.min_stepsize, dt_min=dec.min_stepsize,
                             rtol=dec.rtol, atol=dec.atol, method=dec.method)
    return ys, bm.queries


def run_reference_decoder_heads(dec, ys: torch.Tensor):
    """dec…sde.py:88 tail + :95-99: loc/scale heads on sol_y = ys[1:].permute(1,0,2)."""
    with torch.no_grad():
        sol_y = ys[1:].permute(1, 0, 2)
        loc = dec.decoder(sol_y)
        scale = torch.nn.functional.elu(dec.scale(sol_y), alpha=1.0) + 1.0 + dec.min_scale
    return loc, scale


def run_reference_encoder_loop(lsde, gru, h0, aa_out, actors_mask, nus_mask, dW, minimum_step=0.1, max_past_t=2,
                               historical_steps=21):
    """The loop body of LocalEncoderSDESepPara2.forward (enc…sep2.py:128-182) around the reference's own
    ``sdeint_dual`` and ``GRU_Unit``; only the PyG-dependent producers of aa_out/masks are replaced by arguments."""
    m = load_reference()
    sdeint_dual = m['enc'].sdeint_dual
    past_time_steps = -1 * torch.linspace(-max_past_t, 0, historical_steps)
    prev_t, t_i = past_time_steps[-1] - 0.01, past_time_steps[-1]
    prev_hidden = h0
    latent_ys, gs, queries = [], [], []
    with torch.no_grad():
        for idx, t in enumerate(reversed(range(historical_steps))):
            time_points = torch.tensor([prev_t, t_i])
            bm = m['torchsde'].FixedIncrements(dW[idx:idx + 1])
            pred_y, diff_noise = sdeint_dual(lsde, prev_hidden, time_points, nus_mask, bm=bm, dt=minimum_step,
                                             rtol=0.001, atol=0.001, method='euler')
            queries += bm.queries
            ode_sol = pred_y.permute(1, 2, 0)
            yi = gru(input_tensor=aa_out[t], h_cur=ode_sol[:, :, -1], mask=actors_mask[:, t]).squeeze(0)
            prev_hidden = yi
            if idx + 1 < historical_steps:
                prev_t, t_i = past_time_steps[t], past_time_steps[t - 1]
            latent_ys.append(yi)
            gs.append(diff_noise)
    return torch.stack(latent_ys), torch.stack(gs), queries
