"""YAML-selectable stage classes (the second plugin point — ``encoder.file_path/module_name``, ``decoder.file_path/
module_name`` of configs/nusargo/hivt_nuSArgo_sdesepenc_sdedec.yml:24-26,61-63).

The reference builds its stages with ``SourceFileLoader(module_name, file_path).load_module(module_name)`` and ``getattr(module,
module_name)(**kwargs)`` (models/model_base_mix_sde.py:38-45).  ``trajsde_b200/plugins/*.py`` are such files: they subclass the reference's
own stage classes — same ``__init__``, same parameter names, so reference checkpoints load unchanged — and take their ``forward`` from
the mixins below, which replace

  * the decoder's aggr_embed / sdeint / heads sequence (dec_hivt_nusargo_sde.py:82-99) by the fused prologue, ONE persistent solve and
    the fused heads (differentiable; the gradient returns in the solver's layout, no slice / permute copies);
  * the encoder's 21-iteration Python loop (enc_hivt_nusargo_sde_sep2.py:128-196: 21 x [sdeint_dual + host sync + ``exit()`` check +
    GRU jump + gather]) by ONE launch of the fused recurrence and three gathers; ``forward_ood`` (:204-370) by one launch over
    10 x rows.

What stays with the subclass (``_prepare`` / ``_finish``) is exactly the HiVT graph-attention work the scope table leaves on the
reference path: the AA encoder before the loop and the AL encoder after it.
"""
from typing import Optional

import torch
import torch.nn.functional as F

from . import stage
from .encoder import encoder_recurrence, encoder_recurrence_ood, eos_gather
from .heads import solve_and_heads


class FusedDecoderMixin:
    """``forward`` of ``SDEDecoder`` (dec_hivt_nusargo_sde.py:77-105) over the fused operators.  Expects the reference's attributes:
    ``aggr_embed``, ``lsde_func``, ``decoder``, ``scale`` (if ``uncertain``), ``pi``, ``ts_pred``, ``num_modes``, ``future_steps``,
    ``min_stepsize``, ``min_scale``, ``rtol``, ``atol``, ``method``."""

    solver_kwargs: dict = {}          # e.g. {'mode': 'exact'} or a fixed 'seed'; {'bm': dW} for parity runs

    def forward(self, data, local_embed: torch.Tensor, global_embed: torch.Tensor):
        num_actors = local_embed.shape[0]
        hidden_0 = stage.aggr_embed(self.aggr_embed, local_embed, global_embed)                                   # :82-85
        if self.method != 'euler':
            raise NotImplementedError(f"fused solve implements method='euler' only (reference yml:76), got {self.method!r}")
        # :88, :95-100 as one node: with the scale head the heads kernel writes out['loc'] = cat(loc, elu(scale) + 1 + min_scale) itself
        loc, _ = solve_and_heads(self.lsde_func, self.decoder, self.scale if self.uncertain else None, hidden_0, self.ts_pred,
                                 self.min_stepsize, cat_min_scale=float(self.min_scale) if self.uncertain else None, **self.solver_kwargs)
        pi = stage.pi_head(self.pi, local_embed, global_embed)                                                    # :92-94
        out = {'loc': loc.view(self.num_modes, num_actors, self.future_steps, 4 if self.uncertain else 2), 'pi': pi}
        out['reg_mask'] = ~data['padding_mask'][:, -self.future_steps:]                                           # :104
        return out


class FusedEncoderMixin:
    """``forward`` / ``forward_ood`` of ``LocalEncoderSDESepPara2`` around the fused recurrence.  The subclass provides

        _prepare(data, ood: bool) -> dict(aa_out [21, R, 64], actors_mask [R, 21] bool, nus_mask [R] bool,
                                          agent_index [B] long, n_fake: int, ...anything _finish needs)
        _finish(data, prep, out [N, 64]) -> the stage's first return value (the reference applies its AL encoder here, :198-200)

    where R = N + n_fake rows (actors + one perturbed copy per target agent, :94-103; n_fake = 0 under ``forward_ood``).  Expects the
    reference's attributes ``hidden``, ``lsde_func``, ``gru_unit``, ``ref_time``, ``minimum_step``, ``max_past_t``, ``run_backwards``,
    ``real_label``, ``fake_label``."""

    recurrence_kwargs: dict = {}      # e.g. {'dW': increments} for parity runs, {'seed': ...}

    def _check(self):
        if not getattr(self, 'run_backwards', True):
            raise NotImplementedError("the fused recurrence implements run_backwards=True (the reference configuration, yml:39)")

    def forward(self, data):
        self._check()
        prep = self._prepare(data, ood=False)
        aa_out, actors_mask, nus_mask = prep['aa_out'], prep['actors_mask'], prep['nus_mask']
        agent_index, n_fake = prep['agent_index'], int(prep['n_fake'])
        rows = aa_out.shape[1]
        n_actors = rows - n_fake
        h0 = self.hidden.unsqueeze(0).repeat(rows, 1)                                                             # :78
        latent, g = encoder_recurrence(self.lsde_func, self.gru_unit, h0, aa_out, actors_mask, nus_mask, dt=self.minimum_step,
                                       max_past_t=float(self.max_past_t), **self.recurrence_kwargs)              # :128-182 in one launch
        dev = aa_out.device
        eos_idcs = self.ref_time - torch.argmax(data['bos_mask'].float(), dim=1)                                  # :187
        out = latent[eos_idcs, torch.arange(n_actors, device=dev), :]                                             # :184, :188 (fake rows dropped)
        new_agent_index = torch.cat((agent_index, torch.arange(n_actors, rows, device=dev)))                      # :101
        agent_eos = eos_idcs[agent_index].repeat(2)                                                               # :190
        diff = g[agent_eos, new_agent_index]                                                                      # :171 + :191 -> [2B]
        diff = diff.unsqueeze(-1).expand(-1, aa_out.shape[2])                                                     # the reference's 64 identical columns (:480-481)
        diffusions_in, diffusions_out = torch.chunk(diff, 2, 0)                                                   # :194
        in_labels = torch.full_like(diffusions_in, self.real_label)
        out_labels = torch.full_like(diffusions_out, self.fake_label)
        return self._finish(data, prep, out), diffusions_in, diffusions_out, in_labels, out_labels

    def forward_ood(self, data, eval_iter: int = 10):
        self._check()
        prep = self._prepare(data, ood=True)
        mean, std = encoder_recurrence_ood(self.lsde_func, self.gru_unit, prep['aa_out'], prep['actors_mask'], prep['nus_mask'],
                                           data['bos_mask'], eval_iter=eval_iter, dt=self.minimum_step, max_past_t=float(self.max_past_t),
                                           ref_time=self.ref_time, **{k: v for k, v in self.recurrence_kwargs.items() if k == 'seed'})
        return self._finish(data, prep, mean), std                                                                # :252-313, :315-317
