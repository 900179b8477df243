"""YAML-selectable encoder stage: the reference's ``LocalEncoderSDESepPara2`` (models/encoders/enc_hivt_nusargo_sde_sep2.py:25-370) with
``forward`` / ``forward_ood`` taken from ``FusedEncoderMixin``: same constructor and parameters; the AA encoder before the recurrence
and the AL encoder after it run exactly as in the reference (``_prepare`` / ``_finish`` below follow :68-127, :198-200, :206-250), the
21-iteration loop between them is one fused launch."""
import torch
from torch_geometric.data import Batch, Data
from torch_geometric.utils import subgraph

from models.encoders.enc_hivt_nusargo_sde_sep2 import LocalEncoderSDESepPara2     # the reference repository must be on sys.path

from trajsde_b200.stages import FusedEncoderMixin


class LocalEncoderSDESepPara2Fused(FusedEncoderMixin, LocalEncoderSDESepPara2):

    def _lane_feat(self, data):
        lane_len = (1 - data['lane_paddings']).sum(-1)                                                  # :68-71
        rows = torch.arange(data['lane_positions'].size(0))
        return data['lane_positions'][rows, (lane_len - 1).long(), :] - data['lane_positions'][rows, 0, :]

    def _prepare(self, data, ood: bool):
        if not self.parallel:
            raise NotImplementedError                                                                    # as the reference (:123-124)
        nus_mask = torch.isin(data.batch, torch.where(data.source == 0)[0])                              # :73-74
        actor_num = data.x.shape[0]
        pad = data['padding_mask']
        agent_index = data['agent_index']
        if ood:                                                                                          # :206-250: no perturbed copies
            x_actors, edge_all, positions, bos_mask, rotate_mat, num_nodes, n_fake = \
                data.x, data.edge_index, data['positions'], data['bos_mask'], data['rotate_mat'], data.num_nodes, 0
        else:                                                                                            # :86-103: one noisy copy per target agent
            to_agent = torch.isin(data.edge_index[1], agent_index)
            edge_from, edge_to = data.edge_index[0][to_agent], data.edge_index[1][to_agent]
            _, new_edge_to = torch.unique(edge_to, return_inverse=True)
            edge_all = torch.cat((data.edge_index, torch.stack((edge_from, new_edge_to + actor_num), 0)), -1)
            x_agent = data.x[agent_index]
            x_actors = torch.cat((data.x, x_agent + 2 * torch.randn_like(x_agent)), dim=0)
            pad = torch.cat((pad, pad[agent_index]), dim=0)
            positions = torch.cat((data['positions'], data['positions'][agent_index]), 0)
            bos_mask = torch.cat((data['bos_mask'], data['bos_mask'][agent_index]), 0)
            rotate_mat = torch.cat((data['rotate_mat'], data['rotate_mat'][agent_index]), 0)
            nus_mask = torch.cat((nus_mask, data.source == 0), dim=0)
            n_fake = agent_index.size(0)
            num_nodes = data.num_nodes + n_fake
        actors_mask = ~pad[:, :self.ref_time + 1]                                                        # :100
        snapshots = []
        for t in range(self.historical_steps):                                                           # :107-118
            edge_index_t, _ = subgraph(subset=~pad[:, t], edge_index=edge_all)
            edge_attr_t = positions[edge_index_t[0], t] - positions[edge_index_t[1], t]
            edge_index_t, edge_attr_t = self.drop_edge(edge_index_t, edge_attr_t)
            snapshots.append(Data(x=x_actors[:, t], edge_index=edge_index_t, edge_attr=edge_attr_t, num_nodes=num_nodes))
        batch = Batch.from_data_list(snapshots)
        aa_out = self.aa_encoder(x=batch.x, t=None, edge_index=batch.edge_index, edge_attr=batch.edge_attr, bos_mask=bos_mask,
                                 rotate_mat=rotate_mat)                                                  # :119-121
        aa_out = aa_out.view(self.historical_steps, aa_out.shape[0] // self.historical_steps, -1)
        return {'aa_out': aa_out, 'actors_mask': actors_mask, 'nus_mask': nus_mask, 'agent_index': agent_index, 'n_fake': n_fake}

    def _finish(self, data, prep, out):
        edge_index, edge_attr = self.drop_edge(data['lane_actor_index'], data['lane_actor_vectors'])     # :198-200
        return self.al_encoder(x=(self._lane_feat(data), out), edge_index=edge_index, edge_attr=edge_attr, rotate_mat=data['rotate_mat'])
