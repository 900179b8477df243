"""Functional stand-in for torch_geometric.nn.conv.MessagePassing as the reference's graph-attention stages use it (AAEncoder / ALEncoder
enc_hivt_nusargo_sde_sep2.py:498-614,693-790; GlobalInteractorLayer models/aggregators/agg_hivt.py:61-135): ``aggr='add'``, ``node_dim=0``,
flow source -> target.  ``propagate`` lifts ``<name>_i`` / ``<name>_j`` arguments of ``message`` from the node tensors by target / source
index (tuples: ``_j`` from element 0, ``_i`` from element 1), injects ``index`` / ``ptr`` / ``size_i``, sums the messages per target node
and hands the sum to ``update``.  TEST INFRASTRUCTURE ONLY — restated from PyG's documented contract."""
import inspect

import torch
import torch.nn as nn


class MessagePassing(nn.Module):
    def __init__(self, aggr='add', node_dim=-2, flow='source_to_target', **kwargs):
        super().__init__()
        assert aggr == 'add' and flow == 'source_to_target'
        self.aggr, self.node_dim = aggr, node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        assert self.node_dim == 0
        src, dst = edge_index[0], edge_index[1]

        def n_nodes(which):                                    # 0: source side, 1: target side
            if size is not None and size[which] is not None:
                return int(size[which])
            for v in kwargs.values():
                if isinstance(v, (tuple, list)) and torch.is_tensor(v[which]):
                    return v[which].size(0)
                if torch.is_tensor(v) and v.dim() >= 1 and not isinstance(v, (tuple, list)):
                    return v.size(0)
            raise ValueError("cannot infer the number of nodes")

        n_dst = None
        for v in kwargs.values():
            if isinstance(v, (tuple, list)):
                n_dst = v[1].size(0)
                break
        if n_dst is None:
            n_dst = n_nodes(1)
        msg_args = {}
        for name in inspect.signature(self.message).parameters:
            if name.endswith('_i') or name.endswith('_j'):
                base, tgt = name[:-2], name.endswith('_i')
                if base in kwargs:
                    v = kwargs[base]
                    if isinstance(v, (tuple, list)):
                        v = v[1] if tgt else v[0]
                    msg_args[name] = v.index_select(0, dst if tgt else src)
                    continue
            if name == 'edge_index':
                msg_args[name] = edge_index
            elif name == 'index':
                msg_args[name] = dst
            elif name == 'ptr':
                msg_args[name] = None
            elif name == 'size_i':
                msg_args[name] = n_dst
            elif name in kwargs:
                msg_args[name] = kwargs[name]
        out = self.message(**msg_args)
        agg = torch.zeros((n_dst,) + tuple(out.shape[1:]), dtype=out.dtype, device=out.device).index_add(0, dst, out)
        upd_args = {name: kwargs[name] for name in list(inspect.signature(self.update).parameters)[1:] if name in kwargs}
        return self.update(agg, **upd_args)

    def message(self, x_j):
        return x_j

    def update(self, inputs):
        return inputs
