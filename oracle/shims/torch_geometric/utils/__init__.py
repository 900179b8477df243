"""Functional stand-ins for torch_geometric.utils.softmax / subgraph as the reference calls them.  TEST INFRASTRUCTURE ONLY."""
import torch


def softmax(src, index, ptr=None, num_nodes=None, dim=0):
    """Softmax over the entries that share a target node (segment softmax), numerically stabilised per segment."""
    n = int(num_nodes) if num_nodes is not None else (int(index.max()) + 1 if index.numel() else 0)
    shape = (n,) + tuple(src.shape[1:])
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    mx = torch.full(shape, float('-inf'), dtype=src.dtype, device=src.device).scatter_reduce(0, idx, src, reduce='amax', include_self=True)
    ex = torch.exp(src - mx.gather(0, idx))
    den = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add(0, idx, ex)
    return ex / (den.gather(0, idx) + 1e-16)


def subgraph(subset, edge_index, edge_attr=None, relabel_nodes=False, num_nodes=None):
    """Edges whose two end points both lie in ``subset`` (a boolean node mask here, as in enc…sep2.py:108)."""
    assert subset.dtype == torch.bool and not relabel_nodes
    keep = subset[edge_index[0]] & subset[edge_index[1]]
    return edge_index[:, keep], (edge_attr[keep] if edge_attr is not None else None)
