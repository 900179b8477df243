import torch.nn as nn


class MessagePassing(nn.Module):
    def __init__(self, aggr='add', node_dim=-2, **kwargs):
        super().__init__()
        self.aggr, self.node_dim = aggr, node_dim

    def propagate(self, *a, **k):
        raise NotImplementedError("stub: PyG message passing is out of scope")
