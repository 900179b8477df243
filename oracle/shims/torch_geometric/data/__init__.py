"""Functional stand-in for the parts of torch_geometric.data the reference's encoder stage uses (enc_hivt_nusargo_sde_sep2.py:107-121,
models/utils/util.py:20-75): a dict-with-attributes ``Data`` and ``Batch.from_data_list`` (concatenation with node-offset
``edge_index``).  TEST INFRASTRUCTURE ONLY — restated from PyG 2.2's documented behaviour, not a copy of it."""
import torch


class Data:
    def __init__(self, **kw):
        self.__dict__['_store'] = {}
        for k, v in kw.items():
            self._store[k] = v

    # item and attribute access reach the same store (PyG semantics the reference relies on: data.x, data['padding_mask'], data[f'edge_index_{t}'] = ...)
    def __getattr__(self, k):
        try:
            return self.__dict__['_store'][k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self._store[k] = v

    def __getitem__(self, k):
        return self._store[k]

    def __setitem__(self, k, v):
        self._store[k] = v

    def __contains__(self, k):
        return k in self._store

    def keys(self):
        return self._store.keys()

    def __inc__(self, key, value, *args, **kwargs):
        return self.num_nodes if 'index' in key else 0


class Batch(Data):
    @staticmethod
    def from_data_list(lst):
        """x / edge_attr concatenated along dim 0, edge_index along dim 1 with each graph's node offset added (num_nodes per graph)."""
        xs, eis, eas, off = [], [], [], 0
        for d in lst:
            xs.append(d.x)
            eis.append(d.edge_index + off)
            eas.append(d.edge_attr)
            off += int(d.num_nodes)
        return Batch(x=torch.cat(xs, 0), edge_index=torch.cat(eis, 1), edge_attr=torch.cat(eas, 0), num_nodes=off)


class Dataset:
    def __init__(self, *a, **k):
        pass


class DataLoader:
    def __init__(self, *a, **k):
        raise NotImplementedError("stub")
