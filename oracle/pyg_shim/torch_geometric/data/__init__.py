"""`Data` / `Batch` with the PyG collation rules the reference relies on.

Restated behaviour (PyG 2.x `data/data.py`, `data/collate.py`):
  * Data(x, edge_index, edge_attr, y, pos, **kwargs) stores every non-None field as
    an attribute; `num_nodes` may be given explicitly.
  * __cat_dim__(key, value): -1 when "index" is in the key (or key == "face"), else 0.
  * __inc__(key, value): num_nodes when "index" is in the key (or "face"), else 0.
  * Batch.from_data_list: per key, tensors are offset by the running sum of __inc__
    and concatenated along __cat_dim__; 0-dim tensors are first unsqueezed to (1,)
    so a batch of B scalars becomes a (B,) tensor; python numbers become a tensor of
    length B; other objects become python lists.  `num_nodes` is summed.  `batch`
    (node -> graph id) and `ptr` are added.
"""
import copy
import torch


class Data(object):
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        object.__setattr__(self, "_store", {})
        for k, v in (("x", x), ("edge_index", edge_index), ("edge_attr", edge_attr),
                     ("y", y), ("pos", pos)):
            if v is not None:
                self._store[k] = v
        for k, v in kwargs.items():
            self._store[k] = v

    # ---- attribute plumbing -------------------------------------------------
    def __getattr__(self, key):
        store = object.__getattribute__(self, "_store")
        if key in store:
            return store[key]
        raise AttributeError(key)

    def __setattr__(self, key, value):
        self._store[key] = value

    def __getitem__(self, key):
        return self._store[key]

    def __setitem__(self, key, value):
        self._store[key] = value

    def __contains__(self, key):
        return key in self._store

    @property
    def keys(self):
        return list(self._store.keys())

    def __cat_dim__(self, key, value, *args, **kwargs):
        return -1 if ("index" in key or key == "face") else 0

    def __inc__(self, key, value, *args, **kwargs):
        return self.num_nodes if ("index" in key or key == "face") else 0

    def to(self, device, *args, **kwargs):
        out = copy.copy(self)
        object.__setattr__(out, "_store", {
            k: (v.to(device, *args, **kwargs) if isinstance(v, torch.Tensor) else v)
            for k, v in self._store.items()})
        return out

    def clone(self):
        out = copy.copy(self)
        object.__setattr__(out, "_store", {
            k: (v.clone() if isinstance(v, torch.Tensor) else copy.deepcopy(v))
            for k, v in self._store.items()})
        return out

    def __repr__(self):
        parts = []
        for k, v in self._store.items():
            parts.append("%s=%s" % (k, list(v.shape) if isinstance(v, torch.Tensor) else v))
        return "%s(%s)" % (type(self).__name__, ", ".join(parts))


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list, follow_batch=None, exclude_keys=None):
        assert len(data_list) > 0
        first = data_list[0]
        out = cls()
        object.__setattr__(out, "_data_cls", type(first))
        keys = list(first._store.keys())
        num_nodes = [int(d.num_nodes) for d in data_list]
        for key in keys:
            if key == "num_nodes":
                continue
            vals = [d._store[key] for d in data_list]
            v0 = vals[0]
            if isinstance(v0, torch.Tensor):
                cat_dim = first.__cat_dim__(key, v0)
                pieces, running = [], 0
                for d, v in zip(data_list, vals):
                    if v.dim() == 0:
                        v = v.unsqueeze(0)
                    inc = d.__inc__(key, v)
                    if isinstance(inc, torch.Tensor):
                        inc = int(inc)
                    if running != 0:
                        v = v + running
                    running += inc
                    pieces.append(v)
                out._store[key] = torch.cat(pieces, dim=cat_dim if pieces[0].dim() > 0 else 0)
            elif isinstance(v0, (int, float)) and not isinstance(v0, bool):
                out._store[key] = torch.tensor(vals)
            else:
                out._store[key] = vals
        out._store["num_nodes"] = sum(num_nodes)
        out._store["batch"] = torch.repeat_interleave(
            torch.arange(len(data_list)), torch.tensor(num_nodes))
        ptr = torch.zeros(len(data_list) + 1, dtype=torch.long)
        ptr[1:] = torch.cumsum(torch.tensor(num_nodes), 0)
        out._store["ptr"] = ptr
        object.__setattr__(out, "_num_graphs", len(data_list))
        return out

    @property
    def num_graphs(self):
        return object.__getattribute__(self, "_num_graphs")
