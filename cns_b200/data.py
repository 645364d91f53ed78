"""PyG-free `GraphData` / `Batch` containers with the reference's field names.

Drop-in for `cns.midend.graph_gen.GraphData` (reference graph_gen.py:68-101) and for
`torch_geometric.data.Batch.from_data_list` as the reference uses it
(sim/dataset.py:208).  Only the behaviour the hot path relies on is kept:

* attribute / item access, `hasattr(data, "batch")`, `.to(device)`, tolerating
  non-tensor extras (`intrinsic`, `cur_img`, ...) that the servo pipeline attaches
  (benchmark/pipeline.py:150-166);
* defaults `new_scene=False`, `tPo_norm=1.0` (graph_gen.py:71-74);
* collation offsets (graph_gen.py:78-95): keys containing `l0_dense_edge_index`,
  `l0_to_l1_edge_index_j`, `cluster_centers_index` advance by `num_nodes`; keys
  containing `l1_dense_edge_index`, `l0_to_l1_edge_index_i`, `belong_cluster_index`
  advance by `num_clusters`; any other key containing "index" advances by
  `num_nodes` (PyG default); keys containing "index" concatenate on the last dim,
  everything else on dim 0; 0-d tensors become (B,) vectors.

Unlike PyG the collation is vectorised: one `torch.cat` per field plus one
`repeat_interleave` for the offsets instead of B python-level additions.
"""
import copy
from typing import Any, Dict, Iterable, List

import torch

_INC_BY_NODES = ("l0_dense_edge_index", "l0_to_l1_edge_index_j", "cluster_centers_index")
_INC_BY_CLUSTERS = ("l1_dense_edge_index", "l0_to_l1_edge_index_i", "belong_cluster_index")


class GraphData(object):
    """One servo frame: target/current keypoints plus the two-level graph."""

    def __init__(self, **fields):
        object.__setattr__(self, "_f", {})
        if "new_scene" not in fields:
            fields["new_scene"] = torch.tensor(False)
        if "tPo_norm" not in fields:
            fields["tPo_norm"] = torch.tensor(1.0, dtype=torch.float32)
        for k, v in fields.items():
            if v is not None:
                self._f[k] = v

    # -- attribute plumbing ---------------------------------------------------
    def __getattr__(self, key):
        try:
            return object.__getattribute__(self, "_f")[key]
        except KeyError:
            raise AttributeError(key) from None

    def __setattr__(self, key, value):
        self[key] = value

    def __getitem__(self, key):
        return self._f[key]

    def __setitem__(self, key, value):
        f = self._f
        # `l1_complete` says "the cluster-level edge lists are the complete digraphs GraphGenerator emitted"
        # (cns_batch.flags); an edge list or mask assigned afterwards (custom edge augmentation) voids the claim,
        # and the forward then takes the general CSR path
        if key.startswith("l1_dense_edge_index") or key == "cluster_mask":
            if key in f and f.get("l1_complete") is True:
                f["l1_complete"] = False
        f[key] = value

    def __contains__(self, key):
        return key in self._f

    @property
    def keys(self) -> List[str]:
        return list(self._f.keys())

    def items(self):
        return self._f.items()

    def __repr__(self):
        body = ", ".join(
            "%s=%s" % (k, list(v.shape) if isinstance(v, torch.Tensor) else type(v).__name__)
            for k, v in self._f.items())
        return "%s(%s)" % (type(self).__name__, body)

    # -- reference API --------------------------------------------------------
    def start_new_scene(self, is_new: bool = True):
        self._f["new_scene"] = torch.tensor(bool(is_new))

    def set_distance_scale(self, dist_scale):
        self._f["tPo_norm"] = torch.tensor(float(dist_scale), dtype=torch.float32)

    def to(self, device, non_blocking: bool = False, fields=None):
        """Like PyG's Data.to.  `fields` (optional) restricts the copy to the named tensor fields - e.g.
        `GraphVS.INPUT_FIELDS`, the tensors the policy reads; other tensor fields are dropped."""
        out = copy.copy(self)
        object.__setattr__(out, "_f", {
            k: (v.to(device, non_blocking=non_blocking) if isinstance(v, torch.Tensor) else v)
            for k, v in self._f.items()
            if fields is None or k in fields or not isinstance(v, torch.Tensor)})
        return out

    def pin_memory(self):
        out = copy.copy(self)
        object.__setattr__(out, "_f", {
            k: (v.pin_memory() if isinstance(v, torch.Tensor) else v)
            for k, v in self._f.items()})
        return out

    def clone(self):
        out = copy.copy(self)
        object.__setattr__(out, "_f", {
            k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in self._f.items()})
        return out

    # -- collation rules ------------------------------------------------------
    def __cat_dim__(self, key: str, value=None) -> int:
        return -1 if "index" in key else 0

    def __inc__(self, key: str, value=None):
        for cand in _INC_BY_NODES:
            if cand in key:
                return int(self.num_nodes)
        for cand in _INC_BY_CLUSTERS:
            if cand in key:
                return int(self.num_clusters)
        return int(self.num_nodes) if "index" in key else 0


class DevicePrefetcher(object):
    """Iterates host (ideally pinned) GraphData/Batch objects as device-resident ones, issuing the
    host->device copy of item i+1 on a side stream while the caller computes on item i.

    LIFETIME: by default the tensors of a yielded item are views into a ring of `depth + 2` persistent staging
    slots: an item stays valid for the next `depth + 1` calls of `next()` and is then overwritten.  Streaming
    inference (use, then drop) is the intended consumer.  Pass `owning=True` when items are kept longer - e.g. the
    frames of a TBPTT window whose backward runs after the whole window: every item then gets fresh device tensors."""

    def __init__(self, items, device, depth: int = 1, fields=None, owning: bool = False):
        self.items, self.device, self.depth = iter(items), torch.device(device), max(1, depth)
        self.fields, self.owning = fields, owning
        self.stream = torch.cuda.Stream(self.device)
        self.queue = []
        # persistent staging slots (no allocator traffic in steady state): slot -> {field: byte buffer}
        self.n_slots = self.depth + 2
        self.slots = [dict() for _ in range(self.n_slots)]
        self.slot_free = [None] * self.n_slots          # event: the consumer is done with this slot
        self.next_slot = 0
        self.last = None                                # (slot index) handed out most recently

    def _stage(self, host, slot):
        bufs = self.slots[slot]
        out = {}
        for k, v in host.items():
            if not isinstance(v, torch.Tensor):
                out[k] = v
                continue
            if self.fields is not None and k not in self.fields:
                continue
            if self.owning:
                out[k] = v.to(self.device, non_blocking=True)
                continue
            nbytes = v.numel() * v.element_size()
            buf = bufs.get(k)
            if buf is None or buf.numel() < nbytes:
                buf = torch.empty(max(int(nbytes * 1.25), 16), dtype=torch.uint8, device=self.device)
                bufs[k] = buf
            dst = buf[:nbytes].view(v.dtype).view(v.shape)
            dst.copy_(v, non_blocking=True)
            out[k] = dst
        return type(host)(**out)

    def _fill(self):
        while len(self.queue) < self.depth + 1:
            try:
                host = next(self.items)
            except StopIteration:
                return
            slot = self.next_slot
            self.next_slot = (self.next_slot + 1) % self.n_slots
            with torch.cuda.stream(self.stream):
                if self.slot_free[slot] is not None:
                    self.stream.wait_event(self.slot_free[slot])     # its previous consumer has finished
                dev = self._stage(host, slot)
            ev = torch.cuda.Event()
            ev.record(self.stream)
            self.queue.append((dev, ev, slot))

    def __iter__(self):
        return self

    def __next__(self):
        cur = torch.cuda.current_stream(self.device)
        if self.last is not None:          # everything the caller enqueued on the previous item is ordered before this
            done = torch.cuda.Event()
            done.record(cur)
            self.slot_free[self.last] = done
        self._fill()
        if not self.queue:
            raise StopIteration
        dev, ev, slot = self.queue.pop(0)
        cur.wait_event(ev)
        if self.owning:          # allocated on the copy stream, consumed on the caller's: tell the caching allocator
            for _, v in dev.items():
                if isinstance(v, torch.Tensor):
                    v.record_stream(cur)
        self.last = slot
        self._fill()
        return dev


def _inc_kind(key: str) -> int:
    """0: no offset, 1: by num_nodes, 2: by num_clusters."""
    for cand in _INC_BY_NODES:
        if cand in key:
            return 1
    for cand in _INC_BY_CLUSTERS:
        if cand in key:
            return 2
    return 1 if "index" in key else 0


class Batch(GraphData):
    """B graphs concatenated with index offsets; adds `batch`, `ptr`, `cluster_ptr`."""

    def __init__(self, **fields):
        object.__setattr__(self, "_f", {})
        for k, v in fields.items():
            self._f[k] = v

    @property
    def num_graphs(self) -> int:
        return int(self._f["num_clusters"].numel())

    @classmethod
    def from_data_list(cls, data_list: Iterable[GraphData]) -> "Batch":
        data_list = list(data_list)
        assert len(data_list) > 0
        nb = len(data_list)
        n_nodes = torch.tensor([int(d.num_nodes) for d in data_list], dtype=torch.long)
        n_clu = torch.tensor([int(d.num_clusters) for d in data_list], dtype=torch.long)
        node_off = torch.cumsum(n_nodes, 0) - n_nodes
        clu_off = torch.cumsum(n_clu, 0) - n_clu

        out: Dict[str, Any] = {}
        for key in data_list[0].keys:
            if key == "num_nodes":
                continue
            vals = [d[key] for d in data_list]
            v0 = vals[0]
            if isinstance(v0, torch.Tensor):
                vals = [v.unsqueeze(0) if v.dim() == 0 else v for v in vals]
                cat = torch.cat(vals, dim=-1 if "index" in key else 0)
                kind = _inc_kind(key)
                if kind:
                    lens = torch.tensor([v.shape[-1] for v in vals], dtype=torch.long)
                    off = torch.repeat_interleave(node_off if kind == 1 else clu_off, lens)
                    cat = cat + off  # broadcasts over the leading 2 of (2, E) edge lists
                out[key] = cat
            elif key == "l1_complete":
                out[key] = all(bool(v) for v in vals)
            elif isinstance(v0, (int, float)) and not isinstance(v0, bool):
                out[key] = torch.tensor(vals)
            else:
                out[key] = vals
        out["max_clusters_per_graph"] = int(n_clu.max())
        out["num_nodes"] = int(n_nodes.sum())
        out["batch"] = torch.repeat_interleave(torch.arange(nb), n_nodes)
        ptr = torch.zeros(nb + 1, dtype=torch.long)
        ptr[1:] = torch.cumsum(n_nodes, 0)
        out["ptr"] = ptr
        return cls(**out)
