"""Segment primitives with PyG / torch_scatter semantics (restated, not copied).

scatter(src, index, dim_size, reduce): rows of `src` reduced into `dim_size` rows by
`index`; EMPTY segments yield 0 for every reduce (sum/mean/max) - that is what
torch_scatter and PyG's `scatter` return.
softmax(src, index): exp(src - segmax) / (segsum + 1e-16)   (PyG utils.softmax).
"""
import torch


def _dim_size(index, dim_size):
    if dim_size is not None:
        return int(dim_size)
    return int(index.max()) + 1 if index.numel() > 0 else 0


def scatter(src, index, dim=0, dim_size=None, reduce="sum"):
    assert dim in (0, -2) or src.dim() == 1
    n = _dim_size(index, dim_size)
    shape = (n,) + tuple(src.shape[1:])
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    if reduce in ("sum", "add"):
        return torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(0, idx, src)
    if reduce == "mean":
        s = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(0, idx, src)
        cnt = torch.zeros(n, dtype=src.dtype, device=src.device).scatter_add_(
            0, index, torch.ones_like(index, dtype=src.dtype))
        cnt = cnt.clamp(min=1).view(-1, *([1] * (src.dim() - 1)))
        return s / cnt
    if reduce == "max":
        out = torch.zeros(shape, dtype=src.dtype, device=src.device)
        # include_self=False: rows that receive nothing keep the initial 0.
        return out.scatter_reduce(0, idx, src, reduce="amax", include_self=False)
    raise ValueError(reduce)


def softmax(src, index, ptr=None, num_nodes=None, dim=0):
    n = _dim_size(index, num_nodes)
    seg_max = scatter(src.detach(), index, 0, n, "max")
    out = (src - seg_max.index_select(0, index)).exp()
    seg_sum = scatter(out, index, 0, n, "sum") + 1e-16
    return out / seg_sum.index_select(0, index)
