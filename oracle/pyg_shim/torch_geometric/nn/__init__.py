"""PyG nn symbols used by the reference, restated from PyG's published semantics.

MessagePassing.propagate (flow "source_to_target"):
    x_j = x_src[edge_index[0]], x_i = x_dst[edge_index[1]]; messages are reduced on
    edge_index[1] into x_dst.size(0) rows with `aggr` in {add, mean, max}; rows that
    receive no message are 0.
PointTransformerConv.forward(x=(x_src, x_dst), pos=(pos_src, pos_dst), edge_index):
    alpha = (lin_src(x_src), lin_dst(x_dst)); x = (lin(x_src), x_dst)
    message: delta = pos_nn(pos_i - pos_j)
             a = attn_nn(alpha_i - alpha_j + delta)
             a = softmax(a, index)                 # per destination row, per channel
             return a * (x_j + delta)              # aggr = add
    lin / lin_src / lin_dst are bias-free.
AttentionalAggregation(gate_nn): softmax(gate_nn(x), index) * x summed per index.
"""
import inspect
import torch
import torch.nn as tnn

from ..utils import scatter, softmax
from . import norm  # noqa: F401  (reference does `import torch_geometric.nn.norm as norm`)
from .norm import GraphNorm  # noqa: F401


class Linear(tnn.Linear):
    """PyG's dense Linear: same parameters/state-dict keys as nn.Linear."""

    def __init__(self, in_channels, out_channels, bias=True, **kwargs):
        super().__init__(in_channels, out_channels, bias=bias)

    def _lazy_load_hook(self, *args, **kwargs):      # PyG registers this bound method as a load_state_dict pre-hook
        return None


class MessagePassing(tnn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2, **kwargs):
        super().__init__()
        assert flow == "source_to_target"
        self.aggr = "add" if aggr == "sum" else aggr
        self.node_dim = node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        j, i = edge_index[0], edge_index[1]
        params = list(inspect.signature(self.message).parameters.keys())

        def pick(val, side):  # side 0 = source (j), 1 = destination (i)
            if isinstance(val, (tuple, list)):
                return val[side]
            return val

        dim_size = None
        for val in kwargs.values():
            dst = pick(val, 1)
            if isinstance(dst, torch.Tensor):
                dim_size = dst.size(0)
                break
        if size is not None and size[1] is not None:
            dim_size = size[1]

        msg_kwargs = {}
        for p in params:
            if p.endswith("_j"):
                msg_kwargs[p] = pick(kwargs[p[:-2]], 0).index_select(0, j)
            elif p.endswith("_i"):
                if p == "size_i":
                    msg_kwargs[p] = dim_size
                else:
                    msg_kwargs[p] = pick(kwargs[p[:-2]], 1).index_select(0, i)
            elif p == "index":
                msg_kwargs[p] = i
            elif p == "ptr":
                msg_kwargs[p] = None
            elif p in kwargs:
                msg_kwargs[p] = kwargs[p]
        msg = self.message(**msg_kwargs)
        reduce = {"add": "sum", "mean": "mean", "max": "max"}[self.aggr]
        out = scatter(msg, i, 0, dim_size, reduce)
        return self.update(out)

    def message(self, x_j):
        return x_j

    def update(self, out):
        return out


class PointTransformerConv(MessagePassing):
    def __init__(self, in_channels, out_channels, pos_nn=None, attn_nn=None,
                 add_self_loops=True, **kwargs):
        super().__init__(aggr="add", **kwargs)
        if isinstance(in_channels, int):
            in_channels = (in_channels, in_channels)
        assert not add_self_loops, "reference passes add_self_loops=False (graph_vs.py:200)"
        assert pos_nn is not None and attn_nn is not None
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.add_self_loops = add_self_loops
        self.pos_nn = pos_nn
        self.attn_nn = attn_nn
        self.lin = Linear(in_channels[0], out_channels, bias=False)
        self.lin_src = Linear(in_channels[0], out_channels, bias=False)
        self.lin_dst = Linear(in_channels[1], out_channels, bias=False)

    def forward(self, x, pos, edge_index):
        if isinstance(x, torch.Tensor):
            alpha = (self.lin_src(x), self.lin_dst(x))
            x = (self.lin(x), x)
        else:
            alpha = (self.lin_src(x[0]), self.lin_dst(x[1]))
            x = (self.lin(x[0]), x[1])
        if isinstance(pos, torch.Tensor):
            pos = (pos, pos)
        return self.propagate(edge_index, x=x, pos=pos, alpha=alpha, size=None)

    def message(self, x_j, pos_i, pos_j, alpha_i, alpha_j, index, ptr, size_i):
        delta = self.pos_nn(pos_i - pos_j)
        alpha = alpha_i - alpha_j + delta
        alpha = self.attn_nn(alpha)
        alpha = softmax(alpha, index, ptr, size_i)
        return alpha * (x_j + delta)


class AttentionalAggregation(tnn.Module):
    def __init__(self, gate_nn, nn=None):
        super().__init__()
        self.gate_nn = gate_nn
        self.nn = nn

    def forward(self, x, index=None, ptr=None, dim_size=None, dim=-2):
        if index is None:
            index = torch.zeros(x.size(0), dtype=torch.long, device=x.device)
        gate = self.gate_nn(x)
        if self.nn is not None:
            x = self.nn(x)
        gate = softmax(gate, index, ptr, dim_size)
        return scatter(gate * x, index, 0, dim_size, "sum")


GlobalAttention = AttentionalAggregation


class _ImportOnly(tnn.Module):
    """Names the reference's ablation modules import at module level (cns/ablation/structure/*); the ablations are
    out of scope, so these exist for `import cns.benchmark.pipeline` to succeed and refuse to be built."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("%s is not restated by the PyG shim (ablation models are out of scope)" % type(self).__name__)


class EdgeConv(_ImportOnly):
    pass


class GatedGraphConv(_ImportOnly):
    pass
