"""`GraphVS`: drop-in for `cns.models.graph_vs.GraphVS` whose forward runs on libcns_sm100.

Same constructor, `forward(data, hidden=None) -> (vel_si_vec, vel_si_norm, hidden)`,
`init_hidden`, classmethods `preprocess / postprocess / objectives`,
`get_parameter_groups()` and - because the parameter containers below reuse the
reference's attribute names - the same 43 state-dict keys (graph_vs.py:273-350), so
`load_state_dict(torch.load("checkpoints/cns_state_dict.pth"))` works unchanged.

The modules here only OWN parameters; no torch op computes anything on the hot path.
`forward` hands raw device pointers to `cns_graphvs_forward` (include/cns_b200.h).
There is no CPU implementation: calling it with CPU tensors raises.
"""
import ctypes as C
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from . import functions as F_


class MLP(nn.Module):
    """Parameter container with the reference's layout `fcs.{l}.{weight,bias}` (graph_vs.py:16-24)."""

    def __init__(self, channels, bias=True):
        super().__init__()
        assert len(channels) >= 2
        self.fcs = nn.ModuleList(
            nn.Linear(channels[i - 1], channels[i], bias=bias) for i in range(1, len(channels)))


class _PointTransformer(nn.Module):   # encoder.att_clu_conv
    def __init__(self, dim, pos_dim):
        super().__init__()
        self.pos_nn = MLP([pos_dim * 2, dim, dim])
        self.attn_nn = MLP([dim, dim, dim])
        self.lin = nn.Linear(dim, dim, bias=False)
        self.lin_src = nn.Linear(dim, dim, bias=False)
        self.lin_dst = nn.Linear(dim, dim, bias=False)


class _Encoder(nn.Module):
    def __init__(self, input_dim, dim, pos_dim):
        super().__init__()
        self.in_enc = MLP([input_dim, dim])
        self.att_clu_conv = _PointTransformer(dim, pos_dim)
        self.out_enc = MLP([dim * 2, dim])


class _PointEdgeConv(nn.Module):
    def __init__(self, in_dim, out_dim, pos_dim):
        super().__init__()
        self.msg_nn = MLP([in_dim * 2 + pos_dim, out_dim])


class GraphNormParams(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))
        self.mean_scale = nn.Parameter(torch.ones(dim))


class _PERConv(nn.Module):
    def __init__(self, dim, pos_dim):
        super().__init__()
        self.conv = _PointEdgeConv(dim, dim, pos_dim)
        self.bn = GraphNormParams(dim)
        self.fc = nn.Linear(dim, dim)


class _GRUCell(nn.Module):
    def __init__(self, dim, pos_dim):
        super().__init__()
        self.out_channels = dim
        self.conv_gates = _PointEdgeConv(2 * dim, 2 * dim, pos_dim)
        self.conv_candi = _PointEdgeConv(2 * dim, dim, pos_dim)


class _Backbone(nn.Module):
    def __init__(self, dim, pos_dim):
        super().__init__()
        self.l1_conv0 = _PERConv(dim, pos_dim)
        self.temporal_aggr = _GRUCell(dim, pos_dim)
        self.l1_conv1 = _PERConv(dim, pos_dim)


class _Pool(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.gate_nn = nn.Linear(dim, 1)


class _Decoder(nn.Module):
    def __init__(self, dim, regress_norm):
        super().__init__()
        self.pool = _Pool(dim)
        self.vel_si_vec = MLP([dim, dim, 6])
        if regress_norm:
            self.vel_si_norm = MLP([dim, dim, 1])


class RawPred(tuple):
    """(vel_si_vec, vel_si_norm, hidden) that also carries the fused post-processed velocity."""
    vel = None
    tpo_ptr = None


def _field(data, key, default=None):
    try:
        return data[key] if isinstance(data, dict) else getattr(data, key)
    except (KeyError, AttributeError):
        return default


class TargetCache(object):
    """Owner of a `cns_target_cache` (include/cns_b200.h)."""

    def __init__(self, handle, b, dev, keep):
        self.lib, self.raw = _lib.load(), C.c_void_p()
        _lib.check(self.lib.cns_target_cache_create(handle.raw, C.byref(b), _lib.stream_ptr(dev), C.byref(self.raw)),
                   "cns_target_cache_create")

    def __del__(self):
        try:
            if self.raw:
                self.lib.cns_target_cache_destroy(self.raw)
        except Exception:
            pass


class GraphVS(nn.Module):
    # the GraphData / Batch tensors forward() reads (graph_vs.py:285-307); `vel_si` / `vel` ride along for the loss
    INPUT_FIELDS = ("x_cur", "x_tar", "pos_cur", "pos_tar", "l0_to_l1_edge_index_j_cur", "l0_to_l1_edge_index_i_cur",
                    "l1_dense_edge_index_cur", "l1_dense_edge_index_tar", "cluster_mask", "cluster_centers_index",
                    "num_clusters", "tPo_norm", "new_scene", "vel_si", "vel")

    def __init__(self, input_dim, pos_dim, hidden_dim=128, regress_norm=True, precision="fp32"):
        super().__init__()
        if (input_dim, pos_dim, hidden_dim) != (2, 2, 128) or not regress_norm:
            raise NotImplementedError(
                "libcns_sm100 is specialised for the shipped configuration "
                "GraphVS(2, 2, 128, regress_norm=True) (checkpoints/cns.pth, train_cns.py:56-61)")
        self.regress_norm = regress_norm
        self.encoder = _Encoder(input_dim, hidden_dim, pos_dim)
        self.backbone = _Backbone(hidden_dim, pos_dim)
        self.decoder = _Decoder(hidden_dim, regress_norm)
        self.hidden_dim = hidden_dim
        self.precision = precision
        self._handles = {}       # device index -> (_lib.Handle, weight version)
        self._workspaces = {}    # (device index, key) -> uint8 tensor
        self._retired_workspaces = []
        self._names = None
        self._param_cache = None

    # ------------------------------------------------------------------ reference API
    def init_hidden(self, size0):
        return torch.zeros(int(size0), self.hidden_dim)

    @classmethod
    def preprocess(cls, data):
        return data

    @classmethod
    def postprocess(cls, raw_pred, data):
        vel = getattr(raw_pred, "vel", None)
        tpo = _field(data, "tPo_norm")
        if vel is not None and tpo is not None and raw_pred.tpo_ptr == tpo.data_ptr():
            return vel.clone()
        return F_.postprocess(raw_pred, data, post_scale=True)

    @classmethod
    def objectives(cls, raw_pred, data):
        return F_.objectives(raw_pred, data, gt_si=True, weight=1)

    def get_parameter_groups(self):
        """(decay, no_decay) as graph_vs.py:340-350: Linear weights decay; biases and the
        GraphNorm weight / mean_scale do not."""
        decay, no_decay = [], []
        for m in self.modules():
            if isinstance(getattr(m, "bias", None), nn.Parameter):
                no_decay.append(m.bias)
            if isinstance(m, GraphNormParams):
                no_decay += [m.weight, m.mean_scale]
            elif isinstance(getattr(m, "weight", None), nn.Parameter):
                decay.append(m.weight)
        return decay, no_decay

    # ------------------------------------------------------------------ library plumbing
    def set_precision(self, precision: str):
        assert precision in _lib.PRECISIONS
        self.precision = precision
        for key, (h, _) in self._handles.items():
            if not isinstance(key, tuple):
                h.set_precision(precision)

    def _ordered_params(self):
        """The 43 parameters in the library's canonical order.  Cached: `nn.Module` swaps Parameter objects only in
        `_apply` (.to / .cuda / .float ...), which drops the cache; load_state_dict and optimizers write in place."""
        if self._param_cache is None:
            if self._names is None:
                self._names = _lib.param_names()
            sd = dict(self.named_parameters())
            self._param_cache = [sd[n] for n in self._names]
        return self._param_cache

    def _apply(self, fn, *args, **kwargs):
        self._param_cache = None
        out = super()._apply(fn, *args, **kwargs)
        self._param_cache = None
        return out

    def _handle(self, device: torch.device, training: bool = False) -> "_lib.Handle":
        params = self._ordered_params()
        # in-place writes bump _version, re-allocation changes the storage pointer: one pass over 43 tensors (~8 us)
        version = (sum([p._version for p in params]), params[0].data_ptr(), params[-1].data_ptr(), id(params))
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = (idx, "train") if training else idx       # training always uses the fp32 kernels
        ent = self._handles.get(key)
        if ent is None:
            h = _lib.Handle(torch.device("cuda", idx))
            h.set_precision("fp32" if training else self.precision)
            # (operand slices of the fp32-tolerance kernels: f16 pairs by default, also for training - with bf16 pairs the
            # forward's 1e-5 error decides enough ReLU / max kinks differently from fp64 to move the gradients by 1e-3 in
            # most small batches, scripts/fuzz_backward.py; CNS_SPLIT_BF16=1 trades that for fp32's exponent range)
            ent = [h, None]
            self._handles[key] = ent
        if ent[1] != version:
            on_dev = all(p.is_cuda for p in params)
            if not on_dev and any(p.is_cuda for p in params):
                raise _lib.CnsError("parameters are split between CPU and CUDA")
            ent[0].load_weights(params, on_device=on_dev)
            ent[1] = version
        return ent[0]

    def _workspace(self, idx: int, nbytes: int, key: str = "fwd") -> torch.Tensor:
        ws = self._workspaces.get((idx, key))
        if ws is None or ws.numel() < nbytes:
            if ws is not None:
                # kernels enqueued on another stream than the one this was allocated under may still use the old block
                # (BatchedServo's slots): keep it alive instead of handing it back to the caching allocator.  Sizes grow
                # geometrically, so only a handful of blocks are ever retired.
                self._retired_workspaces.append(ws)
            ws = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=torch.device("cuda", idx))
            self._workspaces[(idx, key)] = ws
        return ws

    @staticmethod
    def make_batch_struct(data) -> Tuple["_lib.cns_batch", list]:
        """cns_batch view of a GraphData/Batch whose tensors live on one CUDA device."""
        keep = []

        def t(key, dtype, required=True):
            v = _field(data, key)
            if v is None:
                if required:
                    raise KeyError("GraphData is missing field %r" % key)
                return None
            if v.dtype != dtype:
                v = v.to(dtype)
            if not v.is_contiguous():
                v = v.contiguous()
            keep.append(v)
            return v

        x_cur, x_tar = t("x_cur", torch.float32), t("x_tar", torch.float32)
        pos_cur, pos_tar = t("pos_cur", torch.float32), t("pos_tar", torch.float32)
        ej, ei = t("l0_to_l1_edge_index_j_cur", torch.int64), t("l0_to_l1_edge_index_i_cur", torch.int64)
        ec, et = t("l1_dense_edge_index_cur", torch.int64), t("l1_dense_edge_index_tar", torch.int64)
        cmask = _field(data, "cluster_mask")
        if cmask.dtype == torch.bool:
            cmask = cmask.view(torch.uint8)
        elif cmask.dtype != torch.uint8:
            cmask = (cmask != 0).view(torch.uint8)
        keep.append(cmask)
        ctr = t("cluster_centers_index", torch.int64)
        batch = None      # the node->graph vector is not needed: graph segments come from num_clusters
        ncl = _field(data, "num_clusters")
        ncl = ncl.reshape(-1).to(torch.int64).contiguous()
        keep.append(ncl)
        tpo = _field(data, "tPo_norm")
        if tpo is not None:
            tpo = tpo.reshape(-1).to(torch.float32).contiguous()
            keep.append(tpo)
        if x_cur.shape[-1] != 2 or pos_cur.shape[-1] != 2:
            raise _lib.CnsError("libcns_sm100 expects 2-D features and positions")
        dev = x_cur.device
        for v in keep:
            if v.device != dev:
                raise _lib.CnsError("all GraphData tensors must live on one device (%s vs %s)" % (v.device, dev))
        b = _lib.cns_batch()
        b.n_nodes, b.n_clusters, b.n_graphs = x_cur.shape[0], ctr.shape[0], ncl.shape[0]
        b.n_e01, b.n_e11_cur, b.n_e11_tar = ej.shape[0], ec.shape[1], et.shape[1]
        p = lambda v: None if v is None else v.data_ptr()
        b.x_cur, b.x_tar, b.pos_cur, b.pos_tar = p(x_cur), p(x_tar), p(pos_cur), p(pos_tar)
        b.l0_to_l1_j, b.l0_to_l1_i = p(ej), p(ei)
        b.l1_edge_cur, b.l1_edge_tar = p(ec), p(et)
        b.l1_cur_stride, b.l1_tar_stride = 0, 0
        b.cluster_mask, b.cluster_centers = p(cmask), p(ctr)
        b.node_graph, b.num_clusters, b.tPo_norm = p(batch), p(ncl), p(tpo)
        complete = data.get("l1_complete", False) if isinstance(data, dict) else getattr(data, "l1_complete", False)
        b.flags = (_lib.BATCH_L1_CUR_COMPLETE | _lib.BATCH_L1_TAR_COMPLETE) if complete is True else 0
        mc = data.get("max_clusters_per_graph", 0) if isinstance(data, dict) else getattr(data, "max_clusters_per_graph", 0)
        if not isinstance(mc, int) or isinstance(mc, bool):
            mc = 0
        if mc == 0 and b.n_graphs == 1:
            mc = int(b.n_clusters)                      # a single graph: its cluster count is the bound
        b.max_clusters_per_graph = mc
        return b, keep

    # ------------------------------------------------------------------ forward
    def forward(self, data, hidden: Optional[torch.Tensor] = None, check: bool = False, target_cache=None):
        """`target_cache` (optional, from `build_target_cache`): reuse what depends on the target side of the batch
        only - the reference builds its target graph once per target image (corr2graph.py:11-22)."""
        x_cur = _field(data, "x_cur")
        if not x_cur.is_cuda:
            raise _lib.CnsError(
                "cns_b200.GraphVS runs only on CUDA (sm_100a); move the GraphData with .to('cuda') "
                "- there is no CPU fallback")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .autograd import graphvs_autograd
            return graphvs_autograd(self, data, hidden)
        return self._forward_inference(data, hidden, check, target_cache)

    def build_target_cache(self, data) -> "TargetCache":
        """Target-side results of `data` (x_tar / pos_tar / cluster_centers_index / num_clusters), valid for every
        batch with the same targets while the weights and the precision mode stay unchanged."""
        b, keep = self.make_batch_struct(data)
        dev = keep[0].device
        belong = _field(data, "belong_cluster_index")
        if belong is not None and belong.numel() == b.n_nodes:
            # every target keypoint with its cluster: the cache then also keeps the target side's per-keypoint attention
            # logits / values (f16 mode), and a forward only re-runs their softmax over the visible keypoints
            belong = belong.to(dev, torch.int64).contiguous()
            ident = torch.arange(b.n_nodes, dtype=torch.int64, device=dev)
            keep += [belong, ident]
            b.n_e01, b.l0_to_l1_j, b.l0_to_l1_i = b.n_nodes, ident.data_ptr(), belong.data_ptr()
        return TargetCache(self._handle(dev), b, dev, keep)

    def _forward_inference(self, data, hidden, check=False, target_cache=None):
        lib = _lib.load()
        b, keep = self.make_batch_struct(data)
        dev = keep[0].device
        h = self._handle(dev)
        if target_cache is not None:
            b.target_cache = target_cache.raw
        nb, nc = b.n_graphs, b.n_clusters
        # one allocation; the (C,128) block first so its rows stay 16-byte aligned
        out = torch.empty(nc * 128 + nb * 13, dtype=torch.float32, device=dev)
        hid = out[:nc * 128].view(nc, 128)
        vec = out[nc * 128:nc * 128 + nb * 6].view(nb, 6)
        vel = out[nc * 128 + nb * 6:nc * 128 + nb * 12].view(nb, 6)
        nrm = out[nc * 128 + nb * 12:].view(nb, 1)
        if hidden is not None:
            hidden = hidden.detach().to(dev, torch.float32).contiguous()
            if hidden.shape != (nc, 128):
                raise _lib.CnsError("hidden has shape %s, expected (%d, 128)" % (tuple(hidden.shape), nc))
        nbytes = lib.cns_forward_workspace_bytes(C.byref(b))
        stream = _lib.stream_ptr(dev)
        # one workspace per stream: forwards enqueued on different streams may overlap on the device
        ws = self._workspace(dev.index, nbytes, ("fwd", stream.value))
        _lib.check(lib.cns_graphvs_forward(
            h.raw, C.byref(b), _lib.ptr(hidden), _lib.ptr(vec), _lib.ptr(nrm), _lib.ptr(hid), _lib.ptr(vel),
            _lib.ptr(ws), ws.numel(), None, 0, stream), "cns_graphvs_forward")
        if check:
            _lib.check(lib.cns_forward_status(_lib.ptr(ws), stream), "cns_forward_status")
        pred = RawPred((vec, nrm, hid))
        pred.vel = vel
        tpo = _field(data, "tPo_norm")
        pred.tpo_ptr = None if tpo is None else tpo.data_ptr()
        return pred
