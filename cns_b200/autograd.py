"""Differentiable `GraphVS.forward`: one `torch.autograd.Function` around the library's fused
forward (`cns_graphvs_forward` with a `save` buffer) and backward (`cns_graphvs_backward`).

Autograd only sees (hidden_in, the 43 parameters) -> (vel_si_vec, vel_si_norm, hidden_out), so
truncated / full BPTT over servo frames (cns/train_gvs_short_seq.py:24-53, train_gvs_long_seq.py:20-53)
chains these nodes through `hidden`.  Training always runs the fp32 kernels.
"""
import ctypes as C

import torch

from . import _lib


class _GraphVSFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, data, hidden, *params):
        lib = _lib.load()
        b, keep = net.make_batch_struct(data)
        dev = keep[0].device
        handle = net._handle(dev, training=True)
        nb, nc = b.n_graphs, b.n_clusters
        out = torch.empty(nc * 128 + nb * 7, dtype=torch.float32, device=dev)
        hid = out[:nc * 128].view(nc, 128)
        vec = out[nc * 128:nc * 128 + nb * 6].view(nb, 6)
        nrm = out[nc * 128 + nb * 6:].view(nb, 1)
        hidden_c = None
        if hidden is not None:
            hidden_c = hidden.detach().to(dev, torch.float32).contiguous()
            if hidden_c.shape != (nc, 128):
                raise _lib.CnsError("hidden has shape %s, expected (%d, 128)" % (tuple(hidden_c.shape), nc))
        ws = net._workspace(dev.index, lib.cns_forward_workspace_bytes(C.byref(b)))
        save = torch.empty(lib.cns_forward_save_bytes(C.byref(b)), dtype=torch.uint8, device=dev)
        stream = _lib.stream_ptr(dev)
        _lib.check(lib.cns_graphvs_forward(
            handle.raw, C.byref(b), _lib.ptr(hidden_c), _lib.ptr(vec), _lib.ptr(nrm), _lib.ptr(hid), None,
            _lib.ptr(ws), ws.numel(), _lib.ptr(save), save.numel(), stream), "cns_graphvs_forward")
        ctx.net, ctx.batch, ctx.keep, ctx.handle = net, b, keep, handle
        ctx.has_hidden = hidden is not None
        ctx.save_for_backward(save, hid, *( [hidden_c] if hidden_c is not None else []))
        return vec, nrm, hid

    @staticmethod
    def backward(ctx, d_vec, d_nrm, d_hid):
        lib = _lib.load()
        saved = ctx.saved_tensors
        save, hid = saved[0], saved[1]
        hidden_c = saved[2] if ctx.has_hidden else None
        net, b, handle = ctx.net, ctx.batch, ctx.handle
        dev = hid.device
        total = lib.cns_param_offset(handle.raw, 43)
        grad = torch.zeros(total, dtype=torch.float32, device=dev)
        d_hidden_in = torch.empty_like(hid) if (ctx.has_hidden and ctx.needs_input_grad[2]) else None
        fix = lambda t: None if t is None else t.to(torch.float32).contiguous()
        d_vec, d_nrm, d_hid = fix(d_vec), fix(d_nrm), fix(d_hid)
        ws = net._workspace(dev.index, lib.cns_backward_workspace_bytes(C.byref(b)), key="bwd")
        _lib.check(lib.cns_graphvs_backward(
            handle.raw, C.byref(b), _lib.ptr(hidden_c), _lib.ptr(hid), _lib.ptr(d_vec), _lib.ptr(d_nrm),
            _lib.ptr(d_hid), _lib.ptr(d_hidden_in), _lib.ptr(grad), _lib.ptr(save), save.numel(), _lib.ptr(ws),
            ws.numel(), _lib.stream_ptr(dev)), "cns_graphvs_backward")
        grads = []
        for k, p in enumerate(net._ordered_params()):
            off = lib.cns_param_offset(handle.raw, k)
            grads.append(grad[off:off + p.numel()].view_as(p) if p.requires_grad else None)
        return (None, None, d_hidden_in) + tuple(grads)


def graphvs_autograd(net, data, hidden):
    from .model import RawPred
    params = net._ordered_params()
    if any(not p.is_cuda for p in params):
        raise _lib.CnsError("training needs the parameters on the GPU (net.to('cuda'))")
    vec, nrm, hid = _GraphVSFn.apply(net, data, hidden, *params)
    return RawPred((vec, nrm, hid))
