"""`postprocess` / `objectives` of the reference API (cns/models/functions.py:26-65) as calls into libcns_sm100.

Thin shim: the arithmetic is in `cns_postprocess` (unit direction x (1 + ELU(norm)), translation x `tPo_norm`) and
`cns_objectives` (1 - cos direction loss + MSE on the inverse-ELU of |gt|, forward and gradient in one launch;
csrc/train_ops.cu).  Like the rest of the package there is no CPU implementation: CPU tensors raise.
"""
import torch

from . import _lib


def _field(data, key):
    return data[key] if isinstance(data, dict) else getattr(data, key)


def _need_cuda(t, what):
    if not t.is_cuda:
        raise _lib.CnsError("%s runs on CUDA tensors only (libcns_sm100 has no CPU path)" % what)


def postprocess(raw_pred, data, post_scale: bool) -> torch.Tensor:
    """functions.py:26-38.  No gradient flows through the result (the reference only uses it to drive the simulator
    / robot, train_gvs_short_seq.py:41-44)."""
    vec, nrm = raw_pred[0].detach(), raw_pred[1]
    _need_cuda(vec, "postprocess")
    vec = vec.contiguous().float()
    nrm = None if nrm is None else nrm.detach().contiguous().float()
    tpo = _field(data, "tPo_norm").to(vec.device, torch.float32).reshape(-1).contiguous() if post_scale else None
    vel = torch.empty_like(vec)
    _lib.check(_lib.load().cns_postprocess(_lib.ptr(vec), _lib.ptr(nrm), _lib.ptr(tpo), vec.shape[0], _lib.ptr(vel),
                                           _lib.stream_ptr(vec.device)), "cns_postprocess")
    return vel


def objectives(raw_pred, data, gt_si: bool, weight=1):
    """functions.py:41-65 -> ({"dir_loss", "norm_loss", "total_loss"}, loss); `loss` carries the gradient
    (`cns_objectives` computes it in the same launch).  `weight` scales both arguments of a cosine similarity and is
    therefore without effect in the regress_norm=True form this library implements."""
    from .train import fused_objectives
    _need_cuda(raw_pred[0], "objectives")
    gt = _field(data, "vel_si" if gt_si else "vel")
    loss, triple = fused_objectives(raw_pred, {"vel_si": gt})
    d, n, t = triple.tolist()          # one host sync instead of the reference's three .item() calls
    return {"dir_loss": d, "norm_loss": n, "total_loss": t}, loss
