"""Training-side callers of the hot path: truncated-BPTT step, gradient clipping, the two AdamW
groups and data-parallel gradient averaging.

Mirrors what `cns/utils/trainer.py:75-100` + `cns/train_gvs_short_seq.py:24-53` do per update
(accumulate the loss over `steps_for_update` recurrent frames, one backward, `clip_grad_norm_(10)`,
step both optimizers, detach the hidden state) and `cns/train_gvs_long_seq.py:20-53` (mean loss over
a long sequence).  The simulator-in-the-loop feedback of the reference (`loader.feedback(pred_vel)`)
is the caller's business: `TBPTTStepper.step` takes the frames of one window.

Data parallelism: graphs are independent, so every rank owns its own scenes and hidden
states; the only collective is ONE all-reduce of a flat fp32 gradient bucket (464 520 floats) per
optimizer step, NCCL on GPUs (gloo in the CPU tests).
"""
from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def make_optimizers(net, lr=5e-4, weight_decay=1e-4) -> Dict[str, torch.optim.Optimizer]:
    """The reference's two groups (train_cns.py:65-69): decayed Linear weights / un-decayed rest."""
    decay, no_decay = net.get_parameter_groups()
    return {"0": torch.optim.AdamW(decay, lr=lr, weight_decay=weight_decay),
            "1": torch.optim.AdamW(no_decay, lr=lr, weight_decay=0)}


class GradBucket(object):
    """One flat fp32 buffer aliasing every parameter's .grad -> a single all-reduce per step."""

    def __init__(self, params: Sequence[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def rebind(self):
        """autograd may have replaced .grad tensors; copy them back into the bucket views."""
        off = 0
        for p in self.params:
            view = self.flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                view.zero_()
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
            p.grad = view
            off += p.numel()

    def allreduce_mean(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))


class _FusedObjectives(torch.autograd.Function):
    """GraphVS.objectives as one kernel (`cns_objectives`): loss triple + gradients in a single launch."""

    @staticmethod
    def forward(ctx, vec, nrm, gt):
        from . import _lib
        lib = _lib.load()
        vec_c, nrm_c, gt_c = vec.detach().contiguous().float(), nrm.detach().contiguous().float(), gt.contiguous().float()
        out = torch.empty(3, dtype=torch.float32, device=vec.device)
        dvec, dnrm = torch.empty_like(vec_c), torch.empty_like(nrm_c)
        _lib.check(lib.cns_objectives(_lib.ptr(vec_c), _lib.ptr(nrm_c), _lib.ptr(gt_c), vec_c.shape[0], 1.0,
                                      _lib.ptr(out), _lib.ptr(dvec), _lib.ptr(dnrm), _lib.stream_ptr(vec.device)),
                   "cns_objectives")
        ctx.save_for_backward(dvec, dnrm)
        ctx.mark_non_differentiable(out)
        return out[2].clone(), out

    @staticmethod
    def backward(ctx, g_loss, _g_out):
        dvec, dnrm = ctx.saved_tensors
        return dvec * g_loss, dnrm * g_loss, None


def fused_objectives(raw_pred, data):
    """(loss, [dir, norm, total] device tensor) - same values as functions.objectives, no host sync."""
    gt = data.vel_si if not isinstance(data, dict) else data["vel_si"]
    return _FusedObjectives.apply(raw_pred[0], raw_pred[1], gt.to(raw_pred[0].device))


class FusedClipAdamW(object):
    """clip_grad_norm_(10) + the reference's two AdamW groups as one fused update (`cns_clip_adamw`) over a
    flat parameter buffer.  Parameters of `net` are re-pointed at views of that buffer."""

    def __init__(self, net, bucket: "GradBucket", lr=5e-4, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8,
                 max_norm=10.0):
        from . import _lib
        self.lib, self._lib = _lib.load(), _lib
        self.net = net
        self.params = bucket.params
        dev = self.params[0].device
        self.flat = torch.cat([p.detach().reshape(-1) for p in self.params]).contiguous()
        decay_ids = {id(p) for p in net.get_parameter_groups()[0]}
        mask = torch.cat([torch.full((p.numel(),), 1 if id(p) in decay_ids else 0, dtype=torch.uint8) for p in self.params])
        self.decay_mask = mask.to(dev)
        off = 0
        for p in self.params:
            p.data = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.bucket = bucket
        self.m, self.v = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self.lr, self.wd, self.betas, self.eps, self.max_norm, self.t = lr, weight_decay, betas, eps, max_norm, 0
        self.ws = torch.empty(self.lib.cns_clip_adamw_workspace_bytes(), dtype=torch.uint8, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)

    def step(self, grad_div: float = 1.0):
        self.t += 1
        L = self._lib
        L.check(self.lib.cns_clip_adamw(L.ptr(self.flat), L.ptr(self.bucket.flat), L.ptr(self.m), L.ptr(self.v),
                                        L.ptr(self.decay_mask), self.flat.numel(), self.max_norm, self.lr, self.betas[0],
                                        self.betas[1], self.eps, self.wd, self.t, float(grad_div), L.ptr(self.grad_norm),
                                        L.ptr(self.ws), self.ws.numel(), L.stream_ptr(self.flat.device)), "cns_clip_adamw")
        # the kernel updated the parameter storage behind torch's back: make GraphVS re-pack its weights
        for ent in self.net._handles.values():
            ent[1] = None


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced split of `n_items` scenes over `world` ranks."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class TBPTTStepper(object):
    """One optimizer update from a window of recurrent frames."""

    def __init__(self, net, optimizers: Optional[Dict[str, torch.optim.Optimizer]] = None, clip_norm: float = 10.0,
                 group=None, fused: bool = True, lr: float = 5e-4, weight_decay: float = 1e-4,
                 hidden_reset: str = "scene"):
        self.net = net
        self.clip_norm = clip_norm
        self.bucket = GradBucket(list(net.parameters()))
        self.fused = fused
        if fused:       # fused loss + fused clip/AdamW kernels (no torch optimizer objects)
            self.optimizers = {}
            self.fused_opt = FusedClipAdamW(net, self.bucket, lr=lr, weight_decay=weight_decay, max_norm=clip_norm)
        else:
            self.optimizers = optimizers if optimizers is not None else make_optimizers(net, lr, weight_decay)
        self.group = group
        self.hidden = None
        self.hidden_reset = hidden_reset
        self._counts = None          # num_clusters (B,) of the frame `hidden` belongs to

    def _carry_hidden(self, data):
        """Hidden state for this frame.  The reference drops the state of the WHOLE batch when any scene is new
        (train_gvs_short_seq.py:28) - harmless there because its loader restarts all scenes together
        (sim/dataset.py:242-246).  Scenes here restart one by one, so with `hidden_reset="scene"` (default) only the
        cluster rows of the restarted scenes are zeroed (a restarted scene may come back with a different cluster
        count: its rows are spliced in as zeros); `hidden_reset="batch"` is the reference's literal rule."""
        if self.hidden is None:
            return None
        new_scene = torch.as_tensor(data.new_scene).reshape(-1)
        counts = torch.as_tensor(data.num_clusters).reshape(-1).to(self.hidden.device)
        fresh = new_scene.to(self.hidden.device, torch.bool)
        if not bool(fresh.any()):
            return self.hidden if int(counts.sum()) == self.hidden.shape[0] else None
        old = self._counts
        if self.hidden_reset != "scene" or old is None or old.numel() != counts.numel() or bool(fresh.all()) \
                or not bool(((old == counts) | fresh).all()):
            return None
        graph = torch.repeat_interleave(torch.arange(counts.numel(), device=counts.device), counts)
        new_ptr, old_ptr = torch.cumsum(counts, 0) - counts, torch.cumsum(old, 0) - old
        keep = ~fresh[graph]
        row = torch.arange(graph.numel(), device=counts.device) - new_ptr[graph] + old_ptr[graph]
        row = torch.where(keep, row, torch.zeros_like(row))
        return self.hidden.index_select(0, row) * keep[:, None].to(self.hidden.dtype)

    def window_loss(self, frames: Iterable, mean: bool = False):
        """Sum (short-seq) or mean (long-seq) of the per-frame objectives, hidden carried with grad."""
        losses, logs = [], []
        for data in frames:
            pred = self.net(data, self._carry_hidden(data))
            self.hidden = pred[2]
            self._counts = torch.as_tensor(data.num_clusters).reshape(-1).to(self.hidden.device).clone()
            self.last_pred = pred          # closed-loop callers read the prediction between frames
            if self.fused:
                loss_t, triple = fused_objectives(pred, data)
                losses.append(loss_t)
                logs.append((triple[0], triple[1]))
                continue
            vec, nrm = pred[0], pred[1]
            gt = data.vel_si.to(vec)
            dir_loss = (1 - torch.nn.functional.cosine_similarity(vec, gt, dim=-1)).mean()
            gt_norm = torch.norm(gt, dim=-1, keepdim=True)
            tgt = torch.where(gt_norm < 1, torch.log(gt_norm + 1e-5), gt_norm - 1)
            norm_loss = torch.nn.functional.mse_loss(nrm, tgt)
            losses.append(dir_loss + norm_loss)
            logs.append((dir_loss.detach(), norm_loss.detach()))
        total = torch.stack(losses).mean() if mean else torch.stack(losses).sum()
        return total, logs

    def step(self, frames: Iterable, mean: bool = False) -> torch.Tensor:
        self.bucket.zero()
        loss, _ = self.window_loss(frames, mean=mean)
        loss.backward()
        self.bucket.rebind()
        if self.fused:
            world = 1
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
                dist.all_reduce(self.bucket.flat, op=dist.ReduceOp.SUM, group=self.group)
                world = dist.get_world_size(self.group)
            self.fused_opt.step(grad_div=float(world))      # averages, clips and steps in one kernel
        else:
            self.bucket.allreduce_mean(self.group)
            torch.nn.utils.clip_grad_norm_(self.bucket.params, self.clip_norm)
            for opt in self.optimizers.values():
                opt.step()
        if self.hidden is not None:
            self.hidden = self.hidden.detach()          # truncate BPTT (train_gvs_short_seq.py:48)
        return loss.detach()
