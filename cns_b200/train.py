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


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced split of `n_items` scenes over `world` ranks."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class TBPTTStepper(object):
    """One optimizer update from a window of recurrent frames."""

    def __init__(self, net, optimizers: Optional[Dict[str, torch.optim.Optimizer]] = None, clip_norm: float = 10.0,
                 group=None):
        self.net = net
        self.optimizers = optimizers if optimizers is not None else make_optimizers(net)
        self.clip_norm = clip_norm
        self.bucket = GradBucket(list(net.parameters()))
        self.group = group
        self.hidden = None

    def window_loss(self, frames: Iterable, mean: bool = False):
        """Sum (short-seq) or mean (long-seq) of the per-frame objectives, hidden carried with grad."""
        losses, logs = [], []
        for data in frames:
            if bool(torch.as_tensor(data.new_scene).any()):
                self.hidden = None
            pred = self.net(data, self.hidden)
            self.hidden = pred[2]
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
        self.bucket.allreduce_mean(self.group)
        torch.nn.utils.clip_grad_norm_(self.bucket.params, self.clip_norm)
        for opt in self.optimizers.values():
            opt.step()
        if self.hidden is not None:
            self.hidden = self.hidden.detach()          # truncate BPTT (train_gvs_short_seq.py:48)
        return loss.detach()
