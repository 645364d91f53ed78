"""Velocity post-processing and training objectives (reference cns/models/functions.py).

`postprocess` (functions.py:26-38): unit direction x (1 + ELU(norm)); translational part
scaled by the scene distance `tPo_norm`.  The inference forward already produces this
fused in the pool/head kernel (`RawPred.vel`); the torch version here is the generic
entry the reference API exposes and what training uses on (B, 6) tensors.
`objectives` (functions.py:41-65): 1 - cos(direction) + MSE on the inverse-ELU norm.
"""
import torch
import torch.nn.functional as F


def _field(data, key):
    return data[key] if isinstance(data, dict) else getattr(data, key)


def alpha_p_elu(x, alpha=1.0):
    return alpha + F.elu(x, alpha=alpha)


def reverse_alpha_p_elu(y: torch.Tensor, alpha=1.0, eps=1e-7):
    return torch.where(y < alpha, torch.log(y / alpha + eps), y - alpha)


def scaled_cosine_similarity_loss(pred, target):
    return 1 - F.cosine_similarity(pred, target, dim=-1)


def postprocess(raw_pred, data, post_scale: bool) -> torch.Tensor:
    vel_vec, vel_norm = raw_pred[0], raw_pred[1]
    if vel_norm is None:
        vel = vel_vec
    else:
        vel = vel_vec / (torch.norm(vel_vec, dim=-1, keepdim=True) + 1e-7) * alpha_p_elu(vel_norm, alpha=1)
    if post_scale:
        vel = vel.clone()
        vel[:, :3] *= _field(data, "tPo_norm").to(vel).view(-1, 1)
    return vel


def objectives(raw_pred, data, gt_si: bool, weight=1):
    vel_vec, vel_norm = raw_pred[0], raw_pred[1]
    gt = _field(data, "vel_si" if gt_si else "vel").to(vel_vec)
    dir_loss = scaled_cosine_similarity_loss(vel_vec * weight, gt * weight).mean()
    if vel_norm is None:
        norm_loss = F.mse_loss(vel_vec * weight, gt * weight, reduction="mean")
    else:
        gt_norm = torch.norm(gt, dim=-1, keepdim=True)
        norm_loss = F.mse_loss(vel_norm, reverse_alpha_p_elu(gt_norm, eps=1e-5), reduction="mean")
    loss = dir_loss + norm_loss
    # one host sync instead of the reference's three .item() calls
    d, n, t = torch.stack([dir_loss.detach(), norm_loss.detach(), loss.detach()]).tolist()
    return {"dir_loss": d, "norm_loss": n, "total_loss": t}, loss
