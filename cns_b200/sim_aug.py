"""Batched, device-resident versions of what the reference's training data loaders do around the hot path every
frame: ground-truth supervisors and observation augmentations.

Reference (one python object per scene, numpy):
  * supervisors   `cns/sim/supervisor.py`: `pbvs_center` :69-89, `pbvs_straight` :92-101, `pbvs_hybrid` :251-260 (with
                  `_pbvs_mine` :226-244, `_move_avg` :133-152, `_cur_ideal_pose` :172-210, `_look_at_center` :213-221),
                  `supervisor_vel` :270-295 (distance-decoupled velocity `vel_si`, `tPo_norm`);
  * augmentations `cns/sim/dataset.py:88-125` / `dataset_long.py:81-118`: keypoint mismatch (10 % replaced by random
                  image positions; long sequences: `ObservationMismatcher`, `sampling.py:330-350`, holds a subset for an
                  exponential interval and rotates their coordinates), position jitter (+-0.002 / +-0.0005), drop-out -
                  `ObservationSampler` ("area shift", `sampling.py:244-327`: every keypoint is a two-state process whose
                  dwell times follow three Perlin-noise walkers over the target image) for scenes with >= 64 keypoints,
                  else an equal-probability 10 % discard - and the 80 % cap on removed keypoints.

Here the B scenes of a batch are ragged tensors on one device (keypoints concatenated, `scene_of` (sum N,) giving the
scene), every function is a batched torch expression in fp64 - host logic moved next to the data, no kernels of its own.
Random numbers come from the caller's `torch.Generator` (the reference draws from numpy's global stream: only the
distributions can agree; the deterministic parts are tested against the reference classes on identical inputs,
tests/test_sim_aug_cpu.py).
"""
import math
from typing import Optional

import torch

from .sim_device import se3_inv, so3_log


# ------------------------------------------------------------------------------------------------ small helpers
def _unit(v):
    return v / v.norm(dim=-1, keepdim=True)


def _nearest_rotation(M):
    """What scipy's Rotation.from_matrix makes of a matrix that is only approximately orthonormal."""
    U, _, Vh = torch.linalg.svd(M)
    R = U @ Vh
    det = torch.linalg.det(R)
    U = U.clone()
    U[:, :, 2] = U[:, :, 2] * det[:, None]
    return U @ Vh


def _rotvec(M):
    return so3_log(_nearest_rotation(M))


def segment_mean(x, scene_of, n_scenes, weight=None):
    """Mean (or weighted sum with normalised weights) of the rows of x per scene."""
    if weight is None:
        weight = torch.ones(x.shape[0], dtype=x.dtype, device=x.device)
        tot = torch.zeros(n_scenes, dtype=x.dtype, device=x.device).index_add_(0, scene_of, weight)
        w = weight / tot.clamp_min(1e-300)[scene_of]
    else:
        tot = torch.zeros(n_scenes, dtype=x.dtype, device=x.device).index_add_(0, scene_of, weight)
        w = weight / (tot + 1e-7)[scene_of]                      # dataset.py:138
    return torch.zeros(n_scenes, x.shape[1], dtype=x.dtype, device=x.device).index_add_(0, scene_of, x * w[:, None])


# ------------------------------------------------------------------------------------------------ supervisors
def pbvs_straight(cur_wcT, tar_wcT):
    """supervisor.py:92-101."""
    tcT = se3_inv(tar_wcT) @ cur_wcT
    R, t = tcT[:, :3, :3], tcT[:, :3, 3]
    return torch.cat([-(R.transpose(1, 2) @ t[:, :, None]).squeeze(-1), -so3_log(R)], dim=-1)


def pbvs_center(cur_wcT, tar_wcT, wPo):
    """supervisor.py:69-89: keeps the projection of the scene centre wPo (B,3) in the middle of the view."""
    tar_cwT, cur_cwT = se3_inv(tar_wcT), se3_inv(cur_wcT)
    tPo = (tar_cwT[:, :3, :3] @ wPo[:, :, None]).squeeze(-1) + tar_cwT[:, :3, 3]
    cPo = (cur_cwT[:, :3, :3] @ wPo[:, :, None]).squeeze(-1) + cur_cwT[:, :3, 3]
    u = so3_log((tar_cwT @ cur_wcT)[:, :3, :3])
    v = -(tPo - cPo + torch.cross(cPo, u, dim=-1))
    return torch.cat([v, -u], dim=-1)


def _move_avg(ax, ay, az_m):
    """supervisor.py:104-152: push the z axis to az_m, averaging the two ways of dragging x / y along."""
    ay1 = torch.cross(az_m, ax, dim=-1)
    ax1 = torch.cross(ay1, az_m, dim=-1)
    rot_yx = torch.stack([_unit(ax1), _unit(ay1), _unit(az_m)], dim=-1)
    ax2 = torch.cross(ay, az_m, dim=-1)
    ay2 = torch.cross(az_m, ax2, dim=-1)
    rot_xy = torch.stack([_unit(ax2), _unit(ay2), _unit(az_m)], dim=-1)
    rot = (rot_xy + rot_yx) / 2.0
    return rot / rot.norm(dim=1, keepdim=True)


def _rotation_from_vectors(a, b):
    """supervisor.py:155-169: rotation that turns a into b."""
    a, b = _unit(a), _unit(b)
    v = torch.cross(a, b, dim=-1)
    c = (a * b).sum(-1)
    K = torch.zeros(a.shape[0], 3, 3, dtype=a.dtype, device=a.device)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -v[:, 2], v[:, 1], v[:, 2]
    K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -v[:, 0], -v[:, 1], v[:, 0]
    eye = torch.eye(3, dtype=a.dtype, device=a.device).expand_as(K)
    return eye + K + (K @ K) / (1 + c)[:, None, None]


def _pbvs_mine(cur_wcT, tar_wcT, wPo):
    """supervisor.py:226-244: roughly straight path that keeps the scene centre in view."""
    cur_R, cur_t, tar_R, tar_t = cur_wcT[:, :3, :3], cur_wcT[:, :3, 3], tar_wcT[:, :3, :3], tar_wcT[:, :3, 3]
    moved_tar_R = _move_avg(tar_R[:, :, 0], tar_R[:, :, 1], wPo - tar_t)
    ideal_R = _rotation_from_vectors(tar_t - wPo, cur_t - wPo) @ moved_tar_R                 # _cur_ideal_pose :172-210
    u_ideal = _rotvec(torch.linalg.inv(ideal_R) @ cur_R)
    moved_cur_R = _move_avg(cur_R[:, :, 0], cur_R[:, :, 1], wPo - cur_t)                     # _look_at_center :213-221
    u_center = _rotvec(torch.linalg.inv(moved_cur_R) @ cur_R)
    w = -(u_center * 7 + u_ideal)
    cam_v = (torch.linalg.inv(cur_R) @ (tar_t - cur_t)[:, :, None]).squeeze(-1)
    return torch.cat([cam_v, w], dim=-1)


def pbvs_hybrid(cur_wcT, tar_wcT, wPo):
    """supervisor.py:251-260: blend of `_pbvs_mine` and the straight-line law by the remaining roll angle."""
    u = _rotvec(torch.linalg.inv(cur_wcT[:, :3, :3]) @ tar_wcT[:, :3, :3])
    uz_deg = u[:, 2].abs() / math.pi * 180
    alpha = (1.0 / (1 + torch.exp(-(uz_deg - 30) * 0.4)))[:, None]
    return alpha * _pbvs_mine(cur_wcT, tar_wcT, wPo) + (1 - alpha) * pbvs_straight(cur_wcT, tar_wcT)


def supervisor_vel(policy: str, cur_wcT, tar_wcT, wPo):
    """supervisor.py:270-295 -> (vel (B,6), tPo_norm (B,), vel_si (B,6)); wPo (B,3) = centre of the points the supervisor
    is shown (dataset.py:136-141: visible keypoints, or their area-shift weighted centre)."""
    if policy == "pbvs_straight":
        vel = pbvs_straight(cur_wcT, tar_wcT)
    elif policy == "pbvs_center":
        vel = pbvs_center(cur_wcT, tar_wcT, wPo)
    elif policy == "pbvs_hybrid":
        vel = pbvs_hybrid(cur_wcT, tar_wcT, wPo)
    else:
        raise ValueError("unknown supervisor policy %r (the reference's IBVS supervisor is not part of this path)" % policy)
    tar_cwT = se3_inv(tar_wcT)
    tPo = (tar_cwT[:, :3, :3] @ wPo[:, :, None]).squeeze(-1) + tar_cwT[:, :3, 3]
    tPo_norm = tPo.norm(dim=-1)
    vel_si = vel.clone()
    vel_si[:, :3] = vel_si[:, :3] / (tPo_norm[:, None] + 1e-7)
    return vel, tPo_norm, vel_si


# ------------------------------------------------------------------------------------------------ Perlin walkers
class BatchedPerlin(object):
    """`Perlin1d` (sampling.py:190-222) for an array of independent generators of any shape."""

    def __init__(self, amp, freq, gen: Optional[torch.Generator] = None):
        self.amp, self.freq, self.gen = amp, freq, gen
        dev, dt = amp.device, amp.dtype
        r = lambda: torch.rand(amp.shape, generator=gen, device=dev, dtype=dt)
        self.offset = r()
        self.g0, self.g1 = r() * 2 - 1, r() * 2 - 1
        self.prev_floor = torch.full(amp.shape, -(2 ** 62), dtype=torch.int64, device=dev)        # "None"
        self.next_grad = lambda mask: torch.rand(amp.shape, generator=gen, device=dev, dtype=dt) * 2 - 1

    def __call__(self, t):
        """t broadcastable to the generator array."""
        t = torch.as_tensor(t, dtype=self.amp.dtype, device=self.amp.device).expand(self.amp.shape)
        bad = torch.isnan(t)
        t = torch.where(bad, torch.zeros_like(t), t)
        self.prev_floor = torch.where(bad, torch.full_like(self.prev_floor, -(2 ** 62)), self.prev_floor)
        t = t * self.freq + self.offset
        s = torch.floor(t)
        si = s.to(torch.int64)
        changed = si != self.prev_floor
        if bool(changed.any()):
            fresh = self.next_grad(changed)
            self.g0 = torch.where(changed, self.g1, self.g0)
            self.g1 = torch.where(changed, fresh, self.g1)
            self.prev_floor = torch.where(changed, si, self.prev_floor)
        w = t - s
        n0, n1 = self.g0 * w, self.g1 * (t - (s + 1))
        n = n0 + (n1 - n0) * (6 * w ** 5 - 15 * w ** 4 + 10 * w ** 3)
        return n * self.amp


# ------------------------------------------------------------------------------------------------ area-shift drop-out
class BatchedObservationSampler(object):
    """`ObservationSampler` (sampling.py:244-327) for B scenes at once.

    points (sum N, 2): target keypoints on the normalised image plane; scene_of (sum N,); center0 (B,2): per-scene
    coordinate median (the caller computes it where the scenes are built)."""
    N_WALKERS = 3

    def __init__(self, points, scene_of, n_scenes, center0, tau_max=5.0, gen: Optional[torch.Generator] = None):
        dev, dt = points.device, torch.float64
        self.points, self.scene_of, self.B, self.gen = points.to(dt), scene_of, n_scenes, gen
        big = torch.full((n_scenes, 2), float("inf"), dtype=dt, device=dev)
        idx = scene_of[:, None].expand(-1, 2)
        lower = big.clone().scatter_reduce(0, idx, self.points, reduce="amin")
        upper = (-big).scatter_reduce(0, idx, self.points, reduce="amax")
        r = lambda *shape: torch.rand(*shape, generator=gen, device=dev, dtype=dt)
        center = center0.to(dt) + (upper - lower) * (r(n_scenes, 2) * 0.6 - 0.3)
        self.sigma = (upper - lower).norm(dim=-1) * 0.15
        self.center = center
        amp = (upper - lower)[:, None, :].expand(n_scenes, self.N_WALKERS, 2).contiguous()
        self.perlin = BatchedPerlin(amp, torch.ones_like(amp), gen)
        self.walker_centers = self.perlin(torch.zeros(n_scenes, 1, 1, dtype=dt, device=dev)) + center[:, None, :]
        n = self.points.shape[0]
        self.state = r(n) < 0.5                              # True = observed
        self.last_time = torch.zeros(n, dtype=dt, device=dev)
        self.tau_total = r(n) * (0.5 * tau_max) + 0.5 * tau_max
        self.no_event_rval = r(n)
        self.dpose_norm_accum = torch.zeros(n_scenes, dtype=dt, device=dev)
        self.prev_wcT = None

    def ratio(self):
        """tau_to_missing_ratio: max over the walkers of a Gaussian of the distance to the walker (sampling.py:300-306)."""
        c = self.walker_centers[self.scene_of]                                      # (sum N, 3, 2)
        dist = (self.points[:, None, :] - c).norm(dim=-1)
        return torch.exp(-dist ** 2 / (2 * self.sigma[self.scene_of][:, None] ** 2)).max(dim=1).values

    def __call__(self, time, current_wcT, dist_scale: float, redraw=None):
        """time (B,) seconds since the scene started.  Returns (observed (sum N,) bool, ratio (sum N,))."""
        if self.prev_wcT is not None:
            dpose = se3_inv(self.prev_wcT) @ current_wcT
            dn = so3_log(dpose[:, :3, :3]).norm(dim=-1) / math.pi * 0.2 + dpose[:, :3, 3].norm(dim=-1) / dist_scale * 0.7
            self.dpose_norm_accum = self.dpose_norm_accum + dn
        self.prev_wcT = current_wcT.clone()
        self.walker_centers = self.perlin(self.dpose_norm_accum[:, None, None]) + self.center[:, None, :]
        ratio = self.ratio()
        tau = torch.where(self.state, self.tau_total * ratio, self.tau_total * (1 - ratio))
        no_event_time = tau * (-torch.log(self.no_event_rval))
        trans = (time.to(tau.dtype)[self.scene_of] - self.last_time) > no_event_time
        if bool(trans.any()):
            self.state = torch.where(trans, ~self.state, self.state)
            fresh = redraw(trans) if redraw is not None else torch.rand(
                trans.shape, generator=self.gen, device=trans.device, dtype=tau.dtype)
            self.no_event_rval = torch.where(trans, fresh, self.no_event_rval)
        return self.state.clone(), ratio


# ------------------------------------------------------------------------------------------------ per-scene subsets
def random_positions(scene_of, n_per_scene, gen=None):
    """pos (sum N,) int64: position of every keypoint in a uniformly random permutation of its scene (the reference's
    `rand_perm = np.random.permutation(N)`, dataset.py:88)."""
    n = scene_of.shape[0]
    dev = scene_of.device
    key = torch.rand(n, generator=gen, device=dev, dtype=torch.float64)
    perm = torch.argsort(scene_of.to(torch.float64) * 2 + key)            # scene-major, random inside a scene
    start = torch.cumsum(n_per_scene, 0) - n_per_scene
    pos = torch.empty(n, dtype=torch.int64, device=dev)
    pos[perm] = torch.arange(n, device=dev) - start[scene_of[perm]]
    return pos


def random_subsets(scene_of, n_per_scene, counts, gen=None, offset=None, pos=None):
    """A uniformly random subset of `counts[b]` keypoints of every scene b (without replacement), as
    (selected (sum N,) bool, order (sum N,) int64): `order` = position of a selected keypoint inside its scene's
    selection (random order), -1 elsewhere.  `offset[b]`: skip that many positions of the scene's random permutation
    `pos` first (the reference takes mismatched and discarded keypoints from disjoint slices of ONE permutation,
    dataset.py:88-121)."""
    if pos is None:
        pos = random_positions(scene_of, n_per_scene, gen)
    lo = torch.zeros_like(counts) if offset is None else offset
    sel = (pos >= lo[scene_of]) & (pos < (lo + counts)[scene_of])
    order = torch.where(sel, pos - lo[scene_of], torch.full_like(pos, -1))
    return sel, order


def rotate_within_subsets(values, scene_of, sel, order, counts):
    """values[indices] = values[np.roll(indices, 1)] per scene (dataset_long.py:86-89): every selected keypoint takes
    the value of its predecessor in the scene's selection order."""
    n = values.shape[0]
    dev = values.device
    idx = torch.nonzero(sel).flatten()
    if idx.numel() == 0:
        return values
    sc = scene_of[idx]
    base = torch.cumsum(counts, 0) - counts
    slot = base[sc] + order[idx]
    table = torch.empty(int(counts.sum()), dtype=torch.int64, device=dev)
    table[slot] = idx
    prev = base[sc] + (order[idx] - 1) % counts[sc]
    out = values.clone()
    out[idx] = values[table[prev]]
    return out


class BatchedMismatcher(object):
    """`ObservationMismatcher(tau, ratio)` (sampling.py:330-350): a subset of int(N * ratio) keypoints per scene, redrawn
    after an exponentially distributed interval."""

    def __init__(self, scene_of, n_per_scene, tau=2.0, ratio=0.1, gen=None):
        self.scene_of, self.n_per_scene, self.tau, self.gen = scene_of, n_per_scene, tau, gen
        self.counts = (n_per_scene.to(torch.float64) * ratio).to(torch.int64)
        B, dev = n_per_scene.shape[0], scene_of.device
        self.accum = torch.zeros(B, dtype=torch.float64, device=dev)
        self.interval = torch.full((B,), -1.0, dtype=torch.float64, device=dev)       # "None"
        self.sel = torch.zeros(scene_of.shape[0], dtype=torch.bool, device=dev)
        self.order = torch.full((scene_of.shape[0],), -1, dtype=torch.int64, device=dev)

    def reset(self, scenes=None):
        if scenes is None:
            self.interval.fill_(-1.0)
        else:
            self.interval[scenes] = -1.0

    def __call__(self, dt: float):
        renew = (self.interval < 0) | (self.accum + dt > self.interval)
        if bool(renew.any()):
            B, dev = renew.shape[0], renew.device
            self.accum = torch.where(renew, torch.zeros_like(self.accum), self.accum + dt)
            draw = -torch.log(torch.rand(B, generator=self.gen, device=dev, dtype=torch.float64)) * self.tau
            self.interval = torch.where(renew, draw, self.interval)
            sel, order = random_subsets(self.scene_of, self.n_per_scene, self.counts, self.gen)
            pt = renew[self.scene_of]
            self.sel = torch.where(pt, sel, self.sel)
            self.order = torch.where(pt, order, self.order)
        else:
            self.accum = self.accum + dt
        return self.sel, self.order
