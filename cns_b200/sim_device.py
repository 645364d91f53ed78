"""Device-resident closed-loop point environment for training.

The reference drives B python `PointEnv`s per step (sim/environment.py:80-127), builds B `GraphData` on the host
and collates them (sim/dataset.py:59-178, 206-209); the supervisor is `pbvs_straight` (sim/supervisor.py:92-101).
At batch 256 that host work costs ~15x the GPU update.  `DeviceSceneBatch` keeps the B scenes on the GPU:

* state      : world points (sum N, 3), target / current camera poses (B, 4, 4), static two-level target graphs;
* observe()  : batched pinhole projection (utils/perception.py:62-68), random keypoint drop-out
               (RandomDiscard, sim/dataset.py:117-121), per-frame graph masking with `cns_graph_mask`
               (bit-exact with GraphGenerator._get_graph + collation), PBVS labels - returns the field dict
               `GraphVS.forward` / `objectives` read;
* feedback() : pose integration at 50 Hz (sim/environment.py:104-121) and re-initialisation of scenes that
               converged, timed out or left the workspace (new initial pose; the scene geometry and its clustering
               are replaced on the host every `refresh_every` calls, amortised).

The pose algebra is plain batched torch (fp64) - host-side logic moved next to the data, not a kernel; it is
checked against the numpy functions of `synth.py` in tests/test_host_cpu.py.
"""
import math
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib, synth
from .data import Batch


# ---------------------------------------------------------------------------------- SO(3) / SE(3), batched
def so3_exp(w: torch.Tensor) -> torch.Tensor:
    """Rotation vectors (B,3) -> matrices (B,3,3) (Rodrigues)."""
    th = w.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    k = w / th
    K = torch.zeros(w.shape[0], 3, 3, dtype=w.dtype, device=w.device)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -k[:, 2], k[:, 1], k[:, 2]
    K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -k[:, 0], -k[:, 1], k[:, 0]
    s, c = torch.sin(th)[..., None], torch.cos(th)[..., None]
    eye = torch.eye(3, dtype=w.dtype, device=w.device).expand_as(K)
    return eye + s * K + (1 - c) * (K @ K)


def so3_log(R: torch.Tensor) -> torch.Tensor:
    """Rotation matrices (B,3,3) -> rotation vectors (B,3), angle in [0, pi]."""
    tr = R.diagonal(dim1=-2, dim2=-1).sum(-1)
    cos = ((tr - 1) / 2).clamp(-1.0, 1.0)
    th = torch.acos(cos)
    v = torch.stack([R[:, 2, 1] - R[:, 1, 2], R[:, 0, 2] - R[:, 2, 0], R[:, 1, 0] - R[:, 0, 1]], dim=-1)
    small = th < 1e-6
    scale = torch.where(small, 0.5 + th * th / 12, th / (2 * torch.sin(th).clamp_min(1e-12)))
    out = v * scale[:, None]
    near_pi = th > math.pi - 1e-4
    if bool(near_pi.any()):                      # axis from the symmetric part (rare)
        Rp = R[near_pi]
        d = (Rp.diagonal(dim1=-2, dim2=-1) + 1) / 2
        ax = d.clamp_min(0).sqrt()
        i = d.argmax(-1)
        rows = torch.arange(Rp.shape[0], device=R.device)
        sgn = torch.sign((Rp + Rp.transpose(1, 2))[rows, i, :])
        sgn[rows, i] = 1
        ax = ax * sgn
        ax = ax / ax.norm(dim=-1, keepdim=True)
        flip = torch.sign((v[near_pi] * ax).sum(-1, keepdim=True))
        flip[flip == 0] = 1
        out[near_pi] = ax * flip * th[near_pi][:, None]
    return out


def se3_inv(T: torch.Tensor) -> torch.Tensor:
    Rt = T[:, :3, :3].transpose(1, 2)
    out = torch.zeros_like(T)
    out[:, :3, :3], out[:, :3, 3], out[:, 3, 3] = Rt, -(Rt @ T[:, :3, 3:4]).squeeze(-1), 1
    return out


def project_norm(wcT: torch.Tensor, points: torch.Tensor, scene: torch.Tensor) -> torch.Tensor:
    """World points (sum N,3) of scenes `scene` (sum N,) -> normalised image plane of cameras wcT (B,4,4)."""
    cwT = se3_inv(wcT)
    pc = (cwT[scene, :3, :3] @ points[:, :, None]).squeeze(-1) + cwT[scene, :3, 3]
    return pc[:, :2] / pc[:, 2:3]


def pbvs_straight(cur_wcT: torch.Tensor, tar_wcT: torch.Tensor) -> torch.Tensor:
    """Straight-line PBVS velocity in the current camera frame (B,6) (sim/supervisor.py:92-101)."""
    tcT = se3_inv(tar_wcT) @ cur_wcT
    R, t = tcT[:, :3, :3], tcT[:, :3, 3]
    return torch.cat([-(R.transpose(1, 2) @ t[:, :, None]).squeeze(-1), -so3_log(R)], dim=-1)


def integrate(cur_wcT: torch.Tensor, vel: torch.Tensor, dt: float = synth.DT) -> torch.Tensor:
    dT = torch.zeros_like(cur_wcT)
    dT[:, :3, :3], dT[:, :3, 3], dT[:, 3, 3] = so3_exp(vel[:, 3:] * dt), vel[:, :3] * dt, 1
    out = cur_wcT @ dT
    out[:, :3, :3] = so3_exp(so3_log(out[:, :3, :3]))          # re-orthonormalise
    return out


def pose_error(cur_wcT: torch.Tensor, tar_wcT: torch.Tensor):
    dT = se3_inv(cur_wcT) @ tar_wcT
    return so3_log(dT[:, :3, :3]).norm(dim=-1) * (180 / math.pi), dT[:, :3, 3].norm(dim=-1) * 1000


# ---------------------------------------------------------------------------------- the environment
class DeviceSceneBatch(object):
    def __init__(self, batch_size: int, rng: np.random.Generator, device="cuda:0", max_points: int = 512,
                 max_steps: int = 200, missing=(0.0, 0.2), refresh_every: int = 0):
        self.B, self.rng, self.dev = batch_size, rng, torch.device(device)
        self.max_points, self.max_steps, self.missing, self.refresh_every = max_points, max_steps, missing, refresh_every
        self.lib = _lib.load()
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(int(rng.integers(0, 2 ** 31 - 1)))
        self.calls = 0
        self._build([synth.Scene(rng, int(rng.integers(4, max_points + 1))) for _ in range(batch_size)])

    # -- static part: geometry, target graphs ------------------------------------------------------------
    def _build(self, scenes):
        dev = self.dev
        self.scenes = scenes                       # host copies of the geometry (poses live on the device)
        gens = [s.generator for s in scenes]
        n_nodes = np.array([g.num_points for g in gens], dtype=np.int64)
        n_clu = np.array([g.num_clusters for g in gens], dtype=np.int64)
        noff, coff = np.cumsum(n_nodes) - n_nodes, np.cumsum(n_clu) - n_clu
        self.N, self.Cn = int(n_nodes.sum()), int(n_clu.sum())
        d = lambda a, dt: torch.as_tensor(np.array(a), dtype=dt).to(dev)
        self.points = d(np.concatenate([s.points for s in scenes]), torch.float64)
        self.scene_of = d(np.repeat(np.arange(self.B), n_nodes), torch.int64)
        self.tar_wcT = d(np.stack([s.tar_wcT for s in scenes]), torch.float64)
        self.cur_wcT = d(np.stack([s.cur_wcT for s in scenes]), torch.float64)
        self.tpo = d([s.tPo_norm for s in scenes], torch.float32)
        self.x_tar = d(np.concatenate([g.target_points for g in gens]), torch.float32)
        self.belong = d(np.concatenate([g.belong_cluster_index + c for g, c in zip(gens, coff)]), torch.int64)
        self.member_order = d(np.concatenate([g.l0_to_l1_edge_index[0] + o for g, o in zip(gens, noff)]), torch.int64)
        self.cluster_size = d(np.concatenate([g.cluster_points_count for g in gens]), torch.int64)
        self.cluster_graph_ptr = d(np.concatenate([[0], np.cumsum(n_clu)]), torch.int64)
        self.centers = d(np.concatenate([g.cluster_centers_index + o for g, o in zip(gens, noff)]), torch.int64)
        self.l1_tar = d(np.concatenate([g.l1_dense_edge_index + c for g, c in zip(gens, coff)], axis=1), torch.int64)
        self.num_clusters = d(n_clu, torch.int64)
        self.node_graph = self.scene_of
        self.max_clusters = int(n_clu.max())
        self.cap = int((n_clu ** 2).sum())
        self.e01_j = torch.empty(max(self.N, 1), dtype=torch.int64, device=dev)
        self.e01_i = torch.empty(max(self.N, 1), dtype=torch.int64, device=dev)
        self.e11 = torch.empty(2, max(self.cap, 1), dtype=torch.int64, device=dev)
        self.cmask = torch.empty(max(self.Cn, 1), dtype=torch.uint8, device=dev)
        self.counts = torch.zeros(2, dtype=torch.int64, device=dev)
        self.ws = torch.empty(self.lib.cns_graph_mask_workspace_bytes(self.N, self.Cn, self.B), dtype=torch.uint8, device=dev)
        self.steps = torch.zeros(self.B, dtype=torch.int64, device=dev)
        self.fresh = torch.ones(self.B, dtype=torch.bool, device=dev)

    # -- per frame ------------------------------------------------------------------------------------------
    def observe(self) -> Batch:
        """One frame of every scene as the `Batch` GraphVS reads (per-frame tensors are fresh copies: the frames of
        a TBPTT window stay valid while the environment moves on)."""
        L, lib, dev = _lib, self.lib, self.dev
        cur_xy = project_norm(self.cur_wcT, self.points, self.scene_of).float()
        frac = torch.rand(self.B, generator=self.gen, device=dev) * (self.missing[1] - self.missing[0]) + self.missing[0]
        visible = (torch.rand(self.N, generator=self.gen, device=dev) >= frac[self.scene_of]).to(torch.uint8)
        L.check(lib.cns_graph_mask(L.ptr(self.belong), L.ptr(self.member_order), L.ptr(self.cluster_size),
                                   L.ptr(self.cluster_graph_ptr), L.ptr(visible), self.N, self.Cn, self.B,
                                   L.ptr(self.e01_j), L.ptr(self.e01_i), L.ptr(self.e11), self.e11.shape[1],
                                   L.ptr(self.cmask), L.ptr(self.counts), L.ptr(self.ws), self.ws.numel(), L.stream_ptr(dev)),
                "cns_graph_mask")
        n01, n11 = self.counts.tolist()                     # the only device -> host read of the step (16 bytes)
        vel = pbvs_straight(self.cur_wcT, self.tar_wcT).float()
        vel_si = vel.clone()
        vel_si[:, :3] /= (self.tpo[:, None] + 1e-7)
        data = {
            "x_cur": cur_xy, "pos_cur": cur_xy, "x_tar": self.x_tar, "pos_tar": self.x_tar,
            "l0_to_l1_edge_index_j_cur": self.e01_j[:n01].clone(), "l0_to_l1_edge_index_i_cur": self.e01_i[:n01].clone(),
            "l1_dense_edge_index_cur": self.e11[:, :n11].clone(), "l1_dense_edge_index_tar": self.l1_tar,
            "cluster_mask": self.cmask[:self.Cn].clone().bool(), "node_mask": visible.bool(),
            "cluster_centers_index": self.centers, "belong_cluster_index": self.belong,
            "num_clusters": self.num_clusters, "tPo_norm": self.tpo, "batch": self.node_graph,
            "new_scene": self.fresh.clone(), "vel": vel, "vel_si": vel_si,
            "l1_complete": True, "max_clusters_per_graph": self.max_clusters,
        }
        self.fresh.zero_()
        return Batch(**data)

    def feedback(self, vel: torch.Tensor):
        """Apply camera velocities (B,6) for one 50 Hz step; scenes that converged (< 1 deg, < 2 mm), timed out or
        left the workspace restart from a new initial pose."""
        self.cur_wcT = integrate(self.cur_wcT, vel.to(self.dev, torch.float64))
        self.steps += 1
        ang, mm = pose_error(self.cur_wcT, self.tar_wcT)
        lost = ~torch.isfinite(self.cur_wcT).all(-1).all(-1) | (self.cur_wcT[:, :3, 3].norm(dim=-1) > 1.8) | \
            (self.cur_wcT[:, 2, 2] > -0.1)
        done = ((ang < 1.0) & (mm < 2.0)) | (self.steps > self.max_steps) | lost
        if bool(done.any()):
            idx = torch.nonzero(done).flatten()
            self.cur_wcT[idx] = self._sample_initial_pose(idx.numel())
            self.steps[idx] = 0
            self.fresh[idx] = True
        self.calls += 1
        if self.refresh_every and self.calls % self.refresh_every == 0:
            self._build([synth.Scene(self.rng, int(self.rng.integers(4, self.max_points + 1))) for _ in range(self.B)])

    def _sample_initial_pose(self, n: int) -> torch.Tensor:
        """Camera on a spherical shell looking at the origin with bounded roll / pitch / yaw perturbation
        (sim/sampling.py:7-34; ranges sim/environment.py:52-62), like synth.sample_initial_pose."""
        dev, g = self.dev, self.gen
        u = lambda lo, hi: torch.rand(n, generator=g, device=dev, dtype=torch.float64) * (hi - lo) + lo
        r, theta = u(0.5, 0.9), u(-math.pi, math.pi)
        z = u(math.sin(math.radians(30)), math.sin(math.radians(90)))
        phi = torch.asin(z)
        d = torch.stack([torch.cos(phi) * torch.cos(theta), torch.cos(phi) * torch.sin(theta), z], dim=-1)
        y_axis = torch.stack([torch.sin(phi) * torch.cos(theta), torch.sin(phi) * torch.sin(theta), -torch.cos(phi)], dim=-1)
        z_axis = -d
        x_axis = torch.cross(y_axis, z_axis, dim=-1)
        rot = torch.stack([x_axis, y_axis, z_axis], dim=-1)
        for axis, lim in ((2, 60.0), (1, 10.0), (0, 10.0)):
            w = torch.zeros(n, 3, dtype=torch.float64, device=dev)
            w[:, axis] = u(-math.radians(lim), math.radians(lim))
            rot = rot @ so3_exp(w)
        T = torch.zeros(n, 4, 4, dtype=torch.float64, device=dev)
        T[:, :3, :3], T[:, :3, 3], T[:, 3, 3] = rot, d * r[:, None], 1
        return T
