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


def synth_intrinsic():
    """(width, height, fx, fy, cx, cy) of CameraIntrinsic.default() (utils/perception.py:62-64)."""
    return (640.0, 480.0, 540.0, 540.0, 320.0, 240.0)


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
    # axis from the symmetric part near pi (rare) - computed for every row and selected, so that no device -> host read
    # (`if near_pi.any()`) stalls the caller four times per frame
    d = (R.diagonal(dim1=-2, dim2=-1) + 1) / 2
    ax = d.clamp_min(0).sqrt()
    i = d.argmax(-1)
    rows = torch.arange(R.shape[0], device=R.device)
    sgn = torch.sign((R + R.transpose(1, 2))[rows, i, :])
    sgn[rows, i] = 1
    ax = ax * sgn
    ax = ax / ax.norm(dim=-1, keepdim=True).clamp_min(1e-300)
    flip = torch.sign((v * ax).sum(-1, keepdim=True))
    flip = torch.where(flip == 0, torch.ones_like(flip), flip)
    return torch.where(near_pi[:, None], ax * flip * th[:, None], out)


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
    """B closed-loop point scenes resident on one GPU.

    aug=False (default)   : the plain environment of round 1 - uniform random keypoint drop-out of a per-scene
                            fraction in `missing`, straight-line PBVS labels, `tPo_norm` of the whole scene.
    aug=True              : the reference's training distribution (sim/dataset.py:88-141 / dataset_long.py:81-140,
                            `ObsAug.Default` = mismatch | area shift | jitter; `discard=True` adds RandomDiscard):
                            10 % mismatched keypoints (N >= 10), position jitter, area-shift drop-out for scenes with
                            >= 64 keypoints (`BatchedObservationSampler`), the 80 % cap on removed keypoints, zeroed
                            coordinates of removed keypoints, and the supervisor shown the visible keypoints (or their
                            area-shift weighted centre), which also sets `tPo_norm` per frame.
    supervisor            : "pbvs_straight" (short sequences, dataset.py:33) | "pbvs_hybrid" (long, dataset_long.py:21)
                            | "pbvs_center".
    long_seq=True         : dataset_long.py's variants - rotating mismatches held for an exponential interval
                            (`BatchedMismatcher`), jitter 0.0005 instead of 0.002.
    reinit                : "scene" - a scene that converged / timed out / left the workspace restarts alone from a new
                            initial pose (its hidden-state rows are reset by TBPTTStepper); "all" - the reference's
                            loaders (dataset.py:242-246, dataset_long.py:215-222): nothing restarts alone, ALL scenes get
                            new geometry, targets and poses once more than half of them need it.
    """
    TARGET_POSE_R_MEAN = 0.7              # np.mean(env.target_pose_r), sim/environment.py:52 (dist_scale of the sampler)

    def __init__(self, batch_size: int, rng: np.random.Generator, device="cuda:0", max_points: int = 512,
                 max_steps: int = 200, missing=(0.0, 0.2), refresh_every: int = 0, aug: bool = False,
                 supervisor: str = "pbvs_straight", long_seq: bool = False, reinit: str = "scene", discard: bool = False,
                 min_points: int = 4, device_noise: bool = True):
        assert reinit in ("scene", "all")
        self.device_noise = device_noise
        self._intr_dev = None
        self.B, self.rng, self.dev = batch_size, rng, torch.device(device)
        self.max_points, self.max_steps, self.missing, self.refresh_every = max_points, max_steps, missing, refresh_every
        self.min_points = min_points
        self.aug, self.supervisor, self.long_seq, self.reinit, self.discard = aug, supervisor, long_seq, reinit, discard
        self.lib = _lib.load()
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(int(rng.integers(0, 2 ** 31 - 1)))
        self.calls = 0
        self.reinit_all_count = 0
        self._build(self._new_scenes())

    def _new_scenes(self):
        """B fresh scenes; with a CUDA device all their target graphs come from ONE batched device clustering call."""
        from .graph_gen import GraphGenerator
        rng = self.rng
        protos = [(synth.sample_points(rng, int(rng.integers(self.min_points, self.max_points + 1))), synth.sample_target_pose(rng))
                  for _ in range(self.B)]
        tars = [synth.project_norm(T, P) for P, T in protos]
        # random training scenes: the clustering's tie-breaking noise is drawn on the device (same distribution as
        # scikit-learn's host draw, which costs 0.25 s per 256 scenes)
        gens = GraphGenerator.build_many(tars, device=self.dev, noise_generator=self.gen if self.device_noise else None) \
            if GraphGenerator.backend() == "device" else [GraphGenerator(t) for t in tars]
        return [synth.Scene(rng, len(P), generator=g, points=P, tar_wcT=T) for (P, T), g in zip(protos, gens)]

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
        self.n_nodes = d(n_nodes, torch.int64)
        self.node_start = d(noff, torch.int64)
        self.tar_wcT = d(np.stack([s.tar_wcT for s in scenes]), torch.float64)
        self.cur_wcT = d(np.stack([s.cur_wcT for s in scenes]), torch.float64)
        self.tpo_static = d([s.tPo_norm for s in scenes], torch.float32)
        self.tpo = self.tpo_static
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
        self.gt_vel = torch.zeros(self.B, 6, dtype=torch.float64, device=dev)
        if self.aug:
            from . import sim_aug
            med = d(np.stack([np.median(g.target_points, axis=0) for g in gens]), torch.float64)
            self._tar_med = med
            self.sampler = sim_aug.BatchedObservationSampler(self.x_tar.double(), self.scene_of, self.B, med, 5.0, self.gen)
            self.mismatcher = sim_aug.BatchedMismatcher(self.scene_of, self.n_nodes, 2.0, 0.1, self.gen)
            # int(N * 0.1) mismatched (N >= 10), int(N * 0.1) discarded (10 <= N < 64, RandomDiscard only)
            ten = (self.n_nodes.double() * 0.1).to(torch.int64)
            self.n_mismatch = torch.where(self.n_nodes >= 10, ten, torch.zeros_like(ten))
            self.n_discard = torch.where((self.n_nodes >= 10) & (self.n_nodes < 64), ten, torch.zeros_like(ten)) \
                if self.discard else torch.zeros_like(ten)
            self.mismatcher.counts = self.n_mismatch

    # -- per frame ------------------------------------------------------------------------------------------
    def _visibility(self, cur_xy):
        """(cur_xy', visible (sum N,) bool, mismatch (sum N,) bool, wPo (B,3)) of this frame."""
        dev, so = self.dev, self.scene_of
        if not self.aug:
            frac = torch.rand(self.B, generator=self.gen, device=dev) * (self.missing[1] - self.missing[0]) + self.missing[0]
            visible = torch.rand(self.N, generator=self.gen, device=dev) >= frac[so]
            from .sim_aug import segment_mean
            return cur_xy, visible, torch.zeros_like(visible), segment_mean(self.points, so, self.B)
        from . import sim_aug
        intr = synth_intrinsic()
        pos = sim_aug.random_positions(so, self.n_nodes, self.gen)
        if self.long_seq:                                  # rotating mismatches held for an exponential interval
            mism, order = self.mismatcher(synth.DT)
            cur_xy = sim_aug.rotate_within_subsets(cur_xy, so, mism, order, self.n_mismatch)
        else:                                              # 10 % of the keypoints land anywhere in the image
            mism = pos < self.n_mismatch[so]
            u = torch.rand(self.N, 2, generator=self.gen, device=dev, dtype=torch.float64)
            if self._intr_dev is None:             # (three host->device constants per frame cost 0.9 ms)
                self._intr_dev = tuple(torch.tensor(v, dtype=torch.float64, device=dev)
                                       for v in ([intr[0], intr[1]], [intr[4], intr[5]], [intr[2], intr[3]]))
            wh, cxy, fxy = self._intr_dev
            rnd = (u * wh - cxy) / fxy
            cur_xy = torch.where(mism[:, None], rnd, cur_xy)
        scale = 0.0005 if self.long_seq else 0.002
        cur_xy = cur_xy + scale * (torch.rand(self.N, 2, generator=self.gen, device=dev, dtype=torch.float64) * 2 - 1)
        # drop-out: area shift for N >= 64, equal-probability discard for 10 <= N < 64 (if enabled)
        observed, ratio = self.sampler(self.steps.double() * synth.DT, self.cur_wcT, self.TARGET_POSE_R_MEAN)
        big = (self.n_nodes >= 64)[so]
        discard = (pos >= self.n_mismatch[so]) & (pos < (self.n_mismatch + self.n_discard)[so])
        removed = torch.where(big, ~observed, discard)
        # "if too much points are removed, preserve some": keep only the first int(0.8 N) removed indices
        rank = torch.cumsum(removed.to(torch.int64), 0) - removed.to(torch.int64)
        rank = rank - rank[self.node_start][so]
        n_removed = torch.zeros(self.B, dtype=torch.int64, device=dev).index_add_(0, so, removed.to(torch.int64))
        cap = (self.n_nodes.double() * 0.8).to(torch.int64)
        over = (n_removed.double() >= 0.8 * self.n_nodes.double())[so]
        removed = removed & (~over | (rank < cap[so]))
        visible = ~removed
        cur_xy = torch.where(removed[:, None], torch.zeros_like(cur_xy), cur_xy)
        # the points the supervisor is shown (dataset.py:136-141)
        w_center = sim_aug.segment_mean(self.points, so, self.B, weight=ratio)
        v_center = sim_aug.segment_mean(self.points[visible], so[visible], self.B)
        wPo = torch.where((self.n_nodes >= 64)[:, None], w_center, v_center)
        return cur_xy, visible, mism, wPo

    def observe(self) -> Batch:
        """One frame of every scene as the `Batch` GraphVS reads (per-frame tensors are fresh copies: the frames of
        a TBPTT window stay valid while the environment moves on)."""
        L, lib, dev = _lib, self.lib, self.dev
        cur_xy = project_norm(self.cur_wcT, self.points, self.scene_of)
        cur_xy, visible_b, mismatch, wPo = self._visibility(cur_xy)
        cur_xy = cur_xy.float()
        visible = visible_b.to(torch.uint8)
        L.check(lib.cns_graph_mask(L.ptr(self.belong), L.ptr(self.member_order), L.ptr(self.cluster_size),
                                   L.ptr(self.cluster_graph_ptr), L.ptr(visible), self.N, self.Cn, self.B,
                                   L.ptr(self.e01_j), L.ptr(self.e01_i), L.ptr(self.e11), self.e11.shape[1],
                                   L.ptr(self.cmask), L.ptr(self.counts), L.ptr(self.ws), self.ws.numel(), L.stream_ptr(dev)),
                "cns_graph_mask")
        n01, n11 = self.counts.tolist()                     # the only device -> host read of the step (16 bytes)
        if self.aug:
            from .sim_aug import supervisor_vel
            vel64, tpo, vel_si64 = supervisor_vel(self.supervisor, self.cur_wcT, self.tar_wcT, wPo)
            self.gt_vel = vel64
            vel, vel_si, self.tpo = vel64.float(), vel_si64.float(), tpo.float()
        else:
            vel64 = pbvs_straight(self.cur_wcT, self.tar_wcT)
            self.gt_vel = vel64
            vel = vel64.float()
            vel_si = vel.clone()
            vel_si[:, :3] /= (self.tpo[:, None] + 1e-7)
        data = {
            "x_cur": cur_xy, "pos_cur": cur_xy, "x_tar": self.x_tar, "pos_tar": self.x_tar,
            "l0_to_l1_edge_index_j_cur": self.e01_j[:n01].clone(), "l0_to_l1_edge_index_i_cur": self.e01_i[:n01].clone(),
            "l1_dense_edge_index_cur": self.e11[:, :n11].clone(), "l1_dense_edge_index_tar": self.l1_tar,
            "cluster_mask": self.cmask[:self.Cn].clone().bool(), "node_mask": visible_b,
            "cluster_centers_index": self.centers, "belong_cluster_index": self.belong,
            "num_clusters": self.num_clusters, "tPo_norm": self.tpo, "batch": self.node_graph,
            "new_scene": self.fresh.clone(), "vel": vel, "vel_si": vel_si, "mismatch_mask": mismatch,
            "l1_complete": True, "max_clusters_per_graph": self.max_clusters,
        }
        self.fresh.zero_()
        return Batch(**data)

    # -- the reference's re-initialisation tests (sim/environment.py:151-203) -----------------------------------
    def abnormal_pose(self) -> torch.Tensor:
        pos = self.cur_wcT[:, :3, 3]
        from .sim_aug import segment_mean
        wPo = segment_mean(self.points, self.scene_of, self.B)
        cur_cwT = se3_inv(self.cur_wcT)
        cPo_z = (cur_cwT[:, 2, :3] * wPo).sum(-1) + cur_cwT[:, 2, 3]
        low = pos[:, 2] < 0.5 * math.sin(30) * 0.5          # sic: environment.py:163 feeds degrees to np.sin
        return ~torch.isfinite(self.cur_wcT).all(-1).all(-1) | low | (self.cur_wcT[:, 2, 2] > -0.1) | \
            (pos.norm(dim=-1) > 0.9 * 2) | (cPo_z < 0)

    def need_reinit(self) -> torch.Tensor:
        ang, mm = pose_error(self.cur_wcT, self.tar_wcT)
        return ((ang < 1.0) & (mm < 2.0)) | (self.steps > self.max_steps) | self.abnormal_pose()

    def reinit_all(self):
        self.reinit_all_count += 1
        self._build(self._new_scenes())

    def feedback(self, vel: torch.Tensor):
        """Apply camera velocities (B,6) for one 50 Hz step.  Scenes in an abnormal pose are driven by the supervisor's
        velocity instead (DataLoader.feedback, dataset.py:211-218).  Then re-initialisation by the `reinit` rule."""
        vel = vel.to(self.dev, torch.float64)
        if self.aug:
            vel = torch.where(self.abnormal_pose()[:, None], self.gt_vel, vel)
        self.cur_wcT = integrate(self.cur_wcT, vel)
        self.steps += 1
        self.calls += 1
        need = self.need_reinit()
        if self.reinit == "all":
            if float(need.double().mean()) > 0.5:
                self.reinit_all()
            return
        if bool(need.any()):
            idx = torch.nonzero(need).flatten()
            self.cur_wcT[idx] = self._sample_initial_pose(idx.numel())
            self.steps[idx] = 0
            self.fresh[idx] = True
            if self.aug:                      # a restarted scene gets a fresh drop-out process (dataset.py:79-83)
                from . import sim_aug
                new = sim_aug.BatchedObservationSampler(self.x_tar.double(), self.scene_of, self.B, self._tar_med, 5.0, self.gen)
                sc, pt = need, need[self.scene_of]
                s = self.sampler
                for name in ("state", "last_time", "tau_total", "no_event_rval"):
                    setattr(s, name, torch.where(pt, getattr(new, name), getattr(s, name)))
                s.center = torch.where(sc[:, None], new.center, s.center)
                s.dpose_norm_accum = torch.where(sc, new.dpose_norm_accum, s.dpose_norm_accum)
                for name in ("offset", "g0", "g1"):
                    setattr(s.perlin, name, torch.where(sc[:, None, None], getattr(new.perlin, name), getattr(s.perlin, name)))
                s.perlin.prev_floor = torch.where(sc[:, None, None], new.perlin.prev_floor, s.perlin.prev_floor)
                if s.prev_wcT is not None:
                    s.prev_wcT = torch.where(sc[:, None, None], self.cur_wcT, s.prev_wcT)
                self.mismatcher.reset(idx)
        if self.refresh_every and self.calls % self.refresh_every == 0:
            self._build(self._new_scenes())

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
