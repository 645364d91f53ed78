"""Synthetic servo scenes: the input distribution of the reference's point simulator.

Restates, as seeded array code, what feeds the hot path in the reference's training /
benchmark loops - it is the data format on the input side of the path, not an
accelerated component:

* scene points: half the time a mix of 20 % uniform + 80 % clustered points, else
  uniform in a cylinder r=0.2, h=0.1 (sim/environment.py:21-25, sim/sampling.py:73-190);
* camera poses on a spherical shell looking at the origin with bounded roll/pitch/yaw
  perturbation (sim/sampling.py:7-34; ranges sim/environment.py:52-62);
* pinhole projection to the normalised camera plane (utils/perception.py:62-68,128-148);
* ground-truth velocity: straight-line PBVS (sim/supervisor.py:92-101) and its
  distance-decoupled form `vel_si`, `tPo_norm` (sim/supervisor.py:286-295);
* pose integration at 50 Hz (sim/environment.py:104-121).

Random streams are numpy `Generator`s seeded by the caller; they are NOT the
reference's global `np.random` stream (only the distribution is shared).
"""
from typing import List, Optional, Tuple

import numpy as np
import torch
from scipy.spatial.transform import Rotation

from .data import Batch, GraphData
from .graph_gen import GraphGenerator

OBS_R, OBS_H = 0.2, 0.1
DT = 1.0 / 50.0


# ----------------------------------------------------------------------------- poses
def look_at_pose(rng, r_rng, phi_rng_deg, drz, dry, drx) -> np.ndarray:
    """World<-camera transform of a camera at radius r, elevation phi, looking at the
    origin, then perturbed by intrinsic z/y/x rotations (degrees)."""
    r = rng.uniform(*r_rng)
    theta = rng.uniform(-np.pi, np.pi)
    z = rng.uniform(np.sin(np.deg2rad(phi_rng_deg[0])), np.sin(np.deg2rad(phi_rng_deg[1])))
    phi = np.arcsin(z)
    d = np.array([np.cos(phi) * np.cos(theta), np.cos(phi) * np.sin(theta), z])
    z_axis = -d
    y_axis = np.array([np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), -np.cos(phi)])
    x_axis = np.cross(y_axis, z_axis)
    rot = np.stack([x_axis, y_axis, z_axis], axis=1)
    for axis, lim in (((0, 0, 1), drz), ((0, 1, 0), dry), ((1, 0, 0), drx)):
        ang = np.deg2rad(rng.uniform(-lim, lim))
        rot = rot @ Rotation.from_rotvec(np.array(axis, dtype=float) * ang).as_matrix()
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = rot, d * r
    return T


def sample_target_pose(rng):
    return look_at_pose(rng, (0.5, 0.9), (70, 90), 15, 5, 5)


def sample_initial_pose(rng):
    return look_at_pose(rng, (0.5, 0.9), (30, 90), 60, 10, 10)


# ---------------------------------------------------------------------------- points
def _ball2(rng, n):
    u = rng.normal(size=(n, 2))
    return u / np.linalg.norm(u, axis=-1, keepdims=True) * np.sqrt(rng.random((n, 1)))


def points_uniform(rng, n) -> np.ndarray:
    n = max(int(n), 3)
    xy = _ball2(rng, n) * OBS_R
    return np.concatenate([xy, (rng.random((n, 1)) - 0.5) * OBS_H], axis=-1)


def _ellipse(rng, n, a, b):
    ang = rng.uniform(0, 2 * np.pi, n)
    # area-uniform inside the ellipse: scale a unit-disc sample
    rad = np.sqrt(rng.random(n))
    return np.stack([a * rad * np.cos(ang), b * rad * np.sin(ang)], axis=-1)


def points_clustered(rng, n, k) -> np.ndarray:
    sizes = np.ones(k, dtype=np.int64) + np.bincount(rng.integers(0, k, n - k), minlength=k)
    centres = points_uniform(rng, k)
    r_max = OBS_R / np.log2(k) * 1.5
    out = []
    for c in range(k):
        a = rng.uniform(0.2 * r_max, 0.8 * r_max)
        p = _ellipse(rng, sizes[c], a, r_max - a)
        p = np.concatenate([p, (rng.random((len(p), 1)) - 0.5) * OBS_H * 0.2], axis=-1)
        rot = np.eye(3)
        for axis, lim in (((0, 0, 1), 180), ((0, 1, 0), 45), ((1, 0, 0), 45)):
            ang = np.deg2rad(rng.uniform(-lim, lim))
            rot = rot @ Rotation.from_rotvec(np.array(axis, dtype=float) * ang).as_matrix()
        out.append(p @ rot.T + centres[c])
    return np.concatenate(out, axis=0)


def sample_points(rng, n) -> np.ndarray:
    n = max(int(n), 4)
    if rng.random() < 0.5 and n >= 20:
        n_uni = int(round(n * 0.2))
        k = int(rng.integers(3, int(np.log2(n))))
        return np.concatenate([points_uniform(rng, n_uni), points_clustered(rng, n - n_uni, k)], 0)
    return points_uniform(rng, n)


# ------------------------------------------------------------------------ projection
def project_norm(wcT: np.ndarray, points: np.ndarray) -> np.ndarray:
    """World points -> normalised image plane (x/z, y/z) of the camera at pose wcT."""
    cwT = np.linalg.inv(wcT)
    pc = points @ cwT[:3, :3].T + cwT[:3, 3]
    return pc[:, :2] / pc[:, 2:3]


def pbvs_straight(cur_wcT, tar_wcT) -> np.ndarray:
    tcT = np.linalg.inv(tar_wcT) @ cur_wcT
    u = Rotation.from_matrix(tcT[:3, :3]).as_rotvec()
    return np.concatenate([-tcT[:3, :3].T @ tcT[:3, 3], -u])


def scene_distance(tar_wcT, points) -> float:
    cwT = np.linalg.inv(tar_wcT)
    return float(np.linalg.norm(points.mean(0) @ cwT[:3, :3].T + cwT[:3, 3]))


def integrate(cur_wcT, vel, dt=DT) -> np.ndarray:
    dT = np.eye(4)
    dT[:3, :3] = Rotation.from_rotvec(np.asarray(vel[3:], dtype=float) * dt).as_matrix()
    dT[:3, 3] = np.asarray(vel[:3], dtype=float) * dt
    out = cur_wcT @ dT
    out[:3, :3] = Rotation.from_matrix(out[:3, :3]).as_matrix()
    return out


def pose_error(cur_wcT, tar_wcT) -> Tuple[float, float]:
    dT = np.linalg.inv(cur_wcT) @ tar_wcT
    ang = np.linalg.norm(Rotation.from_matrix(dT[:3, :3]).as_rotvec()) / np.pi * 180
    return float(ang), float(np.linalg.norm(dT[:3, 3]) * 1000)


# ---------------------------------------------------------------------------- scenes
class Scene(object):
    """One servo scene: points, target pose, current pose and its graph generator."""

    def __init__(self, rng, n_points: int, generator: Optional[GraphGenerator] = None,
                 points: Optional[np.ndarray] = None, tar_wcT: Optional[np.ndarray] = None):
        self.rng = rng
        self.points = sample_points(rng, n_points) if points is None else points
        self.tar_wcT = sample_target_pose(rng) if tar_wcT is None else tar_wcT
        self.cur_wcT = sample_initial_pose(rng)
        self.tar_xy = project_norm(self.tar_wcT, self.points)
        self.generator = GraphGenerator(self.tar_xy) if generator is None else generator
        self.tPo_norm = scene_distance(self.tar_wcT, self.points)

    def frame(self, missing_frac: float = 0.0, with_labels: bool = False) -> GraphData:
        n = len(self.points)
        cur_xy = project_norm(self.cur_wcT, self.points)
        missing = None
        if missing_frac > 0:
            k = int(round(n * missing_frac))
            missing = np.sort(self.rng.choice(n, size=k, replace=False)) if k else None
        data = self.generator.get_data(cur_xy, missing_node_indices=missing)
        data.set_distance_scale(self.tPo_norm)
        if with_labels:
            vel = pbvs_straight(self.cur_wcT, self.tar_wcT)
            vel_si = vel.copy()
            vel_si[:3] /= self.tPo_norm + 1e-7
            data.vel = torch.from_numpy(vel).float().unsqueeze(0)
            data.vel_si = torch.from_numpy(vel_si).float().unsqueeze(0)
        return data

    def step(self, vel):
        self.cur_wcT = integrate(self.cur_wcT, vel)


def make_scenes(n_graphs: int, n_points: int, seed: int = 2022, n_layouts: int = 64) -> List[Scene]:
    """`n_graphs` scenes over `min(n_graphs, n_layouts)` distinct clustered layouts
    (clustering is O(N^2) per layout on the host, so layouts are tiled; every scene
    still gets its own initial camera pose)."""
    rng = np.random.default_rng(seed)
    # AffinityPropagation / KMeans draw from numpy's global RandomState
    state = np.random.get_state()
    np.random.seed(seed)
    try:
        base = [Scene(rng, n_points) for _ in range(min(n_graphs, n_layouts))]
    finally:
        np.random.set_state(state)
    scenes = list(base)
    while len(scenes) < n_graphs:
        src = base[len(scenes) % len(base)]
        scenes.append(Scene(rng, n_points, generator=src.generator, points=src.points,
                            tar_wcT=src.tar_wcT))
    return scenes


def make_batch(n_graphs: int, n_points: int, seed: int = 2022, missing_frac: float = 0.1,
               n_layouts: int = 64, with_labels: bool = False) -> Batch:
    scenes = make_scenes(n_graphs, n_points, seed, n_layouts)
    return Batch.from_data_list([s.frame(missing_frac, with_labels) for s in scenes])
