"""Correspondence -> GraphData, the boundary the servo pipeline calls every frame.

Mirrors `cns/midend/corr2graph.py:11-22` (Midend), the `Correspondence` field layout
of `cns/frontend/utils.py:8-29` (input contract only - detectors/matchers are out of
scope) and the pinhole helper `CameraIntrinsic` of `cns/utils/perception.py:6-74`
(without its open3d / pybullet baggage).
"""
from dataclasses import dataclass
from typing import Optional

import numpy as np

from .graph_gen import GraphGenerator


class CameraIntrinsic(object):
    """Pinhole intrinsics; pixel <-> normalised camera plane."""

    def __init__(self, width, height, fx, fy, cx, cy):
        self.width, self.height = width, height
        self.K = np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]])

    fx = property(lambda self: self.K[0, 0])
    fy = property(lambda self: self.K[1, 1])
    cx = property(lambda self: self.K[0, 2])
    cy = property(lambda self: self.K[1, 2])

    def to_dict(self):
        return {"width": self.width, "height": self.height, "K": self.K.flatten().tolist()}

    @classmethod
    def from_dict(cls, data):
        k = data["K"]
        return cls(width=data["width"], height=data["height"], fx=k[0], fy=k[4], cx=k[2], cy=k[5])

    @classmethod
    def default(cls):
        return cls(width=640, height=480, fx=540, fy=540, cx=320, cy=240)

    def pixel_to_norm_camera_plane(self, uv: np.ndarray) -> np.ndarray:
        return (uv - np.array([self.cx, self.cy])) / np.array([self.fx, self.fy])

    def norm_camera_plane_to_pixel(self, xy: np.ndarray, clip=True, round=False) -> np.ndarray:
        uv = xy * np.array([self.fx, self.fy]) + np.array([self.cx, self.cy])
        if clip:
            uv = np.clip(uv, 0, [self.width - 1, self.height - 1])
        if round:
            uv = np.round(uv).astype(np.int32)
        return uv


@dataclass
class Correspondence(object):
    """Matched keypoints of one frame: `cur_pos_aligned[i]` matches `tar_pos[i]`
    wherever `valid_mask[i]` (pixel units, not normalised)."""
    intrinsic: CameraIntrinsic
    tar_img: Optional[np.ndarray]
    tar_pos: np.ndarray
    cur_img: Optional[np.ndarray]
    cur_pos: Optional[np.ndarray]
    match: Optional[np.ndarray]
    valid_mask: np.ndarray
    cur_pos_aligned: np.ndarray
    detector_name: str = "Unknown"
    tar_img_changed: bool = True

    @classmethod
    def concat(cls, corr):
        """Correspondences of several detectors on the same image pair, stacked (`cns/frontend/utils.py:31-68`);
        None entries are skipped, all None -> None.  `match` rows are shifted by the running keypoint counts (the
        reference shifts the current-side column by `cur_pos.shape[1]`, i.e. always 2 - nothing downstream of the
        frontend reads `match`, the graph path uses `cur_pos_aligned` / `valid_mask`)."""
        valid = [c for c in corr if c is not None]
        if not valid:
            return None
        tar_off = cur_off = 0
        matches = []
        for c in valid:
            m = np.array(c.match, copy=True)
            m[:, 0] += tar_off
            m[:, 1] += cur_off
            tar_off += c.tar_pos.shape[0]
            cur_off += c.cur_pos.shape[0]
            matches.append(m)
        return cls(
            intrinsic=valid[0].intrinsic, tar_img=valid[0].tar_img, cur_img=valid[0].cur_img,
            tar_pos=np.concatenate([c.tar_pos for c in valid], axis=0),
            cur_pos=np.concatenate([c.cur_pos for c in valid], axis=0),
            match=np.concatenate(matches, axis=0),
            valid_mask=np.concatenate([c.valid_mask for c in valid], axis=0),
            cur_pos_aligned=np.concatenate([c.cur_pos_aligned for c in valid], axis=0),
            detector_name=" + ".join(c.detector_name for c in valid),
            tar_img_changed=any(c.tar_img_changed for c in valid))


class Midend(object):
    def __init__(self):
        self.graph_generator_class = GraphGenerator
        self.graph_generator = None

    def get_graph_data(self, corr: Correspondence):
        if corr.tar_img_changed or self.graph_generator is None:
            self.graph_generator = self.graph_generator_class(
                corr.intrinsic.pixel_to_norm_camera_plane(corr.tar_pos))
        return self.graph_generator.get_data(
            current_points=corr.intrinsic.pixel_to_norm_camera_plane(corr.cur_pos_aligned),
            missing_node_indices=np.nonzero(~corr.valid_mask)[0])
