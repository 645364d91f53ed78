"""`CorrespondenceBasedPipeline`: images -> correspondences -> graph -> 6-DoF camera velocity.

Same surface as `cns/benchmark/pipeline.py:49-171`:

    CorrespondenceBasedPipeline(detector, ckpt_path, intrinsic, device="cuda:0", ransac=True, vis=VisOpt.NO)
    CorrespondenceBasedPipeline.from_file("pipeline.json")
    .set_target(image, dist_scale, **kwargs) -> bool
    .get_control_rate(image) -> (vel float32 (6,), GraphData | None, timing dict)

with the reference's timing keys (`total_time` = frontend + backend, midend recorded but not added, :170), the
zero velocity + `data=None` answer when no correspondence is found (:137-139) and the `VisOpt` flags.

The image frontends (OpenCV detectors + FLANN + RANSAC, SuperGlue: `cns/frontend/*`) are out of scope of this
repository.  `detector` is therefore resolved through a registry:

* a frontend OBJECT (anything with the reference's `FrontendBase` interface: `update_target_frame(image)`,
  `process_current_frame(image) -> Correspondence | None`, `.intrinsic`) is used as it is;
* a detector STRING ("SIFT", "AKAZE:02+ORB:13", "SuperGlue", ... - `get_frontend`, pipeline.py:32-46) is split on "+"
  and every part goes to the factory registered for it with `register_frontend` (exact name, then the name before
  ":"); with nothing registered the reference's own `cns.frontend` classes are used when that package is importable
  (the drop-in situation: this pipeline replaces the reference's inside its checkout), else a `CnsError` says so.

Visualisation (`vis`, `cns/utils/visualize.py`) is likewise a hook: `register_visualizer(VisOpt.KP, fn)`; flags
without a hook are accepted and ignored.  `ckpt_path` ending in ".json" selects the reference's classical IBVS
controller, which is not part of this path and raises.
"""
import json
import time
from enum import Flag, auto
from typing import Callable, Dict, List

import numpy as np
import torch

from ._lib import CnsError
from .controller import GraphVSController
from .midend import CameraIntrinsic, Correspondence, Midend


class VisOpt(Flag):
    NO = 0
    KP = auto()
    MATCH = auto()
    GRAPH = auto()
    ALL = KP | MATCH | GRAPH


class FrontendBase(object):
    """Interface of `cns/frontend/utils.py:71-79`."""

    def update_target_frame(self, image, mask=None) -> bool:
        raise NotImplementedError

    def process_current_frame(self, image) -> Correspondence:
        raise NotImplementedError

    def reset_intrinsic(self, intrinsic: CameraIntrinsic):
        self.intrinsic = intrinsic


class FrontendConcatWrapper(object):
    """Several frontends on the same images, correspondences concatenated (`cns/frontend/utils.py:82-101`)."""

    def __init__(self, frontends: List[FrontendBase]):
        self.frontends = frontends

    @property
    def intrinsic(self):
        return self.frontends[0].intrinsic

    def update_target_frame(self, image, mask=None):
        return all([f.update_target_frame(image, mask) for f in self.frontends])

    def process_current_frame(self, image):
        return Correspondence.concat([f.process_current_frame(image) for f in self.frontends])

    def reset_intrinsic(self, intrinsic: CameraIntrinsic):
        for f in self.frontends:
            f.reset_intrinsic(intrinsic)


_FRONTENDS: Dict[str, Callable] = {}
_VISUALIZERS: Dict[VisOpt, Callable] = {}


def register_frontend(name: str, factory: Callable):
    """`factory(intrinsic, config, ransac=True) -> frontend` for detector strings equal to `name` or starting with
    `name + ":"` (case-insensitive).  `name="*"` catches everything."""
    _FRONTENDS[name.lower()] = factory


def register_visualizer(flag: VisOpt, fn: Callable):
    """`fn(data)` for VisOpt.KP / VisOpt.GRAPH, `fn(corr)` for VisOpt.MATCH (pipeline.py:153-158)."""
    _VISUALIZERS[flag] = fn


def _reference_frontend(intrinsic, config: str, ransac: bool):
    try:
        if config.lower().startswith("superglue"):
            from cns.frontend.superglue import SuperGlue
            return SuperGlue(intrinsic, config, ransac=ransac)
        from cns.frontend.classic import Classic
        return Classic(intrinsic, config, ransac=ransac)
    except ImportError as e:
        raise CnsError(
            "no frontend registered for detector %r and the reference's cns.frontend package is not importable (%s). "
            "Image frontends are outside this repository: pass a frontend object as `detector`, or "
            "cns_b200.pipeline.register_frontend(name, factory)." % (config, e)) from None


def get_frontend(intrinsic: CameraIntrinsic, detectors, ransac=True):
    """'AKAZE:02+ORB:13' = AKAZE with rots [0,2] and ORB with rots [1,3] (pipeline.py:32-46)."""
    if not isinstance(detectors, str):
        return detectors                                   # already a frontend object
    frontends = []
    for config in detectors.split("+"):
        config = config.strip()
        key = config.lower()
        factory = _FRONTENDS.get(key) or _FRONTENDS.get(key.split(":")[0]) or _FRONTENDS.get("*")
        frontends.append(factory(intrinsic, config, ransac=ransac) if factory else
                         _reference_frontend(intrinsic, config, ransac))
    return FrontendConcatWrapper(frontends)


class CorrespondenceBasedPipeline(object):
    REQUIRE_IMAGE_FORMAT = "BGR"

    def __init__(self, detector, ckpt_path: str, intrinsic: CameraIntrinsic, device="cuda:0", ransac=True,
                 vis=VisOpt.NO, precision: str = "fp32"):
        self.device = torch.device(device)
        self.vis = vis
        self.frontend = get_frontend(intrinsic, detector, ransac)
        self.intrinsic = intrinsic if intrinsic is not None else getattr(self.frontend, "intrinsic", None)
        self.midend = Midend()
        if ckpt_path.endswith(".json"):
            raise NotImplementedError(
                "a .json checkpoint selects the reference's classical IBVSController (benchmark/controller.py:41-49), "
                "which is not on the path this library implements")
        self.control = GraphVSController(ckpt_path, self.device, precision=precision)
        self.new_scene = True
        self.dist_scale = 1.0
        self.target_kwargs = dict()

    @classmethod
    def from_file(cls, config_path: str, detector=None, **kwargs):
        """`pipeline.json` (pipeline.py:78-99): intrinsic / detector / checkpoint / device / ransac / visualize.
        `detector` (optional) overrides the file's entry, e.g. with a frontend object."""
        with open(config_path, "r") as fp:
            config: dict = json.load(fp)
        intrinsic = config.get("intrinsic", None)
        intrinsic = CameraIntrinsic.default() if intrinsic is None else CameraIntrinsic.from_dict(intrinsic)
        if detector is None:
            detector = config.get("detector", "SIFT")
        ckpt_path: str = config["checkpoint"]
        device: str = config.get("device", "cuda:0")
        ransac = config.get("ransac", True)
        vis_opts = config.get("visualize", "NO").split("|")
        vis = getattr(VisOpt, vis_opts[0])
        for opt in vis_opts[1:]:
            vis = vis | getattr(VisOpt, opt)
        return cls(detector, ckpt_path, intrinsic, device, ransac, vis, **kwargs)

    def set_target(self, image: np.ndarray, dist_scale: float, **kwargs):
        """image: BGR; dist_scale: distance from the camera to the scene centre at the target pose."""
        success = self.frontend.update_target_frame(image)
        self.new_scene = True
        self.dist_scale = dist_scale
        self.target_kwargs = kwargs
        return success

    def get_control_rate(self, image: np.ndarray):
        timing = {"frontend_time": 0, "midend_time": 0, "backend_time": 0, "total_time": 0}
        t0 = time.time()
        corr = self.frontend.process_current_frame(image)
        timing["frontend_time"] = time.time() - t0
        if corr is None:
            vel, data = np.zeros(6, dtype=np.float32), None
        else:
            t0 = time.time()
            data = self.midend.get_graph_data(corr)
            timing["midend_time"] = time.time() - t0
            if self.new_scene:
                data.start_new_scene()
            data.set_distance_scale(self.dist_scale)
            self.new_scene = False
            for k, v in self.target_kwargs.items():
                setattr(data, k, v)
            setattr(data, "intrinsic", self.frontend.intrinsic)
            for flag, arg in ((VisOpt.KP, data), (VisOpt.GRAPH, data), (VisOpt.MATCH, corr)):
                if (self.vis & flag) and flag in _VISUALIZERS:
                    _VISUALIZERS[flag](arg)
            t0 = time.time()
            vel = self.control(data)
            timing["backend_time"] = time.time() - t0
            setattr(data, "cur_img", image)           # for the SSIM stop criterion afterwards
            setattr(data, "tar_img", corr.tar_img)
        timing["total_time"] = timing["frontend_time"] + timing["backend_time"]   # as the reference (:170)
        return vel, data, timing
