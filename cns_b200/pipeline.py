"""`CorrespondenceBasedPipeline`: images -> correspondences -> graph -> 6-DoF camera velocity.

Same surface as `cns/benchmark/pipeline.py:49-171` (`set_target`, `get_control_rate` returning
`(vel, data | None, timing)` with the reference's timing keys, `from_file(json)`).  The image
frontend (OpenCV detectors / SuperGlue, `cns/frontend/*`) is out of scope of this repository: any
object with the reference's `FrontendBase` interface (`update_target_frame`, `process_current_frame
-> Correspondence | None`, `.intrinsic`) is plugged in by the caller.
"""
import json
import time

import numpy as np
import torch

from .controller import GraphVSController
from .midend import CameraIntrinsic, Midend


class CorrespondenceBasedPipeline(object):
    REQUIRE_IMAGE_FORMAT = "BGR"

    def __init__(self, frontend, ckpt_path: str, intrinsic: CameraIntrinsic = None, device="cuda:0",
                 precision: str = "fp32"):
        self.device = torch.device(device)
        self.frontend = frontend
        self.intrinsic = intrinsic if intrinsic is not None else getattr(frontend, "intrinsic", None)
        self.midend = Midend()
        self.control = GraphVSController(ckpt_path, self.device, precision=precision)
        self.new_scene = True
        self.dist_scale = 1.0
        self.target_kwargs = dict()

    @classmethod
    def from_file(cls, config_path: str, frontend):
        """`pipeline.json` of the reference: keys intrinsic / checkpoint / device (detector, ransac and
        visualize configure the frontend the caller builds)."""
        with open(config_path, "r") as fp:
            config = json.load(fp)
        intr = config.get("intrinsic")
        intr = CameraIntrinsic.default() if intr is None else CameraIntrinsic.from_dict(intr)
        device = config.get("device", "cuda:0")
        return cls(frontend, config["checkpoint"], intr, device)

    def set_target(self, image: np.ndarray, dist_scale: float, **kwargs):
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
            data.intrinsic = getattr(self.frontend, "intrinsic", self.intrinsic)
            t0 = time.time()
            vel = self.control(data)
            timing["backend_time"] = time.time() - t0
            data.cur_img = image
            data.tar_img = getattr(corr, "tar_img", None)
        timing["total_time"] = timing["frontend_time"] + timing["backend_time"]   # as the reference (:170)
        return vel, data, timing
