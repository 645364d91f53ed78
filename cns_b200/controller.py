"""`GraphVSController`: checkpoint load + recurrent hidden-state ownership for the servo loop.

Mirror of `cns/benchmark/controller.py:10-38`: `__call__(data) -> np.ndarray (6,)`, hidden state
reset whenever `data.new_scene` is set.  Unlike the reference (whose `hasattr(ckpt, "net")` test is
always False for a dict, controller.py:15) both shipped checkpoint formats load here.
"""
import numpy as np
import torch

from . import ckpt
from .data import GraphData


class GraphVSController(object):
    def __init__(self, ckpt_path: str, device="cuda:0", precision: str = "fp32"):
        self.device = torch.device(device)
        self.net = ckpt.load_model(ckpt_path, self.device, precision=precision)
        self.hidden = None

    def __call__(self, data: GraphData) -> np.ndarray:
        with torch.no_grad():
            data = data.to(self.device)
            data = self.net.preprocess(data)
            if bool(torch.as_tensor(data.new_scene).any()):
                print("[INFO] Got new scene, set hidden state to zero")
                self.hidden = None
            raw_pred = self.net(data, self.hidden, check=True)
            self.hidden = raw_pred[-1]
            vel = self.net.postprocess(raw_pred, data)
        return vel.squeeze(0).cpu().numpy()
