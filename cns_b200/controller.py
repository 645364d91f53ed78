"""`GraphVSController`: checkpoint load + recurrent hidden-state ownership for the servo loop.

Mirror of `cns/benchmark/controller.py:10-38`: `__call__(data) -> np.ndarray (6,)`, hidden state
reset whenever `data.new_scene` is set.  Unlike the reference (whose `hasattr(ckpt, "net")` test is
always False for a dict, controller.py:15) both shipped checkpoint formats load here.

Host-resident GraphData (what the midend produces) goes through `cns_servo_step_host`: one pinned
staging copy in, one copy out, hidden state resident on the device - no per-tensor `.to(device)`.
Device-resident GraphData takes the ordinary `GraphVS.forward` path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, ckpt
from .data import GraphData
from .model import GraphVS


class GraphVSController(object):
    def __init__(self, ckpt_path: str, device="cuda:0", precision: str = "fp32"):
        self.device = torch.device(device)
        self.net = ckpt.load_model(ckpt_path, self.device, precision=precision)
        self.hidden = None
        self._vel = torch.empty(1, 6)

    def reset(self):
        self.hidden = None
        self._fresh = True

    def __call__(self, data: GraphData) -> np.ndarray:
        new_scene = bool(torch.as_tensor(data.new_scene).any())
        if new_scene:
            print("[INFO] Got new scene, set hidden state to zero")
            self.hidden = None
        if not data.x_cur.is_cuda:
            lib = _lib.load()
            b, keep = GraphVS.make_batch_struct(data)          # host pointers
            if self._vel.shape[0] != b.n_graphs:
                self._vel = torch.empty(b.n_graphs, 6)
            h = self.net._handle(self.device)
            _lib.check(lib.cns_servo_step_host(h.raw, C.byref(b), int(new_scene or getattr(self, "_fresh", True)),
                                               _lib.ptr(self._vel), None, None), "cns_servo_step_host")
            self._fresh = False
            return self._vel.squeeze(0).numpy().copy()
        with torch.no_grad():
            data = self.net.preprocess(data)
            raw_pred = self.net(data, self.hidden, check=True)
            self.hidden = raw_pred[-1]
            vel = self.net.postprocess(raw_pred, data)
        return vel.squeeze(0).cpu().numpy()
