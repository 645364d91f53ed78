"""Stop criteria of the servo loop and the match alignment in front of it, on the device.

`ErrorHoldingStopPolicy` / `PixelStopPolicy` keep the reference's interface (cns/benchmark/stop_policy.py:6-126:
constructor arguments, `reset()`, `calculate_error(data)`, `__call__(data, timestamp) -> bool`, attributes
`cur_err`, `min_err`, `accum_time`); the per-keypoint part of `PixelStopPolicy` - witness counts, decayed
weights, weighted error, 90th-percentile cut, weighted mean - runs in one kernel (`cns_pixel_stop_error`) with
its state resident on the GPU, and returns the same float32 the reference computes with numpy 2.x, bit for bit.
The scalar hold-time state machine stays on the host: it is a dozen comparisons per frame, and written with
python operators on that np.float32 it inherits numpy's scalar arithmetic exactly.
`BatchedPixelStop` evaluates B scenes per launch (the companion of `BatchedServo`).
`align_matches` is `Classic._align_cur_to_tar` + `pixel_to_norm_camera_plane` (cns/frontend/classic.py:142-148,
cns/utils/perception.py:66-68) for detections that already live on the GPU.
SSIMStopPolicy is image based and not part of this path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


class ErrorHoldingStopPolicy(object):
    def __init__(self, waiting_time, conduct_thresh, verbose=False):
        self.waiting_time = waiting_time
        self.conduct_thresh = conduct_thresh
        self.verbose = verbose
        self.min_err = None
        self.cur_err = float("inf")
        self.prev_t = None
        self.prev_err = None
        self.accum_time = 0

    def reset(self):
        self.min_err = None
        self.prev_t = None
        self.prev_err = None
        self.accum_time = 0

    def calculate_error(self, data):
        raise NotImplementedError

    def __call__(self, data, timestamp: float) -> bool:
        if data is None:
            self.reset()
            self.cur_err = float("inf")
            return False
        return self.update(self.calculate_error(data), timestamp)

    def update(self, cur_err, timestamp: float) -> bool:
        """The hold-time rule (stop_policy.py:31-66) given this frame's error: stop once the error has stayed
        within 1.1 min_err + 5e-4 of its best value for `waiting_time` seconds (crossings interpolated)."""
        self.cur_err = cur_err
        if cur_err > self.conduct_thresh:
            self.reset()
            return False
        if self.min_err is None or cur_err < 0.9 * self.min_err:
            self.min_err, self.prev_t, self.prev_err, self.accum_time = cur_err, timestamp, cur_err, 0
            return False
        level = 1.1 * self.min_err + 5e-4
        e0, e1, span = self.prev_err, cur_err, timestamp - self.prev_t
        self.prev_t, self.prev_err = timestamp, cur_err
        if e0 < level <= e1:
            held = (level - e0) / (e1 - e0) * span
        elif e1 < level <= e0:
            held = (level - e1) / (e0 - e1) * span
        elif e0 <= level and e1 <= level:
            held = span
        else:
            held = 0
        self.accum_time += held
        if self.verbose:
            print("[INFO] min_err = {:.3e}, cur_err = {:.3e}, dt = {:.3f}, accum = {:.3f}"
                  .format(self.min_err, cur_err, held, self.accum_time))
        return self.accum_time > self.waiting_time


class _PixelErrorState(object):
    """Device state + launch of cns_pixel_stop_error for a fixed keypoint layout."""

    def __init__(self, node_counts, device, decay):
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise _lib.CnsError("the pixel stop error runs on a CUDA device (there is no CPU fallback)")
        self.lib = _lib.load()
        counts = np.asarray(node_counts, dtype=np.int64).reshape(-1)
        self.B, self.N = len(counts), int(counts.sum())
        self.decay, self.maximum = decay, 1.0 / (1 - decay)
        ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        self.node_ptr = torch.from_numpy(ptr).to(self.dev)
        self.weight = torch.zeros(max(self.N, 1), dtype=torch.float32, device=self.dev)
        self.witness = torch.zeros(max(self.N, 1), dtype=torch.int32, device=self.dev)
        self.err = torch.empty(self.B, dtype=torch.float32, device=self.dev)
        self.ws = torch.empty(self.lib.cns_pixel_stop_workspace_bytes(self.N), dtype=torch.uint8, device=self.dev)
        self.reset_flags = torch.zeros(self.B, dtype=torch.uint8, device=self.dev)

    def clear(self):
        self.weight.zero_()
        self.witness.zero_()

    def launch(self, pos_cur, pos_tar, node_mask, reset_flags=None):
        f = lambda t: t.to(self.dev, torch.float32).contiguous()
        pc, pt = f(pos_cur), f(pos_tar)
        m = node_mask.to(self.dev)
        m = (m.view(torch.uint8) if m.dtype == torch.bool else m.to(torch.uint8)).contiguous()
        if pc.shape != (self.N, 2) or pt.shape != (self.N, 2) or m.numel() != self.N:
            raise ValueError("expected %d keypoints, got pos_cur %s node_mask %s" % (self.N, tuple(pc.shape), tuple(m.shape)))
        L = _lib
        L.check(self.lib.cns_pixel_stop_error(L.ptr(pc), L.ptr(pt), L.ptr(m), L.ptr(self.node_ptr), L.ptr(reset_flags),
                                              self.N, self.B, self.decay, self.maximum, L.ptr(self.weight),
                                              L.ptr(self.witness), L.ptr(self.err), None, L.ptr(self.ws),
                                              self.ws.numel(), L.stream_ptr(self.dev)), "cns_pixel_stop_error")
        return self.err


class PixelStopPolicy(ErrorHoldingStopPolicy):
    def __init__(self, waiting_time, conduct_thresh=0.01, device="cuda:0", verbose=False):
        super().__init__(waiting_time, conduct_thresh, verbose)
        self.device = device
        self.decay = 0.95
        self.maximum = 1.0 / (1 - self.decay)
        self.state = None

    @property
    def weight(self):
        return None if self.state is None else self.state.weight[:self.state.N].cpu().numpy()

    @property
    def witness_count(self):
        return None if self.state is None else self.state.witness[:self.state.N].cpu().numpy().astype(np.int64)

    def reset(self):
        if getattr(self, "state", None) is not None:
            self.state.clear()
        return super().reset()

    def calculate_error(self, data):
        mask = getattr(data, "node_mask")
        if self.state is None or self.state.N != mask.numel():
            self.state = _PixelErrorState([mask.numel()], self.device, self.decay)
        err = self.state.launch(getattr(data, "pos_cur"), getattr(data, "pos_tar"), mask)
        return np.float32(err.item())                  # the one scalar the host rule needs (a 4-byte D2H)


class BatchedPixelStop(object):
    """B stop policies evaluated per launch: `__call__(pos_cur, pos_tar, node_mask, timestamp) -> (B,) bool`.
    The tensors are the batched keypoint arrays (sum N rows, scenes back to back as in `BatchedServo`), on the
    device or the host.  Scene resets (error above `conduct_thresh`) are applied by the next launch through
    per-scene flags, so no per-scene device work is issued from python."""

    def __init__(self, node_counts, waiting_time, conduct_thresh=0.01, device="cuda:0"):
        self.state = _PixelErrorState(node_counts, device, 0.95)
        self.rules = [ErrorHoldingStopPolicy(waiting_time, conduct_thresh) for _ in range(self.state.B)]
        self.pending_reset = np.zeros(self.state.B, dtype=np.uint8)
        self.flags_host = torch.zeros(self.state.B, dtype=torch.uint8).pin_memory()
        self.err_host = torch.empty(self.state.B, dtype=torch.float32).pin_memory()

    def reset(self):
        self.state.clear()
        self.pending_reset[:] = 0
        for r in self.rules:
            r.reset()

    @property
    def cur_err(self):
        return np.array([r.cur_err for r in self.rules], dtype=np.float32)

    def errors(self, pos_cur, pos_tar, node_mask) -> torch.Tensor:
        """Device tensor (B,) of this frame's errors (no synchronisation)."""
        self.flags_host.copy_(torch.from_numpy(self.pending_reset))
        self.state.reset_flags.copy_(self.flags_host, non_blocking=True)
        self.pending_reset[:] = 0
        return self.state.launch(pos_cur, pos_tar, node_mask, self.state.reset_flags)

    def __call__(self, pos_cur, pos_tar, node_mask, timestamp: float) -> np.ndarray:
        self.err_host.copy_(self.errors(pos_cur, pos_tar, node_mask), non_blocking=True)
        torch.cuda.current_stream(self.state.dev).synchronize()
        errs = self.err_host.numpy()
        stop = np.zeros(self.state.B, dtype=bool)
        for b, rule in enumerate(self.rules):
            stop[b] = rule.update(errs[b], timestamp)
            if rule.min_err is None:                  # the rule dropped its history: PixelStopPolicy.reset
                self.pending_reset[b] = 1
        return stop


def align_matches(cur_pos: torch.Tensor, matches: torch.Tensor, n_tar: int, intrinsic):
    """cur_pos (n_cur,2) float32 pixel coordinates and matches (M,2) int64 rows [target index, current index],
    both on the GPU -> (cur_xy (n_tar,2) float32 normalized + aligned with the target keypoints, valid (n_tar,)
    bool), bit-exact with the reference's host path (see module docstring)."""
    if cur_pos.device.type != "cuda":
        raise _lib.CnsError("align_matches runs on a CUDA device (there is no CPU fallback)")
    lib, L, dev = _lib.load(), _lib, cur_pos.device
    cur_pos = cur_pos.to(torch.float32).contiguous()
    matches = matches.to(dev, torch.int64).contiguous()
    xy = torch.empty(n_tar, 2, dtype=torch.float32, device=dev)
    valid = torch.empty(n_tar, dtype=torch.uint8, device=dev)
    ws = torch.empty(n_tar + 1, dtype=torch.int32, device=dev)
    L.check(lib.cns_align_matches(L.ptr(cur_pos), cur_pos.shape[0], L.ptr(matches), matches.shape[0], n_tar,
                                  float(intrinsic.fx), float(intrinsic.fy), float(intrinsic.cx), float(intrinsic.cy),
                                  L.ptr(xy), L.ptr(valid), L.ptr(ws), ws.numel() * 4, L.stream_ptr(dev)),
            "cns_align_matches")
    return xy, valid.view(torch.bool)
