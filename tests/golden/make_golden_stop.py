"""Generates tests/golden/stop_policy.npz and align_matches.npz by running the UNMODIFIED reference
(cns/benchmark/stop_policy.py PixelStopPolicy, cns/frontend/classic.py Classic._align_cur_to_tar,
cns/utils/perception.py CameraIntrinsic) in the build container:

    python tests/golden/make_golden_stop.py

stop_policy.npz: a 150-frame servo-like sequence (current keypoints converge on the target ones with
noise, 10 % random misses per frame, a block of keypoints first seen late, one frame without a graph, one
relapse above conduct_thresh) -> per-frame error, stop decision, min_err / accum_time, final weights.
align_matches.npz: random matches with duplicate target indices -> aligned normalized coordinates + mask.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_import  # noqa: E402

ref_import.setup()
from cns.benchmark.stop_policy import PixelStopPolicy  # noqa: E402
from cns.frontend.classic import Classic  # noqa: E402
from cns.utils.perception import CameraIntrinsic  # noqa: E402


def stop_case(seed=0, n=300, frames=150, dt=0.04):
    rng = np.random.default_rng(seed)
    tar = rng.uniform(-0.5, 0.5, size=(n, 2)).astype(np.float32)
    offset = rng.normal(0, 0.08, size=(n, 2)).astype(np.float32)
    pol = PixelStopPolicy(waiting_time=0.8, conduct_thresh=0.01)
    cur_all, mask_all, err, stop, min_err, accum, has = [], [], [], [], [], [], []
    for k in range(frames):
        scale = 0.9 ** k + (0.5 if k == 90 else 0.0)                    # relapse at frame 90
        cur = (tar + offset * np.float32(scale) + rng.normal(0, 2e-4, size=(n, 2))).astype(np.float32)
        mask = rng.random(n) > 0.1
        if k < 40:
            mask[: n // 5] = False                                      # first seen late
        graph = k != 60                                                 # one frame without correspondences
        data = types.SimpleNamespace(pos_cur=torch.from_numpy(cur), pos_tar=torch.from_numpy(tar),
                                     node_mask=torch.from_numpy(mask)) if graph else None
        with contextlib.redirect_stdout(io.StringIO()):
            s = pol(data, k * dt)
        cur_all.append(cur); mask_all.append(mask); has.append(graph)
        err.append(np.float32(pol.cur_err)); stop.append(bool(s))
        min_err.append(np.float32(-1 if pol.min_err is None else pol.min_err))
        accum.append(np.float64(pol.accum_time))
    return dict(tar=tar, cur=np.stack(cur_all), mask=np.stack(mask_all), has_graph=np.array(has),
                dt=np.float64(dt), err=np.array(err), stop=np.array(stop), min_err=np.array(min_err),
                accum=np.array(accum), weight=pol.weight.copy(), witness=pol.witness_count.copy())


def align_case(seed=1, n_tar=200, n_cur=260, n_match=150):
    rng = np.random.default_rng(seed)
    cur_pos = rng.uniform(0, [640, 480], size=(n_cur, 2)).astype(np.float32)
    tar_pos = rng.uniform(0, [640, 480], size=(n_tar, 2)).astype(np.float32)
    matches = np.stack([rng.integers(0, n_tar, n_match), rng.integers(0, n_cur, n_match)], 1).astype(np.int64)
    fake = types.SimpleNamespace(tar_pos=tar_pos)
    aligned, valid = Classic._align_cur_to_tar(fake, cur_pos, matches)
    intr = CameraIntrinsic.default()
    xy = torch.from_numpy(intr.pixel_to_norm_camera_plane(aligned)).float().numpy()     # graph_gen.py:257-263
    return dict(cur_pos=cur_pos, n_tar=np.int64(n_tar), matches=matches, aligned_px=aligned, valid=valid, xy=xy,
                intr=np.array([intr.fx, intr.fy, intr.cx, intr.cy], dtype=np.float64))


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "stop_policy.npz"), **stop_case())
    np.savez_compressed(os.path.join(HERE, "align_matches.npz"), **align_case())
    g = np.load(os.path.join(HERE, "stop_policy.npz"))
    print("stop frames:", np.nonzero(g["stop"])[0][:5], "min err", g["err"][np.isfinite(g["err"])].min())
