"""Generates tests/golden/pipeline_control_rate.pt: the UNMODIFIED reference `CorrespondenceBasedPipeline`
(cns/benchmark/pipeline.py:49-171, with its own Midend / GraphVSController / GraphVS on the PyG shim, CPU) driven by a
scripted frontend for two targets and a frame without correspondences.  Run in the build container only:

    python tests/golden/make_golden_pipeline.py

Recorded per call: the frontend's output (pixel keypoints, aligned matches, validity) and the pipeline's answer
(velocity, whether a graph came back, graph fields).  tests/test_pipeline_*.py replay the same frontend script
through cns_b200's pipeline.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_import  # noqa: E402

ref_import.setup()
import cns.benchmark.pipeline as ref_pipeline  # noqa: E402
from cns.frontend.utils import Correspondence  # noqa: E402
from cns.utils.perception import CameraIntrinsic  # noqa: E402
from cns_b200 import synth  # noqa: E402  (scene sampler only)

SEED = 2023          # demo_sim_Erender.py:9-11,42 seeds numpy right before set_target


def script(seed=SEED):
    """The frontend's answers: [(kind, payload)], kind in {"target", "frame"}."""
    rng = np.random.default_rng(seed)
    intr = CameraIntrinsic.default()
    out = []
    for n_pts, n_frames in ((150, 5), (60, 3)):
        pts = synth.sample_points(rng, n_pts)
        tar_wcT, cur_wcT = synth.sample_target_pose(rng), synth.sample_initial_pose(rng)
        tar_uv = synth.project_norm(tar_wcT, pts) * [intr.fx, intr.fy] + [intr.cx, intr.cy]
        out.append(("target", dict(tar_pos=tar_uv.astype(np.float32), dist_scale=float(synth.scene_distance(tar_wcT, pts)))))
        for t in range(n_frames):
            if n_pts == 150 and t == 2:
                out.append(("frame", None))                 # no matched keypoints in this image
                continue
            cur_uv = synth.project_norm(cur_wcT, pts) * [intr.fx, intr.fy] + [intr.cx, intr.cy]
            valid = rng.random(len(pts)) > (0.15 if t else 0.0)
            out.append(("frame", dict(cur_pos_aligned=cur_uv.astype(np.float32), valid_mask=valid)))
            cur_wcT = synth.integrate(cur_wcT, synth.pbvs_straight(cur_wcT, tar_wcT) * 2)
    return out


from pipeline_script import ScriptedFrontend, drive  # noqa: E402


if __name__ == "__main__":
    steps = script()
    intr = CameraIntrinsic.default()
    real_get_frontend = ref_pipeline.get_frontend
    ref_pipeline.get_frontend = lambda intrinsic, detector, ransac=True: ScriptedFrontend(intrinsic, steps, Correspondence)
    pipe = ref_pipeline.CorrespondenceBasedPipeline(
        "scripted", os.path.join(ROOT, "checkpoints", "cns_state_dict.pth"), intr, device="cpu")
    ref_pipeline.get_frontend = real_get_frontend
    keep = ("x_cur", "l0_to_l1_edge_index_j_cur", "l0_to_l1_edge_index_i_cur", "l1_dense_edge_index_cur", "cluster_mask",
            "cluster_centers_index", "new_scene", "tPo_norm")
    frames = []
    for vel, data, timing in drive(pipe, steps):
        frames.append(dict(vel=torch.from_numpy(np.asarray(vel).copy()), has_data=data is not None,
                           timing_keys=sorted(timing.keys()),
                           fields=None if data is None else {k: getattr(data, k).clone() for k in keep}))
    torch.save(dict(seed=SEED, steps=steps, frames=frames), os.path.join(HERE, "pipeline_control_rate.pt"))
    print("pipeline_control_rate.pt", os.path.getsize(os.path.join(HERE, "pipeline_control_rate.pt")), len(frames), "frames")
    for f in frames:
        print(f["has_data"], f["vel"].numpy().round(4))
