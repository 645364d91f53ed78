"""GPU (-m gpu): `CorrespondenceBasedPipeline.set_target / get_control_rate` (the API north_star names first) on
cuda:0 against (a) the velocities the UNMODIFIED reference pipeline produced for the same frontend script
(tests/golden/pipeline_control_rate.pt, made by tests/golden/make_golden_pipeline.py) and (b) the CPU oracle."""
import os
import sys

import numpy as np
import pytest
import torch

import cns_oracle
from cns_b200 import CameraIntrinsic, Correspondence
from cns_b200.pipeline import CorrespondenceBasedPipeline, VisOpt

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


@pytest.mark.parametrize("ckpt_name", ["cns_state_dict.pth", "cns.pth"])
def test_get_control_rate_matches_reference_pipeline(state_dict, ckpt_name):
    from pipeline_script import ScriptedFrontend, drive
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "pipeline_control_rate.pt"), weights_only=False)
    intr = CameraIntrinsic.default()
    front = ScriptedFrontend(intr, gold["steps"], Correspondence)
    pipe = CorrespondenceBasedPipeline(front, os.path.join(ROOT, "checkpoints", ckpt_name), intr, device="cuda:0",
                                       ransac=True, vis=VisOpt.NO)
    answers = drive(pipe, gold["steps"])
    assert len(answers) == len(gold["frames"]) == 8
    hidden = None
    for t, ((vel, data, timing), g) in enumerate(zip(answers, gold["frames"])):
        assert (data is not None) == g["has_data"]
        ref = g["vel"].numpy()
        if data is None:
            assert vel.dtype == np.float32 and not vel.any() and not ref.any()
            continue
        for k, v in g["fields"].items():
            assert torch.equal(getattr(data, k), v), (t, k)
        assert np.abs(vel - ref).max() < 1e-4 * np.abs(ref).max(), (t, vel, ref)          # fp32 mode, rel 1e-4
        if bool(data.new_scene):
            hidden = None
        ov, on, hidden = cns_oracle.forward(state_dict, data, hidden, dtype=torch.float64)
        ovel = cns_oracle.postprocess((ov, on), data.tPo_norm.reshape(1))[0].numpy()
        assert np.abs(vel - ovel).max() < 1e-4 * np.abs(ovel).max(), t
        assert timing["backend_time"] > 0 and timing["midend_time"] > 0


def test_pipeline_from_file_on_gpu(tmp_path):
    import json
    from pipeline_script import ScriptedFrontend, drive
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "pipeline_control_rate.pt"), weights_only=False)
    cfg = {"detector": "SIFT", "checkpoint": os.path.join(ROOT, "checkpoints", "cns.pth"), "device": "cuda:0"}
    path = tmp_path / "pipeline.json"
    path.write_text(json.dumps(cfg))
    front = ScriptedFrontend(CameraIntrinsic.default(), gold["steps"][:3], Correspondence)
    pipe = CorrespondenceBasedPipeline.from_file(str(path), detector=front, precision="f16")
    (vel, data, _), (vel2, _, _) = drive(pipe, gold["steps"][:3])
    ref = gold["frames"][0]["vel"].numpy()
    assert np.abs(vel - ref).max() < 2e-2 * np.abs(ref).max()                               # 16-bit operand mode
    assert np.abs(vel2 - gold["frames"][1]["vel"].numpy()).max() < 2e-2 * np.abs(gold["frames"][1]["vel"].numpy()).max()
