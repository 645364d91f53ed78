"""CPU: the servo API surface - `CorrespondenceBasedPipeline(detector, ckpt_path, intrinsic, device,
ransac, vis)` / `from_file(path)` as `cns/benchmark/pipeline.py:52-99`, frontend registry, VisOpt."""
import ast
import inspect
import json
import os
import sys

import numpy as np
import pytest

import ref_import
from cns_b200 import CameraIntrinsic, Correspondence, pipeline as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
needs_ref = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")


class _Controller(object):
    def __init__(self, ckpt_path, device, precision="fp32"):
        self.ckpt_path, self.device, self.net = ckpt_path, device, None

    def __call__(self, data):
        return np.full(6, 0.5, np.float32)


def test_constructor_and_from_file_follow_the_reference(tmp_path, monkeypatch):
    sig = inspect.signature(P.CorrespondenceBasedPipeline.__init__)
    assert list(sig.parameters)[:7] == ["self", "detector", "ckpt_path", "intrinsic", "device", "ransac", "vis"]
    assert sig.parameters["device"].default == "cuda:0" and sig.parameters["ransac"].default is True
    assert sig.parameters["vis"].default == P.VisOpt.NO
    assert list(inspect.signature(P.CorrespondenceBasedPipeline.from_file).parameters)[0] == "config_path"
    assert P.VisOpt.ALL == P.VisOpt.KP | P.VisOpt.MATCH | P.VisOpt.GRAPH

    monkeypatch.setattr(P, "GraphVSController", _Controller)
    built = []

    def factory(intrinsic, config, ransac=True):
        built.append((config, ransac, intrinsic.fx))
        return type("F", (P.FrontendBase,), {"intrinsic": intrinsic})()

    monkeypatch.setitem(P._FRONTENDS, "akaze", factory)
    monkeypatch.setitem(P._FRONTENDS, "orb", factory)
    cfg = {"intrinsic": {"width": 640, "height": 480, "K": [615.0, 0, 315.0, 0, 616.0, 248.0, 0, 0, 1]},
           "detector": "AKAZE:02+ORB:13", "ransac": False, "checkpoint": "checkpoints/cns.pth", "device": "cuda:0",
           "visualize": "KP|MATCH"}
    path = tmp_path / "pipeline.json"
    path.write_text(json.dumps(cfg))
    pipe = P.CorrespondenceBasedPipeline.from_file(str(path))
    assert built == [("AKAZE:02", False, 615.0), ("ORB:13", False, 615.0)]
    assert isinstance(pipe.frontend, P.FrontendConcatWrapper) and len(pipe.frontend.frontends) == 2
    assert pipe.vis == P.VisOpt.KP | P.VisOpt.MATCH and pipe.control.ckpt_path == "checkpoints/cns.pth"
    assert pipe.midend.graph_generator_class.__name__ == "GraphGenerator"
    # the reference's shipped pipeline.json parses (detector "SIFT", visualize "MATCH")
    monkeypatch.setitem(P._FRONTENDS, "*", factory)
    ref_cfg = os.path.join(ref_import.REFERENCE_ROOT, "pipeline.json")
    if os.path.isfile(ref_cfg):
        pipe = P.CorrespondenceBasedPipeline.from_file(ref_cfg)
        assert pipe.vis == P.VisOpt.MATCH and built[-1][0] == "SIFT"
    # a classical-IBVS json checkpoint is outside the path
    with pytest.raises(NotImplementedError):
        P.CorrespondenceBasedPipeline("SIFT", "ibvs.json", CameraIntrinsic.default())


def test_unknown_detector_without_a_frontend_is_a_clear_error(monkeypatch):
    monkeypatch.setattr(P, "GraphVSController", _Controller)
    monkeypatch.setattr(P, "_FRONTENDS", {})
    monkeypatch.setitem(sys.modules, "cns.frontend.classic", None)      # as on a machine without the reference
    with pytest.raises(P.CnsError, match="register_frontend"):
        P.CorrespondenceBasedPipeline("SIFT", "checkpoints/cns.pth", CameraIntrinsic.default())


def test_pipeline_flow_with_a_frontend_object(monkeypatch):
    """set_target / get_control_rate bookkeeping without a GPU (controller stubbed): new-scene flag, distance scale,
    target kwargs, intrinsic / images attached, timing keys, the no-correspondence answer, visualiser hooks and a
    custom `midend.graph_generator_class`."""
    from pipeline_script import ScriptedFrontend, drive
    import torch
    monkeypatch.setattr(P, "GraphVSController", _Controller)
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "pipeline_control_rate.pt"), weights_only=False)
    intr = CameraIntrinsic.default()
    front = ScriptedFrontend(intr, gold["steps"], Correspondence)
    shown = []
    monkeypatch.setitem(P._VISUALIZERS, P.VisOpt.GRAPH, lambda d: shown.append("graph"))
    pipe = P.CorrespondenceBasedPipeline(front, "checkpoints/cns.pth", intr, vis=P.VisOpt.GRAPH | P.VisOpt.KP)
    made = []

    class Gen(pipe.midend.graph_generator_class):
        def __init__(self, pts):
            made.append(len(pts))
            super().__init__(pts)

    pipe.midend.graph_generator_class = Gen
    answers = drive(pipe, gold["steps"])
    assert made == [150, 60]                                  # one generator per target image
    assert len(shown) == 7
    for (vel, data, timing), g in zip(answers, gold["frames"]):
        assert sorted(timing.keys()) == g["timing_keys"] and (data is not None) == g["has_data"]
        assert vel.dtype == np.float32 and vel.shape == (6,)
        assert abs(timing["total_time"] - timing["frontend_time"] - timing["backend_time"]) < 1e-12
        if data is None:
            assert not vel.any()
            continue
        for k, v in g["fields"].items():                      # graph tensors bit-exact with the reference pipeline's
            assert torch.equal(getattr(data, k), v), k
        assert data.intrinsic is intr and data.cur_img is not None
    firsts = [bool(d.new_scene) for _, d, _ in answers if d is not None]
    assert firsts == [True, False, False, False, True, False, False]


@needs_ref
def test_signature_matches_reference_source():
    src = open(os.path.join(ref_import.REFERENCE_ROOT, "cns", "benchmark", "pipeline.py")).read()
    cls = [n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "CorrespondenceBasedPipeline"][0]
    ref = {f.name: [a.arg for a in f.args.args] for f in cls.body if isinstance(f, ast.FunctionDef)}
    for name in ("__init__", "from_file", "set_target", "get_control_rate"):
        mine = list(inspect.signature(getattr(P.CorrespondenceBasedPipeline, name)).parameters)
        if name == "from_file":
            mine = ["cls"] + mine
        assert mine[:len(ref[name])] == ref[name], name
