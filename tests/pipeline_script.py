"""Scripted frontend + driver shared by tests/golden/make_golden_pipeline.py (which runs the unmodified reference
pipeline) and the pipeline tests (which replay the recorded script through cns_b200).  No reference import here."""
import numpy as np

SEED = 2023          # demo_sim_Erender.py:9-11,42 seeds numpy right before set_target


class ScriptedFrontend(object):
    """Frontend stand-in with the `FrontendBase` interface; `corr_cls` is the Correspondence dataclass to emit."""

    def __init__(self, intrinsic, steps, corr_cls):
        self.intrinsic, self.steps, self.corr_cls = intrinsic, list(steps), corr_cls
        self.pos, self.tar, self.changed = 0, None, True

    def update_target_frame(self, image, mask=None):
        kind, payload = self.steps[self.pos]
        assert kind == "target"
        self.pos += 1
        self.tar, self.changed = payload, True
        return True

    def process_current_frame(self, image):
        kind, payload = self.steps[self.pos]
        assert kind == "frame"
        self.pos += 1
        if payload is None:
            return None
        changed, self.changed = self.changed, False
        n = len(self.tar["tar_pos"])
        return self.corr_cls(
            intrinsic=self.intrinsic, tar_img=None, tar_pos=self.tar["tar_pos"], cur_img=None,
            cur_pos=payload["cur_pos_aligned"], match=np.stack([np.arange(n), np.arange(n)], axis=1),
            valid_mask=payload["valid_mask"], cur_pos_aligned=payload["cur_pos_aligned"], detector_name="scripted",
            tar_img_changed=changed)


def drive(pipe, steps, seed=SEED):
    """Calls set_target / get_control_rate in script order; returns the per-frame answers."""
    answers = []
    image = np.zeros((4, 4, 3), np.uint8)
    for kind, payload in steps:
        if kind == "target":
            np.random.seed(seed)
            assert pipe.set_target(image, payload["dist_scale"])
            continue
        vel, data, timing = pipe.get_control_rate(image)
        answers.append((vel, data, timing))
    return answers
