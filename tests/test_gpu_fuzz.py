"""GPU (-m gpu): randomised soak of every forward variant (fused / per-layer complete-graph / general CSR / servo)
in every precision mode against the fp64 oracle - scripts/fuzz_forward.py with fixed seeds."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_randomised_forward_soak():
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import fuzz_forward
    res = fuzz_forward.run(n_rounds=8)
    for prec, (ok, worst, detail) in res.items():
        assert ok, "%s: %s (worst case %s)" % (prec, worst, detail)


def test_randomised_clustering_soak():
    """Device graph construction == host (scikit-learn) construction bit for bit over random target point sets, including
    degenerate ones (lines, near-duplicates, tight blobs) that end in the KMeans(4) fallback: scripts/fuzz_clustering.py."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import fuzz_clustering
    assert fuzz_clustering.run(n_sets=60, verbose=False) == 0


def test_randomised_backward_soak():
    """Parameter gradients vs fp64 autograd over random small ragged training batches (1..8 graphs of 4..512 keypoints,
    1..3 recurrent frames, up to 60 % missing keypoints), bar as in test_gpu_backward.py: scripts/fuzz_backward.py.
    (Seeds 5000..5009; over 40 seeds two batches sit on a ReLU / max kink of the network - scripts/kink_scan.py - where
    fp32 and fp64 legitimately take different sides; DESIGN.md section 4, precision policy.)"""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import fuzz_backward
    assert fuzz_backward.run(n_rounds=10, verbose=False) == 0


def test_randomised_controller_soak():
    """Batch-1 servo path (GraphVSController -> cns_servo_step_host) over random sequences: targets switched mid-sequence,
    4..600 keypoints, up to all but one keypoint missing, whole clusters hidden - every frame against the fp64 oracle
    (fp32 mode 1e-4, f16 mode 2e-2): scripts/fuzz_controller.py."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import fuzz_controller
    assert fuzz_controller.run(n_seq=15, verbose=False) == 0
