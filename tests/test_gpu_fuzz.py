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
