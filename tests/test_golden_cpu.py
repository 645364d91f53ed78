"""CPU: the oracle and the host-side graph builder against the committed golden fixtures
(tests/golden/*.pt, produced by the unmodified reference - see make_golden.py)."""
import glob
import os

import numpy as np
import pytest
import torch

import cns_oracle
from cns_b200.graph_gen import GraphGenerator

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "graphgen_*.pt"))))
def test_graph_generator_bit_exact(path):
    """Cluster assignment, centres and every edge list identical to the reference's (bit-exact)."""
    g = torch.load(path, weights_only=False)
    np.random.seed(g["seed"])            # AP / KMeans noise comes from numpy's global stream
    mine = GraphGenerator(g["points"])
    assert np.array_equal(mine.yhat, g["yhat"])
    assert np.array_equal(mine.cluster_ids, g["cluster_ids"])
    assert np.array_equal(mine.cluster_points_count, g["counts"])
    assert np.array_equal(mine.cluster_centers_index, g["centers"])
    assert mine.cluster_centers_index.dtype == np.int64
    assert np.array_equal(mine.belong_cluster_index, g["belong"])
    assert np.array_equal(mine.l0_to_l1_edge_index, g["l0_to_l1"])
    assert np.array_equal(mine.l1_dense_edge_index, g["l1_dense"])
    assert np.array_equal(mine.node2j_index, g["node2j"])
    assert np.array_equal(mine.to_center, g["to_center"])
    for fr in g["frames"]:
        _, l1, l0l1, nmask, cmask = mine._get_graph(fr["missing"])
        assert np.array_equal(l1, fr["l1"])
        assert np.array_equal(l0l1, fr["l0l1"])
        assert np.array_equal(nmask, fr["node_mask"])
        assert np.array_equal(cmask, fr["cluster_mask"])


@pytest.mark.parametrize("name", ["forward_single.pt", "forward_ragged.pt"])
def test_oracle_matches_reference_outputs(name, state_dict):
    """oracle/cns_oracle.py reproduces the reference GraphVS outputs over 3 recurrent frames."""
    case = _load(name)
    hidden = None
    for st in case["steps"]:
        vec, nrm, hidden = cns_oracle.forward(state_dict, st["fields"], hidden)
        vel = cns_oracle.postprocess((vec, nrm), st["fields"]["tPo_norm"])
        for got, ref in ((vec, st["vel_si_vec"]), (nrm, st["vel_si_norm"]), (hidden, st["hidden"]), (vel, st["vel"])):
            assert got.shape == ref.shape
            assert float((got - ref).abs().max()) <= 2e-5 * float(ref.abs().max()) + 1e-7


def test_oracle_fp64_budget(state_dict):
    """fp32 vs fp64 evaluation of the same graph: the rounding floor the 1e-4 tolerance sits above."""
    st = _load("forward_ragged.pt")["steps"][0]
    v32, n32, h32 = cns_oracle.forward(state_dict, st["fields"], None, dtype=torch.float32)
    v64, n64, h64 = cns_oracle.forward(state_dict, st["fields"], None, dtype=torch.float64)
    for a, b in ((v32, v64), (n32, n64), (h32, h64)):
        assert float((a.double() - b).abs().max() / b.abs().max()) < 2e-5
