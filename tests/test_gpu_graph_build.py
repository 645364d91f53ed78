"""GPU (-m gpu): the device graph builder - similarity, median preference, Affinity Propagation,
exemplar refinement, member order, median centres - against scikit-learn / numpy on the host and against the
fixtures recorded from the UNMODIFIED reference GraphGenerator (tests/golden/graphgen_*.pt).  Integer outputs
(labels, centres, edge lists, masks) must be bit-exact."""
import glob
import os

import numpy as np
import pytest
import torch

from cns_b200 import GraphGenerator, _lib, synth
from cns_b200.cluster import build_graphs_device

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _host(points):
    old = GraphGenerator.CLUSTER_BACKEND
    GraphGenerator.CLUSTER_BACKEND = "host"
    try:
        return GraphGenerator(points)
    finally:
        GraphGenerator.CLUSTER_BACKEND = old


def _same_graph(a, b):
    for name in ("yhat", "cluster_ids", "cluster_points_count", "cluster_centers_index", "belong_cluster_index",
                 "l0_to_l1_edge_index", "l1_dense_edge_index", "node2j_index", "to_center"):
        assert np.array_equal(np.asarray(getattr(a, name)), np.asarray(getattr(b, name))), name


def test_default_backend_is_the_device_builder():
    assert GraphGenerator.backend() == "device"
    pts = synth.project_norm(synth.sample_target_pose(np.random.default_rng(1)), synth.sample_points(np.random.default_rng(1), 200))
    np.random.seed(5)
    g = GraphGenerator(pts)
    assert g.cluster_method == "affinity-propagation-device"


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "graphgen_*.pt"))))
def test_device_builder_reproduces_reference_fixtures(path):
    """Every fixture recorded from the reference's cns/midend/graph_gen.py (real scikit-learn, seeded like
    demo_sim_Erender.py:42), rebuilt by the device path: clustering, centres, edge lists and per-frame masks."""
    case = torch.load(path, weights_only=False)
    np.random.seed(case["seed"])
    g = GraphGenerator(case["points"])
    assert np.array_equal(g.yhat, case["yhat"]) and np.array_equal(g.cluster_ids, case["cluster_ids"])
    assert np.array_equal(g.cluster_points_count, case["counts"])
    assert np.array_equal(g.cluster_centers_index, case["centers"])
    assert np.array_equal(g.belong_cluster_index, case["belong"])
    assert np.array_equal(g.l0_to_l1_edge_index, case["l0_to_l1"]) and np.array_equal(g.l1_dense_edge_index, case["l1_dense"])
    assert np.array_equal(g.node2j_index, case["node2j"]) and np.array_equal(g.to_center, case["to_center"])
    for fr in case["frames"]:
        _, l1, l0l1, nmask, cmask = g._get_graph(fr["missing"])
        assert np.array_equal(l1, fr["l1"]) and np.array_equal(l0l1, fr["l0l1"])
        assert np.array_equal(nmask, fr["node_mask"]) and np.array_equal(cmask, fr["cluster_mask"])


def test_similarity_and_preference_match_sklearn():
    """cns_ap_similarity implements, bit for bit, the documented rounding of sklearn's euclidean_distances:
    -max(fl(fl(-2 * fma(x1, y1, fl(x0 * y0)) + |x|^2) + |y|^2), 0) with |x|^2 = fl(fl(x0^2) + fl(x1^2)) - checked with exact
    rational arithmetic; the preference is np.median of that matrix.  scikit-learn on this host may route the K = 2 dot
    product through a BLAS kernel that rounds differently (no FMA / split accumulators): that difference is bounded here
    (a few ulp of the operands) and, as the other tests show, does not change a single label."""
    from fractions import Fraction
    from sklearn.metrics import euclidean_distances
    rng = np.random.default_rng(3)
    sets = [synth.project_norm(synth.sample_target_pose(rng), synth.sample_points(rng, n)) for n in (17, 100, 333)]
    np.random.seed(0)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        layouts, S, pref = build_graphs_device(sets, return_similarity=True)
    S = S.cpu().numpy()
    o = 0
    for k, X in enumerate(sets):
        n = len(X)
        got = S[o:o + n * n].reshape(n, n)
        o += n * n
        assert float(pref[k]) == float(np.median(got))
        ref = -euclidean_distances(X, squared=True)
        assert np.abs(got - ref).max() <= 8 * np.finfo(np.float64).eps * np.abs(X).max() ** 2
        if n > 20:
            continue
        xx = X[:, 0] * X[:, 0] + X[:, 1] * X[:, 1]
        for i in range(n):
            for j in range(n):
                dot = float(Fraction(X[i, 1]) * Fraction(X[j, 1]) + Fraction(X[i, 0] * X[j, 0]))       # one FMA
                d = 0.0 if i == j else max((-2.0 * dot + xx[i]) + xx[j], 0.0)
                assert got[i, j] == -d, (i, j)


def test_segment_median_is_numpy_median():
    lib = _lib.load()
    rng = np.random.default_rng(0)
    segs = [rng.normal(size=n) for n in (1, 2, 3, 10, 11, 1000, 4097)]
    segs.append(np.round(rng.normal(size=501), 1))                 # many ties
    segs.append(-np.abs(rng.normal(size=64)) * 1e-300)             # denormal neighbourhood, all negative
    segs.append(np.array([0.0, -0.0, 0.0, -0.0, 1.0, -1.0]))
    off = np.concatenate([[0], np.cumsum([len(s) for s in segs])]).astype(np.int64)
    v = torch.from_numpy(np.concatenate(segs)).cuda()
    t_off = torch.from_numpy(off).cuda()
    out = torch.empty(len(segs), dtype=torch.float64, device="cuda")
    ws = torch.empty(lib.cns_segment_median_workspace_bytes(len(segs)), dtype=torch.uint8, device="cuda")
    _lib.check(lib.cns_segment_median(_lib.ptr(v), _lib.ptr(t_off), len(segs), int(max(len(s) for s in segs)), _lib.ptr(out),
                                      _lib.ptr(ws), ws.numel(), _lib.stream_ptr(torch.device("cuda:0"))), "median")
    got = out.cpu().numpy()
    for k, s in enumerate(segs):
        assert got[k] == np.median(s), (k, got[k], np.median(s))


def test_batched_build_equals_sequential_host_construction():
    """build_many (one batch of device calls) == constructing host generators one after the other from the same
    numpy RandomState, and both leave that stream in the same state."""
    rng = np.random.default_rng(9)
    sets = [synth.project_norm(synth.sample_target_pose(rng), synth.sample_points(rng, n)) for n in (20, 64, 150, 512, 9, 300)]
    np.random.seed(77)
    host = [_host(p) for p in sets]
    state_host = np.random.get_state()[1].copy()
    np.random.seed(77)
    dev = GraphGenerator.build_many(sets)
    assert np.array_equal(np.random.get_state()[1], state_host)
    for a, b in zip(host, dev):
        _same_graph(a, b)
    # and one by one through the default constructor
    np.random.seed(77)
    for p, a in zip(sets, host):
        _same_graph(a, GraphGenerator(p))


def test_kmeans_fallback_is_reported_and_identical():
    """Two tight blobs: Affinity Propagation returns < 4 clusters, the reference falls back to KMeans(4)
    (graph_gen.py:109-112).  The device path does the same on the host and says so."""
    rng = np.random.default_rng(2)
    pts = np.concatenate([rng.normal(size=(15, 2)) * 1e-3, rng.normal(size=(15, 2)) * 1e-3 + 5.0])
    np.random.seed(3)
    a = _host(pts)
    np.random.seed(3)
    with pytest.warns(UserWarning, match="KMeans"):
        b = GraphGenerator(pts)
    assert b.cluster_method == "kmeans4-host" and b.num_clusters == 4
    _same_graph(a, b)
