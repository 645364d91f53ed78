"""Randomised soak of the device graph construction against the host (scikit-learn) construction the reference uses:
for many random target point sets (17..600 keypoints, projected scenes, uniform clouds, near-duplicate points, points
on a line, tight blobs) the GraphGenerator built on the device must equal the one built on the host bit for bit -
labels, cluster sizes, centres, member order, dense cluster edges.  Exits 1 on the first difference.
usage: python scripts/fuzz_clustering.py [n_sets]"""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cns_b200 import GraphGenerator, synth

FIELDS = ("belong_cluster_index", "cluster_points_count", "cluster_centers_index", "l0_to_l1_edge_index",
          "l1_dense_edge_index")


def point_set(rng, kind, n):
    if kind == "scene":
        return synth.project_norm(synth.sample_target_pose(rng), synth.sample_points(rng, n))
    if kind == "uniform":
        return rng.uniform(-0.6, 0.6, size=(n, 2))
    if kind == "blobs":
        c = rng.uniform(-0.5, 0.5, size=(int(rng.integers(2, 9)), 2))
        return c[rng.integers(0, len(c), n)] + rng.normal(0, 0.01, size=(n, 2))
    if kind == "line":
        t = rng.uniform(-0.5, 0.5, size=n)
        return np.stack([t, 0.3 * t + rng.normal(0, 1e-4, size=n)], 1)
    if kind == "near_duplicates":
        base = rng.uniform(-0.5, 0.5, size=(max(4, n // 3), 2))
        return base[rng.integers(0, len(base), n)] + rng.normal(0, 1e-7, size=(n, 2))
    raise ValueError(kind)


def build(points, backend, seed):
    GraphGenerator.CLUSTER_BACKEND = backend
    np.random.seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return GraphGenerator(points)


def run(n_sets=120, verbose=True):
    rng = np.random.default_rng(77)
    kinds = ["scene", "scene", "uniform", "blobs", "line", "near_duplicates"]
    bad = 0
    methods = {}
    for k in range(n_sets):
        kind = kinds[k % len(kinds)]
        n = int(rng.choice([17, 18, 31, 32, 33, 64, 100, 128, 200, 300, 512, 600]))
        pts = point_set(rng, kind, n)
        try:
            host = build(pts, "host", 1000 + k)
        except Exception as e:          # scikit-learn itself fails on some degenerate sets: the device path must fail too
            try:
                build(pts, "device", 1000 + k)
                print("set %d (%s, n=%d): host raised %r, device did not" % (k, kind, n, e)); bad += 1
            except Exception:
                pass
            continue
        dev = build(pts, "device", 1000 + k)
        methods[dev.cluster_method] = methods.get(dev.cluster_method, 0) + 1
        for f in FIELDS:
            a, b = np.asarray(getattr(host, f)), np.asarray(getattr(dev, f))
            if a.shape != b.shape or not np.array_equal(a, b):
                print("set %d (%s, n=%d): %s differs (host %d clusters via %s, device %d via %s)" %
                      (k, kind, n, f, host.num_clusters, host.cluster_method, dev.num_clusters, dev.cluster_method))
                bad += 1
                break
    GraphGenerator.CLUSTER_BACKEND = None
    if verbose:
        print("%d sets, %d differences; device cluster methods: %s" % (n_sets, bad, methods))
    return bad


if __name__ == "__main__":
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 120) else 0)
