"""Affinity-Propagation clustering of target keypoints with the message passing on the GPU.

`cluster_points_device` returns exactly what the reference's `cluster_points`
(cns/midend/graph_gen.py:104-113 -> sklearn AffinityPropagation(damping=0.9), KMeans(4) fallback)
returns for the same numpy global RNG state: the similarity matrix, its median preference and the
degeneracy noise are built on the host with the same numpy / scikit-learn calls sklearn makes (so the
random stream is consumed identically); the O(iterations * N^2) responsibility / availability loop runs
in `cns_affinity_propagation` (csrc/affinity.cu), bit-exact in fp64; the exemplar refinement at the end
is again the same array code as sklearn's.
"""
import ctypes as C
from typing import List, Sequence

import numpy as np
import torch

from . import _lib


def _prepare_similarity(X: np.ndarray):
    """S as sklearn's AffinityPropagation.fit + _affinity_propagation build it (incl. noise).  Returns
    (S, degenerate labels | None): for inputs whose similarities are all equal sklearn returns before it draws
    any noise (`_equal_similarities_and_preferences`, _affinity_propagation.py:35-55): every sample its own
    cluster when the preference exceeds the common similarity, one cluster otherwise - restated here, no
    scikit-learn call."""
    from sklearn.metrics import euclidean_distances
    X = np.ascontiguousarray(X, dtype=np.float64)
    n = X.shape[0]
    S = -euclidean_distances(X, squared=True)
    preference = np.median(S)
    mask = np.ones(S.shape, dtype=bool)
    np.fill_diagonal(mask, 0)
    if n == 1 or np.all(S[mask].flat == S[mask].flat[0]):
        labels = np.arange(n) if preference > S.flat[n - 1] else np.zeros(n, dtype=np.int64)
        return S, labels
    S.flat[:: n + 1] = preference
    rs = np.random.mtrand._rand                       # check_random_state(None)
    S += (np.finfo(S.dtype).eps * S + np.finfo(S.dtype).tiny * 100) * rs.standard_normal(size=(n, n))
    return S, None


def _labels_from_exemplars(S: np.ndarray, E: np.ndarray):
    """Tail of sklearn's _affinity_propagation: assign, refine exemplars, relabel gapless."""
    n = S.shape[0]
    I = np.flatnonzero(E)
    K = I.size
    if K == 0:
        return np.array([-1] * n), np.array([], dtype=np.int64)
    c = np.argmax(S[:, I], axis=1)
    c[I] = np.arange(K)
    for k in range(K):
        ii = np.flatnonzero(c == k)
        j = np.argmax(np.sum(S[ii[:, np.newaxis], ii], axis=0))
        I[k] = ii[j]
    c = np.argmax(S[:, I], axis=1)
    c[I] = np.arange(K)
    labels = I[c]
    centers = np.unique(labels)
    return np.searchsorted(centers, labels), centers


def affinity_propagation_device(point_sets: Sequence[np.ndarray], damping=0.9, max_iter=200, convergence_iter=15,
                                device="cuda:0"):
    """Labels (one int array per point set) identical to
    `AffinityPropagation(damping=damping).fit_predict(X)` called on the sets in order."""
    lib = _lib.load()
    dev = torch.device(device)
    mats, out, todo = [], [None] * len(point_sets), []
    for k, X in enumerate(point_sets):
        S, degenerate = _prepare_similarity(X)
        if degenerate is not None:
            out[k] = degenerate
        else:
            todo.append(k)
            mats.append(S)
    if not todo:
        return out
    ns = np.array([m.shape[0] for m in mats], dtype=np.int64)
    pt_off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
    mat_off = np.concatenate([[0], np.cumsum(ns * ns)]).astype(np.int64)
    row_graph = np.repeat(np.arange(len(mats), dtype=np.int32), ns)
    S_all = torch.from_numpy(np.concatenate([m.reshape(-1) for m in mats])).to(dev)
    t_pt, t_mat = torch.from_numpy(pt_off).to(dev), torch.from_numpy(mat_off).to(dev)
    t_rg = torch.from_numpy(row_graph).to(dev)
    G, NR, NM = len(mats), int(pt_off[-1]), int(mat_off[-1])
    E = torch.empty(NR, dtype=torch.uint8, device=dev)
    n_iter = torch.zeros(G, dtype=torch.int32, device=dev)
    conv = (C.c_int32 * G)()
    ws = torch.empty(lib.cns_ap_workspace_bytes(NR, NM, G, convergence_iter), dtype=torch.uint8, device=dev)
    _lib.check(lib.cns_affinity_propagation(
        _lib.ptr(S_all), _lib.ptr(t_pt), _lib.ptr(t_mat), _lib.ptr(t_rg), G, NR, NM, float(damping), int(max_iter),
        int(convergence_iter), _lib.ptr(E), _lib.ptr(n_iter), conv, _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)),
        "cns_affinity_propagation")
    E_host = E.cpu().numpy().astype(bool)
    for g, k in enumerate(todo):
        labels, _ = _labels_from_exemplars(mats[g], E_host[pt_off[g]:pt_off[g + 1]])
        out[k] = labels
    affinity_propagation_device.last_n_iter = n_iter.cpu().numpy()
    affinity_propagation_device.last_status = np.array(list(conv))
    return out


def cluster_points_device(points, device="cuda:0"):
    """Drop-in for `cluster_points` (graph_gen.py:104-113) with the AP iterations on the GPU."""
    yhat = affinity_propagation_device([points], device=device)[0]
    ids, counts = np.unique(yhat, return_counts=True)
    if ids.shape[0] < 4:
        from sklearn.cluster import KMeans
        yhat = KMeans(n_clusters=4).fit_predict(points)
        ids, counts = np.unique(yhat, return_counts=True)
    return yhat, ids, counts
