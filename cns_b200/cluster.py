"""Affinity-Propagation clustering of target keypoints and the graph layout derived from it, on the GPU.

`build_graphs_device` runs the whole of `cluster_points` + `GraphGenerator.__init__`
(cns/midend/graph_gen.py:104-196 -> sklearn AffinityPropagation(damping=0.9)) for a batch of target point sets on
the device and returns exactly what the reference computes for the same numpy global RNG state:

    similarity matrix (cns_ap_similarity) -> median preference (cns_segment_median, radix select) -> preference
    diagonal + degeneracy noise (cns_ap_prepare) -> responsibility / availability iterations
    (cns_affinity_propagation, csrc/affinity.cu) -> exemplar refinement + labels (cns_ap_labels) -> member order,
    member counts, median centres (cns_graph_layout)

all in fp64 with numpy's / scikit-learn's operation order, so labels, centres and edge lists are bit-identical.
Only the N x N standard-normal noise is drawn on the host - it has to consume numpy's global RandomState exactly
like sklearn does (`check_random_state(None)`) - and uploaded.  Fewer than 4 clusters falls back to KMeans(4) like
the reference (graph_gen.py:109-112); that fallback runs in scikit-learn on the host and is REPORTED (a warning and
`GraphGenerator.cluster_method == "kmeans4-host"`), never silent.

`affinity_propagation_device` / `cluster_points_device` are the older entry points (host-prepared similarity
matrix, device iterations) kept for A-B comparison.
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
                                device=None):
    with torch.cuda.device(_resolve_device(device)):
        return _affinity_propagation_device(point_sets, damping, max_iter, convergence_iter, _resolve_device(device))


def _affinity_propagation_device(point_sets, damping, max_iter, convergence_iter, device):
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


def cluster_points_device(points, device=None):
    """Drop-in for `cluster_points` (graph_gen.py:104-113) with the AP iterations on the GPU."""
    yhat = affinity_propagation_device([points], device=device)[0]
    ids, counts = np.unique(yhat, return_counts=True)
    if ids.shape[0] < 4:
        from sklearn.cluster import KMeans
        yhat = KMeans(n_clusters=4).fit_predict(points)
        ids, counts = np.unique(yhat, return_counts=True)
    return yhat, ids, counts


# ----------------------------------------------------------------------------------------------- full device path
class DeviceLayout(object):
    """What `GraphGenerator.__init__` derives from one target point set (host numpy arrays)."""
    __slots__ = ("labels", "cluster_size", "centers", "member_order", "n_iter", "converged", "method")


def _resolve_device(device):
    """None -> the calling thread's current CUDA device (one process per GPU: every rank clusters on its own)."""
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    dev = torch.device(device)
    return dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())


def build_graphs_device(point_sets: Sequence[np.ndarray], damping=0.9, max_iter=200, convergence_iter=15,
                        device=None, return_similarity=False, noise_generator=None) -> List[DeviceLayout]:
    """`noise_generator` (a CUDA torch.Generator): draw Affinity Propagation's tie-breaking noise on the device instead of
    from numpy's global RandomState.  Same distribution, not scikit-learn's numbers - for callers that generate random
    training scenes (DeviceSceneBatch), where 17 M host-drawn normals per 256 scenes were the largest part of a re-init."""
    # the library launches on the CURRENT device: make `device` current for the duration of the call
    with torch.cuda.device(_resolve_device(device)):
        return _build_graphs_device(point_sets, damping, max_iter, convergence_iter, _resolve_device(device), return_similarity,
                                    noise_generator)


def _build_graphs_device(point_sets, damping, max_iter, convergence_iter, device, return_similarity, noise_generator=None):
    """Clustering + layout of every point set (each with more than 16 points: smaller sets never reach the
    clustering in the reference, graph_gen.py:137-144) in ONE batch of device calls.  Consumes numpy's global
    RandomState like the sequential host construction (one (n, n) standard-normal draw per set, in order), provided
    no set needs the KMeans(4) fallback (which draws after, not between, the noise draws)."""
    import warnings
    lib = _lib.load()
    dev = torch.device(device)
    sets = [np.ascontiguousarray(p, dtype=np.float64) for p in point_sets]
    G = len(sets)
    if G == 0:
        return []
    ns = np.array([p.shape[0] for p in sets], dtype=np.int64)
    pt_off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
    mat_off = np.concatenate([[0], np.cumsum(ns * ns)]).astype(np.int64)
    NR, NM = int(pt_off[-1]), int(mat_off[-1])
    L = _lib
    stream = L.stream_ptr(dev)
    pts = torch.from_numpy(np.concatenate(sets, axis=0)).to(dev)
    t_pt, t_mat = torch.from_numpy(pt_off).to(dev), torch.from_numpy(mat_off).to(dev)
    t_rg = torch.from_numpy(np.repeat(np.arange(G, dtype=np.int32), ns)).to(dev)
    S = torch.empty(NM, dtype=torch.float64, device=dev)
    L.check(lib.cns_ap_similarity(L.ptr(pts), L.ptr(t_pt), L.ptr(t_mat), L.ptr(t_rg), NR, L.ptr(S), stream), "cns_ap_similarity")
    pref = torch.empty(G, dtype=torch.float64, device=dev)
    ws = torch.empty(lib.cns_segment_median_workspace_bytes(G), dtype=torch.uint8, device=dev)
    L.check(lib.cns_segment_median(L.ptr(S), L.ptr(t_mat), G, int((ns * ns).max()), L.ptr(pref), L.ptr(ws), ws.numel(), stream),
            "cns_segment_median")
    # sklearn returns before drawing any noise when all similarities are equal (degenerate input): those sets are
    # labelled by its rule below and get no noise
    # (in the plane that means: all points coincide, i.e. min(S) == 0)
    smin = torch.segment_reduce(S, "min", lengths=torch.from_numpy(ns * ns).to(dev)).cpu().numpy()
    degenerate = [bool(ns[g] == 1 or smin[g] == 0.0) for g in range(G)]
    if noise_generator is not None:
        noise_d = torch.randn(NM, dtype=torch.float64, device=dev, generator=noise_generator)
        for g in range(G):
            if degenerate[g]:
                noise_d[int(mat_off[g]):int(mat_off[g + 1])] = 0.0
    else:
        rs = np.random.mtrand._rand                                          # check_random_state(None)
        noise = torch.empty(NM, dtype=torch.float64).pin_memory() if NM < (1 << 28) else torch.empty(NM, dtype=torch.float64)
        nz = noise.numpy()
        for g in range(G):
            n = int(ns[g])
            if degenerate[g]:
                nz[int(mat_off[g]):int(mat_off[g + 1])] = 0.0
            else:
                nz[int(mat_off[g]):int(mat_off[g + 1])] = rs.standard_normal(size=(n, n)).reshape(-1)
        noise_d = noise.to(dev, non_blocking=True)
    S_raw = S.clone() if return_similarity else None
    L.check(lib.cns_ap_prepare(L.ptr(S), L.ptr(noise_d), L.ptr(pref), L.ptr(t_pt), L.ptr(t_mat), L.ptr(t_rg), NR, stream),
            "cns_ap_prepare")
    E = torch.empty(NR, dtype=torch.uint8, device=dev)
    n_iter = torch.zeros(G, dtype=torch.int32, device=dev)
    conv = (C.c_int32 * G)()
    ws = torch.empty(lib.cns_ap_workspace_bytes(NR, NM, G, convergence_iter), dtype=torch.uint8, device=dev)
    L.check(lib.cns_affinity_propagation(L.ptr(S), L.ptr(t_pt), L.ptr(t_mat), L.ptr(t_rg), G, NR, NM, float(damping),
                                         int(max_iter), int(convergence_iter), L.ptr(E), L.ptr(n_iter), conv, L.ptr(ws),
                                         ws.numel(), stream), "cns_affinity_propagation")
    del ws
    labels = torch.empty(NR, dtype=torch.int32, device=dev)
    exemplars = torch.empty(NR, dtype=torch.int32, device=dev)
    n_clu = torch.empty(G, dtype=torch.int32, device=dev)
    ws2 = torch.empty(2 * ((NR * 4 + 255) // 256 * 256) + 256, dtype=torch.uint8, device=dev)
    L.check(lib.cns_ap_labels(L.ptr(S), L.ptr(E), L.ptr(t_pt), L.ptr(t_mat), G, NR, L.ptr(labels), L.ptr(exemplars),
                              L.ptr(n_clu), L.ptr(ws2), ws2.numel(), stream), "cns_ap_labels")
    k_host = n_clu.cpu().numpy().astype(np.int64)
    labels_h = labels.cpu().numpy().astype(np.int64)
    n_iter_h = n_iter.cpu().numpy()
    # host-side exceptions of the reference: degenerate inputs and the KMeans(4) fallback for < 4 clusters
    method = ["affinity-propagation-device"] * G
    redo = False
    for g in range(G):
        sl = slice(int(pt_off[g]), int(pt_off[g + 1]))
        if degenerate[g]:
            pf = float(pref[g])
            labels_h[sl] = np.arange(ns[g]) if pf > 0.0 else 0           # preference vs the common similarity (0)
            k_host[g] = ns[g] if pf > 0.0 else 1
            method[g] = "degenerate-rule"
            redo = True
        if k_host[g] < 4:
            from sklearn.cluster import KMeans
            warnings.warn("Affinity Propagation found %d (< 4) clusters for a %d-point target: KMeans(4) fallback on the "
                          "host (scikit-learn), as cns/midend/graph_gen.py:109-112" % (int(k_host[g]), int(ns[g])))
            lab = KMeans(n_clusters=4).fit_predict(sets[g])
            ids = np.unique(lab)
            labels_h[sl] = np.searchsorted(ids, lab)
            k_host[g] = ids.shape[0]
            method[g] = "kmeans4-host"
            redo = True
    if redo:
        labels = torch.from_numpy(labels_h.astype(np.int32)).to(dev)
    clu_off = np.concatenate([[0], np.cumsum(k_host)]).astype(np.int64)
    t_clu = torch.from_numpy(clu_off).to(dev)
    order = torch.empty(NR, dtype=torch.int32, device=dev)
    csize = torch.empty(int(clu_off[-1]), dtype=torch.int32, device=dev)
    centers = torch.empty(int(clu_off[-1]), dtype=torch.int32, device=dev)
    L.check(lib.cns_graph_layout(L.ptr(pts), L.ptr(labels), L.ptr(t_pt), L.ptr(t_clu), G, L.ptr(order), L.ptr(csize),
                                 L.ptr(centers), stream), "cns_graph_layout")
    order_h, csize_h, centers_h = order.cpu().numpy(), csize.cpu().numpy(), centers.cpu().numpy()
    out = []
    for g in range(G):
        d = DeviceLayout()
        sl, cl = slice(int(pt_off[g]), int(pt_off[g + 1])), slice(int(clu_off[g]), int(clu_off[g + 1]))
        d.labels = labels_h[sl].copy()
        d.cluster_size = csize_h[cl].astype(np.int64)
        d.centers = centers_h[cl].astype(np.int64)
        d.member_order = order_h[sl].astype(np.int64)
        d.n_iter, d.converged, d.method = int(n_iter_h[g]), int(conv[g]) == 1, method[g]
        out.append(d)
    if return_similarity:
        return out, S_raw, pref
    return out
