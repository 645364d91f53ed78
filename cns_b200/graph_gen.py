"""Two-level correspondence graph: target keypoints -> clusters -> dense cluster graph.

Host-side mirror of the reference's `cns/midend/graph_gen.py` (same class name,
constructor, `get_data` signature, attribute names and - given the same numpy RNG
state - bit-identical outputs), written as array programs instead of per-cluster
python loops:

* clustering (graph_gen.py:104-113): sklearn AffinityPropagation(damping=0.9) with a
  KMeans(4) fallback when fewer than 4 clusters come back; skipped for <=16 points
  where every point is its own cluster (graph_gen.py:137-144).  Runs once per target
  image on the host (sklearn is the reference's own dependency, so labels are
  identical by construction);
* centre pick (graph_gen.py:156-164): member closest to the per-cluster coordinate
  median, first minimum wins;
* edge lists (graph_gen.py:7-21,166-182): `l0_to_l1` in cluster-major / ascending
  node order, `l1_dense` = every ordered cluster pair incl. self loops, source-major;
* per-frame masking (graph_gen.py:24-65,199-228): order-preserving boolean filters.

The dense intra-cluster keypoint edges (`l0_dense_edge_index`) of the reference are
never packed into GraphData nor read by the model (graph_gen.py:257-275), so they are
not materialised here (`l0_dense_edge_index` is built lazily on request only).
"""
from typing import Optional

import numpy as np
import torch

from .data import GraphData


def gen_dense_connect_edges(node_indices) -> np.ndarray:
    """All ordered pairs over `node_indices`, source-major: (2, n*n) int64."""
    idx = np.asarray(node_indices, dtype=np.int64)
    n = idx.shape[0]
    grid = np.broadcast_to(idx, (n, n))
    return np.stack([grid.T.reshape(-1), grid.reshape(-1)], axis=0)


def subgraph(edge_index, remove_node_indices, on="j", num_nodes=None):
    """Drop edges touching `remove_node_indices` on side `on` (j, i or both).

    Returns (edge_index[:, keep], node_mask, keep) like reference graph_gen.py:24-65.
    """
    on = on.strip().lower()
    assert on in ("j", "i", "both")
    num_nodes = 1 if num_nodes is None else num_nodes
    if edge_index.size == 0:
        return edge_index, np.ones(num_nodes, dtype=bool), np.ones(0, dtype=bool)
    span = {"j": edge_index[0], "i": edge_index[1], "both": edge_index}[on]
    num_nodes = max(int(span.max()) + 1, num_nodes)
    node_mask = np.ones(num_nodes, dtype=bool)
    node_mask[remove_node_indices] = False
    if on == "j":
        keep = node_mask[edge_index[0]]
    elif on == "i":
        keep = node_mask[edge_index[1]]
    else:
        keep = node_mask[edge_index[0]] & node_mask[edge_index[1]]
    return edge_index[:, keep], node_mask, keep


def cluster_points(points):
    """Labels per point + (sorted unique ids, member counts).  reference :104-113."""
    from sklearn.cluster import AffinityPropagation, KMeans
    labels = AffinityPropagation(damping=0.9).fit_predict(points)
    ids, counts = np.unique(labels, return_counts=True)
    if ids.shape[0] < 4:
        labels = KMeans(n_clusters=4).fit_predict(points)
        ids, counts = np.unique(labels, return_counts=True)
    return labels, ids, counts


def try_cluster_points(points, max_tries=1):
    err = None
    for attempt in range(max_tries):
        try:
            return cluster_points(points)
        except Exception as e:  # sklearn convergence / degenerate input
            err = e
            print("[ERR ] Error encountered in cluster points:\n%s" % (e,))
            if attempt < max_tries - 1:
                print("[INFO] Left tries: {}/{}".format(max_tries - attempt - 1, max_tries))
    raise err


def _segment_median(values, order_by_label, starts, sizes):
    """Per-segment median of `values` (1-D) for segments given in grouped order."""
    grouped = values[order_by_label]
    seg_id = np.repeat(np.arange(sizes.shape[0]), sizes)
    srt = grouped[np.lexsort((grouped, seg_id))]
    lo = starts + (sizes - 1) // 2
    hi = starts + sizes // 2
    # numpy's median of an even-length vector is mean([a, b]) = (a + b) / 2
    med = np.where(lo == hi, srt[lo], (srt[lo] + srt[hi]) / 2.0)
    return med


def label_nearest_exemplar(points, exemplar_indices, device=None) -> np.ndarray:
    """1-NN keypoint -> exemplar assignment on the GPU (`cns_label_nearest`): the final labelling step
    of sklearn's AffinityPropagation (argmax_k -||x - e_k||^2, first maximum wins), in fp64."""
    from . import _lib
    from .cluster import _resolve_device
    lib = _lib.load()
    device = _resolve_device(device)
    with torch.cuda.device(device):
        pts = torch.as_tensor(np.ascontiguousarray(points, dtype=np.float64), device=device)
        ex = pts[torch.as_tensor(np.asarray(exemplar_indices, dtype=np.int64), device=device)].contiguous()
        xptr = torch.tensor([0, ex.shape[0]], dtype=torch.int64, device=device)
        labels = torch.empty(pts.shape[0], dtype=torch.int32, device=device)
        _lib.check(lib.cns_label_nearest(_lib.ptr(pts), None, pts.shape[0], _lib.ptr(ex), _lib.ptr(xptr),
                                         _lib.ptr(labels), _lib.stream_ptr(device)), "cns_label_nearest")
        return labels.cpu().numpy().astype(np.int64)


def _cuda_present() -> bool:
    try:
        return torch.cuda.is_available()
    except Exception:
        return False


class GraphGenerator(object):
    # Where the clustering and the layout derived from it run:
    #   "device": everything on the GPU (cns_b200/cluster.py: similarity, median preference, Affinity Propagation,
    #             exemplar refinement, member order, median centres) - bit-identical with the host path;
    #   "host"  : scikit-learn + numpy on the CPU exactly like the reference.
    # None (default) = "device" whenever a CUDA device is present.
    CLUSTER_BACKEND = None
    CLUSTER_DEVICE = None         # None = the current CUDA device of the calling process / thread

    @classmethod
    def backend(cls) -> str:
        return cls.CLUSTER_BACKEND or ("device" if _cuda_present() else "host")

    @classmethod
    def build_many(cls, target_point_sets, device=None, noise_generator=None):
        """Generators for many targets with ONE batched device clustering + layout call.  Equal to the reference's
        sequential construction as long as no set needs the KMeans(4) fallback (the fallback draws from numpy's
        global RNG after, not between, the AP noise draws).  `noise_generator`: see cluster.build_graphs_device."""
        from .cluster import build_graphs_device
        sets = [np.asarray(p) for p in target_point_sets]
        big = [k for k, p in enumerate(sets) if len(p) > 16]
        layouts = dict(zip(big, build_graphs_device([sets[k] for k in big], device=device or cls.CLUSTER_DEVICE,
                                                    noise_generator=noise_generator))) if big else {}
        return [cls(p, layout=layouts.get(k)) for k, p in enumerate(sets)]

    @classmethod
    def from_exemplars(cls, target_points, exemplar_indices, device=None):
        """Graph whose clusters are the Voronoi cells of the given exemplar keypoints (labels computed
        on the device)."""
        return cls(target_points, labels=label_nearest_exemplar(target_points, exemplar_indices, device))

    def __init__(self, target_points, labels=None, layout=None):
        target_points = np.asarray(target_points)
        self.num_points = len(target_points)
        self.target_points = target_points
        n = self.num_points
        self.cluster_method = "given" if labels is not None else "one-cluster-per-point"

        if layout is None and labels is None and n > 16 and self.backend() == "device":
            from .cluster import build_graphs_device
            layout = build_graphs_device([target_points], device=self.CLUSTER_DEVICE)[0]
        if layout is not None:
            # everything below was computed on the device (cns_ap_labels / cns_graph_layout)
            self.cluster_method = layout.method
            self.yhat = layout.labels
            self.cluster_ids = np.arange(layout.cluster_size.shape[0], dtype=np.int64)
            self.cluster_points_count = layout.cluster_size
            self._finish(layout.labels.astype(np.int64), layout.member_order, layout.cluster_size, layout.centers)
            return
        if labels is not None:
            self.yhat = np.asarray(labels)
            self.cluster_ids, self.cluster_points_count = np.unique(self.yhat, return_counts=True)
        elif n > 16:
            self.cluster_method = "affinity-propagation-host"
            self.yhat, self.cluster_ids, self.cluster_points_count = try_cluster_points(target_points, max_tries=3)
        else:
            self.yhat = np.arange(n)
            self.cluster_ids = self.yhat.copy()
            self.cluster_points_count = np.ones_like(self.cluster_ids)

        # cluster rank of every point (cluster_ids is sorted-unique)
        rank = np.searchsorted(self.cluster_ids, self.yhat).astype(np.int64)
        self.belong_cluster_index = rank
        order = np.argsort(rank, kind="stable")          # cluster-major, ascending node id
        sizes = np.bincount(rank, minlength=len(self.cluster_ids)).astype(np.int64)
        starts = np.cumsum(sizes) - sizes
        self._member_order, self._member_starts, self._member_sizes = order, starts, sizes

        # centre = member nearest to the coordinate-wise median (first minimum)
        med = np.stack([_segment_median(target_points[:, d], order, starts, sizes)
                        for d in range(target_points.shape[1])], axis=-1)
        grouped = target_points[order]
        seg_id = np.repeat(np.arange(sizes.shape[0]), sizes)
        diff = grouped - med[seg_id]
        dist = np.sqrt(np.add.reduce(diff * diff, axis=-1))
        seg_min = np.minimum.reduceat(dist, starts) if n else dist
        is_min = dist == seg_min[seg_id]
        pos = np.arange(n)
        first = np.minimum.reduceat(np.where(is_min, pos, n), starts) if n else pos
        self._finish(rank, order, sizes, order[first].astype(np.int64))

    def _finish(self, rank, order, sizes, centers):
        """Edge lists and inverse permutations from (cluster rank per point, cluster-major member order, member counts,
        centre per cluster) - integer bookkeeping shared by the host and the device path (graph_gen.py:166-196)."""
        n, target_points = self.num_points, self.target_points
        sizes = np.asarray(sizes, dtype=np.int64)
        starts = np.cumsum(sizes) - sizes
        self.belong_cluster_index = rank
        self._member_order, self._member_starts, self._member_sizes = order, starts, sizes
        seg_id = np.repeat(np.arange(sizes.shape[0]), sizes)
        self.cluster_centers_index = np.asarray(centers, dtype=np.int64)
        self.cluster_points_index = [order[s:s + c] for s, c in zip(starts, sizes)]
        self.l0_to_l1_edge_index = np.stack([order.astype(np.int64), seg_id.astype(np.int64)], axis=0)
        self.l1_dense_edge_index = gen_dense_connect_edges(np.arange(len(self.cluster_ids)))
        self.to_center = target_points[self.cluster_centers_index][rank] - target_points

        self.node2j_index = np.empty(n, dtype=np.int64)
        self.node2j_index[order] = np.arange(n)
        self._l0_dense = None

    @property
    def l0_dense_edge_index(self) -> np.ndarray:
        """Dense intra-cluster keypoint edges; dead in the model, built on demand."""
        if self._l0_dense is None:
            parts = [gen_dense_connect_edges(m) for m in self.cluster_points_index]
            self._l0_dense = np.concatenate(parts, axis=-1) if parts else np.zeros((2, 0), np.int64)
        return self._l0_dense

    @property
    def num_clusters(self) -> int:
        return len(self.cluster_ids)

    def _get_graph(self, missing_node_indices=None):
        n, c = self.num_points, self.num_clusters
        l0l1 = self.l0_to_l1_edge_index
        l1 = self.l1_dense_edge_index
        node_mask = np.ones(n, dtype=bool)
        cluster_mask = np.ones(c, dtype=bool)
        if missing_node_indices is not None and len(missing_node_indices):
            missing = np.asarray(missing_node_indices, dtype=np.int64)
            node_mask[missing] = False
            l0l1 = l0l1[:, node_mask[l0l1[0]]]
            # np.add.at semantics of the reference: duplicates decrement repeatedly and
            # a cluster is dropped only when its count lands exactly on zero
            hits = np.bincount(self.belong_cluster_index[missing], minlength=c)
            cluster_mask = (self.cluster_points_count - hits) != 0
            if not cluster_mask.all():
                l1 = l1[:, cluster_mask[l1[0]] & cluster_mask[l1[1]]]
                l0l1 = l0l1[:, cluster_mask[l0l1[1]]]
        return None, l1, l0l1, node_mask, cluster_mask

    def get_data(self, current_points, missing_node_indices=None, current_features=None,
                 target_features=None, mismatch_mask=None) -> GraphData:
        if current_features is None:
            current_features = self.default_feature(current_points)
        if target_features is None:
            target_features = self.default_feature(self.target_points)
        _, l1_cur, l0l1, node_mask, cluster_mask = self._get_graph(missing_node_indices)
        f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float()
        i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a)).long()
        return GraphData(
            num_nodes=self.num_points,
            num_clusters=torch.tensor(self.num_clusters).long(),
            x_cur=f32(current_features),
            pos_cur=f32(current_points),
            l1_dense_edge_index_cur=i64(l1_cur),
            l0_to_l1_edge_index_j_cur=i64(l0l1[0]),
            l0_to_l1_edge_index_i_cur=i64(l0l1[1]),
            x_tar=f32(target_features),
            pos_tar=f32(self.target_points),
            l1_dense_edge_index_tar=i64(self.l1_dense_edge_index),
            node_mask=torch.from_numpy(node_mask).bool(),
            cluster_mask=torch.from_numpy(cluster_mask).bool(),
            cluster_centers_index=i64(self.cluster_centers_index),
            belong_cluster_index=i64(self.belong_cluster_index),
            # both cluster-level edge lists are complete digraphs (all clusters / alive clusters) in the
            # source-major order of gen_dense_connect_edges: lets the forward skip the CSR build
            l1_complete=True,
        )

    def default_feature(self, points):
        return points
