"""Device-resident batched servo loop: the target-side graph of every scene lives on the GPU, a frame
only ships what changes per frame.

The reference builds the two-level graph once per target image and, every frame, only masks out the
keypoints that were not matched (`Midend.get_graph_data`, cns/midend/corr2graph.py:11-22 ->
`GraphGenerator.get_data`, graph_gen.py:236-277) before collating B python objects with
`Batch.from_data_list` (sim/dataset.py:206-209).  `BatchedServo` keeps that split for B scenes at once
(SURVEY 8f2): constructor = "target changed" (static tensors uploaded once: target keypoints, cluster
membership, centres); `step(cur_xy, visible)` = one frame: H2D of the current keypoint coordinates +
visibility mask, `cns_frame_mask` (per-frame keypoint->cluster edges, cluster mask and cluster segment
pointers, bit-exact with `_get_graph` + collation), `cns_graphvs_forward` with the hidden state carried on the device, D2H of the
(B,6) velocities.  The cluster-level graphs are complete digraphs (all clusters / alive clusters), so
they are passed as a flag instead of edge lists (CNS_BATCH_L1_*_COMPLETE).
"""
import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import _lib
from .graph_gen import GraphGenerator


class BatchedServo(object):
    def __init__(self, net, generators: Sequence[GraphGenerator], dist_scale, device="cuda:0", depth: int = 2):
        self.net, self.dev = net, torch.device(device)
        self.lib = _lib.load()
        gens = list(generators)
        n_nodes = np.array([g.num_points for g in gens], dtype=np.int64)
        n_clu = np.array([g.num_clusters for g in gens], dtype=np.int64)
        noff, coff = np.cumsum(n_nodes) - n_nodes, np.cumsum(n_clu) - n_clu
        self.B, self.N, self.Cn = len(gens), int(n_nodes.sum()), int(n_clu.sum())
        self.node_off = noff
        self.max_clusters = int(n_clu.max())
        d = lambda a, dt: torch.as_tensor(np.array(a), dtype=dt).to(self.dev)
        tar = np.concatenate([g.target_points for g in gens]).astype(np.float32)
        self.x_tar = d(tar, torch.float32)                                   # x_tar == pos_tar (default_feature)
        self.belong = d(np.concatenate([g.belong_cluster_index + c for g, c in zip(gens, coff)]), torch.int64)
        self.member_order = d(np.concatenate([g.l0_to_l1_edge_index[0] + o for g, o in zip(gens, noff)]), torch.int64)
        csize = np.concatenate([g.cluster_points_count for g in gens]).astype(np.int64)
        self.member_ptr = d(np.concatenate([[0], np.cumsum(csize)]), torch.int32)       # members of cluster c: member_order[ptr[c]:ptr[c+1]]
        self.centers = d(np.concatenate([g.cluster_centers_index + o for g, o in zip(gens, noff)]), torch.int64)
        self.member_cluster = d(np.repeat(np.arange(int(n_clu.sum())), csize), torch.int64)   # cluster of member_order[k]
        self.num_clusters = d(n_clu, torch.int64)
        self.tpo = d(np.broadcast_to(np.asarray(dist_scale, dtype=np.float32), (self.B,)), torch.float32)
        i64 = lambda n: torch.empty(n, dtype=torch.int64, device=self.dev)
        # frames in flight (submit / result): every per-frame buffer exists `depth` times and every slot has its own
        # compute stream, so the kernels of consecutive frames may overlap on the device (the cluster-level backbone of
        # frame t leaves most SMs idle in its second wave; the encoder of frame t+1 fills them).  A carried hidden
        # state orders frame t+1 behind frame t through an event.
        self.depth = max(1, int(depth))
        D = range(self.depth)
        self.e01_j, self.e01_i = [i64(max(self.N, 1)) for _ in D], [i64(max(self.N, 1)) for _ in D]
        self.cmask = [torch.empty(max(self.Cn, 1), dtype=torch.uint8, device=self.dev) for _ in D]
        self.clu_rowptr = [torch.empty(self.Cn + 1, dtype=torch.int32, device=self.dev) for _ in D]
        self.mask_ws = [torch.empty((self.Cn + 1) * 4, dtype=torch.uint8, device=self.dev) for _ in D]
        self.xy_dev = [torch.empty(self.N, 2, dtype=torch.float32, device=self.dev) for _ in D]
        self.vis_dev = [torch.empty(self.N, dtype=torch.uint8, device=self.dev) for _ in D]
        self.vel_host = [torch.empty(self.B, 6).pin_memory() for _ in D]
        self.out = [torch.empty(self.B * 13, dtype=torch.float32, device=self.dev) for _ in D]
        self.copied = [torch.cuda.Event() for _ in D]
        self.done = [None] * self.depth
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.streams = [torch.cuda.Stream(self.dev) for _ in D]
        self.ticket = 0
        self.tcache, self.tcache_handle = None, None       # cns_target_cache: target-side results kept across frames
        self.hidden = [torch.zeros(self.Cn, 128, device=self.dev), torch.zeros(self.Cn, 128, device=self.dev)]
        self.cur, self.have_hidden = 0, False
        self.hidden_ready = None                           # event: the newest hidden state is complete
        self.done_ev = [torch.cuda.Event() for _ in D]
        self._slot_cache = [None] * self.depth
        self._p_member_order, self._p_member_ptr = _lib.ptr(self.member_order), _lib.ptr(self.member_ptr)
        self.h2d_bytes = self.N * 2 * 4 + self.N
        self.d2h_bytes = self.B * 6 * 4

    def reset(self):
        self.have_hidden = False

    def _target_cache(self, h):
        """Everything the forward derives from the target side alone (per-cluster attention term, cluster
        positions, graph pointers, backbone tiles) is computed once per target, like the reference's Midend
        builds its GraphGenerator once per target image."""
        if self.tcache is not None and self.tcache_handle is h:
            return self.tcache
        self.drop_target_cache()
        L = _lib
        b = L.cns_batch()
        b.n_nodes, b.n_clusters, b.n_graphs = self.N, self.Cn, self.B
        b.x_tar, b.pos_tar = self.x_tar.data_ptr(), self.x_tar.data_ptr()
        b.cluster_centers, b.num_clusters = self.centers.data_ptr(), self.num_clusters.data_ptr()
        b.max_clusters_per_graph = self.max_clusters
        # every target keypoint with its cluster (cluster-major): lets the cache keep the target side's per-keypoint
        # attention logits / values, so that a frame only re-runs their softmax over the visible keypoints
        b.n_e01 = self.N
        b.l0_to_l1_j, b.l0_to_l1_i = self.member_order.data_ptr(), self.member_cluster.data_ptr()
        out = C.c_void_p()
        L.check(self.lib.cns_target_cache_create(h.raw, C.byref(b), L.stream_ptr(self.dev), C.byref(out)),
                "cns_target_cache_create")
        self.tcache, self.tcache_handle = out, h
        return out

    def drop_target_cache(self):
        """Call after loading new weights into the policy (the cache is tied to the weights it was built with)."""
        if self.tcache is not None:
            self.lib.cns_target_cache_destroy(self.tcache)
        self.tcache, self.tcache_handle = None, None

    def __del__(self):
        try:
            self.drop_target_cache()
        except Exception:
            pass

    def step(self, cur_xy: torch.Tensor, visible: torch.Tensor) -> torch.Tensor:
        """One closed-loop frame: `result(submit(cur_xy, visible))`."""
        return self.result(self.submit(cur_xy, visible))

    def submit(self, cur_xy: torch.Tensor, visible: torch.Tensor) -> int:
        """Enqueue one frame; returns a ticket for `result`.  cur_xy (sum N, 2) float32 and visible (sum N,)
        uint8/bool are HOST tensors (pinned for full speed) in the scenes' keypoint order.  Up to `depth`
        frames may be in flight: the host->device copy of frame t+1 (copy stream) overlaps the kernels of
        frame t, and so do its kernels when no hidden state is carried (every slot has its own compute stream;
        what the caller enqueued on the current stream before the call is waited for).

        The current-side cluster graph is never materialised: it is the complete digraph over the clusters
        that keep a visible keypoint (graph_gen.py:199-228), which the forward handles from cluster_mask
        alone (CNS_BATCH_L1_CUR_COMPLETE), and the number of keypoint->cluster edges is the number of visible
        keypoints - so nothing is read back from the device before the velocities."""
        lib, L, dev = self.lib, _lib, self.dev
        k = self.ticket % self.depth
        if self.done[k] is not None:
            self.done[k].synchronize()                 # the frame that used this slot has delivered its result
        st = self.streams[k]
        st.wait_stream(torch.cuda.current_stream(dev))  # whatever the caller enqueued (new weights, ...) comes first
        stream = st.cuda_stream
        vis = visible.view(torch.uint8) if visible.dtype == torch.bool else visible
        n01 = int(np.count_nonzero(vis.numpy())) if not vis.is_cuda else int(torch.count_nonzero(vis))
        with torch.cuda.stream(self.copy_stream):
            self.xy_dev[k].copy_(cur_xy, non_blocking=True)
            self.vis_dev[k].copy_(vis, non_blocking=True)
            self.copied[k].record(self.copy_stream)
        st.wait_event(self.copied[k])
        if self.have_hidden and self.hidden_ready is not None:
            st.wait_event(self.hidden_ready)
        h = self.net._handle(dev)
        b, ptrs = self._slot_batch(k)
        b.n_e01 = n01
        b.target_cache = self._target_cache(h)
        L.check(lib.cns_frame_mask(self._p_member_order, self._p_member_ptr, ptrs["vis"], self.N, self.Cn,
                                   ptrs["e01_j"], ptrs["e01_i"], ptrs["cmask"], ptrs["clu_rowptr"],
                                   ptrs["mask_ws"], self.mask_ws[k].numel(), stream), "cns_frame_mask")
        ws = self.net._workspace(dev.index, lib.cns_forward_workspace_bytes(C.byref(b)), ("fwd", stream))
        B = self.B
        hin = self.hidden[self.cur] if self.have_hidden else None
        hout = self.hidden[self.cur ^ 1]
        L.check(lib.cns_graphvs_forward(h.raw, C.byref(b), L.ptr(hin), ptrs["vec"], ptrs["nrm"], L.ptr(hout), ptrs["vel"],
                                        L.ptr(ws), ws.numel(), None, 0, C.c_void_p(stream)), "cns_graphvs_forward")
        self.cur ^= 1
        self.have_hidden = True
        with torch.cuda.stream(st):
            self.vel_host[k].copy_(self.out[k][B * 6:B * 12].view(B, 6), non_blocking=True)
            ev = self.done_ev[k]
            ev.record(st)
        self.done[k] = ev
        self.hidden_ready = ev
        self.ticket += 1
        return self.ticket - 1

    def _slot_batch(self, k):
        """cns_batch of slot k (everything but n_e01 and the target cache is fixed) and the raw pointers of its buffers."""
        if self._slot_cache[k] is None:
            L, B = _lib, self.B
            p = lambda t: t.data_ptr()
            xy, out = self.xy_dev[k], self.out[k]
            b = L.cns_batch()
            b.n_nodes, b.n_clusters, b.n_graphs = self.N, self.Cn, self.B
            b.n_e11_cur, b.n_e11_tar = 0, 0
            b.x_cur, b.pos_cur, b.x_tar, b.pos_tar = p(xy), p(xy), p(self.x_tar), p(self.x_tar)
            b.l0_to_l1_j, b.l0_to_l1_i = p(self.e01_j[k]), p(self.e01_i[k])
            b.l1_edge_cur, b.l1_edge_tar = None, None               # both cluster graphs are implied (complete)
            b.l1_cur_stride, b.l1_tar_stride = 0, 0
            b.cluster_mask, b.cluster_centers = p(self.cmask[k]), p(self.centers)
            b.node_graph, b.num_clusters, b.tPo_norm = None, p(self.num_clusters), p(self.tpo)
            b.flags = L.BATCH_L1_CUR_COMPLETE | L.BATCH_L1_TAR_COMPLETE
            b.max_clusters_per_graph = self.max_clusters
            b.clu_rowptr = p(self.clu_rowptr[k])
            ptrs = {"vis": L.ptr(self.vis_dev[k]), "e01_j": L.ptr(self.e01_j[k]), "e01_i": L.ptr(self.e01_i[k]),
                    "cmask": L.ptr(self.cmask[k]), "clu_rowptr": L.ptr(self.clu_rowptr[k]), "mask_ws": L.ptr(self.mask_ws[k]),
                    "vec": L.ptr(out[:B * 6]), "vel": L.ptr(out[B * 6:B * 12]), "nrm": L.ptr(out[B * 12:])}
            self._slot_cache[k] = (b, ptrs)
        return self._slot_cache[k]

    def result(self, ticket: int) -> torch.Tensor:
        """Velocities (B, 6) of a submitted frame, in pinned host memory (valid until `depth` more submits)."""
        k = ticket % self.depth
        self.done[k].synchronize()
        return self.vel_host[k]
