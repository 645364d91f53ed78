"""GPU (-m gpu): per-op C-ABI entry points vs the oracle's segment primitives / numpy."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import cns_oracle
from cns_b200 import Batch, GraphGenerator, _lib, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _sp():
    return _lib.stream_ptr(torch.device(DEV))


def _csr(edge_index, n_dst):
    lib = _lib.load()
    e = edge_index.shape[1]
    rowptr = torch.empty(n_dst + 1, dtype=torch.int32, device=DEV)
    src = torch.empty(max(e, 1), dtype=torch.int32, device=DEV)
    ws = torch.empty(1 << 20, dtype=torch.uint8, device=DEV)
    ei = edge_index.to(DEV).contiguous()
    _lib.check(lib.cns_csr_by_dst(_lib.ptr(ei), e, n_dst, _lib.ptr(rowptr), _lib.ptr(src), _lib.ptr(ws),
                                  ws.numel(), _sp()), "csr")
    return rowptr, src[:e]


def test_csr_by_dst_is_a_permutation_grouped_by_destination():
    g = torch.Generator().manual_seed(0)
    n, e = 300, 5000
    ei = torch.stack([torch.randint(0, n, (e,), generator=g), torch.randint(0, n - 10, (e,), generator=g)])
    rowptr, src = _csr(ei, n)
    rowptr, src = rowptr.cpu().long(), src.cpu().long()
    deg = torch.bincount(ei[1], minlength=n)
    assert torch.equal(rowptr[1:] - rowptr[:-1], deg) and int(rowptr[0]) == 0
    for i in (0, 5, 123, 289, 299):      # rows hold exactly the sources of that destination (as multisets)
        mine = sorted(src[rowptr[i]:rowptr[i + 1]].tolist())
        assert mine == sorted(ei[0][ei[1] == i].tolist())
    r0, s0 = _csr(torch.zeros(2, 0, dtype=torch.long), 7)       # empty edge list
    assert r0.cpu().tolist() == [0] * 8


@pytest.mark.parametrize("f", [128, 256])
def test_edgeconv_max_factorised_equals_literal_message_passing(f):
    """out_i = P_i + max_j Q_j equals max_j Linear([x_i, x_j - x_i, pos_j - pos_i]) of the oracle
    (graph_vs.py:107-111), including destinations with no in-edge (-> 0)."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(1)
    n, fin = 200, f
    x = torch.randn(n, fin, generator=g, dtype=torch.float64)
    pos = torch.randn(n, 2, generator=g, dtype=torch.float64)
    W = torch.randn(f, 2 * fin + 2, generator=g, dtype=torch.float64) / (fin ** 0.5)
    b = torch.randn(f, generator=g, dtype=torch.float64)
    ei = torch.stack([torch.randint(0, n, (3000,), generator=g), torch.randint(0, n - 20, (3000,), generator=g)])
    ref = cns_oracle.peconv({"c.msg_nn.fcs.0.weight": W, "c.msg_nn.fcs.0.bias": b}, "c", x, pos, ei)
    W1, W2, W3 = W[:, :fin], W[:, fin:2 * fin], W[:, 2 * fin:]
    P = x @ (W1 - W2).T - pos @ W3.T + b
    Q = x @ W2.T + pos @ W3.T
    pq = torch.cat([P, Q], 1).float().to(DEV).contiguous()
    rowptr, src = _csr(ei, n)
    out = torch.empty(n, f, device=DEV)
    _lib.check(lib.cns_edgeconv_max(_lib.ptr(pq), _lib.ptr(rowptr), _lib.ptr(src), n, f, _lib.ptr(out), _sp()), "ec")
    assert float((out.double().cpu() - ref).abs().max() / ref.abs().max()) < 1e-5
    assert float(out[n - 20:].abs().max()) == 0.0


def test_graphnorm_relu_matches_oracle():
    lib = _lib.load()
    g = torch.Generator().manual_seed(2)
    sizes = [1, 7, 18, 40, 3]
    n = sum(sizes)
    x = torch.randn(n, 128, generator=g) * 3 + 1
    x[5] = 0                                   # masked cluster rows are all-zero but counted
    w, b, ms = torch.rand(128, generator=g) + .5, torch.randn(128, generator=g), torch.rand(128, generator=g)
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    ref = torch.relu(cns_oracle.graph_norm({"p.weight": w.double(), "p.bias": b.double(), "p.mean_scale": ms.double()},
                                           "p", x.double(), batch, len(sizes)))
    gp = torch.tensor(np.cumsum([0] + sizes), dtype=torch.int32, device=DEV)
    y = torch.empty(n, 128, device=DEV)
    xd, wd, bd, msd = x.to(DEV), w.to(DEV), b.to(DEV), ms.to(DEV)
    _lib.check(lib.cns_graphnorm_relu(_lib.ptr(xd), _lib.ptr(gp), len(sizes), _lib.ptr(wd), _lib.ptr(bd),
                                      _lib.ptr(msd), _lib.ptr(y), _sp()), "gn")
    assert float((y.double().cpu() - ref).abs().max()) < 2e-5 * float(ref.abs().max())


@pytest.mark.parametrize("m,k1,k2,n", [(1, 128, 0, 128), (77, 128, 128, 256), (1000, 128, 128, 512)])
def test_linear_matches_fp64(m, k1, k2, n):
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    x1, x2 = torch.randn(m, k1, generator=g), torch.randn(m, max(k2, 1), generator=g)
    wt, bias, res = torch.randn(k1 + k2, n, generator=g) / 16, torch.randn(n, generator=g), torch.randn(m, n, generator=g)
    ref = x1.double() @ wt[:k1].double()
    if k2:
        ref = ref + x2.double() @ wt[k1:].double()
    ref = torch.relu(ref + bias.double() + res.double())
    y = torch.empty(m, n, device=DEV)
    a = [t.to(DEV).contiguous() for t in (x1, x2, wt, bias, res)]
    _lib.check(lib.cns_linear(_lib.ptr(a[0]), k1, _lib.ptr(a[1]) if k2 else None, k2, _lib.ptr(a[2]), _lib.ptr(a[3]),
                              _lib.ptr(a[4]), 1, m, n, _lib.ptr(y), _sp()), "linear")
    assert float((y.double().cpu() - ref).abs().max()) < 1e-5 * float(ref.abs().max())


def test_graph_mask_bit_exact_vs_graph_generator():
    """Batched per-frame masking on device == GraphGenerator._get_graph + Batch collation (bit-exact,
    edge order included): reference graph_gen.py:199-228."""
    lib = _lib.load()
    rng = np.random.default_rng(3)
    np.random.seed(3)
    gens, vis, frames = [], [], []
    for n in (12, 40, 130, 16, 300):
        pts = synth.project_norm(synth.sample_target_pose(rng), synth.sample_points(rng, n))
        g = GraphGenerator(pts)
        frac = rng.choice([0.0, 0.3, 0.8, 1.0])
        missing = np.sort(rng.choice(len(pts), size=int(len(pts) * frac), replace=False))
        v = np.ones(len(pts), bool); v[missing] = False
        gens.append(g); vis.append(v); frames.append(g.get_data(pts, missing_node_indices=missing))
    ref = Batch.from_data_list(frames)
    n_nodes = np.array([g.num_points for g in gens]); n_clu = np.array([g.num_clusters for g in gens])
    noff, coff = np.cumsum(n_nodes) - n_nodes, np.cumsum(n_clu) - n_clu
    belong = np.concatenate([g.belong_cluster_index + c for g, c in zip(gens, coff)])
    order = np.concatenate([g.l0_to_l1_edge_index[0] + o for g, o in zip(gens, noff)])
    csize = np.concatenate([g.cluster_points_count for g in gens]).astype(np.int64)
    gptr = np.concatenate([[0], np.cumsum(n_clu)]).astype(np.int64)
    N, Cn, B = int(n_nodes.sum()), int(n_clu.sum()), len(gens)
    cap = int((n_clu ** 2).sum())
    d = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=DEV)
    t_belong, t_order, t_csize, t_gptr = d(belong, torch.int64), d(order, torch.int64), d(csize, torch.int64), d(gptr, torch.int64)
    t_vis = d(np.concatenate(vis), torch.uint8)
    e01_j, e01_i = torch.empty(N, dtype=torch.int64, device=DEV), torch.empty(N, dtype=torch.int64, device=DEV)
    e11 = torch.empty(2, cap, dtype=torch.int64, device=DEV)
    cmask = torch.empty(Cn, dtype=torch.uint8, device=DEV)
    counts = torch.zeros(2, dtype=torch.int64, device=DEV)
    ws = torch.empty(lib.cns_graph_mask_workspace_bytes(N, Cn, B), dtype=torch.uint8, device=DEV)
    _lib.check(lib.cns_graph_mask(_lib.ptr(t_belong), _lib.ptr(t_order), _lib.ptr(t_csize), _lib.ptr(t_gptr),
                                  _lib.ptr(t_vis), N, Cn, B, _lib.ptr(e01_j), _lib.ptr(e01_i), _lib.ptr(e11), cap,
                                  _lib.ptr(cmask), _lib.ptr(counts), _lib.ptr(ws), ws.numel(), _sp()), "graph_mask")
    n01, n11 = counts.cpu().tolist()
    assert n01 == ref.l0_to_l1_edge_index_j_cur.numel() and n11 == ref.l1_dense_edge_index_cur.shape[1]
    assert torch.equal(e01_j[:n01].cpu(), ref.l0_to_l1_edge_index_j_cur)
    assert torch.equal(e01_i[:n01].cpu(), ref.l0_to_l1_edge_index_i_cur)
    assert torch.equal(e11[:, :n11].cpu(), ref.l1_dense_edge_index_cur)
    assert torch.equal(cmask.bool().cpu(), ref.cluster_mask)


def test_frame_mask_bit_exact_vs_graph_generator():
    """Cluster-major per-frame masking (cns_frame_mask): keypoint->cluster edges, cluster mask and the cluster
    segment pointers == GraphGenerator._get_graph + Batch collation, bit-exact (reference graph_gen.py:199-228)."""
    lib = _lib.load()
    rng = np.random.default_rng(5)
    np.random.seed(5)
    gens, vis, frames = [], [], []
    for n in (12, 40, 130, 16, 300, 64):
        pts = synth.project_norm(synth.sample_target_pose(rng), synth.sample_points(rng, n))
        g = GraphGenerator(pts)
        frac = rng.choice([0.0, 0.3, 0.8, 1.0])
        missing = np.sort(rng.choice(len(pts), size=int(len(pts) * frac), replace=False))
        v = np.ones(len(pts), bool); v[missing] = False
        gens.append(g); vis.append(v); frames.append(g.get_data(pts, missing_node_indices=missing))
    ref = Batch.from_data_list(frames)
    n_nodes = np.array([g.num_points for g in gens]); n_clu = np.array([g.num_clusters for g in gens])
    noff = np.cumsum(n_nodes) - n_nodes
    order = np.concatenate([g.l0_to_l1_edge_index[0] + o for g, o in zip(gens, noff)])
    csize = np.concatenate([g.cluster_points_count for g in gens]).astype(np.int64)
    N, Cn = int(n_nodes.sum()), int(n_clu.sum())
    d = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=DEV)
    t_order, t_ptr = d(order, torch.int64), d(np.concatenate([[0], np.cumsum(csize)]), torch.int32)
    t_vis = d(np.concatenate(vis), torch.uint8)
    e01_j, e01_i = torch.full((N,), -1, dtype=torch.int64, device=DEV), torch.full((N,), -1, dtype=torch.int64, device=DEV)
    cmask = torch.empty(Cn, dtype=torch.uint8, device=DEV)
    rowptr = torch.empty(Cn + 1, dtype=torch.int32, device=DEV)
    ws = torch.empty((Cn + 1) * 4, dtype=torch.uint8, device=DEV)
    _lib.check(lib.cns_frame_mask(_lib.ptr(t_order), _lib.ptr(t_ptr), _lib.ptr(t_vis), N, Cn, _lib.ptr(e01_j), _lib.ptr(e01_i),
                                  _lib.ptr(cmask), _lib.ptr(rowptr), _lib.ptr(ws), ws.numel(), _sp()), "frame_mask")
    n01 = ref.l0_to_l1_edge_index_j_cur.numel()
    assert int(rowptr[-1]) == n01 == int(t_vis.sum())
    assert torch.equal(e01_j[:n01].cpu(), ref.l0_to_l1_edge_index_j_cur)
    assert torch.equal(e01_i[:n01].cpu(), ref.l0_to_l1_edge_index_i_cur)
    assert torch.equal(cmask.bool().cpu(), ref.cluster_mask)
    want = torch.zeros(Cn + 1, dtype=torch.int64)
    want[1:] = torch.cumsum(torch.bincount(ref.l0_to_l1_edge_index_i_cur, minlength=Cn), 0)
    assert torch.equal(rowptr.cpu().long(), want)


@pytest.mark.parametrize("n_clusters", [1, 1024, 1025, 70001])
def test_frame_mask_many_clusters(n_clusters):
    """The scan of cns_frame_mask runs 1024 clusters per CTA: sizes around the CTA boundary and far beyond it (ragged
    clusters, some empty, some fully hidden) against the same compaction written in numpy."""
    lib = _lib.load()
    rng = np.random.default_rng(n_clusters)
    csize = rng.integers(0, 13, size=n_clusters)
    csize[0] = max(csize[0], 1)
    N = int(csize.sum())
    order = rng.permutation(N).astype(np.int64)
    ptr = np.concatenate([[0], np.cumsum(csize)]).astype(np.int32)
    vis = rng.random(N) < 0.6
    hidden_clusters = rng.choice(n_clusters, size=max(1, n_clusters // 50), replace=False)
    for c in hidden_clusters:
        vis[order[ptr[c]:ptr[c + 1]]] = False
    owner = np.repeat(np.arange(n_clusters), csize)
    keep = vis[order]
    want_j, want_i = order[keep], owner[keep]
    want_ptr = np.concatenate([[0], np.cumsum(np.bincount(want_i, minlength=n_clusters))])
    d = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=DEV)
    t_order, t_ptr, t_vis = d(order, torch.int64), d(ptr, torch.int32), d(vis, torch.uint8)
    e01_j = torch.full((max(N, 1),), -1, dtype=torch.int64, device=DEV)
    e01_i = torch.full((max(N, 1),), -1, dtype=torch.int64, device=DEV)
    cmask = torch.empty(n_clusters, dtype=torch.uint8, device=DEV)
    rowptr = torch.empty(n_clusters + 1, dtype=torch.int32, device=DEV)
    ws = torch.empty((n_clusters + 1) * 4, dtype=torch.uint8, device=DEV)
    _lib.check(lib.cns_frame_mask(_lib.ptr(t_order), _lib.ptr(t_ptr), _lib.ptr(t_vis), N, n_clusters, _lib.ptr(e01_j),
                                  _lib.ptr(e01_i), _lib.ptr(cmask), _lib.ptr(rowptr), _lib.ptr(ws), ws.numel(), _sp()), "frame_mask")
    n01 = len(want_j)
    assert np.array_equal(rowptr.cpu().numpy(), want_ptr)
    assert np.array_equal(e01_j[:n01].cpu().numpy(), want_j)
    assert np.array_equal(e01_i[:n01].cpu().numpy(), want_i)
    assert np.array_equal(cmask.cpu().numpy().astype(bool), np.diff(want_ptr) > 0)


def test_label_nearest_matches_numpy_argmax():
    lib = _lib.load()
    rng = np.random.default_rng(4)
    sizes, ks = [50, 700, 9], [4, 23, 9]
    pts = [rng.random((n, 2)) for n in sizes]
    ex = [p[rng.choice(len(p), k, replace=False)] for p, k in zip(pts, ks)]
    want = []
    for p, e in zip(pts, ex):
        d2 = (-2 * (p @ e.T) + (p * p).sum(1)[:, None]) + (e * e).sum(1)[None, :]
        want.append(np.argmax(-np.maximum(d2, 0), axis=1))
    P = torch.as_tensor(np.concatenate(pts), device=DEV)
    E = torch.as_tensor(np.concatenate(ex), device=DEV)
    pg = torch.as_tensor(np.repeat(np.arange(3), sizes), device=DEV)
    xp = torch.as_tensor(np.concatenate([[0], np.cumsum(ks)]), device=DEV)
    lab = torch.empty(sum(sizes), dtype=torch.int32, device=DEV)
    _lib.check(lib.cns_label_nearest(_lib.ptr(P), _lib.ptr(pg), sum(sizes), _lib.ptr(E), _lib.ptr(xp), _lib.ptr(lab), _sp()), "label")
    assert np.array_equal(lab.cpu().numpy(), np.concatenate(want))


@pytest.mark.parametrize("seed,sizes", [(0, [60]), (1, [17, 200, 64]), (2, [512]), (3, [300, 33, 128, 450])])
def test_affinity_propagation_device_identical_to_sklearn(seed, sizes):
    """GPU Affinity Propagation == sklearn AffinityPropagation(damping=0.9).fit_predict: labels and
    iteration counts identical (bit-exact fp64 message passing, same numpy random stream)."""
    import warnings
    from sklearn.cluster import AffinityPropagation
    from cns_b200.cluster import affinity_propagation_device, cluster_points_device
    from cns_b200.graph_gen import cluster_points
    rng = np.random.default_rng(seed)
    sets = [synth.project_norm(synth.sample_target_pose(rng), synth.sample_points(rng, n)) for n in sizes]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.random.seed(seed)
        models = [AffinityPropagation(damping=0.9).fit(X) for X in sets]
        np.random.seed(seed)
        mine = affinity_propagation_device(sets, device=DEV)
    for X, m, lab, it in zip(sets, models, mine, affinity_propagation_device.last_n_iter):
        assert np.array_equal(lab, m.labels_)
        assert int(it) == int(m.n_iter_)
    # the reference entry point (with its KMeans fallback) end to end
    np.random.seed(seed + 10); a = cluster_points(sets[0])
    np.random.seed(seed + 10); b = cluster_points_device(sets[0], device=DEV)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_linear_split_matches_fp64():
    """cns_linear_split (tensor cores, three bf16 slices per operand) must be as accurate as an fp32 FMA chain -
    it replaces the FFMA GEMMs of the backward pass, whose gradients are held to 1e-4 of fp64 autograd."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(7)
    ws = torch.empty(lib.cns_linear_split_workspace_bytes(), dtype=torch.uint8, device=DEV)
    for m, scale in [(1, 1.0), (127, 1.0), (128, 1e-6), (129, 1.0), (5000, 1e3), (40000, 1.0)]:
        x = (torch.randn(m, 128, generator=g) * scale).to(DEV)
        x[0, :5] = torch.tensor([1e-30, -3e-38, 1e20, 0.0, -1e-20])[:5].to(DEV) if m > 1 else x[0, :5]
        wt = (torch.randn(128, 128, generator=g) / 11).to(DEV)
        bias = torch.randn(128, generator=g).to(DEV) * scale
        res = (torch.randn(m, 128, generator=g) * scale).to(DEV)
        gate = torch.randn(m, 128, generator=g).to(DEV)
        for use_bias, use_res, relu, use_gate in [(0, 0, 0, 0), (1, 1, 1, 0), (1, 1, 0, 1), (0, 1, 0, 1)]:
            y = torch.full((m, 128), float("nan"), device=DEV)
            _lib.check(lib.cns_linear_split(_lib.ptr(x), m, _lib.ptr(wt), _lib.ptr(bias if use_bias else None),
                                            _lib.ptr(res if use_res else None), _lib.ptr(gate if use_gate else None),
                                            relu, _lib.ptr(y), _lib.ptr(ws), ws.numel(), _sp()), "cns_linear_split")
            ref = x.double() @ wt.double()
            if use_bias:
                ref = ref + bias.double()
            if use_res:
                ref = ref + res.double()
            if relu:
                ref = ref.clamp_min(0)
            if use_gate:
                ref = torch.where(gate > 0, ref, torch.zeros_like(ref))
            fp32 = x @ wt                                           # cuBLAS fp32 as the yardstick for "fp32 accuracy"
            denom = (x.double().abs() @ wt.double().abs()).max().item() + 1e-300
            err = (y.double() - ref).abs().max().item() / denom
            assert torch.isfinite(y).all() and err < 2e-7, (m, scale, use_bias, use_res, relu, use_gate, err)
            if not (use_bias or use_res or relu or use_gate):
                err32 = (fp32.double() - ref).abs().max().item() / denom
                assert err < 4 * err32 + 1e-8, (err, err32)
        # in place on the residual, as the backward pass uses it
        y = res.clone()
        _lib.check(lib.cns_linear_split(_lib.ptr(x), m, _lib.ptr(wt), None, _lib.ptr(y), None, 0, _lib.ptr(y),
                                        _lib.ptr(ws), ws.numel(), _sp()), "cns_linear_split")
        ref = x.double() @ wt.double() + res.double()
        assert (y.double() - ref).abs().max().item() / ((x.double().abs() @ wt.double().abs()).max().item() + res.abs().max().item()) < 2e-7


def test_weight_grad_tensor_core_path_matches_fp64():
    """cns_weight_grad: dW = dY^T X and dbias over many rows, tensor-core (split bf16, MN-major operands) and FFMA
    paths against fp64; both must be fp32-accurate and deterministic."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(11)
    ws = torch.empty(lib.cns_weight_grad_workspace_bytes(), dtype=torch.uint8, device=DEV)
    for m, scale in [(4096, 1.0), (4097, 1e-4), (50000, 1.0), (118000, 3.0), (300, 1.0)]:
        dy = (torch.randn(m, 128, generator=g) * scale).to(DEV)
        x = torch.randn(m, 128, generator=g).to(DEV)
        ref_w = dy.double().t() @ x.double()
        ref_b = dy.double().sum(0)
        den_w = (dy.double().abs().t() @ x.double().abs()).max().item()
        den_b = dy.double().abs().sum(0).max().item()
        outs = []
        for tc in (1, 0, 1):
            dw = torch.full((128, 128), float("nan"), device=DEV)
            db = torch.full((128,), float("nan"), device=DEV)
            _lib.check(lib.cns_weight_grad(_lib.ptr(dy), 128, _lib.ptr(x), 128, m, _lib.ptr(dw), _lib.ptr(db), 0, tc,
                                           _lib.ptr(ws), ws.numel(), _sp()), "cns_weight_grad")
            ew = (dw.double() - ref_w).abs().max().item() / den_w
            eb = (db.double() - ref_b).abs().max().item() / den_b
            assert ew < 3e-7 and eb < 3e-7, (m, scale, tc, ew, eb)
            outs.append((dw.clone(), db.clone()))
        assert torch.equal(outs[0][0], outs[2][0]) and torch.equal(outs[0][1], outs[2][1])       # deterministic
        # accumulate on top of an existing gradient
        dw0 = torch.randn(128, 128, generator=g).to(DEV)
        dw = dw0.clone()
        _lib.check(lib.cns_weight_grad(_lib.ptr(dy), 128, _lib.ptr(x), 128, m, _lib.ptr(dw), None, 1, 1,
                                       _lib.ptr(ws), ws.numel(), _sp()), "cns_weight_grad")
        assert (dw.double() - dw0.double() - ref_w).abs().max().item() / (den_w + 1) < 3e-7
    # strided operands (columns of a wider matrix), as the backward pass passes them
    m = 9000
    wide = torch.randn(m, 256, generator=g).to(DEV)
    x = torch.randn(m, 128, generator=g).to(DEV)
    dw = torch.empty(128, 128, device=DEV)
    _lib.check(lib.cns_weight_grad(C.c_void_p(wide.data_ptr() + 128 * 4), 256, _lib.ptr(x), 128, m, _lib.ptr(dw), None, 0, 1,
                                   _lib.ptr(ws), ws.numel(), _sp()), "cns_weight_grad")
    ref = wide[:, 128:].double().t() @ x.double()
    assert (dw.double() - ref).abs().max().item() / ref.abs().max().item() < 1e-6
