"""CPU: host-side mirror of the reference interface - containers, collation, checkpoints,
parameter groups - and, where /root/reference is present (build container only), a live
cross-check against the reference's own modules running on the PyG shim."""
import os

import numpy as np
import pytest
import torch

import cns_oracle
import ref_import
from cns_b200 import Batch, GraphData, GraphGenerator, GraphVS, Midend, CameraIntrinsic, Correspondence
from cns_b200 import ckpt, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")


def test_both_checkpoints_load_identically():
    a = ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns_state_dict.pth"))
    b = ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns.pth"))     # pickled reference module
    assert list(a.keys()) == list(b.keys()) and len(a) == 43
    assert sum(v.numel() for v in a.values()) == 464520
    assert all(torch.equal(a[k], b[k]) for k in a)
    net = GraphVS(2, 2, 128, regress_norm=True)
    net.load_state_dict(b, strict=True)
    assert list(net.state_dict().keys()) == list(a.keys())


def test_parameter_groups():
    net = GraphVS(2, 2, 128)
    decay, no_decay = net.get_parameter_groups()
    assert len(decay) + len(no_decay) == 43
    assert all(p.dim() == 2 for p in decay) and all(p.dim() == 1 for p in no_decay)


def test_graphdata_defaults_and_batching():
    rng = np.random.default_rng(0)
    np.random.seed(0)
    scenes = [synth.Scene(rng, n) for n in (5, 30, 64)]
    frames = [s.frame(0.2) for s in scenes]
    assert frames[0].new_scene.item() is False and float(frames[0].tPo_norm) > 0
    frames[1].start_new_scene()
    frames[1].intrinsic = CameraIntrinsic.default()      # non-tensor extras must survive
    for f in frames:
        f.intrinsic = CameraIntrinsic.default()
    b = Batch.from_data_list(frames)
    assert b.num_graphs == 3 and b.num_nodes == 5 + 30 + 64
    assert b.new_scene.tolist() == [False, True, False]
    assert b.num_clusters.tolist() == [int(f.num_clusters) for f in frames]
    n_off = np.cumsum([0, 5, 30]); c_off = np.cumsum([0] + [int(f.num_clusters) for f in frames[:2]])
    for g, f in enumerate(frames):
        sel = b.batch == g
        assert torch.equal(b.x_cur[sel], f.x_cur)
    # offsets: j / centres by nodes, i / l1 by clusters
    e0 = 0
    for g, f in enumerate(frames):
        e1 = e0 + f.l0_to_l1_edge_index_j_cur.numel()
        assert torch.equal(b.l0_to_l1_edge_index_j_cur[e0:e1], f.l0_to_l1_edge_index_j_cur + int(n_off[g]))
        assert torch.equal(b.l0_to_l1_edge_index_i_cur[e0:e1], f.l0_to_l1_edge_index_i_cur + int(c_off[g]))
        e0 = e1
    assert b.l1_dense_edge_index_tar.shape[0] == 2
    assert hasattr(b, "batch") and not hasattr(frames[0], "batch")
    assert b["x_cur"] is b.x_cur
    assert isinstance(b.intrinsic, list)
    # structure hints for the C ABI (cns_batch.flags / max_clusters_per_graph): what GraphGenerator built is a
    # complete cluster graph; a frame of unknown origin in the batch clears the claim
    assert all(f.l1_complete is True for f in frames) and b.l1_complete is True
    assert b.max_clusters_per_graph == max(int(f.num_clusters) for f in frames)
    frames[2].l1_complete = False
    assert Batch.from_data_list(frames).l1_complete is False
    moved = b.to("cpu")
    assert moved.l1_complete is True and moved.max_clusters_per_graph == b.max_clusters_per_graph


def test_complete_graph_hint_matches_the_edge_lists():
    """The claim behind CNS_BATCH_L1_*_COMPLETE, checked on the host: l1_dense_edge_index_tar is every ordered
    pair (self loops included) over all clusters, ..._cur over the clusters with cluster_mask, both in
    source-major order (reference graph_gen.py:31-36, 199-228)."""
    rng = np.random.default_rng(4)
    np.random.seed(4)
    s = synth.Scene(rng, 120)
    for frac in (0.0, 0.5, 0.95):
        d = s.frame(frac)
        n = int(d.num_clusters)
        alive = torch.nonzero(d.cluster_mask).flatten()
        full = torch.arange(n)
        for edges, nodes in ((d.l1_dense_edge_index_tar, full), (d.l1_dense_edge_index_cur, alive)):
            m = nodes.numel()
            assert edges.shape == (2, m * m)
            assert torch.equal(edges[0], nodes.repeat_interleave(m)) and torch.equal(edges[1], nodes.repeat(m))


def test_small_graph_every_point_its_own_cluster():
    pts = np.random.default_rng(3).random((9, 2))
    g = GraphGenerator(pts)
    assert g.num_clusters == 9
    assert np.array_equal(g.cluster_centers_index, np.arange(9))
    d = g.get_data(pts, missing_node_indices=np.array([2, 7]))
    assert d.cluster_mask.tolist() == [True, True, False, True, True, True, True, False, True]
    assert d.l1_dense_edge_index_cur.shape == (2, 49) and d.l1_dense_edge_index_tar.shape == (2, 81)
    assert d.l0_to_l1_edge_index_j_cur.tolist() == [0, 1, 3, 4, 5, 6, 8]


def test_midend_pixel_to_graph():
    rng = np.random.default_rng(5)
    intr = CameraIntrinsic.default()
    tar = rng.random((40, 2)) * [640, 480]
    cur = tar + rng.normal(size=(40, 2))
    valid = rng.random(40) > 0.2
    corr = Correspondence(intrinsic=intr, tar_img=None, tar_pos=tar, cur_img=None, cur_pos=None, match=None,
                          valid_mask=valid, cur_pos_aligned=cur, tar_img_changed=True)
    mid = Midend()
    np.random.seed(1)
    d = mid.get_graph_data(corr)
    assert d.x_tar.shape == (40, 2) and d.x_cur.dtype == torch.float32
    assert np.allclose(d.pos_tar.numpy(), (tar - [320, 240]) / 540, atol=1e-6)
    assert d.node_mask.tolist() == valid.tolist()
    gen = mid.graph_generator
    corr.tar_img_changed = False
    mid.get_graph_data(corr)
    assert mid.graph_generator is gen            # clustering happens once per target


@needs_ref
def test_live_reference_cross_check(state_dict):
    """Reference GraphVS / GraphGenerator / Batch (unmodified, on the shim) vs ours, fresh seeds."""
    gv, gg, fn = ref_import.load()
    from torch_geometric.data import Batch as RefBatch
    net = gv.GraphVS(2, 2, 128, regress_norm=True)
    net.load_state_dict(state_dict)
    net.eval()
    rng = np.random.default_rng(77)
    ref_frames, my_frames = [], []
    for n in (17, 60, 150):
        pts = synth.project_norm(synth.sample_target_pose(rng), synth.sample_points(rng, n))
        cur = pts + rng.normal(scale=0.01, size=pts.shape)
        missing = np.sort(rng.choice(len(pts), size=len(pts) // 5, replace=False))
        np.random.seed(n); rg = gg.GraphGenerator(pts)
        np.random.seed(n); mg = GraphGenerator(pts)
        ref_frames.append(rg.get_data(cur, missing_node_indices=missing))
        my_frames.append(mg.get_data(cur, missing_node_indices=missing))
    rb, mb = RefBatch.from_data_list(ref_frames), Batch.from_data_list(my_frames)
    for k in ("x_cur", "pos_tar", "l0_to_l1_edge_index_j_cur", "l0_to_l1_edge_index_i_cur",
              "l1_dense_edge_index_cur", "l1_dense_edge_index_tar", "cluster_mask", "cluster_centers_index",
              "belong_cluster_index", "node_mask", "num_clusters", "batch", "tPo_norm", "new_scene"):
        assert torch.equal(getattr(rb, k), getattr(mb, k)), k
    with torch.no_grad():
        rv, rn, rh = net(rb, None)
    ov, on, oh = cns_oracle.forward(state_dict, mb, None)
    assert float((rv - ov).abs().max()) < 2e-6 and float((rn - on).abs().max()) < 2e-6
    assert float((rh - oh).abs().max()) < 2e-6
    # parameter groups are the same split
    d0, n0 = net.get_parameter_groups()
    d1, n1 = GraphVS(2, 2, 128).get_parameter_groups()
    assert [tuple(p.shape) for p in d0] == [tuple(p.shape) for p in d1]
    assert sorted(tuple(p.shape) for p in n0) == sorted(tuple(p.shape) for p in n1)
    # loss / postprocess restatement
    gt = torch.randn(3, 6)
    mb.vel_si = gt; rb.vel_si = gt
    res, loss = fn.objectives((rv, rn), rb, gt_si=True)
    dl, nl, tl = cns_oracle.objectives((ov, on), gt)
    assert abs(float(loss) - float(tl)) < 1e-5
    assert abs(res["dir_loss"] - float(dl)) < 1e-5 and abs(res["norm_loss"] - float(nl)) < 1e-5
    rvel = fn.postprocess((rv, rn), rb, post_scale=True)
    assert float((rvel - cns_oracle.postprocess((ov, on), mb.tPo_norm.reshape(-1))).abs().max()) < 1e-6


@needs_ref
def test_projection_matches_reference_camera():
    """synth.project_norm == reference Camera.project + pixel_to_norm_camera_plane."""
    ref_import.setup()
    from cns.utils.perception import Camera, CameraIntrinsic as RefIntr
    rng = np.random.default_rng(2)
    pts = synth.sample_points(rng, 50)
    wcT = synth.sample_initial_pose(rng)
    cam = Camera(RefIntr.default())
    uv = cam.project(np.linalg.inv(wcT), pts)
    xy = cam.intrinsic.pixel_to_norm_camera_plane(uv)
    assert np.allclose(xy, synth.project_norm(wcT, pts), atol=1e-9)


def test_batched_pose_algebra_matches_the_scene_functions():
    """cns_b200/sim_device.py (batched torch, fp64) vs the per-scene numpy functions of synth.py."""
    from cns_b200 import sim_device as sd
    rng = np.random.default_rng(11)
    B = 64
    cur = np.stack([synth.sample_initial_pose(rng) for _ in range(B)])
    tar = np.stack([synth.sample_target_pose(rng) for _ in range(B)])
    n = rng.integers(4, 40, B)
    pts = [synth.sample_points(rng, k) for k in n]
    n = np.array([len(p) for p in pts])
    tc, tt = torch.from_numpy(cur), torch.from_numpy(tar)
    P, scene = torch.from_numpy(np.concatenate(pts)), torch.from_numpy(np.repeat(np.arange(B), n))
    xy = sd.project_norm(tc, P, scene).numpy()
    ref = np.concatenate([synth.project_norm(cur[k], pts[k]) for k in range(B)])
    assert np.abs(xy - ref).max() < 1e-12
    v = sd.pbvs_straight(tc, tt).numpy()
    vref = np.stack([synth.pbvs_straight(cur[k], tar[k]) for k in range(B)])
    assert np.abs(v - vref).max() < 1e-9
    vel = rng.normal(size=(B, 6))
    nxt = sd.integrate(tc, torch.from_numpy(vel)).numpy()
    nref = np.stack([synth.integrate(cur[k], vel[k]) for k in range(B)])
    assert np.abs(nxt - nref).max() < 1e-9
    ang, mm = sd.pose_error(tc, tt)
    eref = np.array([synth.pose_error(cur[k], tar[k]) for k in range(B)])
    assert np.abs(ang.numpy() - eref[:, 0]).max() < 1e-7 and np.abs(mm.numpy() - eref[:, 1]).max() < 1e-7
    # rotations close to pi and close to 0 round-trip through log / exp
    w = torch.tensor([[np.pi - 1e-7, 0, 0], [0, 1e-9, 0], [1.2, -2.0, 0.4], [0, 0, -(np.pi - 1e-6)]], dtype=torch.float64)
    R = sd.so3_exp(w)
    assert float((sd.so3_exp(sd.so3_log(R)) - R).abs().max()) < 1e-7


def test_closed_loop_servo_converges(state_dict):
    """Pretrained weights + oracle forward drive a simulated camera to its target pose: this only
    happens when edge direction, softmax, GraphNorm and GRU semantics are the reference's."""
    rng = np.random.default_rng(4)
    np.random.seed(4)
    scene = synth.Scene(rng, 120)
    e0 = synth.pose_error(scene.cur_wcT, scene.tar_wcT)
    hidden = None
    for step in range(150):
        d = scene.frame(0.1 if step % 2 else 0.0)
        vec, nrm, hidden = cns_oracle.forward(state_dict, d, hidden)
        vel = cns_oracle.postprocess((vec, nrm), d.tPo_norm.reshape(1))
        scene.step(vel[0].numpy() * 2)
    e1 = synth.pose_error(scene.cur_wcT, scene.tar_wcT)
    assert e1[0] < 1.0 and e1[1] < 10.0, (e0, e1)   # < 1 degree, < 10 mm
    assert e1[0] < 0.1 * e0[0]


def test_degenerate_similarity_rule_matches_sklearn():
    """cluster._prepare_similarity restates sklearn's early exit for mutually equal similarities (no noise drawn,
    numpy's global stream untouched) instead of calling scikit-learn behind the device path's back."""
    import warnings
    from sklearn.cluster import AffinityPropagation
    from cns_b200.cluster import _prepare_similarity
    for X in (np.tile([[0.3, -0.2]], (20, 1)), np.array([[0.0, 0.0], [1.0, 0.0]]), np.array([[0.5, 0.5]])):
        np.random.seed(3)
        state = np.random.get_state()[1].copy()
        _, labels = _prepare_similarity(X)
        assert np.array_equal(np.random.get_state()[1], state)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = AffinityPropagation(damping=0.9).fit_predict(X)
        assert labels is not None and np.array_equal(labels, ref), (X.shape, labels, ref)
    _, labels = _prepare_similarity(np.random.default_rng(0).random((30, 2)))
    assert labels is None


def test_tbptt_hidden_rows_follow_scenes_that_restart_one_by_one():
    """TBPTTStepper._carry_hidden (host logic, no kernels): with `hidden_reset="scene"` the rows of restarted scenes are
    replaced by zeros - spliced in at the scene's NEW cluster count - while the other scenes keep their rows in order;
    the reference's literal rule (`"batch"`, train_gvs_short_seq.py:28) drops the whole state; a state whose row count
    does not match the frame is never reused."""
    from types import SimpleNamespace
    from cns_b200.train import TBPTTStepper

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w, self.b = torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(1))

        def get_parameter_groups(self):
            return [self.w], [self.b]

    frame = lambda counts, fresh: SimpleNamespace(num_clusters=torch.tensor(counts), new_scene=torch.tensor(fresh))
    old_counts = [2, 3, 1]
    hidden = torch.arange(6, dtype=torch.float32).reshape(6, 1).repeat(1, 128) + 1      # rows 1..6
    for rule in ("scene", "batch"):
        st = TBPTTStepper(Tiny(), fused=False, hidden_reset=rule)
        assert st._carry_hidden(frame(old_counts, [True] * 3)) is None                     # no state yet
        st.hidden, st._counts = hidden.clone(), torch.tensor(old_counts)
        same = st._carry_hidden(frame(old_counts, [False] * 3))
        assert same is st.hidden                                                            # nothing restarted
        got = st._carry_hidden(frame([2, 4, 1], [False, True, False]))                      # scene 1 is back with 4 clusters
        if rule == "batch":
            assert got is None
        else:
            assert got.shape == (7, 128)
            assert got[:, 0].tolist() == [1, 2, 0, 0, 0, 0, 6]
        assert st._carry_hidden(frame([2, 3, 1], [True, True, True])) is None               # everything restarted
        # a cluster count changed WITHOUT a restart flag: the rows cannot be matched, the state is dropped
        assert st._carry_hidden(frame([2, 4, 1], [True, False, False])) is None
        assert st._carry_hidden(frame([2, 2, 1], [False, False, False])) is None
