"""GPU (-m gpu): GraphVS.forward through the C ABI vs the CPU oracle and the golden fixtures.

Tolerance: velocities / hidden features within rel 1e-4 (fp32 path) of the fp64 oracle, where
"rel" is max|got - ref| / max|ref| per tensor (north_star: "rel 1e-4 in fp32")."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import cns_oracle
from cns_b200 import Batch, GraphGenerator, GraphVS, _lib, ckpt, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
# 16-bit modes: SURVEY 8c's reduced-precision bar is 2e-2; f16 is held to half of it (typical 1.3e-3 on velocities /
# 3.4e-3 on the hidden state, worst case 8.5e-3 over the adversarial mixes of scripts/fuzz_forward.py)
TOL = {"fp32": 1e-4, "f16": 1e-2, "bf16": 2e-2}


def rel(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp(min=1e-12))


@pytest.fixture(scope="module")
def net():
    return ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), "cuda:0")


def _check(net, batch, hid_g, hid_o, sd, tol=1e-4):
    ov, on, oh = cns_oracle.forward(sd, batch, hid_o, dtype=torch.float64)
    with torch.no_grad():
        pred = net(batch.to("cuda:0"), hid_g, check=True)
    assert rel(pred[0], ov) < tol and rel(pred[1], on) < tol and rel(pred[2], oh) < tol
    ovel = cns_oracle.postprocess((ov, on), batch.tPo_norm.reshape(-1))
    assert rel(pred.vel, ovel) < tol
    assert rel(net.postprocess(pred, batch.to("cuda:0")), ovel) < tol
    return pred, oh


@pytest.mark.parametrize("name", ["forward_single.pt", "forward_ragged.pt"])
def test_golden_fixtures(net, name):
    """Inputs and outputs recorded from the unmodified reference (tests/golden/make_golden.py)."""
    case = torch.load(os.path.join(GOLD, name), weights_only=False)
    hidden = None
    for st in case["steps"]:
        data = {k: v.to("cuda:0") for k, v in st["fields"].items()}
        with torch.no_grad():
            pred = net(data, hidden, check=True)
        hidden = pred[2]
        assert rel(pred[0], st["vel_si_vec"]) < 1e-4
        assert rel(pred[1], st["vel_si_norm"]) < 1e-4
        assert rel(pred[2], st["hidden"]) < 1e-4
        assert rel(pred.vel, st["vel"]) < 1e-4


@pytest.mark.parametrize("sizes,missing", [([40], 0.0), ([5, 16, 17, 64], 0.2), ([512] * 6, 0.1),
                                           ([24, 100, 300, 512, 33], 0.3), ([2000], 0.05), ([1500, 2048, 900], 0.1)])
def test_forward_matches_oracle_recurrent(net, state_dict, sizes, missing):
    """Ragged batches, missing keypoints, clusters larger than a 64-edge softmax piece (N=2000),
    hidden state carried over 3 frames with the camera moved by the predicted velocity."""
    rng = np.random.default_rng(len(sizes))
    np.random.seed(len(sizes))
    scenes = [synth.Scene(rng, n) for n in sizes]
    hid_g = hid_o = None
    for _ in range(3):
        frames = [s.frame(missing) for s in scenes]
        batch = Batch.from_data_list(frames) if len(frames) > 1 else frames[0]
        pred, hid_o = _check(net, batch, hid_g, hid_o, state_dict)
        hid_g = pred[2]
        for s, v in zip(scenes, pred.vel.cpu().numpy()):
            s.step(v * 2)


@pytest.mark.parametrize("prec", ["f16", "bf16"])
@pytest.mark.parametrize("sizes,missing", [([24, 100, 300, 512, 33], 0.3), ([512] * 6, 0.1), ([7], 0.0), ([2000], 0.05),
                                           ([1500, 2048, 900], 0.1)])     # graphs of 30-45 clusters: few graphs per backbone tile,
                                                                          # more than 32 rows per graph in the pooling kernel
def test_tensor_core_modes_match_oracle(state_dict, prec, sizes, missing):
    """tcgen05 paths (encoder edge kernel + node GEMMs with 16-bit operands, fp32 accumulation in
    TMEM): 5 recurrent frames, hidden carried, within the per-dtype tolerance of the fp64 oracle."""
    net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), "cuda:0", precision=prec)
    rng = np.random.default_rng(len(sizes) + 100)
    np.random.seed(len(sizes) + 100)
    scenes = [synth.Scene(rng, n) for n in sizes]
    hid_g = hid_o = None
    for _ in range(5):
        frames = [s.frame(missing) for s in scenes]
        batch = Batch.from_data_list(frames) if len(frames) > 1 else frames[0]
        pred, hid_o = _check(net, batch, hid_g, hid_o, state_dict, tol=TOL[prec])
        hid_g = pred[2]
        for s, v in zip(scenes, pred.vel.cpu().numpy()):
            s.step(v * 2)


def test_64_edge_tile_variant(state_dict, monkeypatch):
    """CNS_TC_TILE=64: two 64-edge slots per SM instead of four 32-edge ones (MN-major SWIZZLE_128B operand tile)."""
    monkeypatch.setenv("CNS_TC_TILE", "64")
    net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), "cuda:0", precision="f16")
    batch = synth.make_batch(5, 300, seed=13, missing_frac=0.25, n_layouts=5)
    _check(net, batch, None, None, state_dict, tol=TOL["f16"])


def test_all_clusters_of_one_graph_masked(net, state_dict):
    """A graph whose every keypoint is missing: empty segments must give zeros, not NaNs."""
    rng = np.random.default_rng(9)
    np.random.seed(9)
    a, b = synth.Scene(rng, 30), synth.Scene(rng, 50)
    fa = a.generator.get_data(synth.project_norm(a.cur_wcT, a.points), missing_node_indices=np.arange(30))
    fb = b.frame(0.5)
    batch = Batch.from_data_list([fb, fa])      # the fully masked graph is last (PyG would drop its row)
    with torch.no_grad():
        pred = net(batch.to("cuda:0"), None, check=True)
    assert torch.isfinite(pred[0]).all() and torch.isfinite(pred[2]).all()
    ov, on, oh = cns_oracle.forward(state_dict, Batch.from_data_list([fb]), None, dtype=torch.float64)
    assert rel(pred[0][:1], ov) < 1e-4 and rel(pred[2][:oh.shape[0]], oh) < 1e-4


def test_deterministic_and_batch_invariant(net):
    """Same input twice -> bit-identical output (no fp atomics); a graph's result does not depend on
    its neighbours in the batch beyond fp32 re-association of the softmax pieces."""
    batch = synth.make_batch(16, 200, seed=5, n_layouts=4)
    dev = batch.to("cuda:0")
    with torch.no_grad():
        p1 = net(dev, None)
        p2 = net(dev, None)
    assert torch.equal(p1[0], p2[0]) and torch.equal(p1[2], p2[2])
    scenes = synth.make_scenes(16, 200, seed=5, n_layouts=4)
    frames = [s.frame(0.1) for s in scenes]
    with torch.no_grad():
        full = net(Batch.from_data_list(frames).to("cuda:0"), None)
        solo = net(Batch.from_data_list(frames[5:6]).to("cuda:0"), None)
    assert rel(solo[0], full[0][5:6]) < 1e-5


def test_complete_graph_path_matches_general_path_and_is_verified(net):
    """Batches built by GraphGenerator carry l1_complete (cns_batch.flags): the forward then replaces the
    per-row neighbour gathers by one column max per graph.  Same result as the general CSR path, and a
    list that is not the complete digraph it claims to be is reported by check=True."""
    batch = synth.make_batch(6, 200, seed=11, missing_frac=0.4, n_layouts=3)
    assert batch.l1_complete is True
    dev = batch.to("cuda:0")
    hid = torch.randn(dev.cluster_centers_index.shape[0], 128, generator=torch.Generator().manual_seed(3)).to("cuda:0")
    with torch.no_grad():
        a = net(dev, hid, check=True)
        dev.l1_complete = False
        b = net(dev, hid, check=True)
    for x, y in zip(a, b):
        assert rel(x, y) < 1e-5      # same arithmetic up to the fp32 rounding of GraphNorm statistics, seen through the (hi, lo) operand split
    # an edge list assigned after construction voids the hint by itself (custom edge augmentation): the forward
    # then takes the general CSR path and computes what the edge list says
    cut = batch.to("cuda:0")
    cut.l1_dense_edge_index_cur = cut.l1_dense_edge_index_cur[:, 1:].contiguous()      # one edge short
    assert cut.l1_complete is False
    with torch.no_grad():
        c = net(cut, hid, check=True)
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    ov, on, oh = cns_oracle.forward(sd, cut.to("cpu"), hid.cpu(), dtype=torch.float64)
    assert rel(c[0], ov) < 1e-4 and rel(c[2], oh) < 1e-4
    # a hint that is claimed but wrong is reported
    bad = batch.to("cuda:0")
    bad.l1_dense_edge_index_cur = bad.l1_dense_edge_index_cur[:, 1:].contiguous()
    bad.l1_complete = True
    with pytest.raises(_lib.CnsError):
        with torch.no_grad():
            net(bad, None, check=True)
    bad = batch.to("cuda:0")
    e = bad.l1_dense_edge_index_tar.clone()
    e[1, 5] = e[1, 6]                                                                  # right count, wrong pair
    bad.l1_dense_edge_index_tar = e
    bad.l1_complete = True
    with pytest.raises(_lib.CnsError):
        with torch.no_grad():
            net(bad, None, check=True)


def test_unsorted_cluster_edges_are_reported(net):
    d = synth.Scene(np.random.default_rng(1), 40).frame().to("cuda:0")
    d.l0_to_l1_edge_index_i_cur = d.l0_to_l1_edge_index_i_cur.flip(0).contiguous()
    with pytest.raises(_lib.CnsError):
        with torch.no_grad():
            net(d, None, check=True)


def test_cpu_tensors_rejected(net):
    d = synth.Scene(np.random.default_rng(1), 40).frame()
    with pytest.raises(_lib.CnsError):
        net(d, None)


def test_host_pointer_entry_matches_device_entry(net):
    """cns_graphvs_forward_host (host buffers in, host buffers out) == device-pointer call."""
    lib = _lib.load()
    batch = synth.make_batch(8, 128, seed=7, n_layouts=4)
    with torch.no_grad():
        ref = net(batch.to("cuda:0"), None)
    torch.cuda.synchronize()
    b, keep = GraphVS.make_batch_struct(batch)      # CPU tensors -> host pointers
    nb, nc = b.n_graphs, b.n_clusters
    vec, nrm = torch.empty(nb, 6), torch.empty(nb, 1)
    hid, vel = torch.empty(nc, 128), torch.empty(nb, 6)
    h = net._handle(torch.device("cuda:0"))
    _lib.check(lib.cns_graphvs_forward_host(h.raw, C.byref(b), None, _lib.ptr(vec), _lib.ptr(nrm),
                                            _lib.ptr(hid), _lib.ptr(vel)), "forward_host")
    assert torch.equal(vec, ref[0].cpu()) and torch.equal(hid, ref[2].cpu()) and torch.equal(vel, ref.vel.cpu())


def test_servo_controller_host_path(state_dict):
    """GraphVSController on host GraphData (cns_servo_step_host: one staged copy in, hidden kept on the
    device) reproduces the oracle's recurrent velocities, including the new-scene reset."""
    from cns_b200.controller import GraphVSController
    ctl = GraphVSController(os.path.join(ROOT, "checkpoints", "cns_state_dict.pth"), "cuda:0")
    rng = np.random.default_rng(21)
    np.random.seed(21)
    scene = synth.Scene(rng, 150)
    hid = None
    for t in range(6):
        d = scene.frame(0.1)
        if t in (0, 4):
            d.start_new_scene()
            hid = None
        vel = ctl(d)
        ov, on, hid = cns_oracle.forward(state_dict, d, hid, dtype=torch.float64)
        ovel = cns_oracle.postprocess((ov, on), d.tPo_norm.reshape(1))[0].numpy()
        assert vel.shape == (6,) and vel.dtype == np.float32
        assert np.abs(vel - ovel).max() < 1e-4 * np.abs(ovel).max()
        scene.step(vel * 2)


def test_servo_controller_host_path_varying_visibility(state_dict):
    """The carried hidden state must survive frames whose staged input block changes size: the number of
    visible keypoints crosses multiples of 32 (the staging alignment), whole clusters lose all their
    keypoints (fewer cluster-level edges) and come back; every frame is checked against the oracle."""
    from cns_b200.controller import GraphVSController
    ctl = GraphVSController(os.path.join(ROOT, "checkpoints", "cns_state_dict.pth"), "cuda:0")
    rng = np.random.default_rng(33)
    np.random.seed(33)
    scene = synth.Scene(rng, 150)
    gen = scene.generator
    labels = np.asarray(gen.belong_cluster_index)
    n = len(scene.points)
    hid = None
    miss_counts = [0, 15, 31, 33, 64, 1, 97, 40, 0, 70]
    for t, k in enumerate(miss_counts):
        cur = synth.project_norm(scene.cur_wcT, scene.points)
        missing = set(rng.choice(n, size=k, replace=False).tolist()) if k else set()
        if t in (3, 4, 6):      # empty cluster (t % number of clusters) as well
            missing |= set(np.nonzero(labels == (t % gen.num_clusters))[0].tolist())
        missing = np.array(sorted(missing), dtype=np.int64)
        d = gen.get_data(cur, missing_node_indices=missing if len(missing) else None)
        d.set_distance_scale(scene.tPo_norm)
        if t == 0:
            d.start_new_scene()
        if t in (3, 4, 6):
            assert not bool(d.cluster_mask.all())
        vel = ctl(d)
        ov, on, hid = cns_oracle.forward(state_dict, d, hid, dtype=torch.float64)
        ovel = cns_oracle.postprocess((ov, on), d.tPo_norm.reshape(1))[0].numpy()
        assert np.abs(vel - ovel).max() < 1e-4 * np.abs(ovel).max(), "frame %d" % t
        scene.step(vel * 2)


def test_batched_servo_device_graphs_match_host_graph_path(net):
    """BatchedServo (target graphs resident on the device, per-frame masking by cns_graph_mask, hidden
    carried on the device) == GraphGenerator.get_data + Batch.from_data_list + GraphVS.forward on the
    same frames: identical graphs feed identical kernels, so the velocities are bit-equal."""
    from cns_b200.servo import BatchedServo
    rng = np.random.default_rng(5)
    np.random.seed(5)
    scenes = [synth.Scene(rng, n) for n in (30, 200, 64, 512, 12)]
    servo = BatchedServo(net, [s.generator for s in scenes], [s.tPo_norm for s in scenes], "cuda:0")
    hidden = None
    for t in range(4):
        frames, xy, vis = [], [], []
        for s in scenes:
            n = len(s.points)
            cur = synth.project_norm(s.cur_wcT, s.points)
            k = int(rng.integers(0, n))                      # any number of unmatched keypoints, incl. 0
            missing = np.sort(rng.choice(n, size=k, replace=False)) if k else None
            d = s.generator.get_data(cur, missing_node_indices=missing)
            d.set_distance_scale(s.tPo_norm)
            v = np.ones(n, bool)
            if k:
                v[missing] = False
            frames.append(d); xy.append(cur); vis.append(v)
        batch = Batch.from_data_list(frames)
        with torch.no_grad():
            ref = net(batch.to("cuda:0"), hidden, check=True)
        hidden = ref[2]
        vel = servo.step(torch.from_numpy(np.concatenate(xy)).float(), torch.from_numpy(np.concatenate(vis)))
        assert torch.equal(vel, ref.vel.cpu())
        assert torch.equal(servo.hidden[servo.cur], hidden)
        for s, v in zip(scenes, vel.numpy()):
            s.step(v * 2)


def test_target_cache_is_tied_to_weights_and_shapes(net):
    """cns_target_cache: same results as recomputing the target side, rejected once the weights change."""
    from cns_b200.servo import BatchedServo
    rng = np.random.default_rng(8)
    np.random.seed(8)
    scenes = [synth.Scene(rng, n) for n in (40, 90)]
    servo = BatchedServo(net, [s.generator for s in scenes], [s.tPo_norm for s in scenes], "cuda:0")
    xy = torch.from_numpy(np.concatenate([synth.project_norm(s.cur_wcT, s.points) for s in scenes])).float()
    vis = torch.ones(xy.shape[0], dtype=torch.bool)
    v1 = servo.step(xy, vis).clone()
    servo.reset()
    servo.drop_target_cache()
    v2 = servo.step(xy, vis).clone()
    assert torch.equal(v1, v2)
    sd = net.state_dict()
    net.load_state_dict(sd)                      # re-packs the weights: the handle's weights_version moves on
    servo.reset()
    with pytest.raises(_lib.CnsError):
        servo.step(xy, vis)
    servo.drop_target_cache()
    servo.reset()
    assert torch.equal(servo.step(xy, vis), v1)


@pytest.mark.parametrize("depth", [2, 3])
def test_servo_frames_in_flight_keep_their_order(net, depth):
    """submit/result with `depth` frames in flight (every slot on its own compute stream): results equal the one-at-a-time
    loop, hidden state carried from slot to slot."""
    from cns_b200.servo import BatchedServo
    rng = np.random.default_rng(12)
    np.random.seed(12)
    scenes = [synth.Scene(rng, n) for n in (64, 33, 128)]
    gens, tpo = [s.generator for s in scenes], [s.tPo_norm for s in scenes]
    frames = []
    for t in range(5):
        xy = torch.from_numpy(np.concatenate([synth.project_norm(s.cur_wcT, s.points) for s in scenes])).float().pin_memory()
        vis = (torch.rand(xy.shape[0], generator=torch.Generator().manual_seed(t)) > 0.2).pin_memory()
        frames.append((xy, vis))
        for s in scenes:
            s.step(synth.pbvs_straight(s.cur_wcT, s.tar_wcT) * 2)
    a = BatchedServo(net, gens, tpo, "cuda:0")
    one = [a.step(*f).clone() for f in frames]
    b = BatchedServo(net, gens, tpo, "cuda:0", depth=depth)
    tickets, two = [], []
    for i, f in enumerate(frames):
        tickets.append(b.submit(*f))
        if i >= depth - 1:
            two.append(b.result(tickets[i - depth + 1]).clone())
    for t in tickets[len(two):]:
        two.append(b.result(t).clone())
    for x, y in zip(one, two):
        assert torch.equal(x, y)


def test_device_scene_batch_matches_host_graph_path():
    """DeviceSceneBatch.observe(): the same Batch the host path builds for the same poses and
    visibility - graph tensors bit-exact, coordinates / labels to fp32 rounding - over a few closed-loop steps."""
    from cns_b200.sim_device import DeviceSceneBatch
    env = DeviceSceneBatch(6, np.random.default_rng(17), "cuda:0", max_points=200)
    for step in range(3):
        poses = env.cur_wcT.cpu().numpy()
        d = env.observe()
        vis = d.node_mask.cpu().numpy()
        frames, o = [], 0
        for k, s in enumerate(env.scenes):
            n = len(s.points)
            missing = np.nonzero(~vis[o:o + n])[0]
            f = s.generator.get_data(synth.project_norm(poses[k], s.points), missing_node_indices=missing if len(missing) else None)
            f.set_distance_scale(s.tPo_norm)
            frames.append(f); o += n
        ref = Batch.from_data_list(frames)
        for key in ("l0_to_l1_edge_index_j_cur", "l0_to_l1_edge_index_i_cur", "l1_dense_edge_index_cur",
                    "l1_dense_edge_index_tar", "cluster_mask", "cluster_centers_index", "num_clusters", "node_mask"):
            assert torch.equal(d[key].cpu(), ref[key]), key
        assert float((d.x_cur.cpu() - ref.x_cur).abs().max()) < 1e-6 and torch.equal(d.x_tar.cpu(), ref.x_tar)
        want = np.stack([synth.pbvs_straight(poses[k], s.tar_wcT) for k, s in enumerate(env.scenes)])
        assert float((d.vel.cpu() - torch.from_numpy(want).float()).abs().max()) < 1e-6
        env.feedback(d.vel * 2)
        nxt = np.stack([synth.integrate(poses[k], want[k] * 2) for k in range(len(env.scenes))])
        fresh = env.fresh.cpu().numpy()
        assert np.abs(env.cur_wcT.cpu().numpy()[~fresh] - nxt[~fresh]).max() < 1e-6


def test_full_size_batch_properties(net, state_dict):
    """BASELINE config: 1024 graphs x 512 keypoints.  The oracle is too slow for all of it, so:
    (1) spot-check 8 graphs against the oracle, (2) permuting the graphs of the batch permutes the
    outputs, (3) finiteness and hidden-state shape."""
    scenes = synth.make_scenes(1024, 512, seed=2022, n_layouts=16)
    frames = [s.frame(0.1) for s in scenes]
    batch = Batch.from_data_list(frames)
    with torch.no_grad():
        pred = net(batch.to("cuda:0"), None, check=True)
    assert pred[0].shape == (1024, 6) and pred[2].shape == (int(batch.num_clusters.sum()), 128)
    assert torch.isfinite(pred[0]).all() and torch.isfinite(pred[2]).all()
    pick = [0, 1, 17, 255, 256, 700, 1022, 1023]
    ov, on, oh = cns_oracle.forward(state_dict, Batch.from_data_list([frames[i] for i in pick]), None,
                                    dtype=torch.float64)
    assert rel(pred[0][pick], ov) < 1e-4 and rel(pred[1][pick], on) < 1e-4
    perm = np.random.default_rng(0).permutation(1024)
    with torch.no_grad():
        pp = net(Batch.from_data_list([frames[i] for i in perm]).to("cuda:0"), None)
    assert rel(pp[0], pred[0][torch.as_tensor(perm, device="cuda")]) < 1e-5


def test_split_operand_encoder_vs_ffma_encoder(state_dict, monkeypatch):
    """fp32 mode runs the encoder edge chain on tcgen05 with bf16 (hi, lo) operand splits (encoder_split.cu); the
    FFMA kernel it replaces is still there behind CNS_NO_SPLIT_ENCODER=1.  Both must sit within 1e-4 of the fp64
    oracle on ragged batches with missing keypoints over recurrent frames, and close to each other."""
    path = os.path.join(ROOT, "checkpoints", "cns.pth")
    tc = ckpt.load_model(path, "cuda:0")
    monkeypatch.setenv("CNS_NO_SPLIT_ENCODER", "1")
    ffma = ckpt.load_model(path, "cuda:0")
    ffma._handle(torch.device("cuda:0"))            # the switch is read when the handle is created
    monkeypatch.delenv("CNS_NO_SPLIT_ENCODER")
    rng = np.random.default_rng(12)
    np.random.seed(12)
    scenes = [synth.Scene(rng, n) for n in (4, 17, 33, 100, 512, 700, 64, 31)]
    h_tc = h_ff = h_o = None
    for frame in range(3):
        frames = [s.frame(float(rng.uniform(0, 0.6))) for s in scenes]
        batch = Batch.from_data_list(frames)
        ov, on, h_o = cns_oracle.forward(state_dict, batch, h_o, dtype=torch.float64)
        with torch.no_grad():
            a = tc(batch.to("cuda:0"), h_tc, check=True)
            b = ffma(batch.to("cuda:0"), h_ff, check=True)
        h_tc, h_ff = a[2], b[2]
        for got in (a, b):
            assert rel(got[0], ov) < 1e-4 and rel(got[1], on) < 1e-4 and rel(got[2], h_o) < 1e-4
        assert rel(a[2], b[2]) < 5e-5 and rel(a[0], b[0]) < 5e-5
        print("frame %d: split %.2e / %.2e   ffma %.2e / %.2e (vel / hidden vs fp64)" % (
            frame, rel(a[0], ov), rel(a[2], h_o), rel(b[0], ov), rel(b[2], h_o)))
        for s, v in zip(scenes, a.vel.cpu().numpy()):
            s.step(v * 2)


@pytest.mark.parametrize("precision", ["f16", "bf16"])
def test_full_size_batch_16bit_modes(state_dict, precision):
    """The headline shape in the headline mode: 1024 graphs x 512 keypoints through the tcgen05 encoder kernel and
    (f16, complete cluster graphs) the fused backbone kernel - the path bench.py times.  64 graphs spread over the
    batch are checked against the fp64 oracle, with the error of the SIX VELOCITY COMPONENTS stated element-wise:
    every component of the post-processed velocity within  |got - ref| <= atol + rtol * |ref|,
    atol = 2 % of that graph's largest velocity component, rtol = 2e-2 (the reduced-precision bar),
    and the per-tensor figure max|d| / max|ref| within TOL[precision]; the second recurrent frame (hidden carried)
    is held to the same."""
    net16 = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), "cuda:0", precision=precision)
    scenes = synth.make_scenes(1024, 512, seed=2022, n_layouts=16)
    pick = list(range(0, 1024, 16))
    hid_g = hid_o = None
    worst = 0.0
    for frame in range(2):
        frames = [s.frame(0.1) for s in scenes]
        batch = Batch.from_data_list(frames)
        assert batch.l1_complete is True and batch.max_clusters_per_graph <= 128       # fused-backbone preconditions
        with torch.no_grad():
            pred = net16(batch.to("cuda:0"), hid_g, check=True)
        hid_g = pred[2]
        assert torch.isfinite(pred[0]).all() and torch.isfinite(pred[2]).all() and torch.isfinite(pred.vel).all()
        sub = Batch.from_data_list([frames[i] for i in pick])
        # oracle hidden rows of the picked graphs only (graphs are independent)
        ov, on, hid_o = cns_oracle.forward(state_dict, sub, hid_o, dtype=torch.float64)
        ovel = cns_oracle.postprocess((ov, on), sub.tPo_norm.reshape(-1))
        got = pred.vel[pick].double().cpu()
        assert rel(got, ovel) < TOL[precision] and rel(pred[0][pick], ov) < TOL[precision]
        scale = ovel.abs().max(dim=1, keepdim=True).values
        bound = 0.02 * scale + 2e-2 * ovel.abs()
        excess = ((got - ovel).abs() / bound).max()
        worst = max(worst, float(excess))
        assert float(excess) <= 1.0, "frame %d: a velocity component is off by %.2fx its element-wise bound" % (frame, float(excess))
        cptr = torch.cumsum(batch.num_clusters, 0) - batch.num_clusters
        rows = torch.cat([torch.arange(int(cptr[i]), int(cptr[i] + batch.num_clusters[i])) for i in pick])
        assert rel(hid_g[rows.to("cuda:0")], hid_o) < TOL[precision]
        for s, v in zip(scenes, pred.vel.cpu().numpy()):
            s.step(v * 2)
    print("%s: worst element-wise velocity error = %.3f of the bound" % (precision, worst))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_in_the_same_process(state_dict):
    """Kernel attributes (dynamic shared memory opt-in) are per device: a model on cuda:1 after one on cuda:0, all
    precision modes and a training step, must give the same results."""
    batch = synth.make_batch(12, 150, seed=9, n_layouts=3)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        net = GraphVS(2, 2, 128).to(dev)
        net.load_state_dict(state_dict)
        net.eval()
        res = []
        for prec in ("fp32", "f16", "bf16"):
            net.set_precision(prec)
            with torch.no_grad():
                p = net(batch.to(dev), None)
            res.append(p.vel.cpu())
        net.set_precision("fp32")
        net.train()
        pred = net(batch.to(dev), None)
        (pred[0].square().sum() + pred[2].square().sum()).backward()
        res.append(torch.cat([q.grad.flatten().cpu() for q in net.parameters()]))
        outs.append(res)
    for k, (a, b) in enumerate(zip(*outs)):
        assert torch.isfinite(b).all(), k
        assert torch.equal(a, b), (k, float((a - b).abs().max()), float(a.abs().max()))


@pytest.mark.parametrize("long_seq", [False, True])
def test_device_scene_batch_reference_augmentations(net, long_seq):
    """DeviceSceneBatch(aug=True): the reference's training distribution on the device.  Invariants of
    sim/dataset.py:88-141 / dataset_long.py:81-140 per frame, graph tensors bit-exact with the host generator for the
    same visibility, labels equal to the reference-checked batched supervisors, and the loaders' "re-initialise all
    scenes once more than half need it" rule."""
    from cns_b200 import sim_aug
    from cns_b200.sim_device import DeviceSceneBatch
    env = DeviceSceneBatch(12, np.random.default_rng(23), "cuda:0", max_points=300, aug=True, long_seq=long_seq, reinit="all",
                           supervisor="pbvs_hybrid" if long_seq else "pbvs_straight", discard=True)
    n = env.n_nodes.cpu().numpy()
    scene_of = env.scene_of.cpu().numpy()
    removed_seen = mism_seen = 0
    for step in range(6):
        poses = env.cur_wcT.clone()
        d = env.observe()
        vis = d.node_mask.cpu().numpy()
        mism = d.mismatch_mask.cpu().numpy()
        xy = d.x_cur.cpu().numpy()
        clean = synth_project(env, poses)
        frames, o = [], 0
        for k, s in enumerate(env.scenes):
            m = n[k]
            v, mm = vis[o:o + m], mism[o:o + m]
            assert (~v).sum() < max(0.8 * m, 1)                                   # the 80 % cap
            assert mm.sum() == (int(m * 0.1) if m >= 10 else 0)
            if m < 10:
                assert v.all()
            assert not xy[o:o + m][~v].any()                                      # removed keypoints are zeroed
            ok = v & ~mm
            jit = 0.0005 if long_seq else 0.002
            assert np.abs(xy[o:o + m][ok] - clean[o:o + m][ok]).max() <= jit + 1e-6   # jitter only
            missing = np.nonzero(~v)[0]
            f = s.generator.get_data(xy[o:o + m], missing_node_indices=missing if len(missing) else None)
            frames.append(f)
            removed_seen += (~v).sum(); mism_seen += mm.sum()
            o += m
        ref = Batch.from_data_list(frames)
        for key in ("l0_to_l1_edge_index_j_cur", "l0_to_l1_edge_index_i_cur", "l1_dense_edge_index_cur", "cluster_mask"):
            assert torch.equal(d[key].cpu(), ref[key]), key
        # labels: supervisor on the visible keypoints' centre (N < 64) - the batched functions are checked against the
        # reference in tests/test_sim_aug_cpu.py
        small = torch.from_numpy(n < 64).to("cuda:0")
        v_t = d.node_mask
        wPo = sim_aug.segment_mean(env.points[v_t], env.scene_of[v_t], env.B)
        vel, tpo, vel_si = sim_aug.supervisor_vel(env.supervisor, poses, env.tar_wcT, wPo)
        assert float((d.vel - vel.float())[small].abs().max()) < 1e-6 and float((d.tPo_norm - tpo.float())[small].abs().max()) < 1e-6
        assert float((d.vel_si[:, :3] * (d.tPo_norm[:, None] + 1e-7) - d.vel[:, :3]).abs().max()) < 1e-5
        with torch.no_grad():
            pred = net(d, None, check=True)
        assert torch.isfinite(pred.vel).all()
        env.feedback(d.vel * 2)
    assert removed_seen > 0 and mism_seen > 0
    # more than half of the scenes past their step budget: everything is re-initialised together, with new geometry
    before = env.points.clone()
    env.steps[:8] = env.max_steps + 1
    env.feedback(torch.zeros(env.B, 6, device="cuda:0"))
    assert env.reinit_all_count == 1 and bool(env.fresh.all()) and int(env.steps.max()) == 0
    assert env.points.shape != before.shape or not torch.equal(env.points, before)
    assert bool(env.observe().new_scene.all())


def synth_project(env, poses):
    from cns_b200.sim_device import project_norm
    return project_norm(poses, env.points, env.scene_of).float().cpu().numpy()


@pytest.mark.parametrize("precision", ["fp32", "f16"])
def test_cuda_graph_replay_equals_plain_enqueue(state_dict, precision):
    """A forward whose whole argument set repeats (same sizes, pointers, mode) is captured the second time and replayed
    as one CUDA-graph launch from the third: same bits as the plain enqueue, inputs re-read on every replay."""
    net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), "cuda:0", precision=precision)
    batch = synth.make_batch(5, 150, seed=4, missing_frac=0.2, n_layouts=3).to("cuda:0")
    h = net._handle(torch.device("cuda:0"))
    lib = _lib.load()
    b, keep = net.make_batch_struct(batch)
    nb, nc = b.n_graphs, b.n_clusters
    ws = torch.empty(lib.cns_forward_workspace_bytes(C.byref(b)) + 1024, dtype=torch.uint8, device="cuda:0")
    out = torch.zeros(nc * 128 + nb * 13, device="cuda:0")
    hid, vec = out[:nc * 128], out[nc * 128:nc * 128 + nb * 6]
    vel, nrm = out[nc * 128 + nb * 6:nc * 128 + nb * 12], out[nc * 128 + nb * 12:]

    def call(stream):
        _lib.check(lib.cns_graphvs_forward(h.raw, C.byref(b), None, _lib.ptr(vec), _lib.ptr(nrm), _lib.ptr(hid), _lib.ptr(vel),
                                           _lib.ptr(ws), ws.numel(), None, 0, C.c_void_p(stream)), "forward")
        _lib.check(lib.cns_forward_status(_lib.ptr(ws), C.c_void_p(stream)), "status")
        return out.clone()

    side = torch.cuda.Stream("cuda:0")
    for stream in (0, side.cuda_stream):        # the legacy default stream (captured on an internal stream) and a real one
        h.set_option("cuda_graphs", 0)
        ref = call(stream)
        h.set_option("cuda_graphs", 1)
        r0, c0 = h.counter("graph_replays"), h.counter("graph_captures")
        got = [call(stream) for _ in range(4)]
        assert h.counter("graph_captures") == c0 + 1 and h.counter("graph_replays") == r0 + 3
        for g in got:
            assert torch.equal(g, ref)
        # new input values behind the same pointers: a replay reads them
        batch.pos_cur.mul_(1.01)
        batch.x_cur.copy_(batch.pos_cur)
        moved = call(stream)
        h.set_option("cuda_graphs", 0)
        assert torch.equal(moved, call(stream)) and not torch.equal(moved, ref)
        batch.pos_cur.div_(1.01)
        batch.x_cur.copy_(batch.pos_cur)
    h.set_option("cuda_graphs", 1)
    with pytest.raises(_lib.CnsError):
        h.set_option("no_such_option", 1)


def test_forwards_on_two_streams_do_not_interfere(state_dict):
    """GraphVS.forward keeps one workspace per CUDA stream: forwards enqueued on different streams may overlap on the
    device (bench.py alternates two streams) and still give the results of the one-after-the-other loop."""
    net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), "cuda:0", precision="f16")
    batches = [synth.make_batch(24, 256, seed=20 + i, missing_frac=0.1, n_layouts=4).to("cuda:0") for i in range(4)]
    with torch.no_grad():
        ref = [[t.clone() for t in net(b, None)] for b in batches]
        torch.cuda.synchronize()
        streams = [torch.cuda.Stream("cuda:0"), torch.cuda.Stream("cuda:0")]
        for st in streams:
            st.wait_stream(torch.cuda.current_stream())
        got = []
        for rep in range(3):                       # the third pass replays CUDA graphs
            got = []
            for i, b in enumerate(batches):
                with torch.cuda.stream(streams[i % 2]):
                    got.append(net(b, None))
        torch.cuda.synchronize()
    for g, r in zip(got, ref):
        for x, y in zip(g, r):
            assert torch.equal(x, y)


@pytest.mark.parametrize("precision", ["f16"])
def test_target_edge_cache_matches_recomputed_target_side(state_dict, precision):
    """A target cache built with every keypoint's cluster keeps the target side's per-keypoint attention logits (fp32) and
    values (f16); a forward then runs the edge kernel on the current side only and a streaming softmax over the visible
    keypoints for the target side.  Same result as recomputing the target side, up to the f16 rounding of the cached values,
    and still inside the mode's tolerance of the fp64 oracle."""
    net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), "cuda:0", precision=precision)
    h = net._handle(torch.device("cuda:0"))
    rng = np.random.default_rng(21)
    np.random.seed(21)
    scenes = [synth.Scene(rng, n) for n in (300, 40, 512, 17, 128)]
    first = Batch.from_data_list([s.frame(0.0) for s in scenes]).to("cuda:0")
    cache = net.build_target_cache(first)                     # belongs to the targets, not to a frame
    hid_g = hid_c = hid_o = None
    # 30 % / none / 70 % / 10 % of the keypoints missing, different ones every frame: a cluster that keeps most of its
    # keypoints subtracts the missing ones from its cached totals, the others re-sum the visible ones
    for t, missing in enumerate((0.3, 0.0, 0.7, 0.1)):
        frames = [s.frame(missing) for s in scenes]
        batch = Batch.from_data_list(frames)
        dev = batch.to("cuda:0")
        n0 = h.counter("edge_cache_forwards")
        with torch.no_grad():
            plain = net(dev, hid_g, check=True)
            cached = net(dev, hid_c, check=True, target_cache=cache)
        assert h.counter("edge_cache_forwards") == n0 + 1
        hid_g, hid_c = plain[2], cached[2]
        for a, b in zip(plain, cached):
            assert rel(b, a) < 3e-3
        ov, on, hid_o = cns_oracle.forward(state_dict, batch, hid_o, dtype=torch.float64)
        assert rel(cached[0], ov) < TOL[precision] and rel(cached[2], hid_o) < TOL[precision]
        for s, v in zip(scenes, cached.vel.cpu().numpy()):
            s.step(v * 2)


@pytest.mark.parametrize("precision", ["fp32", "f16"])
def test_graphs_with_more_than_128_clusters(state_dict, precision):
    """Clusters given as Voronoi cells of 129 / 200 exemplar keypoints (GraphGenerator.from_exemplars): more cluster rows
    per graph than a whole-graph tile of the fused backbone holds (128), next to an ordinary graph - the batch takes the
    per-layer cluster kernels; two recurrent frames against the oracle."""
    rng = np.random.default_rng(41)
    np.random.seed(41)
    scenes = [synth.Scene(rng, 120)]
    for n, k in ((600, 129), (900, 200)):
        pts, tar = synth.sample_points(rng, n), synth.sample_target_pose(rng)
        gen = GraphGenerator.from_exemplars(synth.project_norm(tar, pts), rng.choice(n, size=k, replace=False))
        assert gen.num_clusters == k
        scenes.append(synth.Scene(rng, n, generator=gen, points=pts, tar_wcT=tar))
    net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), "cuda:0", precision=precision)
    hg = ho = None
    for _ in range(2):
        batch = Batch.from_data_list([s.frame(0.1) for s in scenes])
        pred, ho = _check(net, batch, hg, ho, state_dict, tol=TOL[precision])
        hg = pred[2]
        for s in scenes:
            s.step(synth.pbvs_straight(s.cur_wcT, s.tar_wcT) * 2)


def test_graph_cache_eviction_with_forwards_in_flight(state_dict):
    """More distinct argument sets than the graph cache holds (48), cycled several times over two streams without a host
    synchronisation in between: entries are evicted (their graphs destroyed) while replays of other - and possibly the same -
    entries are still executing.  Every result must equal the plain enqueue of the same batch."""
    net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), "cuda:0", precision="f16")
    h = net._handle(torch.device("cuda:0"))
    n_sets = 60
    lib = _lib.load()
    rng = np.random.default_rng(3)
    batches = [synth.make_batch(int(rng.integers(2, 9)), int(rng.choice([40, 96, 150])), seed=100 + i, missing_frac=0.1,
                                n_layouts=2).to("cuda:0") for i in range(n_sets)]
    streams = [torch.cuda.Stream("cuda:0"), torch.cuda.Stream("cuda:0")]
    # persistent outputs / workspaces per batch so that every argument set repeats exactly (raw C-ABI calls)
    calls = []
    for bt in batches:
        b, keep = net.make_batch_struct(bt)
        ws = torch.empty(lib.cns_forward_workspace_bytes(C.byref(b)) + 1024, dtype=torch.uint8, device="cuda:0")
        out = torch.zeros(b.n_clusters * 128 + b.n_graphs * 13, device="cuda:0")
        calls.append((b, keep, ws, out))

    def run(k, stream):
        b, _, ws, out = calls[k]
        nb, nc = b.n_graphs, b.n_clusters
        hid, vec = out[:nc * 128], out[nc * 128:nc * 128 + nb * 6]
        vel, nrm = out[nc * 128 + nb * 6:nc * 128 + nb * 12], out[nc * 128 + nb * 12:]
        _lib.check(lib.cns_graphvs_forward(h.raw, C.byref(b), None, _lib.ptr(vec), _lib.ptr(nrm), _lib.ptr(hid), _lib.ptr(vel),
                                           _lib.ptr(ws), ws.numel(), None, 0, C.c_void_p(stream.cuda_stream)), "forward")

    h.set_option("cuda_graphs", 0)
    for k in range(n_sets):
        run(k, streams[0])
    torch.cuda.synchronize()
    ref = [c[3].clone() for c in calls]
    h.set_option("cuda_graphs", 1)
    c0, r0 = h.counter("graph_captures"), h.counter("graph_replays")
    for rep in range(4):
        for c in calls:
            c[3].zero_()
        for k in range(n_sets):
            run(k, streams[k % 2])
    torch.cuda.synchronize()
    for k in range(n_sets):
        assert torch.equal(calls[k][3], ref[k]), k
    # the cyclic order defeats an LRU of 48 entries: sets keep being captured again after eviction, none of it may corrupt
    # a result; a second pattern that fits the cache is replayed
    for rep in range(4):
        for k in range(40):
            run(k, streams[k % 2])
    torch.cuda.synchronize()
    assert h.counter("graph_replays") > r0 and h.counter("graph_captures") > c0
    for k in range(40):
        assert torch.equal(calls[k][3], ref[k]), k
