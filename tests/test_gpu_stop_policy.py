"""GPU (-m gpu): stop policy and match alignment on the device vs the reference's own outputs
(tests/golden/stop_policy.npz, align_matches.npz) and the oracle (oracle/servo_edges_oracle.py).  Bar: bit-exact -
float32 error included (the kernel follows numpy's operation order)."""
import os
import types

import numpy as np
import pytest
import torch

import servo_edges_oracle as so
from cns_b200 import _lib
from cns_b200.midend import CameraIntrinsic
from cns_b200.stop_policy import BatchedPixelStop, PixelStopPolicy, align_matches

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_pixel_stop_policy_reproduces_reference_sequence():
    g = np.load(os.path.join(GOLD, "stop_policy.npz"))
    pol = PixelStopPolicy(waiting_time=0.8, conduct_thresh=0.01, device=DEV)
    dt = float(g["dt"])
    tar = torch.from_numpy(g["tar"]).to(DEV)
    for k in range(len(g["cur"])):
        data = None
        if g["has_graph"][k]:
            data = types.SimpleNamespace(pos_cur=torch.from_numpy(g["cur"][k]), pos_tar=tar,
                                         node_mask=torch.from_numpy(g["mask"][k]))
        s = pol(data, k * dt)
        assert np.float32(pol.cur_err).tobytes() == g["err"][k].tobytes(), (k, pol.cur_err, g["err"][k])
        assert bool(s) == bool(g["stop"][k]), k
        assert np.float64(pol.accum_time) == g["accum"][k], k
    assert g["stop"].any()
    np.testing.assert_array_equal(pol.weight, g["weight"])
    np.testing.assert_array_equal(pol.witness_count, g["witness"])


def test_batched_pixel_stop_matches_oracle_per_scene():
    """Ragged scenes incl. one keypoint, fewer than 8 kept values, and more keypoints than the shared-memory
    staging holds (global-memory compaction + the recursive part of the pairwise sum)."""
    rng = np.random.default_rng(3)
    counts = [1, 7, 300, 5000, 12, 129, 1031]
    N, B = sum(counts), len(counts)
    off = np.concatenate([[0], np.cumsum(counts)])
    tar = rng.uniform(-0.5, 0.5, size=(N, 2)).astype(np.float32)
    drift = rng.normal(0, 0.05, size=(N, 2)).astype(np.float32)
    bat = BatchedPixelStop(counts, waiting_time=0.3, conduct_thresh=0.02, device=DEV)
    pix = [so.PixelStopOracle() for _ in counts]
    hold = [so.ErrorHoldingOracle(0.3, 0.02, on_reset=p.reset) for p in pix]
    tar_d = torch.from_numpy(tar).to(DEV)
    stopped = np.zeros(B, dtype=bool)
    for k in range(60):
        scale = np.float32(0.85 ** k + (0.6 if k == 30 else 0.0))
        cur = (tar + drift * scale + rng.normal(0, 1e-4, size=(N, 2))).astype(np.float32)
        mask = rng.random(N) > 0.15
        mask[off[:-1]] = True                               # every scene keeps at least one keypoint
        stop = bat(torch.from_numpy(cur), tar_d, torch.from_numpy(mask), k * 0.05)
        for b in range(B):
            sl = slice(off[b], off[b + 1])
            e = pix[b].calculate_error(cur[sl], tar[sl], mask[sl])
            want = hold[b](e, k * 0.05)
            assert np.float32(bat.rules[b].cur_err).tobytes() == np.float32(e).tobytes(), (k, b, counts[b])
            assert bool(stop[b]) == bool(want), (k, b)
        stopped |= stop
    assert stopped[2:].all()                                # the sequences do reach the stop condition
    w = bat.state.weight[:N].cpu().numpy()
    wc = bat.state.witness[:N].cpu().numpy()
    for b in range(B):
        sl = slice(off[b], off[b + 1])
        if bat.pending_reset[b]:
            continue                                        # reset is applied by the next launch
        np.testing.assert_array_equal(w[sl], pix[b].weight)
        np.testing.assert_array_equal(wc[sl], pix[b].witness)


def test_scene_without_visible_keypoint_gives_nan():
    bat = BatchedPixelStop([4, 6], waiting_time=0.3, device=DEV)
    pos = torch.rand(10, 2, generator=torch.Generator().manual_seed(5))
    mask = torch.tensor([0, 0, 0, 0, 1, 1, 0, 1, 1, 1], dtype=torch.bool)
    err = bat.errors(pos, pos * 0.9, mask).cpu().numpy()
    assert np.isnan(err[0]) and np.isfinite(err[1])


def test_align_matches_golden_and_random():
    g = np.load(os.path.join(GOLD, "align_matches.npz"))
    intr = CameraIntrinsic.default()
    assert [intr.fx, intr.fy, intr.cx, intr.cy] == list(g["intr"])
    xy, valid = align_matches(torch.from_numpy(g["cur_pos"]).to(DEV), torch.from_numpy(g["matches"]).to(DEV),
                              int(g["n_tar"]), intr)
    np.testing.assert_array_equal(valid.cpu().numpy(), g["valid"])
    assert xy.cpu().numpy().tobytes() == g["xy"].tobytes()
    rng = np.random.default_rng(5)
    for n_tar, n_cur, n_match in [(1, 1, 1), (50, 10, 400), (4096, 5000, 3000), (10, 10, 0)]:
        cur = rng.uniform(0, 640, size=(n_cur, 2)).astype(np.float32)
        m = np.stack([rng.integers(0, n_tar, n_match), rng.integers(0, n_cur, n_match)], 1).astype(np.int64)
        xy, valid = align_matches(torch.from_numpy(cur).to(DEV), torch.from_numpy(m).to(DEV), n_tar, intr)
        want_xy, want_valid = so.align_cur_to_tar(n_tar, cur, m, intr.fx, intr.fy, intr.cx, intr.cy)
        np.testing.assert_array_equal(valid.cpu().numpy(), want_valid)
        assert xy.cpu().numpy().tobytes() == want_xy.tobytes(), (n_tar, n_cur, n_match)


def test_pixel_stop_requires_cuda():
    with pytest.raises(_lib.CnsError):
        PixelStopPolicy(0.8, device="cpu").calculate_error(
            types.SimpleNamespace(pos_cur=torch.zeros(3, 2), pos_tar=torch.zeros(3, 2), node_mask=torch.ones(3, dtype=torch.bool)))


def test_batched_closed_loop_servo_stops_at_the_target():
    """Device scenes -> graph masking -> GraphVS (f16, hidden carried) -> camera motion, every scene until its
    PixelStopPolicy fires: the loop must end near the target pose (scripts/servo_closed_loop.py)."""
    import importlib.util
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("servo_closed_loop", os.path.join(root, "scripts", "servo_closed_loop.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    r = mod.run(B=24, N=128, precision="f16", max_frames=500, layouts=6, seed=5)
    assert r["stopped"] >= 20, r
    assert r["final_rot_deg_p50_p90_max"][0] < 0.5 and r["final_trans_mm_p50_p90_max"][0] < 5.0, r
