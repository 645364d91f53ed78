"""CPU: batched supervisors and observation augmentations (cns_b200/sim_aug.py) against the reference's
own per-scene classes (cns/sim/supervisor.py, cns/sim/sampling.py on the stubs) fed identical inputs, plus property
tests of the random subset helpers."""
import numpy as np
import pytest
import torch

import ref_import
from cns_b200 import sim_aug, synth

needs_ref = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")


def _poses(rng, B):
    cur = np.stack([synth.sample_initial_pose(rng) for _ in range(B)])
    tar = np.stack([synth.sample_target_pose(rng) for _ in range(B)])
    return cur, tar


@needs_ref
def test_supervisors_match_reference():
    ref_import.setup()
    from cns.sim import supervisor as S
    rng = np.random.default_rng(0)
    B = 40
    cur, tar = _poses(rng, B)
    pts = [synth.sample_points(rng, int(rng.integers(4, 60))) for _ in range(B)]
    wPo = np.stack([p.mean(0) for p in pts])
    tc, tt, tw = torch.from_numpy(cur), torch.from_numpy(tar), torch.from_numpy(wPo)
    for name, fn, ref in (("pbvs_straight", lambda: sim_aug.pbvs_straight(tc, tt), lambda k: S.pbvs_straight(cur[k], tar[k])),
                          ("pbvs_center", lambda: sim_aug.pbvs_center(tc, tt, tw), lambda k: S.pbvs_center(cur[k], tar[k], pts[k])),
                          ("pbvs_hybrid", lambda: sim_aug.pbvs_hybrid(tc, tt, tw), lambda k: S.pbvs_hybrid(cur[k], tar[k], pts[k]))):
        got = fn().numpy()
        want = np.stack([ref(k) for k in range(B)])
        assert np.abs(got - want).max() < 1e-8 * max(1.0, np.abs(want).max()), name
    for policy, pol in (("pbvs_straight", S.Policy.PBVS_Straight), ("pbvs_hybrid", S.Policy.PBVS_Center_Straight),
                        ("pbvs_center", S.Policy.PBVS_Center)):
        vel, tpo, vel_si = sim_aug.supervisor_vel(policy, tc, tt, tw)
        for k in range(0, B, 7):
            v, (n, vs) = S.supervisor_vel(pol, None, None, None, None, None, cur[k], tar[k], pts[k])
            assert np.abs(vel[k].numpy() - v).max() < 1e-8 and abs(float(tpo[k]) - n) < 1e-12
            assert np.abs(vel_si[k].numpy() - vs).max() < 1e-8


@needs_ref
def test_observation_sampler_matches_reference_given_the_same_draws():
    """Two scenes: the batched sampler, loaded with the state the reference objects drew, walks through the same
    observed / missing transitions and tau ratios when fed the same poses and the same re-draws."""
    ref_import.setup()
    from cns.sim.sampling import ObservationSampler
    rng = np.random.default_rng(1)
    sizes = (80, 150)
    xy = [synth.project_norm(synth.sample_target_pose(rng), synth.sample_points(rng, n)) for n in sizes]
    np.random.seed(5)
    refs = [ObservationSampler(p, tau_max=5) for p in xy]
    pts = torch.from_numpy(np.concatenate(xy))
    scene_of = torch.from_numpy(np.repeat(np.arange(2), [len(p) for p in xy]))
    med = torch.from_numpy(np.stack([np.median(p, axis=0) for p in xy]))
    bs = sim_aug.BatchedObservationSampler(pts, scene_of, 2, med, 5.0, torch.Generator().manual_seed(0))
    # load the reference objects' random state
    bs.center = torch.from_numpy(np.stack([r.walkers[0].center for r in refs]))
    assert all(np.array_equal(r.walkers[0].center, w.center) for r in refs for w in r.walkers)
    assert np.allclose(bs.sigma.numpy(), [r.sigma for r in refs])
    for r in refs:                       # keep every Perlin generator inside its current cell for this test (cell
        for w in r.walkers:              # crossings draw fresh gradients; they are covered by the Perlin test below)
            for pl in w.perlins:
                pl.offset, pl.prev_floor = 0.05, 0
    g = lambda f: torch.tensor([[[f(p) for p in w.perlins] for w in r.walkers] for r in refs], dtype=torch.float64)
    bs.perlin.offset, bs.perlin.g0, bs.perlin.g1 = g(lambda p: p.offset), g(lambda p: p.grad[0]), g(lambda p: p.grad[1])
    bs.perlin.prev_floor = g(lambda p: p.prev_floor).to(torch.int64)
    cat = lambda name: torch.from_numpy(np.concatenate([getattr(r, name) for r in refs]))
    bs.state, bs.tau_total, bs.no_event_rval = cat("state"), cat("tau_total"), cat("no_event_rval")
    cur = [synth.sample_initial_pose(rng) for _ in sizes]
    tar = [synth.sample_target_pose(rng) for _ in sizes]
    flips = 0
    for step in range(40):
        t = 0.02 * step
        np.random.seed(100 + step)
        want = [r(time=t, current_wcT=c.copy(), dist_scale=0.7) for r, c in zip(refs, cur)]
        np.random.seed(100 + step)

        def redraw(trans):           # the reference redraws scene by scene, in keypoint order
            out = torch.zeros(trans.shape, dtype=torch.float64)
            for k in range(2):
                m = trans & (scene_of == k)
                if bool(m.any()):
                    out[m] = torch.from_numpy(np.random.uniform(0, 1, size=int(m.sum())))
            return out

        before = bs.state.clone()
        observed, ratio = bs(torch.full((2,), t, dtype=torch.float64), torch.from_numpy(np.stack(cur)), 0.7, redraw=redraw)
        flips += int((before != observed).sum())
        o = 0
        for (unobs, rat), n in zip(want, sizes):
            assert np.array_equal(np.nonzero(~observed[o:o + n].numpy())[0], unobs)
            assert np.abs(ratio[o:o + n].numpy() - rat).max() < 1e-12
            o += n
        assert np.abs(bs.no_event_rval.numpy() - np.concatenate([r.no_event_rval for r in refs])).max() == 0
        for k in range(2):
            cur[k] = synth.integrate(cur[k], synth.pbvs_straight(cur[k], tar[k]) * 2)
    assert flips > 20                                   # the transitions were actually exercised


@needs_ref
def test_perlin_matches_reference_across_cell_boundaries():
    ref_import.setup()
    from cns.sim.sampling import Perlin1d
    np.random.seed(3)
    refs = [Perlin1d(amp, 1.0) for amp in (0.5, 1.5, 2.0)]
    bp = sim_aug.BatchedPerlin(torch.tensor([0.5, 1.5, 2.0], dtype=torch.float64), torch.ones(3, dtype=torch.float64),
                               torch.Generator().manual_seed(0))
    f64 = lambda v: torch.tensor(v, dtype=torch.float64)
    bp.offset = f64([p.offset for p in refs])
    bp.g0, bp.g1 = f64([p.grad[0] for p in refs]), f64([p.grad[1] for p in refs])
    draws = []
    bp.next_grad = lambda mask: torch.tensor([draws.pop(0) if m else 0.0 for m in mask.tolist()], dtype=torch.float64)
    for t in np.linspace(0, 4.3, 60):
        want = []
        for p in refs:                                   # record what each generator draws at this step
            before = p.grad[1]
            want.append(p(t))
            if p.grad[0] == before and p.prev_floor is not None and p.grad[1] != before:
                draws.append(p.grad[1])
        got = bp(torch.full((3,), t, dtype=torch.float64))
        assert np.abs(got.numpy() - np.array(want)).max() < 1e-12


def test_random_subsets_and_rotation():
    gen = torch.Generator().manual_seed(0)
    n_per = torch.tensor([5, 40, 12, 100])
    scene_of = torch.repeat_interleave(torch.arange(4), n_per)
    counts = (n_per.double() * 0.1).to(torch.int64)
    pos = sim_aug.random_positions(scene_of, n_per, gen)
    for b in range(4):
        assert sorted(pos[scene_of == b].tolist()) == list(range(int(n_per[b])))       # a permutation per scene
    sel, order = sim_aug.random_subsets(scene_of, n_per, counts, pos=pos)
    sel2, _ = sim_aug.random_subsets(scene_of, n_per, counts, offset=counts, pos=pos)
    assert not bool((sel & sel2).any())                                             # disjoint slices of one permutation
    for b in range(4):
        m = sel[scene_of == b]
        assert int(m.sum()) == int(counts[b]) == int(sel2[scene_of == b].sum())
        assert sorted(order[scene_of == b][m].tolist()) == list(range(int(counts[b])))
    # values[indices] = values[np.roll(indices, 1)]
    vals = torch.arange(scene_of.numel(), dtype=torch.float64)[:, None].repeat(1, 2)
    out = sim_aug.rotate_within_subsets(vals, scene_of, sel, order, counts)
    for b in range(4):
        idx = torch.nonzero(sel & (scene_of == b)).flatten()
        idx = idx[torch.argsort(order[idx])].numpy()
        ref = vals.numpy().copy()
        ref[idx] = ref[np.roll(idx, 1)]
        assert np.array_equal(out.numpy()[scene_of.numpy() == b], ref[scene_of.numpy() == b])
    # the mismatcher keeps its subset until the interval has passed
    mm = sim_aug.BatchedMismatcher(scene_of, n_per, tau=2.0, ratio=0.1, gen=gen)
    mm.counts = counts
    s0, _ = mm(0.02)
    s0 = s0.clone()
    mm.interval[:] = 10.0
    s1, _ = mm(0.02)
    assert torch.equal(s0, s1)
    mm.interval[1] = 0.0
    s2, _ = mm(0.02)
    assert torch.equal(s2[scene_of != 1], s0[scene_of != 1]) and int(s2[scene_of == 1].sum()) == int(counts[1])
