"""Randomised soak of the training path: parameter gradients of cns_graphvs_backward (through GraphVS.forward in train
mode + autograd wrapper) against torch autograd on the fp64 oracle, over random ragged batches (4..512 keypoints, 1..8
graphs, 1..3 recurrent frames, 0..60 % missing keypoints).  Bar as in tests/test_gpu_backward.py: per tensor
max|g - ref| <= 1e-4 max|ref| + 2e-6 max over tensors.  Prints the worst ratio per round; exits 1 on a miss.
usage: python scripts/fuzz_backward.py [n_rounds]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import cns_oracle
from cns_b200 import Batch, GraphVS, ckpt, synth


def frames(sizes, n_frames, seed, missing):
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    scenes = [synth.Scene(rng, n) for n in sizes]
    out = []
    for t in range(n_frames):
        fr = [s.frame(missing, with_labels=True) for s in scenes]
        if t == 0:
            for f in fr:
                f.start_new_scene()
        out.append(Batch.from_data_list(fr))
        for s in scenes:
            s.step(synth.pbvs_straight(s.cur_wcT, s.tar_wcT) * 2)
    return out


def run(n_rounds=10, verbose=True):
    sd = ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns_state_dict.pth"))
    net = GraphVS(2, 2, 128).to("cuda:0")
    net.load_state_dict(sd)
    net.train()
    worst_all, bad = 0.0, 0
    for rnd in range(n_rounds):
        rng = np.random.default_rng(5000 + rnd)
        sizes = [int(rng.choice([4, 9, 16, 17, 40, 128, 300, 512])) for _ in range(int(rng.integers(1, 9)))]
        n_frames = int(rng.integers(1, 4))
        missing = float(rng.choice([0.0, 0.15, 0.4, 0.6]))
        batches = frames(sizes, n_frames, 5000 + rnd, missing)
        W = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
        hid, loss_o = None, 0
        for bt in batches:
            vec, nrm, hid = cns_oracle.forward(W, bt, hid, dtype=torch.float64)
            loss_o = loss_o + cns_oracle.objectives((vec, nrm), bt.vel_si)[2]
        loss_o.backward()
        net.zero_grad(set_to_none=True)
        hid, loss_g = None, 0
        for bt in batches:
            d = bt.to("cuda:0")
            pred = net(d, hid)
            hid = pred[2]
            loss_g = loss_g + net.objectives(pred, d)[1]
        loss_g.backward()
        scale = max(float(w.grad.abs().max()) for w in W.values())
        worst, who = 0.0, None
        for name, p in net.named_parameters():
            g, r = p.grad.double().cpu(), W[name].grad
            ratio = float((g - r).abs().max()) / (1e-4 * float(r.abs().max()) + 2e-6 * scale)
            if not torch.isfinite(g).all():
                ratio = float("inf")
            if ratio > worst:
                worst, who = ratio, name
        dl = abs(float(loss_g) - float(loss_o)) / max(1.0, abs(float(loss_o)))
        ok = worst <= 1.0 and dl < 1e-5
        bad += 0 if ok else 1
        worst_all = max(worst_all, worst)
        if verbose:
            print("round %3d  %-4s sizes %s frames %d missing %.2f  loss diff %.1e  worst |dg| / bar %.3f (%s)" %
                  (rnd, "ok" if ok else "FAIL", sizes, n_frames, missing, dl, worst, who), flush=True)
    print("%d rounds, %d over the bar; worst |dg| / bar %.3f" % (n_rounds, bad, worst_all))
    return bad


if __name__ == "__main__":
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 10) else 0)
