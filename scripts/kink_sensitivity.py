"""CPU only: is a round of scripts/fuzz_backward.py decided by rounding?  The fp64 oracle's parameter gradients are recomputed
with the keypoint coordinates of every frame perturbed by a relative 1e-6 (the size of an fp32 forward's error; 4 trials)
and compared with the unperturbed fp64 gradients under the same bar the kernels are held to (|dg| <= 1e-4 max|g| + 2e-6 max
over tensors).  A smooth batch moves by ~1e-2 of the bar; a batch that sits on a ReLU / max kink jumps by a multiple of it -
no fp32 implementation can be expected to match fp64 autograd there.
usage: python scripts/kink_sensitivity.py <round> [...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import copy
import numpy as np, torch
import cns_oracle
from cns_b200 import GraphGenerator, ckpt
from fuzz_backward import frames

GraphGenerator.CLUSTER_BACKEND = "host"
sd = ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns_state_dict.pth"))


def grads(batches):
    W = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    hid, loss = None, 0
    for bt in batches:
        vec, nrm, hid = cns_oracle.forward(W, bt, hid, dtype=torch.float64)
        loss = loss + cns_oracle.objectives((vec, nrm), bt.vel_si)[2]
    loss.backward()
    return {k: v.grad for k, v in W.items()}


for rnd in [int(a) for a in sys.argv[1:]]:
    rng = np.random.default_rng(5000 + rnd)
    sizes = [int(rng.choice([4, 9, 16, 17, 40, 128, 300, 512])) for _ in range(int(rng.integers(1, 9)))]
    n_frames = int(rng.integers(1, 4)); missing = float(rng.choice([0.0, 0.15, 0.4, 0.6]))
    batches = frames(sizes, n_frames, 5000 + rnd, missing)
    ref = grads(batches)
    scale = max(float(g.abs().max()) for g in ref.values())
    g = torch.Generator().manual_seed(rnd)
    worst = []
    for trial in range(4):
        pert = []
        for bt in batches:
            b2 = copy.copy(bt)
            noise = 1 + 1e-6 * torch.randn(bt.pos_cur.shape, generator=g)
            b2.pos_cur = (bt.pos_cur.double() * noise).float()
            b2.x_cur = b2.pos_cur.clone()
            pert.append(b2)
        gp = grads(pert)
        worst.append(max(float((gp[k] - ref[k]).abs().max()) / (1e-4 * float(ref[k].abs().max()) + 2e-6 * scale) for k in ref))
    print("round %3d sizes %s frames %d: fp64 gradient change under 1e-6 input noise, in units of the bar: %s" %
          (rnd, sizes, n_frames, ", ".join("%.2f" % w for w in worst)), flush=True)
