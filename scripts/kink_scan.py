"""CPU only: how close one round of scripts/fuzz_backward.py comes to a kink of the network - the smallest |pre-activation|
of any ReLU and the smallest gap between the two largest messages of any max aggregation, relative to the layer's largest
value, in the fp64 oracle.  A margin below the forward error of an fp32 computation (1e-6..1e-5) means the gradient of
that batch is decided by rounding: any fp32 implementation may differ from fp64 autograd there by one unit's worth.
usage: python scripts/kink_scan.py <round> [...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import numpy as np, torch
import cns_oracle
from cns_b200 import GraphGenerator, ckpt
from fuzz_backward import frames

GraphGenerator.CLUSTER_BACKEND = "host"
sd = ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns_state_dict.pth"))
W = {k: v.double() for k, v in sd.items()}
found = []
_relu, _segmax = torch.relu, cns_oracle.seg_max


def relu_spy(x):
    if x.numel():
        a = x.detach().abs()
        a = torch.where(a == 0, torch.full_like(a, float("inf")), a)       # exact zeros are structural (empty clusters)
        found.append(("relu", float(a.min() / x.detach().abs().max().clamp(min=1e-30)), tuple(x.shape)))
    return _relu(x)


def segmax_spy(src, index, n):
    if src.dim() == 2 and src.shape[0] > 1:
        s = src.detach()
        top = _segmax(s, index, n)
        masked = torch.where(s == top[index], torch.full_like(s, -float("inf")), s)   # drop the winner(s)
        second = torch.full_like(top, -float("inf")).scatter_reduce(0, index.view(-1, 1).expand_as(s), masked, reduce="amax")
        gap = (top - second)
        gap = gap[torch.isfinite(gap)]
        if gap.numel():
            found.append(("max", float(gap.min() / s.abs().max().clamp(min=1e-30)), tuple(src.shape)))
    return _segmax(src, index, n)


for rnd in [int(a) for a in sys.argv[1:]]:
    rng = np.random.default_rng(5000 + rnd)
    sizes = [int(rng.choice([4, 9, 16, 17, 40, 128, 300, 512])) for _ in range(int(rng.integers(1, 9)))]
    n_frames = int(rng.integers(1, 4)); missing = float(rng.choice([0.0, 0.15, 0.4, 0.6]))
    batches = frames(sizes, n_frames, 5000 + rnd, missing)
    found.clear()
    torch.relu, cns_oracle.seg_max = relu_spy, segmax_spy
    try:
        hid = None
        with torch.no_grad():
            for bt in batches:
                _, _, hid = cns_oracle.forward(W, bt, hid, dtype=torch.float64)
    finally:
        torch.relu, cns_oracle.seg_max = _relu, _segmax
    r = min((f for f in found if f[0] == "relu"), key=lambda f: f[1])
    m = min((f for f in found if f[0] == "max"), key=lambda f: f[1], default=("max", float("inf"), ()))
    print("round %3d sizes %s frames %d: closest ReLU margin %.1e (layer output %s), closest max-aggregation gap %.1e (%s)" %
          (rnd, sizes, n_frames, r[1], r[2], m[1], m[2]))
