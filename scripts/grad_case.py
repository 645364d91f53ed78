"""One round of scripts/fuzz_backward.py in detail: the GPU gradients against fp64 autograd AND against the same oracle run
in fp32 (a deviation both fp32 computations share is a kink of the network - a max / ReLU decided differently under
rounding - not an error of the kernels).  usage: python scripts/grad_case.py <round> [tensor-name-substring]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import numpy as np, torch
import cns_oracle
from cns_b200 import GraphVS, ckpt
from fuzz_backward import frames

rnd = int(sys.argv[1]); pick = sys.argv[2] if len(sys.argv) > 2 else None
sd = ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns_state_dict.pth"))
rng = np.random.default_rng(5000 + rnd)
sizes = [int(rng.choice([4, 9, 16, 17, 40, 128, 300, 512])) for _ in range(int(rng.integers(1, 9)))]
n_frames = int(rng.integers(1, 4)); missing = float(rng.choice([0.0, 0.15, 0.4, 0.6]))
batches = frames(sizes, n_frames, 5000 + rnd, missing)
print("round", rnd, "sizes", sizes, "frames", n_frames, "missing", missing)


def oracle(dtype):
    W = {k: v.to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    hid, loss = None, 0
    for bt in batches:
        vec, nrm, hid = cns_oracle.forward(W, bt, hid, dtype=dtype)
        loss = loss + cns_oracle.objectives((vec, nrm), bt.vel_si)[2]
    loss.backward()
    return {k: v.grad.double() for k, v in W.items()}


g64, g32 = oracle(torch.float64), oracle(torch.float32)
net = GraphVS(2, 2, 128).to("cuda:0"); net.load_state_dict(sd); net.train()
rel = lambda a, b: float((a.double().cpu() - b).abs().max() / b.abs().max().clamp(min=1e-30))
with torch.no_grad():                      # forward outputs frame by frame: oracle fp64 vs train-mode and eval-mode kernels
    W = {k: v.double() for k, v in sd.items()}
    ho = ht = he = None
    for t, bt in enumerate(batches):
        ov, on, ho = cns_oracle.forward(W, bt, ho, dtype=torch.float64)
        d = bt.to("cuda:0")
        net.train(); pt = net(d, ht); ht = pt[2]
        net.eval(); pe = net(d, he); he = pe[2]
        print("frame %d  train-mode: vec %.2e nrm %.2e hidden %.2e   eval-mode: vec %.2e nrm %.2e hidden %.2e" %
              (t, rel(pt[0], ov), rel(pt[1], on), rel(pt[2], ho), rel(pe[0], ov), rel(pe[1], on), rel(pe[2], ho)))
net.train()
hid, loss = None, 0
for bt in batches:
    d = bt.to("cuda:0"); pred = net(d, hid); hid = pred[2]; loss = loss + net.objectives(pred, d)[1]
loss.backward()
scale = max(float(g.abs().max()) for g in g64.values())
print("%-52s %10s %12s %12s %12s" % ("tensor", "max|ref|", "gpu-fp64", "fp32-fp64", "gpu-fp32"))
for name, p in net.named_parameters():
    g = p.grad.double().cpu()
    bar = 1e-4 * float(g64[name].abs().max()) + 2e-6 * scale
    a, b, c = [float(x.abs().max()) / bar for x in (g - g64[name], g32[name] - g64[name], g - g32[name])]
    if max(a, b) > 0.5 or (pick and pick in name):
        print("%-52s %10.3e %12.3f %12.3f %12.3f" % (name, float(g64[name].abs().max()), a, b, c))
    if pick and pick in name:
        dlt = (g - g64[name]).abs()
        idx = torch.nonzero(dlt > 0.2 * dlt.max())
        print("   elements above 20 %% of the largest deviation: %d of %d; rows %s cols %s" %
              (len(idx), dlt.numel(), sorted(set(idx[:, 0].tolist()))[:12], sorted(set(idx[:, -1].tolist()))[:12]))
