"""Gradients of two fresh models on the same inputs must be bit-identical (fixed-order reductions, no fp atomics)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cns_b200 import GraphVS, ckpt, synth
sd = ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns.pth"))
devs = sys.argv[1:] or ["cuda:0", "cuda:0"]
for n_graphs, n_pts in ((12, 150), (40, 512)):
    batch = synth.make_batch(n_graphs, n_pts, seed=9, n_layouts=3)
    outs = []
    for dev in devs:
        net = GraphVS(2, 2, 128).to(dev); net.load_state_dict(sd); net.train()
        pred = net(batch.to(dev), None)
        (pred[0].square().sum() + pred[2].square().sum()).backward()
        d = {n: p.grad.detach().cpu().clone() for n, p in net.named_parameters()}
        d["_vec"], d["_nrm"], d["_hidden"] = pred[0].detach().cpu(), pred[1].detach().cpu(), pred[2].detach().cpu()
        outs.append(d)
    bad = [(n, float((outs[0][n] - outs[1][n]).abs().max())) for n in outs[0] if not torch.equal(outs[0][n], outs[1][n])]
    print(n_graphs, n_pts, devs, "differing:", len(bad), "equal:", [n for n in outs[0] if torch.equal(outs[0][n], outs[1][n])])
