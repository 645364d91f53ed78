"""Development check: gradients of cns_graphvs_backward vs torch autograd on the fp64 oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import cns_oracle
from cns_b200 import GraphVS, ckpt, synth
from cns_b200.data import Batch

sd = ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns_state_dict.pth"))
rng = np.random.default_rng(1); np.random.seed(1)
sizes = [int(x) for x in sys.argv[1:]] or [40, 120, 300]
scenes = [synth.Scene(rng, n) for n in sizes]
frames = [[s.frame(0.15, with_labels=True) for s in scenes]]
for s in scenes: s.step(synth.pbvs_straight(s.cur_wcT, s.tar_wcT) * 2)
frames.append([s.frame(0.15, with_labels=True) for s in scenes])
batches = [Batch.from_data_list(f) for f in frames]

# oracle: fp64 autograd through two recurrent frames
W = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
hid = None; loss_o = 0
for bt in batches:
    vec, nrm, hid = cns_oracle.forward(W, bt, hid, dtype=torch.float64)
    loss_o = loss_o + cns_oracle.objectives((vec, nrm), bt.vel_si)[2] + 0.01 * (hid ** 2).mean()
loss_o.backward()

net = GraphVS(2, 2, 128).to("cuda:0"); net.load_state_dict(sd); net.train()
hid = None; loss_g = 0
for bt in batches:
    d = bt.to("cuda:0")
    pred = net(d, hid)
    hid = pred[2]
    loss_g = loss_g + net.objectives(pred, d)[1] + 0.01 * (hid ** 2).mean()
loss_g.backward()
print("loss oracle %.8f  gpu %.8f" % (float(loss_o), float(loss_g)))
worst = 0
for name, p in net.named_parameters():
    g, r = p.grad.double().cpu(), W[name].grad
    err = float((g - r).abs().max() / r.abs().max().clamp(min=1e-12))
    worst = max(worst, err)
    flag = "" if err < 1e-4 else "   <-- BAD"
    print("%-58s |ref| %.3e  rel %.2e%s" % (name, float(r.abs().max()), err, flag))
print("worst rel err %.2e" % worst)
