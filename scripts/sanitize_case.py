"""Tiny forward (fp32 + f16) and backward used under compute-sanitizer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cns_b200 import GraphVS, ckpt, synth
from cns_b200.data import Batch
rng = np.random.default_rng(0); np.random.seed(0)
scenes = [synth.Scene(rng, n) for n in (9, 40, 130)]
batch = Batch.from_data_list([s.frame(0.2, with_labels=True) for s in scenes]).to("cuda:0")
for prec in ("fp32", "f16", "bf16"):
    net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), "cuda:0", prec)
    with torch.no_grad():
        p = net(batch, None); p2 = net(batch, p[2])
    torch.cuda.synchronize()
net = GraphVS(2, 2, 128).to("cuda:0"); net.load_state_dict(ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns.pth"))); net.train()
p = net(batch, None); p2 = net(batch, p[2])
(net.objectives(p2, batch)[1] + net.objectives(p, batch)[1]).backward()
torch.cuda.synchronize()
print("done")
