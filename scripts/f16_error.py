"""Relative error (max|d|/max|ref|) of the f16 / bf16 modes against the fp64 oracle over recurrent frames."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import cns_oracle
from cns_b200 import Batch, ckpt, synth
sd = ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns.pth"))
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-12))
for prec in sys.argv[1:] or ["f16"]:
    net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), "cuda:0", precision=prec)
    rng = np.random.default_rng(3); np.random.seed(3)
    scenes = [synth.Scene(rng, n) for n in (24, 100, 300, 512, 33, 512, 200, 64)]
    hg = ho = None; worst = [0.0, 0.0]
    for f in range(8):
        batch = Batch.from_data_list([s.frame(0.2) for s in scenes])
        ov, on, ho = cns_oracle.forward(sd, batch, ho, dtype=torch.float64)
        ovel = cns_oracle.postprocess((ov, on), batch.tPo_norm.reshape(-1))
        with torch.no_grad():
            pred = net(batch.to("cuda:0"), hg, check=True)
        hg = pred[2]
        worst = [max(worst[0], rel(pred.vel, ovel)), max(worst[1], rel(pred[2], ho))]
        for s, v in zip(scenes, pred.vel.cpu().numpy()): s.step(v * 2)
    print("%s: vel %.2e hidden %.2e" % (prec, worst[0], worst[1]))
