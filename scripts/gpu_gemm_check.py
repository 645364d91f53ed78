"""Development check: bf16/f16 forward vs fp32 forward, stage by stage error (GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import cns_oracle
from cns_b200 import ckpt, synth
from cns_b200.data import Batch
sd = ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns_state_dict.pth"))
rng = np.random.default_rng(1); np.random.seed(1)
scenes = [synth.Scene(rng, n) for n in (40, 300, 512)]
batch = Batch.from_data_list([s.frame(0.1) for s in scenes])
ov, on, oh = cns_oracle.forward(sd, batch, None, dtype=torch.float64)
hid = torch.randn(oh.shape[0], 128, dtype=torch.float64) * 0.3
ov2, on2, oh2 = cns_oracle.forward(sd, batch, hid, dtype=torch.float64)
rel = lambda a, b: float((a.double().cpu() - b).abs().max() / b.abs().max())
for prec in ("fp32", "bf16", "f16"):
    net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), "cuda:0", prec)
    with torch.no_grad():
        p1 = net(batch.to("cuda:0"), None, check=True)
        p2 = net(batch.to("cuda:0"), hid.float().cuda(), check=True)
    print(prec, "hidden=None: vec %.2e nrm %.2e hid %.2e | hidden given: vec %.2e nrm %.2e hid %.2e" % (
        rel(p1[0], ov), rel(p1[1], on), rel(p1[2], oh), rel(p2[0], ov2), rel(p2[1], on2), rel(p2[2], oh2)))
