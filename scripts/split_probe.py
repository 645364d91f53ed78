"""fp32-mode encoder: operand-split tcgen05 kernel vs the FFMA kernel vs the fp64 oracle on one small batch."""
import faulthandler, os, sys
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import cns_oracle
from cns_b200 import Batch, ckpt, synth
print("imports ok", flush=True)
sd = ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns_state_dict.pth"))
net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), "cuda:0", precision=os.environ.get("CNS_PROBE_PRECISION", "fp32"))
print("model ok", flush=True)
h = net._handle(torch.device("cuda:0"))
torch.cuda.synchronize()
print("handle ok", flush=True)
rng = np.random.default_rng(0); np.random.seed(0)
sizes = [int(x) for x in (sys.argv[1:] or ["40", "200"])]
scenes = [synth.Scene(rng, n) for n in sizes]
batch = Batch.from_data_list([s.frame(0.1) for s in scenes])
ov, on, oh = cns_oracle.forward(sd, batch, None, dtype=torch.float64)
with torch.no_grad():
    pred = net(batch.to("cuda:0"), None, check=True)
torch.cuda.synchronize()
print("forward ok", flush=True)
rel = lambda a, b: float((a.double().cpu() - b).abs().max() / b.abs().max())
print("vec %.3e  nrm %.3e  hidden %.3e" % (rel(pred[0], ov), rel(pred[1], on), rel(pred[2], oh)))
