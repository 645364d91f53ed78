"""Quick GPU check used during development: forward vs the CPU oracle + a timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import cns_oracle
from cns_b200 import GraphVS, ckpt, synth
from cns_b200.data import Batch

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
sd = ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns_state_dict.pth"))
net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), "cuda:0", prec)

def relerr(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-12))

for sizes in ([40], [24, 100, 300, 512], [512] * 8):
    scenes = []
    rng = np.random.default_rng(1)
    np.random.seed(1)
    scenes = [synth.Scene(rng, n) for n in sizes]
    hid_o, hid_g = None, None
    for step in range(3):
        batch = Batch.from_data_list([s.frame(0.1) for s in scenes])
        ov, on, oh, inter = cns_oracle.forward(sd, batch, hid_o, dtype=torch.float64, return_intermediates=True)
        hid_o = oh
        dev = batch.to("cuda:0")
        with torch.no_grad():
            pred = net(dev, hid_g, check=True)
        gv, gn, gh = pred
        hid_g = gh
        ovel = cns_oracle.postprocess((ov, on), batch.tPo_norm)
        print("sizes=%s step=%d  vec %.2e  nrm %.2e  hid %.2e  vel %.2e" % (
            sizes if len(sizes) < 5 else "%dx%d" % (len(sizes), sizes[0]), step,
            relerr(gv, ov), relerr(gn, on), relerr(gh, oh), relerr(pred.vel, ovel)))
        vel = pred.vel.cpu().numpy()
        for s, v in zip(scenes, vel):
            s.step(v * 2)

# timing: B graphs x 512
B = int(os.environ.get("SMOKE_B", "256"))
batch = synth.make_batch(B, 512, seed=3, n_layouts=8).to("cuda:0")
with torch.no_grad():
    for _ in range(3):
        net(batch, None)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        net(batch, None)
    t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / 10
print("B=%d N=512 forward %.3f ms -> %.0f graphs/s (%s)" % (B, ms, B / ms * 1e3, prec))
