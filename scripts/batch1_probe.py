import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from cns_b200 import ckpt, synth
from cns_b200.controller import GraphVSController
ROOT="/root/repo"
rng = np.random.default_rng(0); np.random.seed(0)
scene = synth.Scene(rng, 512)
net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), "cuda:0", precision="f16")
d = scene.frame(0.1).to("cuda:0")
with torch.no_grad():
    hid = net(d, None)[2]
    for _ in range(20): hid = net(d, hid)[2]
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200): hid = net(d, hid)[2]
    b.record(); torch.cuda.synchronize()
    print("device-resident batch-1 forward: %.1f us per frame (GPU+launch, async)" % (a.elapsed_time(b) * 1000 / 200))
    t0 = time.perf_counter()
    for _ in range(200): hid = net(d, hid)[2]
    t1 = time.perf_counter(); torch.cuda.synchronize()
    print("host enqueue time per frame: %.1f us" % ((t1 - t0) * 1e6 / 200))
