"""One (or a few) short-sequence training update(s): used for profiling / timing the fwd+bwd path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cns_b200 import GraphVS, ckpt, synth
from cns_b200.data import Batch
from cns_b200.train import TBPTTStepper

bs, n_frames, npts, iters = [int(x) for x in (sys.argv[1:5] + ["256", "2", "512", "2"][len(sys.argv) - 1:])]
rng = np.random.default_rng(3); np.random.seed(3)
base = [synth.Scene(rng, npts) for _ in range(8)]
scenes = [synth.Scene(rng, npts, generator=base[i % 8].generator, points=base[i % 8].points, tar_wcT=base[i % 8].tar_wcT)
          for i in range(bs)]
frames = []
for t in range(n_frames):
    fr = [s.frame(0.1, with_labels=True) for s in scenes]
    frames.append(Batch.from_data_list(fr).to("cuda:0"))
    for s in scenes:
        s.step(synth.pbvs_straight(s.cur_wcT, s.tar_wcT) * 2)
net = GraphVS(2, 2, 128).to("cuda:0")
net.load_state_dict(ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns.pth")))
net.train()
st = TBPTTStepper(net)
for _ in range(iters):
    st.hidden = None
    loss = st.step(frames)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); st.hidden = None; loss = st.step(frames); b.record(); torch.cuda.synchronize()
print("bs %d frames %d N %d: %.2f ms per update, loss %.4f" % (bs, n_frames, npts, a.elapsed_time(b), float(loss)))
