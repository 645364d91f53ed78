"""Is the training update GPU-bound or launch-bound?  Host enqueue time vs CUDA-event time of one update."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cns_b200 import GraphVS, ckpt, synth
from cns_b200.data import Batch
from cns_b200.train import TBPTTStepper
bs, n_frames, npts = 256, 8, 512
rng = np.random.default_rng(3); np.random.seed(3)
base = [synth.Scene(rng, npts) for _ in range(8)]
scenes = [synth.Scene(rng, npts, generator=base[i % 8].generator, points=base[i % 8].points, tar_wcT=base[i % 8].tar_wcT) for i in range(bs)]
frames = []
for t in range(n_frames):
    frames.append(Batch.from_data_list([s.frame(0.1, with_labels=True) for s in scenes]).to("cuda:0"))
    for s in scenes: s.step(synth.pbvs_straight(s.cur_wcT, s.tar_wcT) * 2)
net = GraphVS(2, 2, 128).to("cuda:0"); net.load_state_dict(ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns.pth"))); net.train()
st = TBPTTStepper(net, lr=1e-5)
for _ in range(3):
    st.hidden = None; st.step(frames)
torch.cuda.synchronize()
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record(); st.hidden = None
    # forward only part
    loss, _ = st.window_loss(frames)
    t1 = time.perf_counter(); f = torch.cuda.Event(enable_timing=True); f.record()
    st.bucket.zero(); loss.backward(); st.bucket.rebind(); st.fused_opt.step()
    b.record(); t2 = time.perf_counter(); torch.cuda.synchronize(); t3 = time.perf_counter()
    print("host: fwd enqueue %.1f ms, bwd enqueue %.1f ms, drain %.1f ms | gpu: fwd %.1f ms, bwd+opt %.1f ms, total %.1f ms" % (
        (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, a.elapsed_time(f), f.elapsed_time(b), a.elapsed_time(b)))
