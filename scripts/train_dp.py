"""Data-parallel training check (run under torchrun, one rank per GPU, NCCL):
every rank trains on its own shard of the scenes; gradients are averaged with ONE all-reduce of the
flat bucket per update.  Rank 0 compares the averaged gradient against a single-process pass over the
whole global batch and reports update time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from cns_b200 import GraphVS, ckpt, synth
from cns_b200.data import Batch
from cns_b200.train import TBPTTStepper, shard_range

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
per_rank, n_frames, npts = (int(os.environ.get(k, d)) for k, d in (('PER_RANK', 32), ('FRAMES', 4), ('NPTS', 256)))
check = os.environ.get('CHECK', '1') == '1'       # compare against a single-process pass over the whole global batch
rng = np.random.default_rng(5); np.random.seed(5)          # identical scenes on every rank
base = [synth.Scene(rng, npts) for _ in range(4)]
scenes = [synth.Scene(rng, npts, generator=base[i % 4].generator, points=base[i % 4].points, tar_wcT=base[i % 4].tar_wcT)
          for i in range(per_rank * world)]
all_frames = []
for t in range(n_frames):
    all_frames.append([s.frame(0.1, with_labels=True) for s in scenes])
    for s in scenes:
        s.step(synth.pbvs_straight(s.cur_wcT, s.tar_wcT) * 2)
lo, hi = shard_range(len(scenes), rank, world)
mine = [Batch.from_data_list(fr[lo:hi]).to(dev) for fr in all_frames]
sd = ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns.pth"))
net = GraphVS(2, 2, 128).to(dev); net.load_state_dict(sd); net.train()
st = TBPTTStepper(net)
st.bucket.zero()
loss, _ = st.window_loss(mine)
loss.backward(); st.bucket.rebind(); st.bucket.allreduce_mean()
g_dp = st.bucket.flat.clone()
if rank == 0 and check:
    ref = GraphVS(2, 2, 128).to(dev); ref.load_state_dict(sd); ref.train()
    sr = TBPTTStepper(ref)
    sr.bucket.zero()
    full = [Batch.from_data_list(fr).to(dev) for fr in all_frames]
    l2, _ = sr.window_loss(full)
    l2.backward(); sr.bucket.rebind()
    err = float((g_dp - sr.bucket.flat).abs().max() / sr.bucket.flat.abs().max())
    print("DP world=%d: |grad_dp - grad_full| / |grad_full| = %.2e (equal shards: mean of means == global mean)" % (world, err))
dist.barrier()
# timed updates
for _ in range(2):
    st.hidden = None; st.step(mine)
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    st.hidden = None; st.step(mine)
b.record(); torch.cuda.synchronize()
t = torch.tensor([a.elapsed_time(b) / 5], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("DP world=%d: %.2f ms per update, %.0f graph-frames/s (global batch %d x %d frames)" % (
        world, float(t), per_rank * world * n_frames / float(t) * 1e3, per_rank * world, n_frames))
dist.destroy_process_group()
