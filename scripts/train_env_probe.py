"""Host profile of the closed-loop training loop with scenes on the GPU (train_cns.py --device-env)."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cns_b200 import GraphVS, ckpt
from cns_b200.sim_device import DeviceSceneBatch
from cns_b200.train import TBPTTStepper

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
net = GraphVS(2, 2, 128).to(dev)
net.load_state_dict(ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns.pth")))
net.train()
stepper = TBPTTStepper(net, fused=True, lr=1e-5)
envs = DeviceSceneBatch(B, np.random.default_rng(0), dev, 512, aug=True, reinit="all", supervisor="pbvs_straight")


def update():
    def rollout():
        for _ in range(8):
            data = envs.observe()
            yield data
            with torch.no_grad():
                pred = stepper.last_pred
                vel = net.postprocess((pred[0].detach(), pred[1].detach()), data)
                n_gt = max(1, int(vel.shape[0] * 0.1))
                vel[:n_gt] = data.vel[:n_gt].to(vel)
            envs.feedback(vel * 2)
    return stepper.step(rollout())


for _ in range(40):
    update()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(40):
    update()
torch.cuda.synchronize()
print("%.2f ms per update (%d scenes, 8 frames)" % ((time.perf_counter() - t0) / 40 * 1e3, B))
pr = cProfile.Profile()
pr.enable()
for _ in range(40):
    update()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
