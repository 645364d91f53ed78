import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cns_b200 import ckpt
from cns_b200.data import DevicePrefetcher
import bench
dev = torch.device("cuda:0")
net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), dev, precision="f16")
host = [b.pin_memory() for b in bench.build_batches(1024, 512, 2022, 4)]
K = 40
def timeit(name, fn):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter(); fn(); torch.cuda.synchronize()
    print("%-44s %.3f ms/step" % (name, (time.perf_counter() - t) / K * 1e3))
with torch.no_grad():
    res = [b.to(dev) for b in host]
    timeit("compute only (resident)", lambda: [net(res[i % 4], None) for i in range(K)])
    timeit("compute + .cpu() each step", lambda: [net(res[i % 4], None).vel.cpu() for i in range(K)])
    timeit("copy only (.to non_blocking)", lambda: [host[i % 4].to(dev, non_blocking=True) for i in range(K)])
    def serial():
        for i in range(K):
            d = host[i % 4].to(dev, non_blocking=True); net(d, None).vel.cpu()
    timeit("serial copy+compute+read", serial)
    def pref():
        for d in DevicePrefetcher((host[i % 4] for i in range(K)), dev):
            net(d, None).vel.cpu()
    timeit("prefetcher + read each step", pref)
    def pref_deferred():
        prev = None
        for d in DevicePrefetcher((host[i % 4] for i in range(K)), dev):
            p = net(d, None)
            if prev is not None: prev.cpu()
            prev = p.vel
        prev.cpu()
    timeit("prefetcher + read deferred one step", pref_deferred)
