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
K = 100
def timeit(name, fn):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter(); a.record(); fn(); b.record(); torch.cuda.synchronize()
    print("%-52s wall %.3f  events %.3f ms/step" % (name, (time.perf_counter() - t) / K * 1e3, a.elapsed_time(b) / K))
F = net.INPUT_FIELDS
with torch.no_grad():
    def pref(fields, post):
        def run():
            for d in DevicePrefetcher((host[i % 4] for i in range(K)), dev, fields=fields):
                p = net(d, None)
                (net.postprocess(p, d) if post else p.vel).cpu()
        return run
    timeit("prefetch all fields, .vel.cpu()", pref(None, False))
    timeit("prefetch INPUT_FIELDS, .vel.cpu()", pref(F, False))
    timeit("prefetch INPUT_FIELDS, postprocess().cpu()", pref(F, True))
    timeit("prefetch depth=2 INPUT_FIELDS", lambda: [net(d, None).vel.cpu() for d in DevicePrefetcher((host[i % 4] for i in range(K)), dev, depth=2, fields=F)])
    def no_thread_copy():
        for i in range(K):
            d = host[i % 4].to(dev, non_blocking=True, fields=F); net(d, None).vel.cpu()
    timeit("serial INPUT_FIELDS", no_thread_copy)
    # how long does the host side of .to() take?
    t = time.perf_counter()
    for i in range(K): host[i % 4].to(dev, non_blocking=True, fields=F)
    print("host time of .to(): %.3f ms" % ((time.perf_counter() - t) / K * 1e3)); torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(K): net(res, None) if False else None
    res = host[0].to(dev)
    torch.cuda.synchronize(); t = time.perf_counter()
    for i in range(K): net(res, None)
    print("host time of forward launch: %.3f ms" % ((time.perf_counter() - t) / K * 1e3)); torch.cuda.synchronize()
