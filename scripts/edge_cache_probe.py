"""Step time with / without the per-target edge cache (one stream), and whether the cached path is taken."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from cns_b200 import ckpt
dev = torch.device("cuda:0")
net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), dev, precision=os.environ.get("CNS_PRECISION", "f16"))
b = bench.build_batches(1024, 512, 2022, 1)[0].to(dev)
h = net._handle(dev)
cache = net.build_target_cache(b)


def t(n=30, **kw):
    with torch.no_grad():
        for _ in range(6):
            net(b, None, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            net(b, None, **kw)
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("no cache      %.4f ms" % t())
c0 = h.counter("edge_cache_forwards")
print("target cache  %.4f ms   (edge-cache forwards enqueued: %d)" % (t(target_cache=cache), h.counter("edge_cache_forwards") - c0))
