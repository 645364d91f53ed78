"""Where the host time of BatchedServo.submit/result goes (cProfile) at the headline shape."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cns_b200 import ckpt  # noqa: E402
from cns_b200.servo import BatchedServo  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), dev, precision="f16")
host, scenes = bench.build_batches(1024, 512, 2022, bench.N_ROTATE, with_scenes=True)
host = [b.pin_memory() for b in host]
servo = BatchedServo(net, [s.generator for s in scenes], [s.tPo_norm for s in scenes], dev)
frames = [(b.pos_cur, b.node_mask.view(torch.uint8)) for b in host]


def loop(n):
    prev = None
    for i in range(n):
        servo.reset()
        t = servo.submit(*frames[i % len(frames)])
        if prev is not None:
            servo.result(prev)
        prev = t
    servo.result(prev)


with torch.no_grad():
    loop(30)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loop(200)
    torch.cuda.synchronize()
    print("wall per frame: %.4f ms" % ((time.perf_counter() - t0) / 200 * 1e3))
    # host-only cost: never wait for results (slots are reused, so done[k].synchronize still bounds the queue depth)
    pr = cProfile.Profile()
    pr.enable()
    loop(200)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)
