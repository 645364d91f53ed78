"""Step time of the headline workload: one stream vs two alternating streams, CUDA-graph replay on/off.  python scripts/two_stream_probe.py [steps]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cns_b200 import ckpt  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), dev, precision="f16")
host = bench.build_batches(1024, 512, 2022, bench.N_ROTATE)
res = [b.to(dev) for b in host]
h = net._handle(dev)
streams = [torch.cuda.Stream(dev) for _ in range(4)]


def loop(n_streams, n):
    outs = []
    if n_streams == 1:
        for i in range(n):
            outs.append(net(res[i % len(res)], None).vel)
            if len(outs) > 2:
                outs.pop(0)
        return
    cur = torch.cuda.current_stream(dev)
    for st in streams:
        st.wait_stream(cur)
    for i in range(n):
        with torch.cuda.stream(streams[i % n_streams]):
            outs.append(net(res[i % len(res)], None).vel)
        if len(outs) > 8:
            outs.pop(0)
    for st in streams:
        cur.wait_stream(st)


def timed(n_streams):
    with torch.no_grad():
        loop(n_streams, 24)
        torch.cuda.synchronize()
        loop(n_streams, 4)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        loop(n_streams, steps)
        t1.record()
        torch.cuda.synchronize()
    return t0.elapsed_time(t1) / steps


out = {}
with torch.no_grad():
    ref = net(res[0], None)
    ref = [t.clone() for t in (ref[0], ref[1], ref[2], ref.vel)]
for graphs in (0, 1):
    for dyn in (0,):
        h.set_option("cuda_graphs", graphs)
        with torch.no_grad():
            got = net(res[0], None, check=True)
            err = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip((got[0], got[1], got[2], got.vel), ref))
        r = {"one_stream_ms": timed(1), "two_streams_ms": timed(2), "three_streams_ms": timed(3), "max_rel_diff_vs_static": err,
             "replays": h.counter("graph_replays"), "captures": h.counter("graph_captures")}
        out["graphs%d_dyn%d" % (graphs, dyn)] = r
        print("graphs", graphs, "dyn", dyn, r, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "two_stream_probe.json"), "w") as f:
    json.dump(out, f, indent=1)
