import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cns_b200 import ckpt, _lib, synth
import bench
dev = torch.device("cuda:0")
net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), dev, precision="f16")
for name, B, N in (("B=1024", 1024, 512), ("B=1", 1, 512)):
    batch = bench.build_batches(B, N, 2022, 1)[0].to(dev)
    lib = _lib.load()
    K = 200
    def T(label, fn):
        fn(); torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(K): fn()
        dt = (time.perf_counter() - t) / K * 1e6
        torch.cuda.synchronize()
        print("%s %-28s %.1f us" % (name, label, dt))
    with torch.no_grad():
        T("make_batch_struct", lambda: net.make_batch_struct(batch))
        T("_handle", lambda: net._handle(dev))
        b, keep = net.make_batch_struct(batch)
        h = net._handle(dev)
        nb, nc = b.n_graphs, b.n_clusters
        out = torch.empty(nc * 128 + nb * 13, dtype=torch.float32, device=dev)
        ws = net._workspace(dev.index, lib.cns_forward_workspace_bytes(C.byref(b)))
        hid = out[:nc * 128]; vec = out[nc * 128:]; 
        stream = _lib.stream_ptr(dev)
        T("workspace_bytes", lambda: lib.cns_forward_workspace_bytes(C.byref(b)))
        T("ctypes forward (C++ launches)", lambda: lib.cns_graphvs_forward(h.raw, C.byref(b), None, _lib.ptr(vec), _lib.ptr(vec), _lib.ptr(hid), None, _lib.ptr(ws), ws.numel(), None, 0, stream))
        T("net(batch) total host", lambda: net(batch, None))
