"""Time the encoder edge kernel alone (cns_profile_begin/end) on the bench workload."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from cns_b200 import _lib, ckpt
lib = _lib.load()
dev = torch.device("cuda:0")
net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), dev, precision=os.environ.get("CNS_PRECISION", "f16"))
b = bench.build_batches(1024, 512, 2022, 1)[0].to(dev)
h = net._handle(dev)
with torch.no_grad():
    for _ in range(5): net(b, None)
    torch.cuda.synchronize()
    _lib.check(lib.cns_profile_begin(h.raw), "b")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): net(b, None)
    e1.record(); torch.cuda.synchronize()
    kms, kn, ln = C.c_float(), C.c_int(), C.c_longlong()
    _lib.check(lib.cns_profile_end(h.raw, C.byref(kms), C.byref(kn), C.byref(ln)), "e")
print("dbg=%s step %.4f ms  enc_edge %.4f ms" % (os.environ.get("CNS_TC_DBG"), e0.elapsed_time(e1) / 30, kms.value / kn.value))
