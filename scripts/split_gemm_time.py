"""Times cns_linear_split / cns_weight_grad alone (CUDA events), E = keypoint->cluster edges of a 256-graph batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cns_b200 import _lib
lib = _lib.load(); dev = "cuda:0"
m = int(sys.argv[1]) if len(sys.argv) > 1 else 118000
flags = [int(a) for a in sys.argv[2:]] or [0]
x = torch.randn(m, 128, device=dev); wt = torch.randn(128, 128, device=dev) / 11
res = torch.randn(m, 128, device=dev); gate = torch.randn(m, 128, device=dev); y = torch.empty(m, 128, device=dev)
ws = torch.empty(lib.cns_linear_split_workspace_bytes(), dtype=torch.uint8, device=dev)
ws2 = torch.empty(lib.cns_weight_grad_workspace_bytes(), dtype=torch.uint8, device=dev)
dw = torch.empty(128, 128, device=dev); db = torch.empty(128, device=dev)
sp = _lib.stream_ptr(torch.device(dev))
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for fl in flags:
    for name, r, g in (("plain", None, None), ("res", res, None), ("res+gate", res, gate)):
        us = t(lambda: lib.cns_linear_split(_lib.ptr(x), m, _lib.ptr(wt), None, _lib.ptr(r), _lib.ptr(g), fl, _lib.ptr(y), _lib.ptr(ws), ws.numel(), sp))
        print("linear_split m=%d flags=%d %-9s %.1f us (incl. 3.7 us weight pack)" % (m, fl, name, us))
for tc in (1, 0):
    us = t(lambda: lib.cns_weight_grad(_lib.ptr(x), 128, _lib.ptr(res), 128, m, _lib.ptr(dw), _lib.ptr(db), 0, tc, _lib.ptr(ws2), ws2.numel(), sp))
    print("weight_grad m=%d tensor_cores=%d %.1f us (incl. reduce)" % (m, tc, us))
