"""Stage timeline of tile 0 of backbone_fused_kernel (needs a -DBF_TIMELINE build: scripts/build_variant.sh bftl -DBF_TIMELINE
cns_b200/csrc/backbone_fused.cu; CNS_LIB=build/variants/bftl.so python scripts/bf_timeline.py)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from cns_b200 import _lib, ckpt
lib = _lib.load()
dev = torch.device("cuda:0")
net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), dev, precision="f16")
b = bench.build_batches(1024, 512, 2022, 1)[0].to(dev)
with torch.no_grad():
    for _ in range(6):
        pred = net(b, None)
    hid = pred[2]
    for _ in range(3):
        net(b, hid)
    torch.cuda.synchronize()
buf = (C.c_longlong * 64)()
assert lib.cns_bf_timeline(buf) == 0
t = list(buf)
names = {0: "start", 1: "inputs loaded + A written", 31: "publish 0 (out_enc A)"}
for g in range(8):
    names[2 + 2 * g] = "GEMM %d done (wait returned)" % g
    names[3 + 2 * g] = "epilogue %d done -> publish" % g
names[17] = "end"
order = [0, 1, 31] + [i for g in range(8) for i in (2 + 2 * g, 3 + 2 * g)]
order[-1] = 17
prev = t[0]
for i in order:
    print("%-34s +%7d cycles  (t=%7d)" % (names[i], t[i] - prev, t[i] - t[0]))
    prev = t[i]

sub = {40: "perconv(c1): start", 41: "Q -> staging", 42: "barrier", 43: "column max + broadcast", 44: "barrier",
       45: "P + max -> staging", 46: "barrier", 15: "GraphNorm + A tile (-> publish)"}
prev = t[40]
for i in (40, 41, 42, 43, 44, 45, 46, 15):
    print("  %-32s +%7d cycles" % (sub[i], t[i] - prev))
    prev = t[i]

print("tile 0: G = %d graphs, R = %d rows" % (t[60], t[61]))
for base, name in ((47, "unit 0"), (50, "unit 16")):
    if t[base]:
        print("  GraphNorm %s: stats pass done +%d (from barrier), scale/shift +%d, normalise + A +%d" % (
            name, t[base] - t[46], t[base + 1] - t[base], t[base + 2] - t[base + 1]))
