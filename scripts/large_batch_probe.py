"""Sizes far above the benchmark batch: B graphs x 512 keypoints (B = 4096, 8192, 16384 by default - 8.4 M keypoints, 300 k
clusters in the last one) in every precision mode.  Checks: finite outputs; graphs [k*1024, (k+1)*1024) of the big batch
equal the same 1024 graphs run as their own batch (softmax-piece re-association only: 1e-5 fp32 / 5e-3 16-bit); 6 graphs spread
over the batch against the fp64 oracle.  usage: python scripts/large_batch_probe.py [B ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import cns_oracle
from cns_b200 import Batch, ckpt, synth

sizes = [int(a) for a in sys.argv[1:]] or [4096, 8192, 16384]
sd = ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns.pth"))
rel = lambda a, b: float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max().clamp(min=1e-30))
bad = 0
scenes = synth.make_scenes(max(sizes), 512, seed=7, n_layouts=32)
frames = [s.frame(0.1) for s in scenes]
for prec, tol_o, tol_s in (("fp32", 1e-4, 1e-5), ("f16", 2e-2, 5e-3), ("bf16", 2e-2, 2e-2)):
    net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), "cuda:0", precision=prec)
    for B in sizes:
        big = Batch.from_data_list(frames[:B]).to("cuda:0")
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with torch.no_grad():
            p = net(big, None, check=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        ok = bool(torch.isfinite(p[0]).all() and torch.isfinite(p[2]).all())
        worst_sub = 0.0
        for k in sorted({0, B // 2048, B // 1024 - 1}):
            with torch.no_grad():
                q = net(Batch.from_data_list(frames[k * 1024:(k + 1) * 1024]).to("cuda:0"), None)
            worst_sub = max(worst_sub, rel(q[0], p[0][k * 1024:(k + 1) * 1024]))
        pick = [0, 1, B // 3, B // 2 + 1, B - 2, B - 1]
        ov, on, oh = cns_oracle.forward(sd, Batch.from_data_list([frames[i] for i in pick]), None, dtype=torch.float64)
        e_o = max(rel(p[0][pick], ov), rel(p[1][pick], on))
        good = ok and worst_sub < tol_s and e_o < tol_o
        bad += 0 if good else 1
        print("%-5s B=%6d (%d keypoints, %d clusters): %s  first call %.1f ms  sub-batch %.1e  oracle %.1e" %
              (prec, B, big.x_cur.shape[0], int(big.num_clusters.sum()), "ok" if good else "FAIL", dt * 1e3, worst_sub, e_o), flush=True)
        del big, p
sys.exit(1 if bad else 0)
