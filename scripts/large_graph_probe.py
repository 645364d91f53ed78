"""Graphs far larger than the benchmark's: a batch of targets with 3000 / 8192 / 40 / 5000 keypoints (several hundred
clusters each, so they do not fit the fused backbone's 128-row whole-graph tiles and take the per-layer path), two
recurrent frames, 10 % missing keypoints, every precision mode against the fp64 oracle; plus graphs with 129 / 150 / 400
given clusters."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import cns_oracle
from cns_b200 import Batch, GraphGenerator, ckpt, synth

sd = ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns.pth"))
rel = lambda a, b: float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max().clamp(min=1e-30))
rng = np.random.default_rng(11); np.random.seed(11)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    scenes = [synth.Scene(rng, n) for n in (3000, 8192, 40, 5000)]
    # clusters given as Voronoi cells of many exemplar keypoints: 150 / 400 / 129 clusters per graph, beyond the 128 rows of
    # a whole-graph backbone tile (the affinity-propagation layouts above stay below 100 clusters)
    for n, k in ((2000, 150), (4000, 400), (600, 129)):
        pts = synth.sample_points(rng, n)
        tar = synth.sample_target_pose(rng)
        gen = GraphGenerator.from_exemplars(synth.project_norm(tar, pts), rng.choice(n, size=k, replace=False))
        scenes.append(synth.Scene(rng, n, generator=gen, points=pts, tar_wcT=tar))
print("clusters per graph:", [s.generator.num_clusters for s in scenes])
seq = []
for t in range(2):
    seq.append(Batch.from_data_list([s.frame(0.1) for s in scenes]))
    for s in scenes:
        s.step(synth.pbvs_straight(s.cur_wcT, s.tar_wcT) * 2)
ho, ref = None, []
for b in seq:
    ov, on, ho = cns_oracle.forward(sd, b, ho, dtype=torch.float64)
    ref.append((ov, on, ho))
bad = 0
for prec, tol in (("fp32", 1e-4), ("f16", 2e-2), ("bf16", 2e-2)):
    net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), "cuda:0", precision=prec)
    h, worst = None, 0.0
    for b, (ov, on, oh) in zip(seq, ref):
        with torch.no_grad():
            p = net(b.to("cuda:0"), h, check=True)
        h = p[2]
        worst = max(worst, rel(p[0], ov), rel(p[1], on), rel(p[2], oh))
    bad += 0 if worst < tol else 1
    print("%-5s worst %.2e (bar %.0e) %s" % (prec, worst, tol, "ok" if worst < tol else "FAIL"), flush=True)
sys.exit(1 if bad else 0)
