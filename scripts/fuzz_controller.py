"""Randomised soak of the batch-1 servo path (GraphVSController on host GraphData -> cns_servo_step_host: staged copy in,
hidden state and per-target terms kept on the device between frames): random targets of 4..600 keypoints, the target
switched in the middle of a sequence (new scene, possibly a different size), 0..100 % of the keypoints missing per frame,
whole clusters hidden and back.  Every frame against the fp64 oracle carried alongside.  Exits 1 on a miss.
usage: python scripts/fuzz_controller.py [n_sequences]"""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import cns_oracle
from cns_b200 import ckpt, synth
from cns_b200.controller import GraphVSController

TOL = {"fp32": 1e-4, "f16": 2e-2}


def run(n_seq=20, precisions=("fp32", "f16"), verbose=True):
    path = os.path.join(ROOT, "checkpoints", "cns_state_dict.pth")
    sd = ckpt.read_state_dict(path)
    bad = 0
    for prec in precisions:
        ctl = GraphVSController(path, "cuda:0", precision=prec)
        worst, where = 0.0, None
        for q in range(n_seq):
            rng = np.random.default_rng(9000 + q); np.random.seed(9000 + q)
            hid, scene = None, None
            for t in range(int(rng.integers(4, 12))):
                if scene is None or rng.random() < 0.2:              # a new target (maybe of another size)
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        scene = synth.Scene(rng, int(rng.choice([4, 9, 16, 17, 33, 64, 150, 300, 600])))
                    hid, fresh = None, True
                gen, n = scene.generator, len(scene.points)
                labels = np.asarray(gen.belong_cluster_index)
                frac = float(rng.choice([0.0, 0.1, 0.3, 0.7, 0.95]))
                missing = set(rng.choice(n, size=int(n * frac), replace=False).tolist())
                if rng.random() < 0.3:                               # hide one whole cluster as well
                    missing |= set(np.nonzero(labels == int(rng.integers(0, gen.num_clusters)))[0].tolist())
                if len(missing) >= n:                                # the reference needs at least one visible keypoint
                    missing.discard(int(rng.integers(0, n)))
                miss = np.array(sorted(missing), dtype=np.int64)
                d = gen.get_data(synth.project_norm(scene.cur_wcT, scene.points), missing_node_indices=miss if len(miss) else None)
                d.set_distance_scale(scene.tPo_norm)
                if fresh:
                    d.start_new_scene(); fresh = False
                vel = ctl(d)
                ov, on, hid = cns_oracle.forward(sd, d, hid, dtype=torch.float64)
                ovel = cns_oracle.postprocess((ov, on), d.tPo_norm.reshape(1))[0].numpy()
                err = float(np.abs(vel - ovel).max() / max(np.abs(ovel).max(), 1e-12))
                if not np.isfinite(vel).all() or err > TOL[prec]:
                    bad += 1
                    print("%s sequence %d frame %d: n=%d missing %d  err %.2e" % (prec, q, t, n, len(miss), err))
                if err > worst:
                    worst, where = err, "sequence %d frame %d: %d keypoints, %d missing, max|vel| %.2f" % (q, t, n, len(miss), float(np.abs(ovel).max()))
                scene.step(ovel * 2)                                 # both modes walk the same trajectory
        if verbose:
            print("%-5s %d sequences, worst %.2e (bar %.0e) at %s" % (prec, n_seq, worst, TOL[prec], where), flush=True)
    return bad


if __name__ == "__main__":
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 20) else 0)
