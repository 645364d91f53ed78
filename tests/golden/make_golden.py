"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on the
PyG shim (oracle/pyg_shim).  Run in the build container only:

    python tests/golden/make_golden.py

Fixtures:
  graphgen_*.pt : target points (f64) + everything GraphGenerator derives from them + per-frame
                  masked graphs, produced by the reference's cns/midend/graph_gen.py with real
                  scikit-learn, `np.random.seed(s)` right before construction
                  (mirrors demo_sim_Erender.py:42).
  forward_*.pt  : batched GraphData fields (from the reference GraphGenerator + shim Batch),
                  then reference GraphVS.forward / postprocess outputs with
                  checkpoints/cns_state_dict.pth over 3 recurrent frames (hidden carried).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_import  # noqa: E402

gv, gg, fn = ref_import.load()
from torch_geometric.data import Batch  # noqa: E402  (the shim)
from cns_b200 import synth  # noqa: E402  (scene sampler only; graphs come from the reference)

FIELDS = ["x_cur", "x_tar", "pos_cur", "pos_tar", "l0_to_l1_edge_index_j_cur", "l0_to_l1_edge_index_i_cur",
          "l1_dense_edge_index_cur", "l1_dense_edge_index_tar", "cluster_mask", "cluster_centers_index",
          "belong_cluster_index", "node_mask", "num_clusters", "new_scene", "tPo_norm", "batch"]


def graphgen_case(seed, n):
    rng = np.random.default_rng(seed)
    pts3 = synth.sample_points(rng, n)
    xy = synth.project_norm(synth.sample_target_pose(rng), pts3)
    np.random.seed(seed)
    g = gg.GraphGenerator(xy)
    frames = []
    for frac in (0.0, 0.1, 0.5, 0.9):
        k = int(round(len(xy) * frac))
        missing = np.sort(rng.choice(len(xy), size=k, replace=False)) if k else None
        _, l1, l0l1, nmask, cmask = g._get_graph(missing)
        frames.append(dict(missing=missing, l1=l1.copy(), l0l1=l0l1.copy(), node_mask=nmask.copy(),
                           cluster_mask=cmask.copy()))
    return dict(seed=seed, points=xy, yhat=g.yhat, cluster_ids=g.cluster_ids,
                counts=g.cluster_points_count, centers=g.cluster_centers_index,
                belong=g.belong_cluster_index, l0_to_l1=g.l0_to_l1_edge_index,
                l1_dense=g.l1_dense_edge_index, node2j=g.node2j_index, to_center=g.to_center, frames=frames)


def forward_case(seed, sizes, missing_frac, steps=3):
    sd = torch.load(os.path.join(ROOT, "checkpoints", "cns_state_dict.pth"), weights_only=True)
    net = gv.GraphVS(2, 2, 128, regress_norm=True)
    net.load_state_dict(sd)
    net.eval()
    rng = np.random.default_rng(seed)
    scenes = []
    for n in sizes:
        pts3 = synth.sample_points(rng, n)
        tar = synth.sample_target_pose(rng)
        cur = synth.sample_initial_pose(rng)
        xy = synth.project_norm(tar, pts3)
        np.random.seed(seed)
        scenes.append(dict(pts=pts3, tar=tar, cur=cur, gen=gg.GraphGenerator(xy),
                           dist=synth.scene_distance(tar, pts3)))
    hidden, out = None, []
    for step in range(steps):
        datas = []
        for sc in scenes:
            n = len(sc["pts"])
            k = int(round(n * missing_frac))
            missing = np.sort(rng.choice(n, size=k, replace=False)) if k else None
            d = sc["gen"].get_data(synth.project_norm(sc["cur"], sc["pts"]), missing_node_indices=missing)
            d.set_distance_scale(sc["dist"])
            if step == 0:
                d.start_new_scene()
            datas.append(d)
        batch = Batch.from_data_list(datas) if len(datas) > 1 else datas[0]
        with torch.no_grad():
            raw = net(batch, hidden)
            vel = net.postprocess(raw, batch)
        hidden = raw[2]
        fields = {k: getattr(batch, k) for k in FIELDS if k in batch}
        out.append(dict(fields=fields, num_nodes=int(batch.num_nodes), vel_si_vec=raw[0].clone(),
                        vel_si_norm=raw[1].clone(), hidden=raw[2].clone(), vel=vel.clone()))
        for sc, v in zip(scenes, vel.numpy()):
            sc["cur"] = synth.integrate(sc["cur"], v * 2)
    return dict(seed=seed, sizes=list(sizes), steps=out)


if __name__ == "__main__":
    for seed, n in ((0, 12), (1, 40), (2, 128), (3, 300)):
        torch.save(graphgen_case(seed, n), os.path.join(HERE, "graphgen_s%d_n%d.pt" % (seed, n)))
    torch.save(forward_case(11, (48,), 0.0), os.path.join(HERE, "forward_single.pt"))
    torch.save(forward_case(12, (10, 24, 100, 260), 0.15), os.path.join(HERE, "forward_ragged.pt"))
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
