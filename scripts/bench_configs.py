"""Measures the BASELINE.json configurations other than the bench.py headline (configs[1]):
  C1  batch-1 recurrent servo latency (p50/p90), N=512, through GraphVSController; CPU oracle beside it
  C3  short-sequence training update (8 frames TBPTT + clip + 2x AdamW), bs 256, N~U{4..512} and N=512
  C4  long-sequence update (200 recurrent frames, one backward), bs 16
  C5  keypoint-count sweep 128..8192, 64 k graphs per keypoint count, forward only, Affinity Propagation on the device for
      every layout; under torchrun the graphs are split over the ranks
Writes one JSON document to stdout / profiles/configs_<tag>.json.  GPU box only."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
from cns_b200 import GraphVS, ckpt, synth
from cns_b200.controller import GraphVSController
from cns_b200.data import Batch
from cns_b200.graph_gen import GraphGenerator
from cns_b200.train import TBPTTStepper

CK = os.path.join(ROOT, "checkpoints", "cns.pth")
which = set(sys.argv[1].split(",")) if len(sys.argv) > 1 else {"c1", "c2", "c3", "c4", "c5"}
out = {}


def ev_time(fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


if "c1" in which:
    import cns_oracle
    res = {}
    for prec in ("fp32", "f16"):
        rng = np.random.default_rng(0); np.random.seed(0)
        scene = synth.Scene(rng, 512)
        ctl = GraphVSController(CK, "cuda:0", precision=prec)
        lat = []
        for t in range(220):
            d = scene.frame(0.1)
            if t == 0:
                d.start_new_scene()
            t0 = time.perf_counter()
            vel = ctl(d)                      # host GraphData in, numpy velocity out (sync)
            lat.append((time.perf_counter() - t0) * 1e3)
            scene.step(vel * 2)
        lat = np.array(lat[20:])
        res[prec] = {"p50_ms": float(np.percentile(lat, 50)), "p90_ms": float(np.percentile(lat, 90)),
                     "final_pose_error_deg_mm": synth.pose_error(scene.cur_wcT, scene.tar_wcT)}
    sd = ckpt.read_state_dict(CK)
    rng = np.random.default_rng(0); np.random.seed(0)
    scene = synth.Scene(rng, 512)
    torch.set_num_threads(os.cpu_count())
    lat, hid = [], None
    with torch.no_grad():
        for t in range(70):
            d = scene.frame(0.1)
            t0 = time.perf_counter()
            vec, nrm, hid = cns_oracle.forward(sd, d, hid)
            vel = cns_oracle.postprocess((vec, nrm), d.tPo_norm.reshape(1))[0].numpy()
            lat.append((time.perf_counter() - t0) * 1e3)
            scene.step(vel * 2)
    lat = np.array(lat[20:])
    res["cpu_oracle"] = {"p50_ms": float(np.percentile(lat, 50)), "p90_ms": float(np.percentile(lat, 90)),
                         "cores": torch.get_num_threads()}
    out["C1_batch1_latency_N512"] = res

if "c0" in which:
    import warnings
    from sklearn.cluster import AffinityPropagation
    from cns_b200.cluster import affinity_propagation_device
    res = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for n, reps in ((128, 5), (512, 3), (1024, 2), (2048, 1)):
            rng = np.random.default_rng(n)
            X = synth.project_norm(synth.sample_target_pose(rng), synth.sample_points(rng, n))
            np.random.seed(0); t0 = time.perf_counter()
            for _ in range(reps): ref = AffinityPropagation(damping=0.9).fit(X)
            t_sk = (time.perf_counter() - t0) / reps
            np.random.seed(0); affinity_propagation_device([X]); torch.cuda.synchronize()
            np.random.seed(0); t0 = time.perf_counter()
            for _ in range(reps): lab = affinity_propagation_device([X])[0]
            t_dev = (time.perf_counter() - t0) / reps
            np.random.seed(0); lab = affinity_propagation_device([X])[0]
            np.random.seed(0); ref = AffinityPropagation(damping=0.9).fit(X)
            res[str(n)] = {"sklearn_s": t_sk, "device_s": t_dev, "speedup": t_sk / t_dev, "n_iter": int(ref.n_iter_),
                           "labels_identical": bool(np.array_equal(lab, ref.labels_))}
        rng = np.random.default_rng(7)
        sets = [synth.project_norm(synth.sample_target_pose(rng), synth.sample_points(rng, 512)) for _ in range(64)]
        np.random.seed(1); t0 = time.perf_counter(); labs = affinity_propagation_device(sets); t_b = time.perf_counter() - t0
        np.random.seed(1); t0 = time.perf_counter(); refs = [AffinityPropagation(damping=0.9).fit_predict(X) for X in sets[:8]]
        t_s8 = time.perf_counter() - t0
        res["batch_64x512"] = {"device_s_total": t_b, "sklearn_s_per_set": t_s8 / 8, "speedup": (t_s8 / 8 * 64) / t_b,
                               "labels_identical_first8": bool(all(np.array_equal(a, b) for a, b in zip(labs[:8], refs)))}
    out["C0_affinity_propagation"] = res

if "c3" in which or "c4" in which:
    def make_windows(bs, n_frames, n_points, seed):
        rng = np.random.default_rng(seed); np.random.seed(seed)
        sizes = [n_points] * bs if n_points else list(rng.integers(4, 513, bs))
        base = {}
        scenes = []
        for n in sizes:                      # cluster each distinct layout size once (AP is host-side)
            key = int(n) if len(base) >= 32 else None
            if key is not None and key in base:
                src = base[key]
                scenes.append(synth.Scene(rng, int(n), generator=src.generator, points=src.points, tar_wcT=src.tar_wcT))
            else:
                sc = synth.Scene(rng, int(n)); base[int(n)] = sc; scenes.append(sc)
        frames = []
        for t in range(n_frames):
            fr = [s.frame(0.1, with_labels=True) for s in scenes]
            if t == 0:
                for f in fr: f.start_new_scene()
            frames.append(Batch.from_data_list(fr).to("cuda:0"))
            for s in scenes:
                s.step(synth.pbvs_straight(s.cur_wcT, s.tar_wcT) * 2)
        return frames

if "c3" in which:
    res = {}
    for name, npts in (("N_uniform_4_512", 0), ("N512", 512)):
        frames = make_windows(256, 8, npts, 11)
        net = GraphVS(2, 2, 128).to("cuda:0"); net.load_state_dict(ckpt.read_state_dict(CK)); net.train()
        stepper = TBPTTStepper(net)
        ms = ev_time(lambda: stepper.step(frames), 5, warm=2)
        with torch.no_grad():
            net.eval(); fwd_ms = ev_time(lambda: [net(f, None) for f in frames], 5, warm=2); net.train()
        res[name] = {"ms_per_update": ms, "graph_frames_per_s_fwd_bwd": 256 * 8 / ms * 1e3, "fwd_only_ms_8_frames_fp32": fwd_ms,
                     "keypoints_per_frame": int(frames[0].x_cur.shape[0])}
    out["C3_short_seq_train_bs256_tbptt8"] = res

if "c4" in which:
    frames = make_windows(16, 200, 0, 13)
    net = GraphVS(2, 2, 128).to("cuda:0"); net.load_state_dict(ckpt.read_state_dict(CK)); net.train()
    stepper = TBPTTStepper(net)
    ms = ev_time(lambda: stepper.step(frames, mean=True), 3, warm=1)
    out["C4_long_seq_train_bs16_200frames"] = {"ms_per_update": ms, "graph_frames_per_s_fwd_bwd": 16 * 200 / ms * 1e3}

if "c2" in which:
    # headline workload (bench.py) with the recurrent hidden state carried instead of hidden=None
    import bench
    net = ckpt.load_model(CK, "cuda:0", precision="f16")
    batches = [b.to("cuda:0") for b in bench.build_batches(1024, 512, 2022, 2)]
    tc = net.build_target_cache(batches[0])
    res = {}
    with torch.no_grad():
        hid = net(batches[0], None, target_cache=tc)[2]
        st = {"h": hid, "i": 0}
        def carried():
            st["h"] = net(batches[st["i"] & 1], st["h"], target_cache=tc)[2]; st["i"] += 1
        def first():
            net(batches[st["i"] & 1], None, target_cache=tc); st["i"] += 1
        for name, fn in (("hidden_none", first), ("hidden_carried", carried)):
            ms = ev_time(fn, 50)
            res[name] = {"ms": ms, "graphs_per_s": 1024 / ms * 1e3}
    out["C2_batched_inference_1024x512_f16"] = res

if "c5" in which:
    # BASELINE configs[4] as written: keypoint-count sweep 128..8192, 64 k graphs per keypoint count, forward only, split
    # over the ranks (torchrun: every rank takes 64 k / world graphs).  Every layout is clustered by Affinity Propagation
    # on the device (cns_b200/cluster.py), also above 1024 keypoints where scikit-learn needs minutes per target.
    # A rank holds one resident chunk of graphs (<= ~2 M keypoints, distinct camera poses, 4 layouts) and passes over it
    # until its share of the 64 k graphs is done; all passes are timed.
    import torch.distributed as dist
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    TOTAL = 65536
    res = {}
    net = ckpt.load_model(CK, dev, precision=os.environ.get("CNS_PRECISION", "f16"))
    sweep = [int(x) for x in os.environ.get("CNS_C5_N", "128,256,512,1024,2048,4096,8192").split(",")]
    for n in sweep:
        share = TOTAL // world
        chunk = max(8, min(share, (1 << 21) // n))
        passes = -(-share // chunk)
        rng = np.random.default_rng(n + rank); np.random.seed(n + rank)
        t0 = time.perf_counter()
        base = []
        for _ in range(4):
            pts3 = synth.sample_points(rng, n)
            tar = synth.sample_target_pose(rng)
            base.append((pts3, tar, GraphGenerator(synth.project_norm(tar, pts3))))
        t_cluster = (time.perf_counter() - t0) / 4
        frames = []
        for g in range(chunk):
            pts3, tar, gen = base[g % 4]
            frames.append(synth.Scene(rng, n, generator=gen, points=pts3, tar_wcT=tar).frame(0.1))
        batch = Batch.from_data_list(frames).to(dev)
        del frames
        with torch.no_grad():
            for _ in range(2):
                net(batch, None)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(passes):
                net(batch, None)
            b.record()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        done = passes * chunk * world
        res[str(n)] = {"graphs": done, "graphs_per_rank_chunk": chunk, "passes": passes, "ms_total": ms,
                       "graphs_per_s": done / ms * 1e3, "keypoints_per_s": done * n / ms * 1e3,
                       "clusters_per_graph": [int(b_[2].num_clusters) for b_ in base],
                       "clustering": base[0][2].cluster_method, "cluster_seconds_per_target": t_cluster}
        del batch
        torch.cuda.empty_cache()
    if rank == 0:
        out["C5_keypoint_sweep_64k_graphs_fwd_%s_%dgpu" % (os.environ.get("CNS_PRECISION", "f16"), world)] = res
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        sys.exit(0)

print(json.dumps(out, indent=1))
