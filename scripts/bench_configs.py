"""Measures the BASELINE.json configurations other than the bench.py headline (configs[1]):
  C1  batch-1 recurrent servo latency (p50/p90), N=512, through GraphVSController; CPU oracle beside it
  C3  short-sequence training update (8 frames TBPTT + clip + 2x AdamW), bs 256, N~U{4..512} and N=512
  C4  long-sequence update (200 recurrent frames, one backward), bs 16
  C5  keypoint-count sweep 128..8192, forward only, ~512k keypoints per batch
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
    res = {}
    net = ckpt.load_model(CK, "cuda:0", precision="f16")
    for n in (128, 256, 512, 1024, 2048, 4096, 8192):
        nb = max(8, 524288 // n)
        rng = np.random.default_rng(n); np.random.seed(n)
        n_layouts = 4
        base = []
        for _ in range(n_layouts):
            pts3 = synth.sample_points(rng, n)
            tar = synth.sample_target_pose(rng)
            xy = synth.project_norm(tar, pts3)
            if n <= 1024:
                gen = GraphGenerator(xy)
                how = "affinity propagation (host sklearn)"
            else:
                k = int(round(0.8 * np.sqrt(n)))
                gen = GraphGenerator.from_exemplars(xy, rng.choice(n, size=k, replace=False), device="cuda:0")
                how = "random exemplars + cns_label_nearest (device 1-NN)"
            base.append((pts3, tar, gen))
        frames = []
        for g in range(nb):
            pts3, tar, gen = base[g % n_layouts]
            sc = synth.Scene(rng, n, generator=gen, points=pts3, tar_wcT=tar)
            frames.append(sc.frame(0.1))
        batch = Batch.from_data_list(frames).to("cuda:0")
        with torch.no_grad():
            ms = ev_time(lambda: net(batch, None), 10)
        res[str(n)] = {"graphs": nb, "clusters_per_graph": gen.num_clusters, "ms": ms, "graphs_per_s": nb / ms * 1e3,
                       "keypoints_per_s": nb * n / ms * 1e3, "clustering": how}
    out["C5_keypoint_sweep_fwd_f16"] = res

print(json.dumps(out, indent=1))
