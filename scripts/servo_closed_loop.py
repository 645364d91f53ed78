"""Batched closed-loop servo entirely on the GPU: B simulated scenes (cns_b200.sim_device) -> per-frame graph masking
-> GraphVS forward (hidden state carried) -> camera motion, each scene running until ITS PixelStopPolicy fires
(cns_b200.stop_policy.BatchedPixelStop, the reference's stop rule: waiting_time 0.8 s, conduct_thresh 0.01,
cns/benchmark/benchmark.py:220).  Reports how many scenes stopped, the pose error at the stop and the frame rate.

    python scripts/servo_closed_loop.py [B=256] [N=512] [precision=f16] [max_frames=500] [cache]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from cns_b200 import BatchedPixelStop, GraphVS, ckpt, sim_device, synth


def run(B=256, N=512, precision="f16", max_frames=500, layouts=32, missing=0.1, gain=2.0, seed=11, device="cuda:0",
        cache=False):
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    env = sim_device.DeviceSceneBatch(1, rng, device=device, max_points=8, missing=(missing, missing))
    base = [synth.Scene(rng, N) for _ in range(min(layouts, B))]
    env.B = B
    env._build([synth.Scene(rng, N, generator=base[i % len(base)].generator, points=base[i % len(base)].points,
                            tar_wcT=base[i % len(base)].tar_wcT) for i in range(B)])
    net = GraphVS(2, 2, 128).to(device)
    net.load_state_dict(ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns.pth")))
    net.eval()
    net.set_precision(precision)
    counts = [len(s.points) for s in env.scenes]
    stop_rule = BatchedPixelStop(counts, waiting_time=0.8, conduct_thresh=0.01, device=device)
    active = torch.ones(B, dtype=torch.bool, device=device)
    stop_frame = np.full(B, -1)
    ang0, mm0 = sim_device.pose_error(env.cur_wcT, env.tar_wcT)
    hidden, tcache = None, None
    torch.cuda.synchronize()
    t0 = time.time()
    frames = 0
    with torch.no_grad():
        for k in range(max_frames):
            batch = env.observe()
            if cache and tcache is None:      # per-target terms (and, in f16 mode, per-keypoint target-side logits / values) once
                tcache = net.build_target_cache(batch)
            vec, nrm, hidden = net(batch, hidden, target_cache=tcache)
            vel = net.postprocess((vec, nrm), batch)
            stop = stop_rule(batch.pos_cur, batch.pos_tar, batch.node_mask, k * synth.DT)
            newly = stop & (stop_frame < 0)
            stop_frame[newly] = k
            active &= ~torch.from_numpy(stop).to(device)
            env.cur_wcT = sim_device.integrate(env.cur_wcT, (vel.double() * gain) * active[:, None])
            frames += 1
            if not bool(active.any()):
                break
    torch.cuda.synchronize()
    wall = time.time() - t0
    ang, mm = sim_device.pose_error(env.cur_wcT, env.tar_wcT)
    ang, mm = ang.cpu().numpy(), mm.cpu().numpy()
    done = stop_frame >= 0
    q = lambda a: [float(np.percentile(a, p)) for p in (50, 90, 100)] if len(a) else None
    return {"scenes": B, "keypoints": N, "precision": precision, "target_cache": bool(cache), "missing": missing, "gain": gain,
            "frames_run": frames,
            "stopped": int(done.sum()), "stop_frame_p50_p90_max": q(stop_frame[done]),
            "initial_rot_deg_p50_p90_max": q(ang0.cpu().numpy()), "initial_trans_mm_p50_p90_max": q(mm0.cpu().numpy()),
            "final_rot_deg_p50_p90_max": q(ang[done]), "final_trans_mm_p50_p90_max": q(mm[done]),
            "wall_s": wall, "scene_frames_per_s": B * frames / wall}


if __name__ == "__main__":
    a = sys.argv[1:]
    B = int(a[0]) if len(a) > 0 else 256
    N = int(a[1]) if len(a) > 1 else 512
    prec = a[2] if len(a) > 2 else "f16"
    mf = int(a[3]) if len(a) > 3 else 500
    print(json.dumps(run(B, N, prec, mf, cache=len(a) > 4 and a[4] == "cache")))
