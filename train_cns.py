#!/usr/bin/env python
"""Training entry point with the reference's CLI (`train_cns.py:17-31`: --name --hdim --batch-size
--epochs --early-stop --init-lr --weight-decay --device --load --save --long) on top of the B200 path:
GraphVS forward/backward through libcns_sm100, closed-loop synthetic point scenes (cns_b200/synth.py
restates the reference's PointEnv distribution), short-sequence truncated BPTT
(cns/train_gvs_short_seq.py:24-53: window 8, 10 % of the scenes driven by the ground-truth policy) or
long-sequence BPTT (cns/train_gvs_long_seq.py:20-53: mean loss over --long-steps frames), gradient
clipping at 10 and the two AdamW groups.

Data parallel: launch with `python -m torch.distributed.run --nproc-per-node N train_cns.py ...`; every
rank simulates and trains on its own `batch-size` scenes, gradients are averaged with one NCCL
all-reduce of the flat bucket per update.
"""
import argparse
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from cns_b200 import Batch, GraphVS, ckpt, synth
from cns_b200.train import TBPTTStepper


def get_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--name", type=str, default="cns_b200")
    ap.add_argument("--hdim", type=int, default=128)
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--epochs", type=int, default=50)
    ap.add_argument("--early-stop", type=int, default=40)
    ap.add_argument("--init-lr", type=float, default=5e-4)
    ap.add_argument("--weight-decay", type=float, default=1e-4)
    ap.add_argument("--device", type=str, default="cuda:0")
    ap.add_argument("--load", type=str, default=None)
    ap.add_argument("--save", action="store_true")
    ap.add_argument("--long", action="store_true", help="long-sequence stage (one backward per --long-steps frames)")
    ap.add_argument("--long-steps", type=int, default=200)
    ap.add_argument("--updates-per-epoch", type=int, default=50)
    ap.add_argument("--max-points", type=int, default=512)
    ap.add_argument("--fused", dest="fused", action="store_true", default=True,
                    help="fused loss and clip+AdamW kernels (cns_objectives / cns_clip_adamw); the default")
    ap.add_argument("--no-fused", dest="fused", action="store_false", help="torch loss + torch AdamW objects")
    ap.add_argument("--seed", type=int, default=2022)
    ap.add_argument("--device-env", action="store_true",
                    help="keep the closed-loop scenes on the GPU (cns_b200.sim_device.DeviceSceneBatch): batched projection, "
                         "supervisor, pose integration and per-frame graph masking instead of B python scenes per step")
    ap.add_argument("--refresh-every", type=int, default=0, help="--device-env --reinit scene: new scene geometry every this many frames")
    ap.add_argument("--no-aug", dest="aug", action="store_false", default=True,
                    help="--device-env: plain scenes (uniform drop-out, whole-scene tPo_norm) instead of the reference's "
                         "training augmentations (mismatch, jitter, area-shift drop-out; sim/dataset.py:88-141)")
    ap.add_argument("--reinit", choices=["all", "scene"], default="all",
                    help="--device-env: 'all' = the reference's loaders (every scene re-initialised together once more than "
                         "half need it, sim/dataset.py:242-246), 'scene' = scenes restart one by one")
    return ap.parse_args()


class SceneBatch(object):
    """`batch_size` closed-loop point scenes: observation -> GraphData, action -> pose integration,
    re-initialisation when a scene converges, times out or leaves the workspace
    (sim/environment.py:80-127,189-214; sim/dataset.py:59-166)."""

    def __init__(self, batch_size, rng, max_points=512, max_steps=200):
        self.rng, self.max_points, self.max_steps = rng, max_points, max_steps
        self.scenes = [self._new() for _ in range(batch_size)]
        self.steps = [0] * batch_size
        self.fresh = [True] * batch_size

    def _new(self):
        return synth.Scene(self.rng, int(self.rng.integers(4, self.max_points + 1)))

    def get(self, device) -> Batch:
        frames = []
        for k, s in enumerate(self.scenes):
            d = s.frame(missing_frac=float(self.rng.uniform(0.0, 0.2)), with_labels=True)
            if self.fresh[k]:
                d.start_new_scene()
                self.fresh[k] = False
            frames.append(d)
        return Batch.from_data_list(frames).to(device)

    def feedback(self, vel: np.ndarray):
        for k, s in enumerate(self.scenes):
            s.step(vel[k])
            self.steps[k] += 1
            ang, mm = synth.pose_error(s.cur_wcT, s.tar_wcT)
            lost = not np.isfinite(s.cur_wcT).all() or np.linalg.norm(s.cur_wcT[:3, 3]) > 1.8 or s.cur_wcT[2, 2] > -0.1
            if (ang < 1.0 and mm < 2.0) or self.steps[k] > self.max_steps or lost:
                self.scenes[k], self.steps[k], self.fresh[k] = self._new(), 0, True


def main():
    args = get_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        args.device = "cuda:%d" % local
        dist.init_process_group("nccl", device_id=torch.device(args.device))
    if args.hdim != 128:
        raise SystemExit("--hdim %d: libcns_sm100's kernels are specialised for hidden_dim 128, the only configuration the "
                         "reference ships weights for (checkpoints/cns.pth, train_cns.py:56-61); see INTEGRATION.md" % args.hdim)
    dev = torch.device(args.device)
    np.random.seed(args.seed + rank)
    torch.manual_seed(0)                       # identical initial weights on every rank
    net = GraphVS(2, 2, args.hdim, regress_norm=True).to(dev)
    load = args.load if args.load and args.load.lower() != "none" else None        # the reference spells "none"
    if load:
        net.load_state_dict(ckpt.read_state_dict(load))       # state dict, cns.pth or a Trainer checkpoint
    net.train()
    stepper = TBPTTStepper(net, fused=args.fused, lr=args.init_lr, weight_decay=args.weight_decay)
    if load:
        try:          # resume the optimizer state when the file is a Trainer checkpoint (trainer.py:333-348)
            ckpt.load_checkpoint(load, net, stepper.optimizers, getattr(stepper, "fused_opt", None))
        except ValueError:
            pass
    if args.device_env:
        from cns_b200.sim_device import DeviceSceneBatch
        # short sequences: straight-line PBVS supervisor (sim/dataset.py:33); long: the hybrid law, rotating mismatches and
        # the smaller jitter of sim/dataset_long.py:21,81-97
        envs = DeviceSceneBatch(args.batch_size, np.random.default_rng(args.seed + rank), dev, args.max_points,
                                refresh_every=args.refresh_every, aug=args.aug, reinit=args.reinit, long_seq=args.long,
                                supervisor="pbvs_hybrid" if (args.long and args.aug) else "pbvs_straight")
    else:
        envs = SceneBatch(args.batch_size, np.random.default_rng(args.seed + rank), args.max_points)
    window = args.long_steps if args.long else 8
    best, not_improved = float("inf"), 0
    for epoch in range(args.epochs):
        t0, losses = time.time(), []
        for _ in range(args.updates_per_epoch):
            frames = []

            def rollout():
                for _ in range(window):
                    data = envs.observe() if args.device_env else envs.get(dev)
                    frames.append(data)
                    yield data
                    # closed loop: drive the simulators with the predicted velocity, 10 % with the teacher
                    with torch.no_grad():
                        pred = stepper.last_pred
                        vel = net.postprocess((pred[0].detach(), pred[1].detach()), data)
                        n_gt = max(1, int(vel.shape[0] * 0.1))
                        vel[:n_gt] = data.vel[:n_gt].to(vel)
                    envs.feedback(vel * 2 if args.device_env else (vel * 2).cpu().numpy())
            losses.append(float(stepper.step(rollout(), mean=args.long)) / (1 if args.long else window))
        mean_loss = float(np.mean(losses))
        if world > 1:
            t = torch.tensor([mean_loss], device=dev)
            dist.all_reduce(t)
            mean_loss = float(t) / world
        if rank == 0:
            print("[INFO] epoch %d | loss/frame %.5f | %.1f s | %.0f graph-frames/s" % (
                epoch, mean_loss, time.time() - t0,
                args.updates_per_epoch * window * args.batch_size * world / (time.time() - t0)))
            if args.save:
                # Trainer format (trainer.py:138-162): loadable by the unmodified reference
                ckpt.save_checkpoint(epoch, net, stepper.optimizers, mean_loss, mean_loss < best, not_improved,
                                     os.path.join("checkpoints", args.name), fused_opt=getattr(stepper, "fused_opt", None))
        if mean_loss < best:
            best, not_improved = mean_loss, 0
        else:
            not_improved += 1
            if not_improved % 5 == 0:          # halve the learning rate (train_gvs_short_seq.py:79-95)
                for opt in stepper.optimizers.values():
                    for gp in opt.param_groups:
                        gp["lr"] = max(gp["lr"] * 0.5, 1e-5)
                if args.fused:
                    stepper.fused_opt.lr = max(stepper.fused_opt.lr * 0.5, 1e-5)
            if not_improved >= args.early_stop:
                break
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
