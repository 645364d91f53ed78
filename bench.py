#!/usr/bin/env python
"""bench.py - servo graphs/s of the CNS hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision fp32|f16|bf16]

Workload (BASELINE.json configs[1]): pretrained checkpoints/cns.pth, batched inference,
1024 synthetic servo graphs x 512 keypoints per GPU (weak scaling: every rank owns its own 1024
graphs; inference needs no collective).  One "step" = one GraphVS forward (+ fused postprocess)
over one 1024-graph batch.

  value : graphs/s with the batches already resident in HBM (CUDA events, max over ranks).
  e2e   : the same through the public API from HOST (pinned) memory, servo form (BatchedServo.step):
          the target-side graphs are per-scene state resident on the device (the reference's Midend
          builds them once per target), a step copies the current keypoint coordinates + visibility
          in, masks the graphs on the device, runs the forward and reads the (B,6) velocities back.
          e2e.full_graph: GraphVS.forward on a prebuilt host Batch copied whole every step.
  roofline     : the dominant kernel (encoder edge kernel), timed live with CUDA events on its
                 launching stream (cns_profile_begin/end).
  cpu_baseline : the oracle (CPU port of the reference forward) on a bounded sample of the same
                 workload, on this box's host cores.
`--impl reference` times that CPU implementation alone (the reference itself is Python on PyG,
which cannot be installed here; the port in oracle/ is the stand-in, kind "port").
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GRAPHS, N_POINTS, MISSING = 1024, 512, 0.1
N_ROTATE = 6          # distinct input batches cycled through so that the bytes a step reads exceed L2 (126 MB)
# algorithmic FLOPs of the encoder edge kernel per (edge, side): 5 GEMMs 128x128 + in_enc + pos_nn.0
FLOP_PER_EDGE_SIDE = 2 * (5 * 128 * 128 + 2 * 128 + 4 * 128)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("CNS_PRECISION", "f16"), choices=["fp32", "f16", "bf16"])
    ap.add_argument("--graphs", type=int, default=N_GRAPHS)
    ap.add_argument("--points", type=int, default=N_POINTS)
    ap.add_argument("--cpu-sample", type=int, default=64, help="graphs per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (NVML, ~2 ms period;
    falls back to polling nvidia-smi when pynvml is unavailable)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag = index, False
        self.sm, self.sm_max, self.reasons, self.source = [], None, set(), "nvml"
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv, self.source = None, "nvidia-smi"

    def _sample(self):
        if self.nv is not None:
            self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
            bits = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            self.reasons |= {n for n, b in self.BITS.items() if bits & b}
            time.sleep(0.002)
            return
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        parts = [x.strip() for x in out.strip().split(",")]
        if len(parts) >= 6:
            self.sm.append(float(parts[0]))
            self.sm_max = float(parts[1])
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            self.reasons |= {n for i, n in enumerate(names) if parts[2 + i].lower().startswith("active")}

    def run(self):
        while not self.stop_flag:
            try:
                self._sample()
            except Exception:
                time.sleep(0.01)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unavailable"], "samples": 0}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": self.source}


def build_batches(n_graphs, n_points, seed, n_batches, with_scenes=False):
    """`n_batches` host batches: the same scenes (clustered once) observed at successive frames."""
    from cns_b200 import synth
    from cns_b200.data import Batch
    scenes = synth.make_scenes(n_graphs, n_points, seed=seed, n_layouts=64)
    out = []
    for _ in range(n_batches):
        out.append(Batch.from_data_list([s.frame(MISSING) for s in scenes]))
        for s in scenes:    # move every camera a little so each batch holds different coordinates
            s.step(synth.pbvs_straight(s.cur_wcT, s.tar_wcT) * 2)
    return (out, scenes) if with_scenes else out


def tensor_bytes(batch):
    import torch
    return sum(v.numel() * v.element_size() for _, v in batch.items() if isinstance(v, torch.Tensor))


def cpu_reference_rate(n_graphs, n_points, steps, warmup, budget_s=25.0):
    """graphs/s of the CPU port of the reference forward (oracle/cns_oracle.py), all host threads."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cns_oracle
    from cns_b200 import ckpt
    torch.set_num_threads(os.cpu_count() or 1)
    sd = ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns.pth"))
    batch = build_batches(n_graphs, n_points, 2022, 1)[0]
    with torch.no_grad():
        for _ in range(max(1, warmup)):
            cns_oracle.forward(sd, batch, None)
        t0, done = time.perf_counter(), 0
        for _ in range(steps):
            cns_oracle.forward(sd, batch, None)
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
    return n_graphs * done / dt, dt / done * 1e3, done, torch.get_num_threads()


_JSON_FD = None


def claim_stdout():
    """stdout carries exactly ONE line, the JSON result: libraries that print to fd 1 (NCCL's version banner, ...)
    are sent to stderr for the rest of the run."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    args = parse()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    workload = "cns.pth batched inference, %d synthetic servo graphs x %d keypoints per GPU" % (args.graphs, args.points)

    if args.impl == "reference":
        if rank != 0:
            return
        rate, ms, done, cores = cpu_reference_rate(args.cpu_sample, args.points, args.steps, args.warmup, 120.0)
        sample = "%d of the %d graphs per step (N=%d keypoints, 10%% missing), fp32 torch CPU" % (
            args.cpu_sample, args.graphs, args.points)
        emit({
            "impl": "reference", "metric": "servo graphs/sec (fwd)", "value": rate, "unit": "graphs/s",
            "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "sample": sample},
            "cpu_baseline": {"value": rate, "unit": "graphs/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    import torch
    import torch.distributed as dist
    from cns_b200 import _lib, ckpt

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    net = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), dev, precision=args.precision)

    host, scenes = build_batches(args.graphs, args.points, 2022 + rank, N_ROTATE, with_scenes=True)
    host = [b.pin_memory() for b in host]
    resident = [b.to(dev) for b in host]
    n_edges = [int(b.l0_to_l1_edge_index_j_cur.numel()) for b in host]
    h2d = sum(v.numel() * v.element_size() for k, v in host[0].items()
              if isinstance(v, torch.Tensor) and k in net.INPUT_FIELDS)
    # per-batch tensors the forward reads
    read_per_step = sum(host[0][k].numel() * host[0][k].element_size() for k in (
        "x_cur", "pos_cur", "pos_tar", "x_tar", "l0_to_l1_edge_index_j_cur", "l0_to_l1_edge_index_i_cur",
        "l1_dense_edge_index_cur", "l1_dense_edge_index_tar", "cluster_mask"))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    handle = net._handle(dev)
    with torch.no_grad():
        # ---------------- device-resident throughput ("value") + live kernel timing ----------------
        # the stateless call of the reference API, GraphVS.forward(Batch, hidden=None): nothing is kept from one
        # step to the next
        for i in range(args.warmup):
            net(resident[i % N_ROTATE], None)
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        _lib.check(lib.cns_profile_begin(handle.raw), "profile_begin")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            net(resident[i % N_ROTATE], None)
        e1.record()
        barrier()
        kms, kn, launches = C.c_float(), C.c_int(), C.c_longlong()
        _lib.check(lib.cns_profile_end(handle.raw, C.byref(kms), C.byref(kn), C.byref(launches)), "profile_end")
        sampler.stop_flag = True
        total_ms = max_over_ranks(e0.elapsed_time(e1))
        value = args.graphs * world * args.steps / (total_ms * 1e-3)

        # the same loop with a per-target cache (GraphVS.build_target_cache): the rotating batches are the same
        # scenes at successive frames, and what depends on the targets alone (per-cluster attention term, cluster
        # positions, tile packing) is built once, as the reference's Midend keeps one GraphGenerator per target
        # image.  Reported beside `value`, not as `value`.
        tcache = net.build_target_cache(resident[0])
        for i in range(args.warmup):
            net(resident[i % N_ROTATE], None, target_cache=tcache)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for i in range(args.steps):
            net(resident[i % N_ROTATE], None, target_cache=tcache)
        c1.record()
        barrier()
        cached_ms = max_over_ranks(c0.elapsed_time(c1))

        # ---------------- end to end: pinned host GraphData in, velocities out ----------------------
        for i in range(args.warmup):
            net(host[i % N_ROTATE].to(dev, non_blocking=True, fields=net.INPUT_FIELDS), None).vel.cpu()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        from cns_b200.data import DevicePrefetcher
        f0.record()
        # every step copies its own pinned host Batch to the device (on a side stream, overlapping the
        # previous step's kernels) and reads its velocities back to the host
        # software pipeline: step i's kernels are enqueued before step i-1's velocities are read back, so the
        # GPU never idles on the host; all K results reach the host inside the timed region
        pending = None
        for data in DevicePrefetcher((host[i % N_ROTATE] for i in range(args.steps)), dev, fields=net.INPUT_FIELDS):
            pred = net(data, None)
            if pending is not None:
                vel = pending.cpu()                                    # device -> host read of a step's result
            pending = net.postprocess(pred, data)
        vel = pending.cpu()
        f1.record()
        barrier()
        full_ms = max_over_ranks(f0.elapsed_time(f1))
        e2e_full = args.graphs * world * args.steps / (full_ms * 1e-3)
        d2h_full = vel.numel() * vel.element_size()

        # ---------------- end to end, servo form: target graphs resident, a frame ships keypoints ---
        # BatchedServo = the reference's Midend split (graph built once per target, masked per frame):
        # per step H2D of the current keypoint coordinates + visibility, per-frame graph masking on the
        # device, forward, D2H of the velocities; each step waits for its own result (closed loop).
        from cns_b200.servo import BatchedServo
        servo = BatchedServo(net, [s.generator for s in scenes], [s.tPo_norm for s in scenes], dev)
        frames = [(b.pos_cur, b.node_mask.view(torch.uint8)) for b in host]     # pinned host tensors
        for i in range(args.warmup):
            servo.reset()
            vel = servo.step(*frames[i % N_ROTATE])
        ref = net(resident[(args.warmup - 1) % N_ROTATE], None).vel.cpu()
        assert torch.equal(vel, ref), "servo path and batch path disagree"
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        # two frames in flight: frame i+1's inputs are copied while frame i computes; every frame's velocities
        # are read back on the host inside the timed region
        prev = None
        for i in range(args.steps):
            servo.reset()                                 # hidden=None, the workload `value` is quoted on
            t = servo.submit(*frames[i % N_ROTATE])
            if prev is not None:
                vel = servo.result(prev)
            prev = t
        vel = servo.result(prev)
        g1.record()
        barrier()
        e2e_ms = max_over_ranks(g0.elapsed_time(g1))
        e2e = args.graphs * world * args.steps / (e2e_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_src = peaks()
    kernel_ms = kms.value / max(kn.value, 1)
    flops_per_launch = FLOP_PER_EDGE_SIDE * 2 * (sum(n_edges[i % N_ROTATE] for i in range(args.steps)) / args.steps)
    achieved = flops_per_launch / (kernel_ms * 1e-3) / 1e12 if kernel_ms > 0 else 0.0
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "enc_edge_traffic.json")) as f:
            traffic = json.load(f).get(args.precision, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    line = {
        "metric": "servo graphs/sec (fwd)", "value": value, "unit": "graphs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "f16": "f16", "bf16": "bf16"}[args.precision], "data": "synthetic",
        "config": {"workload": workload, "missing_keypoints": MISSING, "weights": "checkpoints/cns.pth",
                   "hidden": "None (first frame)", "call": "GraphVS.forward(Batch, None), stateless (no target cache)", "l2": "inputs rotate over %d distinct resident batches: a step reads %.0f MB of its batch (keypoints, "
                         "keypoint->cluster edges, the cluster edge lists being verified), %.0f MB pass before a batch "
                         "repeats > L2 (126 MB)" % (N_ROTATE, read_per_step / 1e6, N_ROTATE * read_per_step / 1e6),
                   "graphs_per_gpu": args.graphs},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e, "unit": "graphs/s", "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": servo.h2d_bytes, "d2h_bytes_per_step": servo.d2h_bytes,
                "api": "BatchedServo.submit/result(cur_xy, visible): target graphs resident on the device, per-frame "
                       "graph masking on the device, two frames in flight (H2D of frame i+1 overlaps frame i), "
                       "every frame's velocities read back on the host",
                "full_graph": {"value": e2e_full, "ms_per_step": full_ms / args.steps, "h2d_bytes_per_step": h2d,
                               "d2h_bytes_per_step": d2h_full,
                               "api": "GraphVS.forward(Batch.to(device)): the whole prebuilt host Batch is copied "
                                      "every step (DevicePrefetcher), read-back one step behind"}},
        "gpu_launches": int(launches.value),
        "with_target_cache": {"value": args.graphs * world * args.steps / (cached_ms * 1e-3), "unit": "graphs/s",
                              "ms_per_step": cached_ms / args.steps,
                              "note": "same loop with GraphVS.build_target_cache (per-target terms built once outside "
                                      "the timed region); not the headline"},
        "roofline": {"kernel": "enc_edge (%s)" % args.precision, "bound": "tensor", "achieved": achieved,
                     "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": pk_src + " bf16 sustained", "kernel_ms": kernel_ms,
                     "kernel_share_of_step": kernel_ms / (total_ms / args.steps),
                     "flops_per_launch": flops_per_launch},
    }
    if world == 1 and not args.no_cpu_baseline:
        rate, ms, done, cores = cpu_reference_rate(args.cpu_sample, args.points, 50, 1, 20.0)
        line["cpu_baseline"] = {"value": rate, "unit": "graphs/s", "cores": cores, "kind": "port",
                                "sample": "%d graphs x %d keypoints per step, %d steps (%.1f ms/step)" % (
                                    args.cpu_sample, args.points, done, ms)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
