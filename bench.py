#!/usr/bin/env python
"""bench.py - servo graphs/s of the CNS hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision fp32|f16|bf16]

Workload (BASELINE.json configs[1]): pretrained checkpoints/cns.pth, batched inference,
1024 synthetic servo graphs x 512 keypoints per GPU (weak scaling: every rank owns its own 1024
graphs; inference needs no collective).  One "step" = one GraphVS forward (+ fused postprocess)
over one 1024-graph batch.

  value : graphs/s with the batches already resident in HBM (CUDA events, max over ranks): K stateless
          GraphVS.forward(Batch, None) calls, consecutive calls rotating over N_STREAMS (3) CUDA streams so that
          up to three forwards are in flight (the encoder of step i+1 fills the SMs that the latency-bound
          cluster-level kernels of step i leave idle); a call whose argument set repeats is replayed by the library
          as ONE CUDA-graph launch.  `one_stream` = the same calls strictly one after the other.
  e2e   : the same through the public API from HOST (pinned) memory, servo form (BatchedServo.submit/result):
          the target-side graphs are per-scene state resident on the device (the reference's Midend
          builds them once per target), a step copies the current keypoint coordinates + visibility
          in, masks the graphs on the device, runs the forward and reads the (B,6) velocities back; up to three
          frames in flight, every frame's velocities reach the host inside the timed region.
          e2e.full_graph: GraphVS.forward on a prebuilt host Batch copied whole every step.
  roofline     : the dominant kernel (encoder edge kernel), timed live with CUDA events on its
                 launching stream (cns_profile_begin/end; inside a replayed graph the pair is a pair of external
                 event-record nodes) in the ONE-stream loop, where the kernel has the GPU to itself.
                 `frac` is quoted against the BURST bf16 peak while the K-step loop is shorter than 1 s
                 (it runs at boost clocks); `roofline.sustained` repeats the loop for >= 2 s and quotes
                 the same kernel against the sustained peak.
  cpu_baseline : the oracle (CPU port of the reference forward) on the same workload (full 1024-graph
                 batch, bounded number of steps; batch-size sweep beside it), on this box's host cores.
  fp32_mode    : the same device-resident loop in the fp32-tolerance mode (rel 1e-4 parity).
  train        : BASELINE configs[2] - short-sequence training update (8-frame truncated BPTT, 256 graphs
                 x 512 keypoints per GPU, fused loss, ONE NCCL all-reduce of the flat gradient bucket
                 inside the timed region when world > 1, fused clip + AdamW): graph-frames/s fwd+bwd.
  latency      : BASELINE configs[0] - batch-1 recurrent servo frames through GraphVSController
                 (host GraphData in, velocity out): p50 / p90, the CPU port beside it.
  per_rank_ms  : min / max over ranks of the per-step time of the `value` loop.
`--impl reference` times that CPU implementation alone on the full workload (the reference itself is
Python on PyG, which cannot be installed here; the port in oracle/ is the stand-in, kind "port").
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
PREROLL = 2           # untimed steps enqueued between the barrier and the start event of a timed loop
N_STREAMS = int(os.environ.get("CNS_BENCH_STREAMS", "3"))   # CUDA streams the `value` loop alternates between (must divide N_ROTATE)
PRIME = 12            # untimed steps after the warm-up: every (batch, stream) argument set is seen twice -> CUDA-graph replay
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
    ap.add_argument("--cpu-sample", type=int, default=0, help="graphs per CPU-baseline step (0 = the full batch)")
    ap.add_argument("--cpu-chunk", type=int, default=64, help="reference arm: graphs per forward call (0 = one call per step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the fp32 / train / latency / sustained legs")
    ap.add_argument("--train-graphs", type=int, default=256)
    ap.add_argument("--train-frames", type=int, default=8)
    ap.add_argument("--train-updates", type=int, default=5)
    ap.add_argument("--latency-frames", type=int, default=200)
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


def cpu_reference_rate(n_graphs, n_points, steps, warmup, budget_s=25.0, batch=None, chunk=0):
    """graphs/s of the CPU port of the reference forward (oracle/cns_oracle.py), all host threads.  A step covers all
    `n_graphs` graphs; with `chunk` > 0 it does so as ceil(n_graphs / chunk) forward calls of `chunk` graphs (the CPU
    port is fastest at moderate batch sizes - cache footprint - so this is its best way through the workload)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cns_oracle
    from cns_b200 import ckpt
    torch.set_num_threads(os.cpu_count() or 1)
    sd = ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns.pth"))
    if batch is None and chunk and chunk < n_graphs:
        from cns_b200 import synth
        from cns_b200.data import Batch
        scenes = synth.make_scenes(n_graphs, n_points, seed=2022, n_layouts=64)
        frames = [s.frame(MISSING) for s in scenes]
        batches = [Batch.from_data_list(frames[i:i + chunk]) for i in range(0, n_graphs, chunk)]
    else:
        batches = [batch if batch is not None else build_batches(n_graphs, n_points, 2022, 1)[0]]
    with torch.no_grad():
        for _ in range(max(1, warmup)):
            for b in batches:
                cns_oracle.forward(sd, b, None)
        t0, done = time.perf_counter(), 0
        for _ in range(steps):
            for b in batches:
                cns_oracle.forward(sd, b, None)
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
    return n_graphs * done / dt, dt / done * 1e3, done, torch.get_num_threads()


def sub_batch(host_batch_scenes, n):
    """The first `n` scenes' current frames as one host Batch (CPU-baseline sweep)."""
    from cns_b200.data import Batch
    return Batch.from_data_list([s.frame(MISSING) for s in host_batch_scenes[:n]])


def cpu_latency(scene, frames, warm=5):
    """Batch-1 recurrent frames through the CPU port: per-frame wall times (ms)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cns_oracle
    from cns_b200 import ckpt
    sd = ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns.pth"))
    hid, out = None, []
    with torch.no_grad():
        for t in range(frames + warm):
            d = scene.frame(MISSING)
            t0 = time.perf_counter()
            vec, nrm, hid = cns_oracle.forward(sd, d, hid)
            vel = cns_oracle.postprocess((vec, nrm), d.tPo_norm.reshape(1))[0].numpy()
            if t >= warm:
                out.append((time.perf_counter() - t0) * 1e3)
            scene.step(vel * 2)
    return out


def workload_config(args):
    """`config` of the JSON line - identical in the b200 and the reference arm."""
    return {"workload": "cns.pth batched inference, %d synthetic servo graphs x %d keypoints per GPU" % (args.graphs, args.points),
            "graphs_per_gpu": args.graphs, "keypoints_per_graph": args.points, "missing_keypoints": MISSING,
            "weights": "checkpoints/cns.pth", "hidden": "None (first frame)",
            "call": "GraphVS.forward(Batch, None), stateless (no target cache); consecutive steps rotate over a few CUDA "
                    "streams (<= %d forwards in flight), repeated argument sets replay as CUDA graphs" % N_STREAMS}


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

    if args.impl == "reference":
        if rank != 0:
            return
        n_cpu = args.cpu_sample or args.graphs
        rate, ms, done, cores = cpu_reference_rate(n_cpu, args.points, args.steps, args.warmup, 150.0, chunk=args.cpu_chunk)
        sample = "all %d of the %d graphs per step (N=%d keypoints, 10%% missing) as %d forward calls of %d graphs (the port's " \
                 "fastest batch size, see cpu_baseline.batch_sweep of the b200 arm), fp32 torch CPU, %d steps" % (
                     n_cpu, args.graphs, args.points, -(-n_cpu // max(args.cpu_chunk, 1)) if args.cpu_chunk else 1,
                     args.cpu_chunk or n_cpu, done)
        emit({
            "impl": "reference", "metric": "servo graphs/sec (fwd)", "value": rate, "unit": "graphs/s",
            "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": rate, "unit": "graphs/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    import torch
    import torch.distributed as dist
    from cns_b200 import _lib, ckpt, synth

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

    def gather_ms(ms):
        """(max over ranks, [min, max]) of a per-rank elapsed time."""
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            lo, hi = t.clone(), t.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            return float(hi.item()), [float(lo.item()), float(hi.item())]
        return ms, [ms, ms]

    streams = [torch.cuda.Stream(dev) for _ in range(N_STREAMS)]

    def enqueue(model, first, n, multi_stream, **kw):
        """`n` stateless forwards over the rotating resident batches, step i on batch (first + i) % N_ROTATE.  With
        `multi_stream` consecutive steps rotate over N_STREAMS CUDA streams (<= N_STREAMS forwards in flight: the encoder
        of step i+1 fills the SMs the second wave of step i's cluster-level backbone leaves idle); the streams fork from
        and join the current stream, so events recorded on it bracket all of the work."""
        if not multi_stream:
            for i in range(n):
                model(resident[(first + i) % N_ROTATE], None, **kw)
            return
        cur = torch.cuda.current_stream(dev)
        for st in streams:
            st.wait_stream(cur)
        for i in range(n):
            with torch.cuda.stream(streams[(first + i) % N_STREAMS]):
                model(resident[(first + i) % N_ROTATE], None, **kw)
        for st in streams:
            cur.wait_stream(st)

    def forward_loop(model, steps, warmup, profile=True, multi_stream=False, **kw):
        """`steps` stateless forwards: (elapsed ms on this rank, mean encoder-edge kernel ms, launches counted by the
        library).  Untimed before the start event: `warmup` steps, then PRIME more so that the library has seen every
        (batch, stream) argument set twice and replays it as a CUDA graph, then a barrier and PREROLL steps."""
        hd = model._handle(dev)
        enqueue(model, 0, warmup, multi_stream, **kw)
        if profile:
            _lib.check(lib.cns_profile_begin(hd.raw), "profile_begin")     # timed graphs are captured with their event pair
        enqueue(model, warmup, PRIME, multi_stream, **kw)
        barrier()
        # pre-roll: the barrier leaves the GPU idle; two more untimed steps are enqueued right in front of the start event
        # so the timed steps run back to back with work (an idle -> busy transition and a host hiccup of up to a step
        # would otherwise land inside a 9 ms region; at 8 ranks that showed as one rank 9 % slow)
        first = warmup + PRIME
        enqueue(model, first, PREROLL, multi_stream, **kw)
        if profile:
            _lib.check(lib.cns_profile_begin(hd.raw), "profile_begin")
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        enqueue(model, first + PREROLL, steps, multi_stream, **kw)
        t1.record()
        barrier()
        kms, kn, nl = C.c_float(), C.c_int(), C.c_longlong()
        if profile:
            _lib.check(lib.cns_profile_end(hd.raw, C.byref(kms), C.byref(kn), C.byref(nl)), "profile_end")
        return t0.elapsed_time(t1), kms.value / max(kn.value, 1), int(nl.value)

    extras = not args.no_extras
    fp32 = train = latency = sustained = None
    with torch.no_grad():
        # ---------------- device-resident throughput ("value") + live kernel timing ----------------
        # the stateless call of the reference API, GraphVS.forward(Batch, hidden=None): nothing is kept from one
        # step to the next
        # (a) one stream: the calls run strictly one after the other; the encoder edge kernel is timed here, where it
        #     has the GPU to itself
        one_ms, kernel_ms, launches = forward_loop(net, args.steps, args.warmup)
        one_total, one_rank = gather_ms(one_ms)
        # (b) `value`: the same calls rotating over N_STREAMS CUDA streams (<= N_STREAMS forwards in flight)
        sampler = ClockSampler(local)
        sampler.start()
        my_ms, _, _ = forward_loop(net, args.steps, args.warmup, profile=False, multi_stream=True)
        sampler.stop_flag = True
        total_ms, rank_ms = gather_ms(my_ms)
        value = args.graphs * world * args.steps / (total_ms * 1e-3)
        hd = net._handle(dev)
        graph_stats = {"replays": hd.counter("graph_replays"), "captures": hd.counter("graph_captures")}

        # the same loop with a per-target cache (GraphVS.build_target_cache): the rotating batches are the same
        # scenes at successive frames, and what depends on the targets alone (per-cluster attention term, cluster
        # positions, tile packing) is built once, as the reference's Midend keeps one GraphGenerator per target
        # image.  Reported beside `value`, not as `value`.
        tcache = net.build_target_cache(resident[0])
        cached_ms = max_over_ranks(forward_loop(net, args.steps, args.warmup, profile=False, multi_stream=True, target_cache=tcache)[0])

        if extras:
            # ---------------- the same loop for >= 2 s: clocks settle at their sustained level ---------------
            n_sus = max(args.steps, int(2200.0 / max(my_ms / args.steps, 1e-3)))
            sus_sampler = ClockSampler(local)
            sus_sampler.start()
            sus_ms = forward_loop(net, n_sus, 0, profile=False, multi_stream=True)[0]
            # ... and, right behind it, the one-stream loop that times the encoder edge kernel at those clocks
            _, sus_kernel_ms, _ = forward_loop(net, 64, 0)
            sus_sampler.stop_flag = True
            sus_total = max_over_ranks(sus_ms)
            sustained = {"steps": n_sus, "seconds": sus_total * 1e-3, "ms_per_step": sus_total / n_sus,
                         "value": args.graphs * world * n_sus / (sus_total * 1e-3), "kernel_ms": sus_kernel_ms,
                         "clocks": sus_sampler.summary()}
            # ---------------- fp32-tolerance mode (rel 1e-4 parity), same loop -------------------------------
            if args.precision != "fp32":
                net32 = ckpt.load_model(os.path.join(ROOT, "checkpoints", "cns.pth"), dev, precision="fp32")
                f1_ms, f_kernel_ms, f_launches = forward_loop(net32, args.steps, args.warmup)
                f_ms = forward_loop(net32, args.steps, args.warmup, profile=False, multi_stream=True)[0]
                f_total = max_over_ranks(f_ms)
                fp32 = {"value": args.graphs * world * args.steps / (f_total * 1e-3), "unit": "graphs/s",
                        "ms_per_step": f_total / args.steps, "one_stream_ms_per_step": max_over_ranks(f1_ms) / args.steps,
                        "kernel_ms": f_kernel_ms, "gpu_launches": f_launches,
                        "tolerance": "rel 1e-4 vs the fp64 oracle (tests/test_gpu_forward.py)"}
                del net32

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
        frames = [(b.pos_cur, b.node_mask.view(torch.uint8)) for b in host]     # pinned host tensors

        def servo_e2e():
            servo = BatchedServo(net, [s.generator for s in scenes], [s.tPo_norm for s in scenes], dev, depth=N_STREAMS)
            for i in range(args.warmup + PRIME):       # PRIME: every frame's argument set seen twice -> graph replay
                servo.reset()
                vel = servo.step(*frames[i % N_ROTATE])
            ref = net(resident[(args.warmup + PRIME - 1) % N_ROTATE], None).vel.cpu()
            # (with the per-target cache the target side comes from values kept in f16, so the two paths agree to the
            # operand rounding of the mode, not bit for bit)
            err = float((vel - ref).abs().max() / ref.abs().max())
            assert err < (1e-5 if args.precision == "fp32" else 5e-3), "servo path and batch path disagree (%.2e)" % err
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            # N_STREAMS frames in flight: frame i+1's inputs are copied while frame i computes; every frame's
            # velocities are read back on the host inside the timed region
            pending = []
            for i in range(args.steps):
                servo.reset()                             # hidden=None, the workload `value` is quoted on
                pending.append(servo.submit(*frames[i % N_ROTATE]))
                if len(pending) == servo.depth:
                    vel = servo.result(pending.pop(0))
            while pending:
                vel = servo.result(pending.pop(0))
            g1.record()
            barrier()
            ms = max_over_ranks(g0.elapsed_time(g1))
            return ms, servo.h2d_bytes, servo.d2h_bytes

        e2e_ms, h2d_servo, d2h_servo = servo_e2e()
        e2e = args.graphs * world * args.steps / (e2e_ms * 1e-3)
        # the same loop with the target side recomputed every frame (no per-keypoint logits / values kept per target)
        net._handle(dev).set_option("target_edge_cache", 0)
        try:
            e2e_nc_ms, _, _ = servo_e2e()
        finally:
            net._handle(dev).set_option("target_edge_cache", 1)
        e2e_nc = args.graphs * world * args.steps / (e2e_nc_ms * 1e-3)

    if extras:
        # ---------------- training update (BASELINE configs[2]): fwd + bwd, all-reduce inside -------------
        from cns_b200 import GraphVS
        from cns_b200.data import Batch
        from cns_b200.train import TBPTTStepper
        tb, tf = min(args.train_graphs, args.graphs), args.train_frames
        tscenes = scenes[:tb]
        window = []
        for t in range(tf):
            fr = [s.frame(MISSING, with_labels=True) for s in tscenes]
            if t == 0:
                for f in fr:
                    f.start_new_scene()
            window.append(Batch.from_data_list(fr).to(dev))
            for s in tscenes:
                s.step(synth.pbvs_straight(s.cur_wcT, s.tar_wcT) * 2)
        tnet = GraphVS(2, 2, 128).to(dev)
        tnet.load_state_dict(ckpt.read_state_dict(os.path.join(ROOT, "checkpoints", "cns.pth")))
        tnet.train()
        stepper = TBPTTStepper(tnet, fused=True, lr=1e-5)
        for _ in range(2):
            stepper.hidden = None
            stepper.step(window)
        barrier()
        th = tnet._handle(dev, training=True)
        _lib.check(lib.cns_profile_begin(th.raw), "profile_begin")
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        u0.record()
        for _ in range(args.train_updates):
            stepper.hidden = None
            loss = stepper.step(window)
        u1.record()
        barrier()
        tk, tn, tl = C.c_float(), C.c_int(), C.c_longlong()
        _lib.check(lib.cns_profile_end(th.raw, C.byref(tk), C.byref(tn), C.byref(tl)), "profile_end")
        upd_ms = max_over_ranks(u0.elapsed_time(u1)) / args.train_updates
        ar_ms = None
        if world > 1:        # the collective alone (1.86 MB flat bucket), for the record
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.all_reduce(stepper.bucket.flat)
            barrier()
            a0.record()
            for _ in range(20):
                dist.all_reduce(stepper.bucket.flat)
            a1.record()
            barrier()
            ar_ms = max_over_ranks(a0.elapsed_time(a1)) / 20
        train = {"workload": "short-sequence training update: %d-frame truncated BPTT, %d graphs x %d keypoints per GPU, fused "
                             "loss, gradient all-reduce (world %d) + fused clip/AdamW inside the timed region" % (tf, tb, args.points, world),
                 "ms_per_update": upd_ms, "value": tb * tf * world / (upd_ms * 1e-3), "unit": "graph-frames/s (fwd+bwd)",
                 "updates": args.train_updates, "loss": float(loss), "allreduce_ms": ar_ms,
                 "allreduce_bytes": int(stepper.bucket.flat.numel() * 4), "gpu_launches_per_update": int(tl.value) // max(args.train_updates, 1),
                 "enc_edge_fwd_kernel_ms": tk.value / max(tn.value, 1), "precision": "fp32-tolerance kernels (training)"}
        del stepper, tnet, window

        # ---------------- batch-1 latency (BASELINE configs[0]) through GraphVSController, rank 0 --------
        if rank == 0:
            from cns_b200.controller import GraphVSController
            import numpy as np
            lat = {}
            for prec in (args.precision, "fp32") if args.precision != "fp32" else ("fp32",):
                ctl = GraphVSController(os.path.join(ROOT, "checkpoints", "cns.pth"), dev, precision=prec)
                sc = synth.Scene(np.random.default_rng(7), args.points, generator=scenes[0].generator, points=scenes[0].points,
                                 tar_wcT=scenes[0].tar_wcT)
                times = []
                for t in range(args.latency_frames + 20):
                    d = sc.frame(MISSING)
                    if t == 0:
                        d.start_new_scene()
                    t0 = time.perf_counter()
                    v = ctl(d)                                            # host GraphData in, np velocity out (synchronous)
                    if t >= 20:
                        times.append((time.perf_counter() - t0) * 1e3)
                    sc.step(v * 2)
                times.sort()
                lat[prec] = {"p50_ms": times[len(times) // 2], "p90_ms": times[int(len(times) * 0.9)], "frames": len(times)}
            latency = {"workload": "batch 1, %d keypoints, recurrent (hidden carried), GraphVSController(host GraphData) -> velocity" % args.points,
                       "unit": "ms per frame (host wall clock, synchronous)", **lat}
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_src = peaks()
    edges_per_step = sum(n_edges[i % N_ROTATE] for i in range(args.steps)) / args.steps
    flops_per_launch = FLOP_PER_EDGE_SIDE * 2 * edges_per_step
    achieved = flops_per_launch / (kernel_ms * 1e-3) / 1e12 if kernel_ms > 0 else 0.0
    # a K-step loop shorter than 1 s runs at boost clocks: the burst peak is its denominator; the >= 2 s loop
    # (roofline.sustained) is quoted against the sustained peak
    burst = total_ms < 1000.0
    peak = pk["bf16_tflops"] if burst else pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "enc_edge_traffic.json")) as f:
            traffic = json.load(f).get(args.precision, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    config = workload_config(args)
    config["preroll"] = ("after the %d warm-up steps: %d more untimed steps (every argument set seen twice -> graph replay), a "
                         "barrier, then %d untimed steps directly in front of the start event" % (args.warmup, PRIME, PREROLL))
    config["l2"] = ("inputs rotate over %d distinct resident batches: a step reads %.0f MB of its batch (keypoints, "
                    "keypoint->cluster edges, the cluster edge lists being verified), %.0f MB pass before a batch "
                    "repeats > L2 (126 MB)" % (N_ROTATE, read_per_step / 1e6, N_ROTATE * read_per_step / 1e6))
    line = {
        "metric": "servo graphs/sec (fwd)", "value": value, "unit": "graphs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "f16": "f16", "bf16": "bf16"}[args.precision], "data": "synthetic",
        "config": config,
        "clocks": sampler.summary(),
        "per_rank_ms": {"min": rank_ms[0] / args.steps, "max": rank_ms[1] / args.steps, "what": "ms per step of the `value` loop"},
        "e2e": {"value": e2e, "unit": "graphs/s", "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": h2d_servo, "d2h_bytes_per_step": d2h_servo,
                "api": "BatchedServo.submit/result(cur_xy, visible): target graphs resident on the device, per-frame "
                       "graph masking on the device, %d frames in flight, each slot on its own compute stream (H2D and "
                       "kernels of later frames overlap frame i), every frame's velocities read back on the host; what depends on "
                       "the target alone (per-keypoint target-side logits and values, per-cluster softmax totals) is kept "
                       "per target, as the reference's set_target / get_control_rate split allows - "
                       "`recomputed_target_side` is the same loop without it" % N_STREAMS,
                "recomputed_target_side": {"value": e2e_nc, "unit": "graphs/s", "ms_per_step": e2e_nc_ms / args.steps},
                "full_graph": {"value": e2e_full, "ms_per_step": full_ms / args.steps, "h2d_bytes_per_step": h2d,
                               "d2h_bytes_per_step": d2h_full,
                               "api": "GraphVS.forward(Batch.to(device)): the whole prebuilt host Batch is copied "
                                      "every step (DevicePrefetcher), read-back one step behind"}},
        "gpu_launches": int(launches),
        "one_stream": {"value": args.graphs * world * args.steps / (one_total * 1e-3), "unit": "graphs/s",
                       "ms_per_step": one_total / args.steps,
                       "per_rank_ms": {"min": one_rank[0] / args.steps, "max": one_rank[1] / args.steps},
                       "note": "the same calls on ONE stream (each forward starts when the previous one has finished); "
                               "roofline.kernel_ms and gpu_launches are measured in this loop"},
        "cuda_graphs": graph_stats,
        "with_target_cache": {"value": args.graphs * world * args.steps / (cached_ms * 1e-3), "unit": "graphs/s",
                              "ms_per_step": cached_ms / args.steps,
                              "note": "same loop with GraphVS.build_target_cache (per-target terms built once outside "
                                      "the timed region); not the headline"},
        "roofline": {"kernel": "enc_edge (%s)" % args.precision, "bound": "tensor", "achieved": achieved,
                     "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": pk_src + (" bf16 burst (timed loop %.1f ms < 1 s)" % total_ms if burst else " bf16 sustained"),
                     "kernel_ms": kernel_ms,
                     "kernel_share_of_step": kernel_ms / (one_ms / args.steps),
                     "timed_in": "the one-stream loop (one_stream.ms_per_step): with two forwards in flight the event pair "
                                 "around a kernel also spans the time it waits for SMs",
                     "flops_per_launch": flops_per_launch},
    }
    if sustained is not None:
        sp = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        sa = flops_per_launch / (sustained["kernel_ms"] * 1e-3) / 1e12 if sustained["kernel_ms"] > 0 else 0.0
        line["roofline"]["sustained"] = {"achieved": sa, "peak": sp, "frac": sa / sp, **sustained}
    if fp32 is not None:
        fa = flops_per_launch / (fp32["kernel_ms"] * 1e-3) / 1e12 if fp32["kernel_ms"] > 0 else 0.0
        fp32["roofline"] = {"kernel": "enc_edge (fp32 mode)", "bound": "tensor", "achieved": fa, "peak": pk["bf16_tflops"],
                            "unit": "TFLOP/s", "frac": fa / pk["bf16_tflops"],
                            "note": "algorithmic fp32 flops of the reference chain (not the operand-split products) over the bf16 burst peak"}
        line["fp32_mode"] = fp32
    if train is not None:
        line["train"] = train
    if latency is not None:
        line["latency"] = latency
    if world == 1 and not args.no_cpu_baseline:
        n_cpu = args.cpu_sample or args.graphs
        rate, ms, done, cores = cpu_reference_rate(n_cpu, args.points, 30, 1, 15.0, batch=sub_batch(scenes, n_cpu))
        sweep = {}
        for nb in (64, 256):
            if nb < n_cpu:
                r2, m2, d2, _ = cpu_reference_rate(nb, args.points, 30, 1, 5.0, batch=sub_batch(scenes, nb))
                sweep[str(nb)] = {"graphs_per_s": r2, "ms_per_step": m2, "steps": d2}
        sweep[str(n_cpu)] = {"graphs_per_s": rate, "ms_per_step": ms, "steps": done}
        best = max(v["graphs_per_s"] for v in sweep.values())
        line["cpu_baseline"] = {"value": best, "unit": "graphs/s", "cores": cores, "kind": "port",
                                "sample": "best of a batch-size sweep of the oracle port, fp32 torch CPU, %d keypoints per graph; "
                                          "full batch: %d graphs per step, %d steps (%.0f ms/step)" % (args.points, n_cpu, done, ms),
                                "batch_sweep": sweep}
        if latency is not None:
            import numpy as np
            sc = synth.Scene(np.random.default_rng(7), args.points, generator=scenes[0].generator, points=scenes[0].points,
                             tar_wcT=scenes[0].tar_wcT)
            ct = sorted(cpu_latency(sc, 30))
            line["latency"]["cpu_port"] = {"p50_ms": ct[len(ct) // 2], "p90_ms": ct[int(len(ct) * 0.9)], "frames": len(ct),
                                           "cores": cores}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
