#!/usr/bin/env python
"""Headline benchmark of the Rot-MV hot path (BASELINE.json: "multi-view samples/sec (224^2, fwd &
fwd+bwd) at 1/2/4/8 B200; roofline %").

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--mode both|infer|train]

The default run (`--mode both`) measures BOTH halves of the metric and prints ONE JSON line (rank 0):

  top level  configs[1]: forward of `FeatRotationSymm` (ResNet-50 trunk -> rotation-constrained
             cross-view fusion x3 -> gaze heads) over 256 two-view samples per GPU (512 images,
             224x224), random-init weights, synthetic inputs, bf16 tcgen05 engine
     value      whole-job multi-view samples/s, inputs resident in HBM, CUDA-graph replay, CUDA events
     e2e        same metric through the host-buffer entry (pinned host -> HBM copy of images+rotations
                and device -> host read of pred_gaze inside the timed region, every step); variants:
                fp32 NCHW tensors (the reference's input) and raw uint8 HWC images
     roofline   tcgen05 implicit-GEMM kernel family, measured live: per-class (tensor-/HBM-bound) split
     cpu_baseline  the CPU oracle (port of the reference's PyTorch path) on the host cores
  "train"    configs[3]: the training step (forward + angular loss + backward + Adam with coupled L2,
             reference trainer.py:119-123,141-143) at 128 two-view samples per GPU, CUDA-graph
             captured; with N > 1 the flat fp32 gradient is all-reduced over NCCL every step
     value / ms_per_step / e2e / roofline{frac, bn_ms, conv_ms, wgrad_ms} / allreduce{bytes,
     exposed_ms} / clocks / cpu_baseline

`--impl reference` times the reference's own CPU path (oracle port; /root/reference does not travel
to the GPU box) on a bounded sample of the same two workloads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rot-mvgaze_b200"))

import torch  # noqa: E402

METRIC = "multi-view samples/sec (224^2, fwd)"
METRIC_TRAIN = "multi-view samples/sec (224^2, fwd+bwd+Adam)"
UNIT = "samples/s"
# SURVEY 8d / BASELINE.md: forward FLOPs per view (2*MAC, conv + linear), default config
FLOPS_PER_VIEW = 8.306399e9


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "src": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "src": "fallback (B200_PROFILING.md)"}


def infer_workload(args, batch, views):
    return {"workload": "configs[1]: Rot-MV 2-view inference, synthetic batch 256 bf16 on 1xB200 "
                        "(xgaze2mpiinv_known shape)" if (batch, views) == (256, 2)
            else f"Rot-MV {views}-view inference, synthetic batch {batch}",
            "batch_per_gpu": batch, "views": views, "image": "3x224x224 fp32 NCHW",
            "backbone": "resnet50", "num_iter": 3, "weights": "random-init (seed 0)",
            "mode": "eval forward (no_grad)", "precision": args.precision,
            "parallelism": f"dp{args.gpus} (batch sharded, no collective)",
            "l2": "inputs %.0f MB/step per GPU exceed the 126 MB L2; no flush needed"
                  % (batch * views * 3 * 224 * 224 * 4 / 1e6)}


def train_workload(args, batch, views, world, grad_bytes):
    return {"workload": "configs[3]: Rot-MV 2-view training step (fwd+bwd+AdamW), data-parallel, "
                        "batch 128 per GPU" if (batch, views) == (128, 2)
            else f"Rot-MV {views}-view training step, batch {batch} per GPU",
            "batch_per_gpu": batch, "views": views, "backbone": "resnet50", "num_iter": 3,
            "optimizer": "AdamW (decoupled)" if args.adamw else "Adam + coupled L2 (reference trainer.py:54)",
            "parallelism": (f"dp{world}, fp32 gradient all-reduce of {grad_bytes / 1e6:.0f} MB per step "
                            "(NCCL, bucketed per trunk stage, overlapped with the backward)")
            if world > 1 else "dp1",
            "precision": args.precision,
            "l2": "inputs %.0f MB/step per GPU exceed the 126 MB L2" % (batch * views * 602112 / 1e6)}


# --------------------------------------------------------------------------------------------
# CPU oracle legs (the only place bench.py executes oracle/)
# --------------------------------------------------------------------------------------------
def time_cpu_oracle(views: int, sample_batch: int, min_seconds: float, max_iters: int,
                    warmup: int = 1, train: bool = False):
    from oracle import rotmv_oracle as O

    torch.set_num_threads(os.cpu_count())
    model = O.build_model(num_iter=3, depth=50, seed=0)
    images, pose, gt = O.synthetic_batch(sample_batch, views, seed=1)
    rot = O.pairwise_rotations(pose)
    times = []
    if train:
        model.train()
        opt = O.make_adam(model, lr=1e-6)
        fn = lambda: O.train_step(model, opt, images, rot, gt)  # noqa: E731
    else:
        model.eval()

        def fn():
            with torch.no_grad():
                model.forward_views(images, rot)
    for _ in range(warmup):
        fn()
    t_all = time.perf_counter()
    while len(times) < max_iters and (time.perf_counter() - t_all < min_seconds or len(times) < 2):
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 8
    from oracle import rotmv_oracle as O

    torch.set_num_threads(os.cpu_count())
    cores = torch.get_num_threads()

    def leg(train: bool, views: int):
        model = O.build_model(num_iter=3, depth=50, seed=0)
        images, pose, gt = O.synthetic_batch(sample, views, seed=1)
        rot = O.pairwise_rotations(pose)
        if train:
            model.train()
            opt = O.make_adam(model, lr=1e-6)
            fn = lambda: O.train_step(model, opt, images, rot, gt)  # noqa: E731
        else:
            model.eval()

            def fn():
                with torch.no_grad():
                    model.forward_views(images, rot)
        for _ in range(args.warmup):
            fn()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        dt = time.perf_counter() - t0
        return sample * args.steps / dt, dt / args.steps * 1e3

    def block(value, ms, metric, config, what):
        desc = (f"oracle port of the reference PyTorch CPU path; each step = {sample} of the "
                f"{config['batch_per_gpu']} samples of the workload batch ({config['views']} views, fp32, {what})")
        return {"impl": "reference", "metric": metric, "value": value, "unit": UNIT,
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}

    line = None
    if args.mode in ("both", "infer"):
        b = args.batch or 256
        v, ms = leg(False, args.views)
        line = block(v, ms, METRIC, infer_workload(args, b, args.views), "eval forward")
    if args.mode in ("both", "train"):
        b = args.train_batch or args.batch or 128
        v, ms = leg(True, args.views)
        tr = block(v, ms, METRIC_TRAIN, train_workload(args, b, args.views, 1, 0), "fwd+loss+bwd+Adam step")
        if line is None:
            line = tr
        else:
            line["train"] = tr
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                rows.append((float(f[0]), float(f[1]), float(f[6]), f[2:6]))
            except ValueError:
                continue
        loaded = [r for r in rows if r[2] >= 50]   # samples taken under load (the timed loops)
        sm, mx, reasons = [], [], set()
        for clk, cmax, _, flags in (loaded or rows):
            sm.append(clk); mx.append(cmax)
            for nm, val in zip(names, flags):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# process context
# --------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self):
        import torch.distributed as dist

        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                             "(use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("NCCL_DEBUG", "INFO")           # keep the communicator log on
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = torch.tensor([x], device=self.dev, dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item()

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def timed_wall(ctx, fn, steps):
    ctx.barrier()
    t0 = time.perf_counter()
    fn(steps)
    ctx.barrier()
    return ctx.max_over_ranks(time.perf_counter() - t0)


def load_traffic():
    """DRAM bytes per launch of the dominant kernel from the round's `ncu --set full` capture
    (profiles/igemm_traffic.json; the file names the capture and its date)."""
    tpath = os.path.join(ROOT, "profiles", "igemm_traffic.json")
    if not os.path.exists(tpath):
        return None, None
    t = json.load(open(tpath))
    return t.get("dram_bytes_per_launch"), {k: t[k] for k in ("source", "captured", "note") if k in t}


# --------------------------------------------------------------------------------------------
# configs[1]: inference
# --------------------------------------------------------------------------------------------
def bench_infer(args, ctx):
    from rotmv_b200 import functional as RF
    from rotmv_b200.engine import GraphedForward
    from rotmv_b200.module import FeatRotationSymm

    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    B, V = args.batch or 256, args.views
    W = max(args.warmup, 3)
    torch.manual_seed(0)
    model = FeatRotationSymm(50, 3, precision=args.precision, trunk_chunk=args.chunk).to(dev).eval()
    g = torch.Generator().manual_seed(1 + rank)
    images_host = torch.randn((B, V, 3, 224, 224), generator=g).pin_memory()
    pose_host = torch.rand((B, V, 2), generator=g) - 0.5
    rot_dev = RF.pose_to_rotations(pose_host.to(dev))
    rot_host = rot_dev.cpu().pin_memory()
    sess = GraphedForward(model, B, V, precision=args.precision, copy_chunks=args.copy_chunks)
    sess.images.copy_(images_host.to(dev))
    sess.rotations.copy_(rot_dev)

    # ---- device-resident throughput: W warm-up replays, then K timed replays ------------------
    sampler = ClockSampler(ctx.local) if rank == 0 else None   # runs through every timed loop below
    for _ in range(W):
        sess()
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record()
    for _ in range(args.steps):
        sess()
        launches += sess.launches_per_replay
    e1.record()
    ctx.barrier()
    ms_total = ctx.max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    pred_check = sess.pred.float().abs().sum().item()
    if not (pred_check == pred_check):
        raise SystemExit("bench.py: non-finite predictions")

    # ---- end to end: host buffers in, prediction out, every step -----------------------------
    def blocking(n):
        for _ in range(n):
            sess.run_host(images_host, rot_host)

    def make_pipelined(s, img):
        def run(n):
            prev = None
            for _ in range(n):
                tk = s.submit(img, rot_host)
                if prev is not None:
                    s.result(prev)
                prev = tk
            return s.result(prev)
        return run

    blocking(2)
    e2e_blocking_s = timed_wall(ctx, blocking, args.steps)
    # the same host entry used asynchronously (two calls in flight): call k+1's host->HBM copy
    # overlaps call k's kernels; every step's H2D and its D2H read are inside the timed region
    pipelined = make_pipelined(sess, images_host)
    pipelined(3)
    e2e_s = timed_wall(ctx, pipelined, args.steps)
    h2d = images_host.numel() * 4 + rot_host.numel() * 4
    d2h = B * 2 * 4
    # the limiter of the fp32 entry at N > 2: aggregate pinned-host -> HBM copy bandwidth of the box.
    # All ranks copy their batch at the same time, nothing else running.
    stage = torch.empty_like(sess.images)

    def copies(n):
        for _ in range(n):
            stage.copy_(images_host, non_blocking=True)
        torch.cuda.synchronize()

    copies(2)
    copy_s = timed_wall(ctx, copies, 10)
    h2d_gbs = images_host.numel() * 4 * 10 / copy_s / 1e9
    del stage
    fp32_var = {"value": world * B * args.steps / e2e_s, "unit": UNIT, "ms_per_step": e2e_s / args.steps * 1e3,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "h2d_gbs_per_gpu_needed": h2d / (ms_step * 1e-3) / 1e9,
                "h2d_gbs_per_gpu_copy_only": h2d_gbs,
                "limiter": ("host->HBM copy bandwidth (all ranks copying at once: %.1f GB/s per GPU, the "
                            "device-resident rate needs %.1f)" % (h2d_gbs, h2d / (ms_step * 1e-3) / 1e9))
                if h2d_gbs < 1.05 * h2d / (ms_step * 1e-3) / 1e9 else "kernels (copy hidden behind compute)",
                "blocking_call_ms_per_step": e2e_blocking_s / args.steps * 1e3,
                "input": "fp32 NCHW images [B,V,3,224,224] (the reference's model input, trainer.py:99-106)"}
    variants = {"fp32": fp32_var}
    if not args.no_u8:
        # raw uint8 HWC images, ToTensor + Normalize folded into the stem loader:
        # 4x fewer bytes over PCIe, same copy accounting
        sess8 = GraphedForward(model, B, V, precision=args.precision, copy_chunks=args.copy_chunks,
                               input_dtype=torch.uint8)
        raw_host = torch.randint(0, 256, (B, V, 224, 224, 3), dtype=torch.uint8, generator=g).pin_memory()
        pipelined8 = make_pipelined(sess8, raw_host)
        pipelined8(3)
        u8_s = timed_wall(ctx, pipelined8, args.steps)

        def blocking8(n):
            for _ in range(n):
                sess8.run_host(raw_host, rot_host)

        blocking8(2)
        u8_block_s = timed_wall(ctx, blocking8, args.steps)
        h2d8 = raw_host.numel() + rot_host.numel() * 4
        variants["uint8"] = {
            "value": world * B * args.steps / u8_s, "unit": UNIT, "ms_per_step": u8_s / args.steps * 1e3,
            "h2d_bytes_per_step": h2d8, "d2h_bytes_per_step": d2h,
            "h2d_gbs_per_gpu_needed": h2d8 / (ms_step * 1e-3) / 1e9,
            "blocking_call_ms_per_step": u8_block_s / args.steps * 1e3,
            "input": "uint8 HWC images [B,V,224,224,3] (what the decoder delivers); normalisation "
                     "(main.py:38-56) runs inside the stem kernel"}
        del sess8
    clocks = sampler.stop() if sampler else None

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": infer_workload(args, B, V), "clocks": clocks,
            "e2e": {"value": fp32_var["value"], "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": fp32_var["ms_per_step"],
                    "h2d_gbs_per_gpu": h2d_gbs,
                    "note": "GraphedForward.submit/result, two calls in flight: the pinned-host "
                            "images of call k+1 are copied on a copy stream while call k computes; "
                            "`value` is the fp32 entry (the reference's tensors), `variants` lists both "
                            "host entries with the same copy accounting",
                    "variants": variants},
            "gpu_launches": launches}

    # ---- roofline of the dominant kernel family (tcgen05 implicit GEMM), measured live ---------
    if rank == 0:
        RF.PROFILE = []
        with torch.no_grad():
            model.engine(args.precision).run(sess.images, sess.rotations, want_all=False)
        torch.cuda.synchronize()
        recs, RF.PROFILE = RF.PROFILE, None
        pk = peaks()
        tc = [(r[1], r[2].elapsed_time(r[3]), r[5].get("bytes", 0.0)) for r in recs if r[0] == "tcgen05"]
        if tc:
            flops = sum(f for f, _, _ in tc)
            ms = sum(t for _, t, _ in tc)
            achieved = flops / (ms * 1e-3) / 1e12
            # per-launch roofline bound (tensor or HBM, whichever is slower), split into the two classes
            cls = {"tensor": [0, 0.0, 0.0, 0.0, 0.0], "hbm": [0, 0.0, 0.0, 0.0, 0.0]}  # n, ms, flops, bytes, bound_ms
            for f, t, by in tc:
                tb, hb = f / (pk["tflops"] * 1e12) * 1e3, by / (pk["hbm_gbs"] * 1e9) * 1e3
                c = cls["tensor" if tb >= hb else "hbm"]
                c[0] += 1; c[1] += t; c[2] += f; c[3] += by; c[4] += max(tb, hb)
            bound_ms = cls["tensor"][4] + cls["hbm"][4]
            traffic, tsrc = load_traffic()
            line["roofline"] = {
                "kernel": "igemm_kernel (tcgen05.mma + TMA implicit GEMM: all 53 convs + 14 linears)",
                "bound": "tensor", "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                "frac": achieved / pk["tflops"], "traffic": traffic, "traffic_source": tsrc,
                "peak_source": pk["src"],
                "launches_per_step": len(tc), "flops_per_launch": flops / len(tc),
                "avg_launch_ms": ms / len(tc), "kernel_share_of_step": ms / ms_step,
                "mixed_bound_frac": bound_ms / ms,   # HEADLINE fraction: sum over launches of max(flop, HBM) bound / measured
                "per_class": {
                    "tensor_bound": {"launches": cls["tensor"][0], "ms": cls["tensor"][1],
                                     "achieved_tflops": cls["tensor"][2] / max(cls["tensor"][1], 1e-9) / 1e9,
                                     "frac": cls["tensor"][4] / max(cls["tensor"][1], 1e-9)},
                    "hbm_bound": {"launches": cls["hbm"][0], "ms": cls["hbm"][1],
                                  "achieved_gbs": cls["hbm"][3] / max(cls["hbm"][1], 1e-9) / 1e6,
                                  "peak_gbs": pk["hbm_gbs"],
                                  "frac": cls["hbm"][4] / max(cls["hbm"][1], 1e-9)}},
                "note": "the family mixes tensor-bound launches (3x3 / deep 1x1) and HBM-bound ones (the "
                        "expanding/reducing 1x1 convs, SURVEY 8d): `frac` is against the tensor peak alone, "
                        "`mixed_bound_frac` against each launch's own roofline, `per_class` splits them",
                "whole_step_frac": (B * V * FLOPS_PER_VIEW / (ms_step * 1e-3) / 1e12) / pk["tflops"]}
        # ---- CPU baseline: the oracle on this box's host cores, bounded sample -----------------
        if world == 1 and not args.no_cpu:
            times = time_cpu_oracle(V, 8, min_seconds=12.0, max_iters=400)
            cpu_value = 8 / (sum(times) / len(times))
            line["cpu_baseline"] = {
                "value": cpu_value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                "sample": f"{len(times)} eval forwards of 8 of the {B} samples ({V} views, fp32, "
                          f"oracle port of the reference PyTorch CPU path), {sum(times):.1f} s"}
    del sess, model
    torch.cuda.empty_cache()
    return line


# --------------------------------------------------------------------------------------------
# configs[3]: training step
# --------------------------------------------------------------------------------------------
def bench_train(args, ctx):
    """BASELINE config 4: Rot-MV 2-view training step (fwd + loss + bwd + Adam), data-parallel,
    batch 128 per GPU; gradients all-reduced over NCCL when N > 1."""
    from rotmv_b200 import functional as RF
    from rotmv_b200.module import FeatRotationSymm
    from rotmv_b200.train import GraphedTrainStep, TrainEngine

    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    B, V = args.train_batch or (args.batch if args.mode == "train" and args.batch else 128), args.views
    W = max(args.warmup, 3)
    torch.manual_seed(0)
    model = FeatRotationSymm(50, 3, precision=args.precision).to(dev).train()
    eng = TrainEngine(model, precision=args.precision, lr=1e-6, weight_decay=1e-6, decoupled=args.adamw)
    g = torch.Generator().manual_seed(1 + rank)
    images_host = torch.randn((B, V, 3, 224, 224), generator=g).pin_memory()
    pose_host = (torch.rand((B, V, 2), generator=g) - 0.5).pin_memory()
    gt_host = (torch.rand((B, V, 2), generator=g) - 0.5).pin_memory()
    images = images_host.to(dev)
    rot = RF.pose_to_rotations(pose_host.to(dev))
    gt = gt_host.to(dev)
    gstep = GraphedTrainStep(eng, B, V)   # forward+backward graphs, [NCCL all-reduces], Adam graph
    gstep.step(images, rot, gt)
    sampler = ClockSampler(ctx.local) if rank == 0 else None
    for _ in range(W - 1):
        gstep.step()
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record()
    for _ in range(args.steps):
        gstep.step()
        launches += gstep.launches_per_step
    e1.record()
    ctx.barrier()
    ms_total = ctx.max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    loss = eng.loss.item()
    if not (loss == loss):
        raise SystemExit("bench.py: non-finite loss")

    # end to end: host batch (images, head poses, labels) in, loss out, every step
    pose_d = torch.empty((B, V, 2), device=dev)
    loss_h = torch.empty((1,), dtype=torch.float32).pin_memory()

    def blocking(n):
        for _ in range(n):
            gstep.images.copy_(images_host, non_blocking=True)
            pose_d.copy_(pose_host, non_blocking=True)
            gstep.gt.copy_(gt_host, non_blocking=True)
            RF.pose_to_rotations(pose_d, out=gstep.rotations)
            loss_h.copy_(gstep.step(), non_blocking=True)
            torch.cuda.current_stream().synchronize()

    # the same host batch through the asynchronous entry (two steps in flight): step k+1's
    # host->HBM copies overlap step k; every step's H2D and its loss read-back are timed
    last = [0.0]

    def pipelined(n):
        prev = None
        for _ in range(n):
            tk = gstep.submit(images_host, pose_host, gt_host)
            if prev is not None:
                gstep.result(prev)
            prev = tk
        last[0] = float(gstep.result(prev))

    blocking(2)
    e2e_blocking_s = timed_wall(ctx, blocking, args.steps)
    pipelined(3)
    e2e_s = timed_wall(ctx, pipelined, args.steps)
    if not (last[0] == last[0]):
        raise SystemExit("bench.py: non-finite loss (host path)")
    clocks = sampler.stop() if sampler else None

    # exposed all-reduce time = step time with the collectives - step time without them (measured
    # last: without the all-reduce the replicas drift apart)
    ar = None
    grad_bytes = eng.n_trained * 4
    if world > 1:
        gstep.skip_allreduce = True
        for _ in range(3):
            gstep.step()
        ctx.barrier()
        e0.record()
        for _ in range(args.steps):
            gstep.step()
        e1.record()
        ctx.barrier()
        ms_nocomm = ctx.max_over_ranks(e0.elapsed_time(e1)) / args.steps
        gstep.skip_allreduce = False
        ar = {"bytes": eng.flat_g.numel() * 4, "collective": gstep.collective_desc(),
              "ms_per_step_without_collectives": ms_nocomm,
              "exposed_ms": ms_step - ms_nocomm}

    train_flops = 3 * V * FLOPS_PER_VIEW - V * 0.236e9
    pk = peaks()
    achieved = B * train_flops / (ms_step * 1e-3) / 1e12
    roof = {"kernel": "whole step (conv fwd/dgrad/wgrad on tcgen05 + HBM-bound BN/elementwise)",
            "bound": "tensor", "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
            "frac": achieved / pk["tflops"], "traffic": None, "peak_source": pk["src"]}
    if rank == 0:
        # per-class device time of one eager step (CUDA events around every launch, warm)
        RF.PROFILE = []
        eng.forward_backward(gstep.images, gstep.rotations, gstep.gt)
        torch.cuda.synchronize()
        recs, RF.PROFILE = RF.PROFILE, None
        split = {"bn_ms": 0.0, "conv_ms": 0.0, "wgrad_ms": 0.0, "other_ms": 0.0}
        for engn, _, a0, a1, what, _ in recs:
            t = a0.elapsed_time(a1)
            if what.startswith("rmv_bn") or engn == "tcgen05-bnstat":   # incl. the recomputed-conv reductions
                split["bn_ms"] += t
            elif engn in ("tcgen05", "tcgen05-stem"):
                split["conv_ms"] += t
            elif engn == "tcgen05-wgrad":
                split["wgrad_ms"] += t
            else:
                split["other_ms"] += t
        roof.update(split)
        roof["note"] = ("bn_ms/conv_ms/wgrad_ms/other_ms: per-launch CUDA-event times of one eager "
                        "forward+backward (Adam excluded), summed per class; bn_ms = the BatchNorm statistics / "
                        "apply kernels incl. the recomputed-conv reductions, conv_ms = forward convs and data "
                        "gradients incl. the convs that carry a BatchNorm-apply epilogue")
    block = {"metric": METRIC_TRAIN, "value": value, "unit": UNIT,
             "n_gpus": world, "steps": args.steps, "warmup": W,
             "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
             "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
             "config": train_workload(args, B, V, world, grad_bytes),
             "clocks": clocks, "loss": loss,
             "e2e": {"value": world * B * args.steps / e2e_s, "unit": UNIT,
                     "h2d_bytes_per_step": images_host.numel() * 4 + pose_host.numel() * 4 + gt_host.numel() * 4,
                     "d2h_bytes_per_step": 4, "ms_per_step": e2e_s / args.steps * 1e3,
                     "note": "GraphedTrainStep.submit/result, two steps in flight (host batch = fp32 "
                             "images + head poses + labels; loss read back every step)",
                     "blocking_call": {"value": world * B * args.steps / e2e_blocking_s, "unit": UNIT,
                                       "ms_per_step": e2e_blocking_s / args.steps * 1e3}},
             "gpu_launches": launches, "roofline": roof, "allreduce": ar}
    if rank == 0 and world == 1 and not args.no_cpu:
        times = time_cpu_oracle(V, 8, min_seconds=10.0, max_iters=200, train=True)
        block["cpu_baseline"] = {
            "value": 8 / (sum(times) / len(times)), "unit": UNIT, "cores": torch.get_num_threads(),
            "kind": "port",
            "sample": f"{len(times)} training steps (fwd+loss+bwd+Adam) on 8 of the {B} samples ({V} views, "
                      f"fp32, oracle port of the reference PyTorch CPU path), {sum(times):.1f} s"}
    del gstep, eng, model
    torch.cuda.empty_cache()
    return block


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="both", choices=["both", "infer", "train"],
                    help="both (default) = configs[1] headline + configs[3] in the `train` block; "
                         "infer / train = one of them alone")
    ap.add_argument("--batch", type=int, default=None, help="multi-view samples per GPU per step "
                    "(default 256 for infer, 128 for train)")
    ap.add_argument("--train-batch", type=int, default=None, help="training batch per GPU (default 128)")
    ap.add_argument("--views", type=int, default=2)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--chunk", type=int, default=int(os.environ.get("ROTMV_CHUNK", "512")),
                    help="images per trunk micro-batch")
    ap.add_argument("--copy-chunks", type=int, default=4,
                    help="infer e2e: batch slices whose host->device copy overlaps the trunk")
    ap.add_argument("--adamw", action="store_true", help="train: decoupled weight decay")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-u8", action="store_true", help="skip the uint8-input end-to-end variant")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    ctx = Ctx()
    line = None
    if args.mode in ("both", "infer"):
        line = bench_infer(args, ctx)
    if args.mode in ("both", "train"):
        tr = bench_train(args, ctx)
        if line is None:
            line = tr
        else:
            line["train"] = tr
    if ctx.rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
