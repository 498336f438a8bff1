#!/usr/bin/env python
"""Headline benchmark: Rot-MV multi-view inference throughput (BASELINE.json configs[1]).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

A "step" is one forward pass of `FeatRotationSymm` (ResNet-50 trunk -> rotation-constrained
cross-view fusion x3 -> gaze heads) over one batch of 256 two-view samples (512 images, 224x224),
random-init weights, synthetic inputs. Prints ONE JSON line (rank 0):
  value      whole-job multi-view samples/s, inputs resident in HBM, CUDA-graph replay, CUDA events
  e2e        same metric through the host-buffer entry (pinned host -> HBM copy of images+rotations
             and device -> host read of pred_gaze inside the timed region, every step)
  roofline   tcgen05 implicit-GEMM kernel: algorithmic FLOPs / measured launch time vs measured peak
  cpu_baseline  the CPU oracle (port of the reference's PyTorch path) on the host cores
`--impl reference` times the reference's own CPU path (oracle port; /root/reference does not travel
to the GPU box) on a bounded sample of the same workload.
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


def workload(args):
    return {"workload": "configs[1]: Rot-MV 2-view inference, synthetic batch 256 bf16 on 1xB200 "
                        "(xgaze2mpiinv_known shape)" if (args.batch, args.views) == (256, 2)
            else f"Rot-MV {args.views}-view inference, synthetic batch {args.batch}",
            "batch_per_gpu": args.batch, "views": args.views, "image": "3x224x224 fp32 NCHW",
            "backbone": "resnet50", "num_iter": 3, "weights": "random-init (seed 0)",
            "mode": "eval forward (no_grad)", "precision": args.precision,
            "parallelism": f"dp{args.gpus} (batch sharded, no collective)",
            "l2": "inputs %.0f MB/step per GPU exceed the 126 MB L2; no flush needed"
                  % (args.batch * args.views * 3 * 224 * 224 * 4 / 1e6)}


# --------------------------------------------------------------------------------------------
# CPU oracle legs
# --------------------------------------------------------------------------------------------
def time_cpu_oracle(views: int, sample_batch: int, min_seconds: float, max_iters: int,
                    warmup: int = 1):
    from oracle import rotmv_oracle as O

    torch.set_num_threads(os.cpu_count())
    model = O.build_model(num_iter=3, depth=50, seed=0).eval()
    images, pose, _ = O.synthetic_batch(sample_batch, views, seed=1)
    rot = O.pairwise_rotations(pose)
    times = []
    with torch.no_grad():
        for _ in range(warmup):
            model.forward_views(images, rot)
        t_all = time.perf_counter()
        while len(times) < max_iters and (time.perf_counter() - t_all < min_seconds or len(times) < 2):
            t0 = time.perf_counter()
            model.forward_views(images, rot)
            times.append(time.perf_counter() - t0)
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 8
    from oracle import rotmv_oracle as O

    torch.set_num_threads(os.cpu_count())
    model = O.build_model(num_iter=3, depth=50, seed=0).eval()
    images, pose, gt = O.synthetic_batch(sample, args.views, seed=1)
    rot = O.pairwise_rotations(pose)
    metric = METRIC
    if args.mode == "train":
        metric = "multi-view samples/sec (224^2, fwd+bwd+Adam)"
        opt = O.make_adam(model, lr=1e-6)
        for _ in range(args.warmup):
            O.train_step(model, opt, images, rot, gt)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            O.train_step(model, opt, images, rot, gt)
        dt = time.perf_counter() - t0
    else:
        with torch.no_grad():
            for _ in range(args.warmup):
                model.forward_views(images, rot)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                model.forward_views(images, rot)
            dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    desc = (f"oracle port of the reference PyTorch CPU path; each step = {sample} of the "
            f"{args.batch} samples of the workload batch ({args.views} views, fp32, {args.mode})")
    line = {"impl": "reference", "metric": metric, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload(args),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(),
                             "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    from rotmv_b200 import _lib as L
    from rotmv_b200 import functional as RF
    from rotmv_b200.engine import GraphedForward
    from rotmv_b200.module import FeatRotationSymm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    torch.manual_seed(0)
    model = FeatRotationSymm(50, 3, precision=args.precision, trunk_chunk=args.chunk)
    model = model.to(dev).eval()
    g = torch.Generator().manual_seed(1 + rank)
    B, V = args.batch, args.views
    images_host = torch.randn((B, V, 3, 224, 224), generator=g).pin_memory()
    pose_host = torch.rand((B, V, 2), generator=g) - 0.5
    rot_dev = RF.pose_to_rotations(pose_host.to(dev))
    rot_host = rot_dev.cpu().pin_memory()
    images_dev = images_host.to(dev)

    sess = GraphedForward(model, B, V, precision=args.precision, copy_chunks=args.copy_chunks)
    sess.images.copy_(images_dev)
    sess.rotations.copy_(rot_dev)
    del images_dev

    # ---- device-resident throughput: warm-up, then K timed replays ---------------------------
    # clocks are sampled from the last warm-up replays (the same load) through the timed region, so
    # that a short timed region (K x 7 ms) still yields samples
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(args.warmup, 3) + 10):
        sess()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = 0
    e0.record()
    for _ in range(args.steps):
        sess()
        launches0 += sess.launches_per_replay
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / args.steps
    value = n_gpus * B * args.steps / (ms_total * 1e-3)
    pred_check = sess.pred.float().abs().sum().item()
    if not (pred_check == pred_check):
        raise SystemExit("bench.py: non-finite predictions")

    # ---- end to end: host buffers in, prediction out, every step -----------------------------
    for _ in range(2):
        sess.run_host(images_host, rot_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sess.run_host(images_host, rot_host)
    barrier()
    e2e_blocking_s = max_over_ranks(time.perf_counter() - t0)
    # the same host entry used asynchronously (two calls in flight): call k+1's host->HBM copy
    # overlaps call k's kernels; every step's H2D and its D2H read are inside the timed region
    def pipelined(n):
        prev = None
        for _ in range(n):
            tk = sess.submit(images_host, rot_host)
            if prev is not None:
                sess.result(prev)
            prev = tk
        return sess.result(prev)

    pipelined(3)
    barrier()
    t0 = time.perf_counter()
    pred_last = pipelined(args.steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    if not bool(torch.isfinite(pred_last).all()):
        raise SystemExit("bench.py: non-finite predictions (host path)")
    e2e_value = n_gpus * B * args.steps / e2e_s
    h2d = images_host.numel() * 4 + rot_host.numel() * 4
    d2h = B * 2 * 4

    # extra (not the headline): the same batch as raw uint8 HWC images, ToTensor + Normalize folded
    # into the stem loader -- 4x fewer bytes over PCIe
    u8 = None
    if not args.no_u8:
        sess8 = GraphedForward(model, B, V, precision=args.precision, copy_chunks=args.copy_chunks,
                               input_dtype=torch.uint8)
        raw_host = torch.randint(0, 256, (B, V, 224, 224, 3), dtype=torch.uint8, generator=g).pin_memory()

        def pipelined8(n):
            prev = None
            for _ in range(n):
                tk = sess8.submit(raw_host, rot_host)
                if prev is not None:
                    sess8.result(prev)
                prev = tk
            return sess8.result(prev)

        pipelined8(3)
        barrier()
        t0 = time.perf_counter()
        pipelined8(args.steps)
        barrier()
        u8_s = max_over_ranks(time.perf_counter() - t0)
        for _ in range(2):
            sess8.run_host(raw_host, rot_host)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            sess8.run_host(raw_host, rot_host)
        barrier()
        u8_block_s = max_over_ranks(time.perf_counter() - t0)
        u8 = {"value": n_gpus * B * args.steps / u8_s, "unit": UNIT, "ms_per_step": u8_s / args.steps * 1e3,
              "h2d_bytes_per_step": raw_host.numel() + rot_host.numel() * 4,
              "blocking_call_ms_per_step": u8_block_s / args.steps * 1e3,
              "note": "input = uint8 HWC images [B,V,224,224,3] (what the decoder delivers); "
                      "normalisation (main.py:38-56) runs inside the stem kernel"}
        del sess8

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3) + 10, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": workload(args), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3,
                    "note": "GraphedForward.submit/result, two calls in flight: the pinned-host fp32 "
                            "images of call k+1 are copied on a copy stream while call k computes",
                    "blocking_call": {"value": n_gpus * B * args.steps / e2e_blocking_s, "unit": UNIT,
                                      "ms_per_step": e2e_blocking_s / args.steps * 1e3,
                                      "note": f"GraphedForward.run_host (one call at a time, returns "
                                              f"the prediction): images copied in {len(sess.slices)} "
                                              "slices overlapped with the trunk of the previous slice"},
                    "uint8_input": u8},
            "gpu_launches": launches0}

    # ---- roofline of the dominant kernel (tcgen05 implicit GEMM), measured live ----------------
    if rank == 0:
        RF.PROFILE = []
        with torch.no_grad():
            model.engine(args.precision).run(sess.images, sess.rotations, want_all=False)
        torch.cuda.synchronize()
        recs, RF.PROFILE = RF.PROFILE, None
        tc = [(r[1], r[2].elapsed_time(r[3])) for r in recs if r[0] == "tcgen05"]
        pk = peaks()
        # per-launch roofline bound (tensor or HBM, whichever is slower) summed over the igemm launches
        bound_ms = sum(max(r[1] / (pk["tflops"] * 1e12), r[5].get("bytes", 0.0) / (pk["hbm_gbs"] * 1e9)) * 1e3
                       for r in recs if r[0] == "tcgen05")
        if tc:
            flops = sum(f for f, _ in tc)
            ms = sum(t for _, t in tc)
            achieved = flops / (ms * 1e-3) / 1e12
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "igemm_traffic.json")
            if os.path.exists(tpath):
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            line["roofline"] = {
                "kernel": "igemm_kernel (tcgen05.mma + TMA implicit GEMM: all 53 convs + 14 linears)",
                "bound": "tensor", "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                "frac": achieved / pk["tflops"], "traffic": traffic, "peak_source": pk["src"],
                "launches_per_step": len(tc), "flops_per_launch": flops / len(tc),
                "avg_launch_ms": ms / len(tc), "kernel_share_of_step": ms / ms_step,
                "mixed_bound_frac": bound_ms / ms,   # sum over launches of max(flop, HBM) bound / measured
                "note": "30 of the 53 convs are HBM-bound at bf16 (SURVEY 8d): frac is against the "
                        "tensor peak alone, mixed_bound_frac against each launch's own roofline",
                "whole_step_frac": (B * V * FLOPS_PER_VIEW / (ms_step * 1e-3) / 1e12) / pk["tflops"]}
        # ---- CPU baseline: the oracle on this box's host cores, bounded sample -----------------
        if n_gpus == 1 and not args.no_cpu:
            times = time_cpu_oracle(V, 8, min_seconds=12.0, max_iters=400)
            cpu_value = 8 / (sum(times) / len(times))
            line["cpu_baseline"] = {
                "value": cpu_value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                "sample": f"{len(times)} eval forwards of 8 of the {B} samples ({V} views, fp32, "
                          f"oracle port of the reference PyTorch CPU path), {sum(times):.1f} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_train(args):
    """BASELINE config 4: Rot-MV 2-view training step (fwd + loss + bwd + Adam), data-parallel,
    batch 128 per GPU; gradients all-reduced over NCCL when N > 1."""
    import torch.distributed as dist

    from rotmv_b200 import _lib as L
    from rotmv_b200 import functional as RF
    from rotmv_b200.module import FeatRotationSymm
    from rotmv_b200.train import GraphedTrainStep, TrainEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    B, V = args.batch, args.views
    torch.manual_seed(0)
    model = FeatRotationSymm(50, 3, precision=args.precision).to(dev).train()
    eng = TrainEngine(model, precision=args.precision, lr=1e-6, weight_decay=1e-6,
                      decoupled=args.adamw)
    g = torch.Generator().manual_seed(1 + rank)
    images_host = torch.randn((B, V, 3, 224, 224), generator=g).pin_memory()
    pose_host = (torch.rand((B, V, 2), generator=g) - 0.5).pin_memory()
    gt_host = (torch.rand((B, V, 2), generator=g) - 0.5).pin_memory()
    images = images_host.to(dev)
    rot = RF.pose_to_rotations(pose_host.to(dev))
    gt = gt_host.to(dev)
    gstep = GraphedTrainStep(eng, B, V)   # forward+backward graph, [NCCL all-reduce], Adam graph
    gstep.step(images, rot, gt)
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(args.warmup, 3) + 4):
        gstep.step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record()
    for _ in range(args.steps):
        gstep.step()
        launches += gstep.launches_per_step
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    value = world * B * args.steps / (ms_total * 1e-3)
    loss = eng.loss.item()
    if not (loss == loss):
        raise SystemExit("bench.py: non-finite loss")
    # end to end: host batch (images, head poses, labels) in, loss out, every step
    pose_d = torch.empty((B, V, 2), device=dev)
    loss_h = torch.empty((1,), dtype=torch.float32).pin_memory()

    def host_step():
        gstep.images.copy_(images_host, non_blocking=True)
        pose_d.copy_(pose_host, non_blocking=True)
        gstep.gt.copy_(gt_host, non_blocking=True)
        gstep.rotations.copy_(RF.pose_to_rotations(pose_d), non_blocking=True)
        loss_h.copy_(gstep.step(), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(2):
        host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_step()
    barrier()
    e2e_blocking_s = max_over_ranks(time.perf_counter() - t0)

    # the same host batch through the asynchronous entry (two steps in flight): step k+1's
    # host->HBM copies overlap step k; every step's H2D and its loss read-back are timed
    def pipelined(n):
        prev = None
        for _ in range(n):
            tk = gstep.submit(images_host, pose_host, gt_host)
            if prev is not None:
                gstep.result(prev)
            prev = tk
        return gstep.result(prev)

    pipelined(3)
    barrier()
    t0 = time.perf_counter()
    last_loss = float(pipelined(args.steps))
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    if not (last_loss == last_loss):
        raise SystemExit("bench.py: non-finite loss (host path)")
    train_flops = 3 * V * FLOPS_PER_VIEW - V * 0.236e9
    pk = peaks()
    line = {"metric": "multi-view samples/sec (224^2, fwd+bwd+Adam)", "value": value, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3) + 4,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "configs[3]: Rot-MV 2-view training step (fwd+bwd+Adam), "
                                   "data-parallel, batch 128 per GPU" if (B, V) == (128, 2)
                       else f"Rot-MV {V}-view training step, batch {B} per GPU",
                       "batch_per_gpu": B, "views": V, "backbone": "resnet50", "num_iter": 3,
                       "optimizer": "AdamW (decoupled)" if args.adamw else
                       "Adam + coupled L2 (reference trainer.py:54)",
                       "parallelism": f"dp{world}, one fp32 gradient all-reduce of "
                                      f"{eng.flat_g.numel() * 4 / 1e6:.0f} MB per step" if world > 1
                       else "dp1", "precision": args.precision,
                       "l2": "inputs %.0f MB/step per GPU exceed the 126 MB L2" % (B * V * 602112 / 1e6)},
            "clocks": clocks, "loss": loss,
            "e2e": {"value": world * B * args.steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": images_host.numel() * 4 + pose_host.numel() * 4 + gt_host.numel() * 4,
                    "d2h_bytes_per_step": 4, "ms_per_step": e2e_s / args.steps * 1e3,
                    "note": "GraphedTrainStep.submit/result, two steps in flight (host batch = fp32 "
                            "images + head poses + labels; loss read back every step)",
                    "blocking_call": {"value": world * B * args.steps / e2e_blocking_s, "unit": UNIT,
                                      "ms_per_step": e2e_blocking_s / args.steps * 1e3}},
            "gpu_launches": launches,
            "roofline": {"kernel": "whole step (conv fwd/dgrad/wgrad on tcgen05 + HBM-bound BN/elementwise)",
                         "bound": "tensor", "achieved": B * train_flops / (ms_total / args.steps * 1e-3) / 1e12,
                         "peak": pk["tflops"], "unit": "TFLOP/s",
                         "frac": B * train_flops / (ms_total / args.steps * 1e-3) / 1e12 / pk["tflops"],
                         "traffic": None, "peak_source": pk["src"]}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="infer = BASELINE configs[1] (headline); train = configs[3] (fwd+bwd+Adam)")
    ap.add_argument("--batch", type=int, default=None, help="multi-view samples per GPU per step "
                    "(default 256 for infer, 128 for train)")
    ap.add_argument("--views", type=int, default=2)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--chunk", type=int, default=int(os.environ.get("ROTMV_CHUNK", "512")),
                    help="images per trunk micro-batch")
    ap.add_argument("--copy-chunks", type=int, default=4,
                    help="infer e2e: batch slices whose host->device copy overlaps the trunk")
    ap.add_argument("--adamw", action="store_true", help="train: decoupled weight decay")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-u8", action="store_true", help="skip the extra uint8-input end-to-end leg")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 256 if args.mode == "infer" else 128
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "train":
        run_train(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
