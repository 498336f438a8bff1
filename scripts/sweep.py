"""BASELINE configs[4]: throughput sweep over views {2,4,8} x batch {64..2048} (per GPU), forward
(eval, CUDA graph) and training step (fwd+bwd+Adam, CUDA graph), device-resident inputs, CUDA events.
One process per GPU under torchrun (weak scaling: every rank runs the same per-GPU batch; the
training step all-reduces the gradients). Prints one JSON line per point (rank 0).

    python scripts/sweep.py [--views 2,4,8] [--batches 64,128,...] [--modes infer,train] [--steps 5]
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rot-mvgaze_b200"))
import torch
import torch.distributed as dist
from rotmv_b200 import functional as RF
from rotmv_b200.engine import GraphedForward
from rotmv_b200.module import FeatRotationSymm
from rotmv_b200.train import GraphedTrainStep, TrainEngine

ap = argparse.ArgumentParser()
ap.add_argument("--views", default="2,4,8")
ap.add_argument("--batches", default="64,128,256,512,1024,2048")
ap.add_argument("--modes", default="infer,train")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--global-batches", default="", help="GLOBAL batch sizes (configs[4]: 64..2048 over all GPUs); "
                "each GPU runs global/world samples; overrides --batches")
ap.add_argument("--cpu-only", action="store_true", help="time ONLY the CPU oracle (reference PyTorch CPU path port) "
                "per view count and mode, B=8 samples, all host cores -- the CPU column of configs[4]. Run it "
                "as its own single process: with 8 ranks spinning in NCCL the host cores are not free "
                "(measured: 1.5 instead of 61 samples/s)")
ap.add_argument("--cpu-json", default="", help="JSON line written by a --cpu-only run: adds cpu_samples_per_s / "
                "gpu_over_cpu to every point")
ap.add_argument("--max-images", type=int, default=4096, help="skip points with batch*views above this (memory)")
ap.add_argument("--max-train-images", type=int, default=1024)
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
cpu = {}
if args.cpu_only:
    # the CPU column comes from `bench.py --impl reference` (the one program besides tests/ and smoke()
    # that executes oracle/): one run per view count, inference + training step, B = 8 samples per step
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for v in [int(x) for x in args.views.split(",")]:
        res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--views", str(v),
                              "--steps", str(max(3, args.steps)), "--warmup", "1", "--mode", "both"],
                             check=True, capture_output=True, text=True)
        line = json.loads(res.stdout.strip().splitlines()[-1])
        cpu[("infer", v)] = line["value"]
        cpu[("train", v)] = line["train"]["value"]
        cores = line["cpu_baseline"]["cores"]
    print(json.dumps({"cpu_reference": {f"{m}_v{v}": round(x, 2) for (m, v), x in cpu.items()
                                        if m in args.modes.split(",")},
                      "cores": cores, "sample": "bench.py --impl reference (oracle port of the reference CPU path), "
                                                "B=8 samples per step, fp32"}), flush=True)
    sys.exit(0)
if args.cpu_json:
    rec = json.loads(open(args.cpu_json).read().strip().splitlines()[-1])["cpu_reference"]
    cpu = {tuple([k.rsplit("_v", 1)[0], int(k.rsplit("_v", 1)[1])]): x for k, x in rec.items()}
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def timed(fn, steps):
    for _ in range(3): fn()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(steps): fn()
    e1.record()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


batches = [int(x) for x in args.batches.split(",")]
if args.global_batches:
    batches = [int(x) // world for x in args.global_batches.split(",") if int(x) // world >= 1]
torch.manual_seed(0)
models = {}
for mode in args.modes.split(","):
    for v in [int(x) for x in args.views.split(",")]:
        for b in batches:
            cap = args.max_images if mode == "infer" else args.max_train_images
            if b * v > cap:
                continue
            if mode == "infer":
                if "infer" not in models:
                    models["infer"] = FeatRotationSymm(50, 3).to(dev)
                model = models["infer"].eval()
                sess = GraphedForward(model, b, v)
                sess.images.normal_(); sess.rotations.copy_(RF.pose_to_rotations(torch.rand((b, v, 2), device=dev) - 0.5))
                ms = timed(sess, args.steps)
                del sess
            else:
                if "train" not in models:
                    models["train"] = FeatRotationSymm(50, 3).to(dev)
                model = models["train"].train()
                eng = TrainEngine(model, precision="bf16", lr=1e-6, weight_decay=1e-6)
                g = GraphedTrainStep(eng, b, v)
                g.step(torch.randn((b, v, 3, 224, 224), device=dev), RF.pose_to_rotations(torch.rand((b, v, 2), device=dev) - 0.5),
                       torch.rand((b, v, 2), device=dev) - 0.5)
                ms = timed(g.step, args.steps)
                loss = eng.loss.item()
                assert loss == loss, "non-finite loss"
                del g, eng
            torch.cuda.empty_cache()
            if rank == 0:
                rec = {"mode": mode, "views": v, "batch_per_gpu": b, "global_batch": b * world, "n_gpus": world,
                       "ms_per_step": round(ms, 3), "samples_per_s": round(world * b / ms * 1e3, 1),
                       "images_per_s": round(world * b * v / ms * 1e3, 1)}
                if (mode, v) in cpu:
                    rec["cpu_samples_per_s"] = round(cpu[(mode, v)], 2)
                    rec["gpu_over_cpu"] = round(rec["samples_per_s"] / cpu[(mode, v)], 1)
                print(json.dumps(rec), flush=True)
if world > 1:
    dist.destroy_process_group()
