"""2-GPU correctness check of the data-parallel training step (torchrun --nproc-per-node 2):
both ranks get the SAME batch, so the averaged gradient equals the single-GPU gradient and the
parameters after K graphed steps (overlapped two-slice NCCL all-reduce + fused Adam) must match a
single-process run of the same steps up to the bf16/split-K summation noise."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rot-mvgaze_b200"))
import torch
import torch.distributed as dist
from rotmv_b200 import functional as RF
from rotmv_b200.module import FeatRotationSymm
from rotmv_b200.train import GraphedTrainStep, TrainEngine

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
B, V, K = 8, 2, 3
g = torch.Generator().manual_seed(7)
images = torch.randn((B, V, 3, 224, 224), generator=g).to(dev)
pose = (torch.rand((B, V, 2), generator=g) - 0.5).to(dev)
gt = (torch.rand((B, V, 2), generator=g) - 0.5).to(dev)
rot = RF.pose_to_rotations(pose)


def run(tag):
    torch.manual_seed(0)
    model = FeatRotationSymm(50, 3).to(dev).train()
    eng = TrainEngine(model, precision="bf16", lr=1e-3, weight_decay=1e-6)
    p0 = eng.flat_p.clone()
    step = GraphedTrainStep(eng, B, V)
    step.step(images, rot, gt)
    torch.cuda.synchronize()
    g1 = eng.flat_g.clone() / eng.world      # gradient of step 1 (summed over the ranks when dp)
    for _ in range(K - 1):
        step.step()
    torch.cuda.synchronize()
    return eng, p0, eng.flat_p.clone(), eng.loss.item(), g1

eng1, p0, p_single, loss_single, g_single = run("single")      # before init_process_group: world == 1
_, _, p_again, _, g_again = run("single again")                  # run-to-run noise floor (atomics order)
assert eng1.world == 1
dist.init_process_group("nccl", device_id=dev)
eng2, p0b, p_dp, loss_dp, g_dp = run("dp")
assert eng2.world == dist.get_world_size() == 2 and torch.equal(p0, p0b)
d_single, d_dp = (p_single - p0), (p_dp - p0)
split = eng2.grad_split
rel = ((g_dp - g_single).norm() / g_single.norm()).item()
rel_trunk = ((g_dp[:split] - g_single[:split]).norm() / g_single[:split].norm()).item()
rel_fusion = ((g_dp[split:] - g_single[split:]).norm() / g_single[split:].norm()).item()
noise = ((g_again - g_single).norm() / g_single.norm()).item()
upd = ((d_dp - d_single).norm() / d_single.norm()).item()
upd_noise = (((p_again - p0) - d_single).norm() / d_single.norm()).item()
# both ranks must hold identical parameters after the step (same reduced gradient everywhere)
other = p_dp.clone()
dist.broadcast(other, src=0)
same = torch.equal(other, p_dp)
if rank == 0:
    print(f"loss single {loss_single:.5f} dp {loss_dp:.5f}; step-1 gradient dp vs single: all {rel:.3e}, trunk slice "
          f"{rel_trunk:.3e}, fusion slice {rel_fusion:.3e}; single vs single again (noise floor) {noise:.3e}; "
          f"parameter update after {K} Adam steps dp vs single {upd:.3e}, single vs single {upd_noise:.3e}; "
          f"ranks identical: {same}")
ok = torch.tensor([1 if (same and rel <= max(3 * noise, 1e-3)) else 0], device=dev)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if ok.item() == 1 else 1)
