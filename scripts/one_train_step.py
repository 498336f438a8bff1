"""Three eager (graph-free) training steps of the configs[3] workload (B=128, V=2, bf16): the ncu
target for the training path (every launch is a plain stream launch). Prints each step's device time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rot-mvgaze_b200"))
import torch
from rotmv_b200.module import FeatRotationSymm
from rotmv_b200.train import TrainEngine
from rotmv_b200 import functional as RF
from rotmv_b200 import _lib as L

B = int(os.environ.get("B", 128)); V = int(os.environ.get("V", 2))
torch.manual_seed(0)
model = FeatRotationSymm(50, 3).cuda().train()
eng = TrainEngine(model, precision="bf16", lr=1e-6)
images = torch.randn((B, V, 3, 224, 224), device="cuda")
rot = RF.pose_to_rotations(torch.rand((B, V, 2), device="cuda") - 0.5)
gt = torch.rand((B, V, 2), device="cuda") - 0.5
for i in range(3):
    n0 = L.STATS["launches"]
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    if i == 2:
        torch.cuda.profiler.start()     # ncu --profile-from-start off: only the third (warm) step
    e0.record()
    eng.step(images, rot, gt)
    e1.record(); torch.cuda.synchronize()
    if i == 2:
        torch.cuda.profiler.stop()
    print(f"step {i}: {L.STATS['launches'] - n0} launches, {e0.elapsed_time(e1):.3f} ms, loss {eng.loss.item():.4f}")
