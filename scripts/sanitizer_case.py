"""Smallest end-to-end case for compute-sanitizer memcheck: one bf16 forward (B=2, V=2) and one bf16
training step (B=4, V=2) through every kernel family of the library."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rot-mvgaze_b200"))
import torch
from rotmv_b200.module import FeatRotationSymm
from rotmv_b200.train import TrainEngine
from rotmv_b200 import functional as RF

torch.manual_seed(0)
model = FeatRotationSymm(50, 2).cuda().eval()
images = torch.randn((2, 2, 3, 224, 224), device="cuda")
rot = RF.pose_to_rotations(torch.rand((2, 2, 2), device="cuda") - 0.5)
with torch.no_grad():
    p = model(images, rot)
    raw = torch.randint(0, 256, (2, 2, 224, 224, 3), dtype=torch.uint8, device="cuda")
    p8 = model(raw, rot)
torch.cuda.synchronize()
print("forward ok", p[0].tolist(), p8[0].tolist())
model.train()
eng = TrainEngine(model, precision="bf16", lr=1e-6)
images = torch.randn((4, 2, 3, 224, 224), device="cuda")
rot = RF.pose_to_rotations(torch.rand((4, 2, 2), device="cuda") - 0.5)
gt = torch.rand((4, 2, 2), device="cuda") - 0.5
for _ in range(2):
    loss = eng.step(images, rot, gt)
torch.cuda.synchronize()
print("train ok", loss.item())
