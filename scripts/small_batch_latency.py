"""Latency of small batches: eager module call (ctypes launches + tensor-map encodes per call) vs
CUDA-graph session (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rot-mvgaze_b200"))
import torch
from rotmv_b200.module import FeatRotationSymm
from rotmv_b200.engine import GraphedForward
from rotmv_b200 import functional as RF
torch.manual_seed(0)
model = FeatRotationSymm(50, 3).cuda().eval()
for B in (1, 2, 8, 32):
    images = torch.randn((B, 2, 3, 224, 224), device="cuda")
    rot = RF.pose_to_rotations(torch.rand((B, 2, 2), device="cuda") - 0.5)
    with torch.no_grad():
        for _ in range(3): model(images, rot)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): p = model(images, rot)
        torch.cuda.synchronize(); eager = (time.perf_counter() - t0) / 20 * 1e3
    sess = GraphedForward(model, B, 2)
    sess(images, rot); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): sess(images, rot)
    torch.cuda.synchronize(); graph = (time.perf_counter() - t0) / 20 * 1e3
    print(f"B={B}: eager {eager:.3f} ms, graph {graph:.3f} ms per call")
