"""Front chunking of the inference trunk (ROTMV_FRONT_CHUNK): same predictions as the unchunked
forward (bit-identical: every output row is computed from its own rows in the same order)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rot-mvgaze_b200"))
import torch
from rotmv_b200 import functional as RF
from rotmv_b200.module import FeatRotationSymm

torch.manual_seed(0)
model = FeatRotationSymm(50, 3).cuda().eval()
model.auto_graph = False
images = torch.randn((40, 2, 3, 224, 224), device="cuda")
rot = RF.pose_to_rotations(torch.rand((40, 2, 2), device="cuda") - 0.5)
with torch.no_grad():
    ref = model(images, rot).clone()
    for g in (16, 24, 32):
        eng = model.engine("bf16")
        eng.front_chunk = g
        out = model(images, rot)
        print(f"front_chunk={g}: max |diff| {float((out - ref).abs().max()):.3e}, equal {torch.equal(out, ref)}")
        assert torch.isfinite(out).all() and (out - ref).abs().max().item() <= 1e-5
