"""Quick device-side timing of the inference forward (development aid, not the bench)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rot-mvgaze_b200"))
import torch
from rotmv_b200.module import FeatRotationSymm
from rotmv_b200 import functional as RF

B = int(os.environ.get("B", 256)); V = int(os.environ.get("V", 2))
chunks = [int(c) for c in os.environ.get("CHUNKS", "32").split(",")]
torch.manual_seed(0)
model = FeatRotationSymm(50, 3).cuda().eval()
images = torch.randn((B, V, 3, 224, 224), device="cuda")
pose = torch.rand((B, V, 2), device="cuda") - 0.5
rot = RF.pose_to_rotations(pose)
for chunk in chunks:
    model.trunk_chunk = chunk; model.invalidate()
    with torch.no_grad():
        for _ in range(2): model(images, rot)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        t0 = time.perf_counter(); e0.record()
        n = 3
        for _ in range(n): model(images, rot)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"chunk {chunk}: B={B} V={V} fwd {ms:.2f} ms  -> {B / ms * 1e3:.0f} samples/s "
              f"({B * V / ms * 1e3:.0f} img/s), host wall {(time.perf_counter() - t0) / n * 1e3:.2f} ms")
