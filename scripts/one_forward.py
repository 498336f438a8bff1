"""Three eager (graph-free) forwards of the configs[1] workload: the ncu target (every launch is a
plain stream launch, so -s/-c select the third, warm forward). Prints the device time of the third."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rot-mvgaze_b200"))
import torch
from rotmv_b200.module import FeatRotationSymm
from rotmv_b200 import functional as RF
from rotmv_b200 import _lib as L

B = int(os.environ.get("B", 256)); V = int(os.environ.get("V", 2))
torch.manual_seed(0)
model = FeatRotationSymm(50, 3, trunk_chunk=B * V).cuda().eval()
images = torch.randn((B, V, 3, 224, 224), device="cuda")
rot = RF.pose_to_rotations(torch.rand((B, V, 2), device="cuda") - 0.5)
eng = model.engine("bf16")
with torch.no_grad():
    for i in range(3):
        n0 = L.STATS["launches"]
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        if i == 2:
            torch.cuda.profiler.start()     # ncu --profile-from-start off: only the third (warm) pass
        e0.record()
        eng.run(images, rot, want_all=False)
        e1.record(); torch.cuda.synchronize()
        if i == 2:
            torch.cuda.profiler.stop()
        print(f"forward {i}: {L.STATS['launches'] - n0} launches, {e0.elapsed_time(e1):.3f} ms")
