"""Device-side timing + per-launch breakdown of one training step (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rot-mvgaze_b200"))
import torch
from rotmv_b200.module import FeatRotationSymm
from rotmv_b200.train import TrainEngine
from rotmv_b200 import functional as RF

B = int(os.environ.get("B", 128)); V = int(os.environ.get("V", 2)); prec = os.environ.get("PREC", "bf16")
torch.manual_seed(0)
model = FeatRotationSymm(50, 3).cuda().train()
eng = TrainEngine(model, precision=prec, lr=1e-4)
images = torch.randn((B, V, 3, 224, 224), device="cuda")
rot = RF.pose_to_rotations(torch.rand((B, V, 2), device="cuda") - 0.5)
gt = torch.rand((B, V, 2), device="cuda") - 0.5
for _ in range(2): eng.step(images, rot, gt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
n = 3
for _ in range(n): loss = eng.step(images, rot, gt)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"{prec} train step B={B} V={V}: {ms:.2f} ms -> {B / ms * 1e3:.0f} samples/s; loss {loss.item():.4f}; "
      f"{eng.launches_last_step} launches; mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
RF.PROFILE = []
eng.step(images, rot, gt); torch.cuda.synchronize()
recs, RF.PROFILE = RF.PROFILE, None
agg = {}; tot = 0
for eng_name, flops, a, b, what, meta in recs:
    t = a.elapsed_time(b); tot += t
    key = what if what != "rmv_conv2d_fwd" else "rmv_conv2d_fwd[" + eng_name + "]"
    if os.environ.get("DETAIL"): key = meta.get("desc", key)
    x = agg.setdefault(key, [0, 0.0, 0.0]); x[0] += 1; x[1] += t; x[2] += flops
print(f"sum of launch times {tot:.2f} ms over {len(recs)} launches")
for k, (c, t, f) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:56s} {c:4d} {t:9.3f} ms {100 * t / tot:5.1f}%  {f / t / 1e9 if t else 0:8.1f} TF/s")
