"""One tensor-bound layer (3x3, 256->256, 14x14, 512 images) on the single-CTA and the CTA-pair
kernel: device time of each (development aid / ncu target)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rot-mvgaze_b200"))
import torch
from rotmv_b200 import functional as RF, _lib as L
x = torch.randn((512, 14, 14, 256), device="cuda").bfloat16()
w = (torch.randn((256, 3, 3, 256), device="cuda") / 48).bfloat16()
y = torch.empty((512, 14, 14, 256), device="cuda", dtype=torch.bfloat16)
for mode in (0, 1):
    L.check(L.load().rmv_set_tuning(b"CTA2", mode), "tune")
    for _ in range(3): RF.conv2d(x, w, stride=1, pad=1, out=y, engine=L.ENGINE_TC)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): RF.conv2d(x, w, stride=1, pad=1, out=y, engine=L.ENGINE_TC)
    e1.record(); torch.cuda.synchronize()
    print(f"CTA2={mode}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per launch")
