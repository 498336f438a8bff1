"""Where does the end-to-end (host buffers in, prediction out) inference time go? Times the pure
pinned H2D copy, each trunk-slice graph alone, the fusion graph, and run_host for several slicings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rot-mvgaze_b200"))
import torch
from rotmv_b200.module import FeatRotationSymm
from rotmv_b200.engine import GraphedForward
from rotmv_b200 import functional as RF

B, V = 256, 2
torch.manual_seed(0)
model = FeatRotationSymm(50, 3, trunk_chunk=int(os.environ.get("CHUNK", 512))).cuda().eval()
images_host = torch.randn((B, V, 3, 224, 224)).pin_memory()
rot = RF.pose_to_rotations(torch.rand((B, V, 2), device="cuda") - 0.5)
rot_host = rot.cpu().pin_memory()
dev = torch.empty_like(images_host, device="cuda")

def ev_time(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

ms = ev_time(lambda: dev.copy_(images_host, non_blocking=True))
print(f"H2D {images_host.numel() * 4 / 1e6:.0f} MB pinned: {ms:.3f} ms -> {images_host.numel() * 4 / ms / 1e6:.1f} GB/s")
for fr in ([1.0], [0.0625, 0.25, 0.5625, 1.0], [0.03125, 0.125, 0.28125, 0.5, 0.75, 1.0],
           [0.0625, 0.1875, 0.375, 0.625, 1.0], [0.125, 0.25, 0.375, 0.5, 0.625, 0.75, 0.875, 1.0]):
    sess = GraphedForward(model, B, V, slice_fracs=fr)
    sess.images.copy_(images_host); sess.rotations.copy_(rot)
    parts = [ev_time(g.replay) for g in sess.trunk_graphs]
    fus = ev_time(sess.fusion_graph.replay)
    full = ev_time(lambda: sess())
    for _ in range(3): sess.run_host(images_host, rot_host)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); n = 20
    for _ in range(n): sess.run_host(images_host, rot_host)
    torch.cuda.synchronize()
    e2e = (time.perf_counter() - t0) / n * 1e3
    print(f"slices {sess.slices}: trunk graphs {['%.2f' % p for p in parts]} (sum {sum(parts):.2f}) fusion {fus:.2f} "
          f"device-resident {full:.2f} ms; e2e {e2e:.2f} ms")
