"""Device time of the graph-captured fusion stage (lifter + 3 x (rotate, fuser, head) + loss tail) and of
the trunk graph of one inference forward at configs[1] (B=256, V=2) -- per-launch CUDA events are
host-bound for these 5-25 us kernels in eager mode, graph replays are not. Development aid.
RMV_SPLITK=0/1 switches the split-K path of the small-M GEMMs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rot-mvgaze_b200"))
import torch
from rotmv_b200.engine import GraphedForward
from rotmv_b200.module import FeatRotationSymm

B = int(os.environ.get("B", 256)); V = int(os.environ.get("V", 2))
torch.manual_seed(0)
model = FeatRotationSymm(50, 3).cuda().eval()
sess = GraphedForward(model, B, V)
sess.images.normal_()
for name, g in (("fusion", sess.fusion_graph), ("trunk", sess.full_trunk_graph)):
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    n = 50
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"B={B} V={V} SPLITK={os.environ.get('RMV_SPLITK', '1')}: {name} graph {e0.elapsed_time(e1) / n * 1e3:.1f} us per replay")
