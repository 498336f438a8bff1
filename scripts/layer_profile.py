"""Per-launch device times of one eager forward (CUDA events around every C-ABI launch, warm):
shape, ms, TFLOP/s, GB/s of algorithmic bytes, and the per-layer roofline bound. Development aid."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rot-mvgaze_b200"))
import torch
from rotmv_b200.module import FeatRotationSymm
from rotmv_b200 import functional as RF

B = int(os.environ.get("B", 256)); V = int(os.environ.get("V", 2)); chunk = int(os.environ.get("CHUNK", 512))
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650, "bf16_tflops_sustained": 1400}
torch.manual_seed(0)
model = FeatRotationSymm(50, 3, trunk_chunk=chunk).cuda().eval()
model.auto_graph = False   # per-launch events need the eager launches, not graph replays
images = torch.randn((B, V, 3, 224, 224), device="cuda")
rot = RF.pose_to_rotations(torch.rand((B, V, 2), device="cuda") - 0.5)
with torch.no_grad():
    for _ in range(2): model(images, rot)
    torch.cuda.synchronize()
    RF.PROFILE = []
    model(images, rot)
    torch.cuda.synchronize()
recs, RF.PROFILE = RF.PROFILE, None
agg = {}
tot = 0.0
for eng, flops, e0, e1, what, meta in recs:
    ms = e0.elapsed_time(e1); tot += ms
    key = meta.get("desc", what)
    a = agg.setdefault(key, [0, 0.0, 0.0, 0.0]); a[0] += 1; a[1] += ms; a[2] += flops; a[3] += meta.get("bytes", 0.0)
print(f"B={B} V={V} chunk={chunk}: {len(recs)} launches, sum of launch times {tot:.2f} ms")
print(f"{'kernel':52s} {'n':>3s} {'ms':>8s} {'%':>5s} {'TF/s':>7s} {'GB/s':>7s} {'bound_ms':>8s} {'eff':>5s}")
for key, (n, ms, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    bound = max(fl / (pk["bf16_tflops_sustained"] * 1e12), by / (pk["hbm_gbs"] * 1e9)) * 1e3
    print(f"{key:52s} {n:3d} {ms:8.3f} {100 * ms / tot:5.1f} {fl / ms / 1e9 if ms else 0:7.1f} {by / ms / 1e6 if ms else 0:7.0f} {bound:8.3f} {bound / ms if ms else 0:5.2f}")
