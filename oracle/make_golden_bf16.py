"""Generate tests/golden/rotmv_r50_bf16_autocast_b64.npz from the UNMODIFIED reference (build
container only) -- TEST INFRASTRUCTURE.

    python oracle/make_golden_bf16.py

The stated bf16 tolerance of the tcgen05 engine is "no worse than the reference's OWN bf16 path on
the same inputs" (BASELINE.md section 4.6). This script pins that statement at a sample size that
gives a distribution instead of a handful of values: the imported reference (models/rot_mv.py,
random-init seed 0, BatchNorm running statistics calibrated with the recipe of
rotmv_oracle.calibrate_bn) is run on 64 seeded two-view samples in fp32 and under
`torch.autocast("cpu", dtype=torch.bfloat16)`; the fixture holds both sets of predictions
(3 iterations x 2 views x 64 samples x (pitch, yaw)). tests/test_fullsize_gpu.py runs the CUDA bf16
engine on the same inputs and weights and compares the two angular-delta distributions.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle import rotmv_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "rotmv_r50_bf16_autocast_b64.npz")
B, V, SEED = 64, 2, 21


def ang(a, b):
    def vec(p):
        return torch.stack([torch.cos(p[:, 0]) * torch.sin(p[:, 1]), torch.sin(p[:, 0]),
                            torch.cos(p[:, 0]) * torch.cos(p[:, 1])], 1)
    s = (vec(a.double()) * vec(b.double())).sum(1).clamp(-1, 1)
    return torch.acos(s) * 180 / math.pi


def main():
    torch.set_num_threads(os.cpu_count())
    ns = ref_loader.load()
    # weights: oracle seed 0 == reference seed 0 (asserted bit-exact by oracle/make_golden.py);
    # calibration through the oracle's trunk, then the state goes into the imported reference
    ora = O.build_model(num_iter=3, depth=50, seed=0)
    images8, _, _ = O.synthetic_batch(8, 2, seed=1)
    O.calibrate_bn(ora, images8)
    O.seed_all(0)
    ref = ns.FeatRotationSymm(backbone_depth=50, num_iter=3)
    ref.load_state_dict(ora.state_dict(), strict=True)
    ref.eval()
    images, pose, _ = O.synthetic_batch(B, V, seed=SEED)

    def run(autocast: bool):
        preds = torch.empty((3, V, B, 2))
        for s in range(0, B, 8):   # batches of 8: eval mode, so slicing changes nothing
            d = {"img_0": images[s:s + 8, 0].clone(), "img_1": images[s:s + 8, 1].clone(),
                 "rot_0": ns.rotation_matrix_2d(pose[s:s + 8, 0]), "rot_1": ns.rotation_matrix_2d(pose[s:s + 8, 1])}
            with torch.no_grad():
                if autocast:
                    with torch.autocast("cpu", dtype=torch.bfloat16):
                        out = ref(d)
                else:
                    out = ref(d)
            for i in range(3):
                for v in range(V):
                    preds[i, v, s:s + 8] = out[f"iter_{i}"][f"pred_gaze_{v}"].float()
        return preds

    p32 = run(False)
    # the oracle restatement must agree with the reference on these inputs as well
    with torch.no_grad():
        o = ora.forward_views(images[:8], O.pairwise_rotations(pose[:8]))
    assert torch.equal(o["iter_2"]["pred_gaze_0"], p32[2, 0, :8]), "oracle != reference"
    p16 = run(True)
    d = ang(p16.reshape(-1, 2), p32.reshape(-1, 2))
    q = torch.quantile(d, torch.tensor([0.5, 0.9, 0.99], dtype=torch.float64))
    print(f"reference bf16-autocast vs fp32, {d.numel()} predictions: mean {d.mean():.3f} "
          f"p50 {q[0]:.3f} p90 {q[1]:.3f} p99 {q[2]:.3f} max {d.max():.3f} deg")
    np.savez_compressed(OUT, pred_fp32=p32.numpy(), pred_bf16_autocast=p16.numpy(),
                        seed=np.int64(SEED), batch=np.int64(B))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
