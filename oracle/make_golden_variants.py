"""Golden vectors for the non-default constructor flags from the UNMODIFIED reference
-- TEST INFRASTRUCTURE, run in the build container:

    python oracle/make_golden_variants.py   ->  tests/golden/rotmv_variants_r18.npz

For every flag set: same seed -> bit-identical init between the imported reference and the oracle
(asserted here), BatchNorm running statistics calibrated on the batch, IntensityBatchNorm running
std set to a non-trivial value, then the reference's eval forward on a seeded (B=4, V=2) batch.
Stored: the reference's predictions / features; the tests rebuild the weights from the seed.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle import rotmv_oracle as O  # noqa: E402

VARIANTS = {"encode_rotmat": dict(encode_rotmat=True),
            "share_feature": dict(share_feature=True),
            "encode_rotmat_shared": dict(encode_rotmat=True, share_weights=True),
            "ignore_rotmat": dict(ignore_rotmat=True)}
B, V, DEPTH, ITERS = 4, 2, 18, 2


def prepare(model, images):
    """Deterministic, non-trivial buffers: calibrated BatchNorm + a perturbed IntensityBatchNorm."""
    O.calibrate_bn(model, images, passes=2)
    g = torch.Generator().manual_seed(11)
    for name, buf in model.named_buffers():
        if name.endswith("_batchnorm.running_mean"):
            buf.copy_(0.5 + torch.rand(buf.shape, generator=g))
    model.eval()


def main():
    torch.set_num_threads(os.cpu_count())
    ns = ref_loader.load()
    images, pose, _ = O.synthetic_batch(B, V, seed=5)
    gold = {}
    for name, flags in VARIANTS.items():
        O.seed_all(0)
        ref = ns.FeatRotationSymm(backbone_depth=DEPTH, num_iter=ITERS, **flags)
        ora = O.build_model(num_iter=ITERS, depth=DEPTH, seed=0, **flags)
        rsd, osd = ref.state_dict(), ora.state_dict()
        assert list(rsd.keys()) == list(osd.keys()), name
        for k in rsd:
            assert torch.equal(rsd[k], osd[k]), (name, k)
        prepare(ora, images)
        ref.load_state_dict(ora.state_dict(), strict=True)
        ref.eval()
        data = {"img_0": images[:, 0].clone(), "img_1": images[:, 1].clone(),
                "rot_0": ns.rotation_matrix_2d(pose[:, 0]), "rot_1": ns.rotation_matrix_2d(pose[:, 1])}
        with torch.no_grad():
            out_r = ref(dict(data))
            out_o = ora.forward_views(images, O.pairwise_rotations(pose))
        for i in range(ITERS):
            for key in ("pred_gaze_0", "pred_gaze_1", "feat_0", "feat_1"):
                a, b = out_r[f"iter_{i}"][key], out_o[f"iter_{i}"][key]
                assert torch.equal(a, b), f"oracle != reference: {name} iter_{i}.{key} " \
                    f"({(a - b).abs().max().item():.3e})"
                gold[f"{name}.iter_{i}.{key}"] = a.numpy()
        assert torch.equal(out_r["img_feat_0"], out_o["img_feat_0"])
        gold[f"{name}.img_feat_0"] = out_r["img_feat_0"].numpy()
        gold[f"{name}.pred_gaze"] = out_r["pred_gaze"].numpy()
        print(name, "oracle == reference (bit-exact); pred_gaze[0] =", out_r["pred_gaze"][0].tolist())
    path = os.path.join(ROOT, "tests", "golden", "rotmv_variants_r18.npz")
    np.savez_compressed(path, **gold)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
