"""Golden vectors for the TRAINING step of the non-default constructor flags from the
UNMODIFIED reference -- TEST INFRASTRUCTURE, run in the build container:

    python oracle/make_golden_variants_train.py   ->  tests/golden/rotmv_variants_train_r18.npz

For every flag set: same seed -> bit-identical init (asserted), then ONE train-mode forward + loss
(main.py:239-240) + backward of the imported reference and of the oracle on a seeded (B=6, V=2)
batch. Asserted bit-exact here: loss, every gradient, every buffer afterwards (BatchNorm running
statistics, IntensityBatchNorm running std after its four train-mode calls per iteration). Stored:
the reference's loss, gradient norms and, for share_feature, the IntensityBatchNorm buffers; the
tests rebuild the weights from the seed and check the oracle against them on any machine.
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
            "ignore_rotmat": dict(ignore_rotmat=True),
            "share_weights": dict(share_weights=True)}
B, V, DEPTH, ITERS, SEED = 6, 2, 18, 2, 2


def main():
    torch.set_num_threads(os.cpu_count())
    ns = ref_loader.load()
    metrics = ref_loader.make_loss(ns)
    images, pose, gt = O.synthetic_batch(B, V, seed=SEED)
    rotations = O.pairwise_rotations(pose)
    gold = {}
    for name, flags in VARIANTS.items():
        O.seed_all(0)
        ref = ns.FeatRotationSymm(backbone_depth=DEPTH, num_iter=ITERS, **flags)
        ora = O.build_model(num_iter=ITERS, depth=DEPTH, seed=0, **flags)
        rsd, osd = ref.state_dict(), ora.state_dict()
        assert list(rsd.keys()) == list(osd.keys()), name
        for k in rsd:
            assert torch.equal(rsd[k], osd[k]), (name, k)
        ref.train(); ora.train()
        data = {"img_0": images[:, 0].clone(), "img_1": images[:, 1].clone(),
                "rot_0": ns.rotation_matrix_2d(pose[:, 0]), "rot_1": ns.rotation_matrix_2d(pose[:, 1]),
                "gt_gaze": gt[:, 0].clone(), "gt_gaze_1": gt[:, 1].clone()}
        rloss = metrics(ref(data))
        rloss.backward()
        oloss = O.iteration_loss(ora.forward_views(images, rotations), [gt[:, 0], gt[:, 1]])
        oloss.backward()
        assert torch.equal(rloss.detach(), oloss.detach()), (name, rloss.item(), oloss.item())
        rg, og = dict(ref.named_parameters()), dict(ora.named_parameters())
        assert list(rg.keys()) == list(og.keys()), name
        norms = []
        for k in rg:
            if rg[k].grad is None:
                assert og[k].grad is None, (name, k)
                norms.append(-1.0)
                continue
            assert torch.equal(rg[k].grad, og[k].grad), f"oracle != reference: {name} grad {k} " \
                f"({(rg[k].grad - og[k].grad).abs().max().item():.3e})"
            norms.append(rg[k].grad.double().norm().item())
        rsd, osd = ref.state_dict(), ora.state_dict()
        for k in rsd:
            assert torch.equal(rsd[k], osd[k]), f"oracle != reference: {name} buffer {k} after the step"
        gold[f"{name}.loss"] = np.float32(rloss.item())
        gold[f"{name}.grad_norms"] = np.array(norms, dtype=np.float64)
        gold[f"{name}.bn1_running_mean"] = rsd["_feat_extractor.0.bn1.running_mean"].numpy().copy()
        for k in rsd:
            if k.endswith("_batchnorm.running_mean"):
                gold[f"{name}.{k}"] = rsd[k].numpy().copy()
        print(f"{name}: oracle == reference (bit-exact) for loss {rloss.item():.6f}, "
              f"{sum(n >= 0 for n in norms)} gradients, all buffers")
    path = os.path.join(ROOT, "tests", "golden", "rotmv_variants_train_r18.npz")
    np.savez_compressed(path, **gold)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
