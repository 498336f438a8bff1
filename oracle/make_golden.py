"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container, where
/root/reference exists) -- TEST INFRASTRUCTURE.

    python oracle/make_golden.py

Also asserts, at generation time, that the oracle restatement (oracle/rotmv_oracle.py) is
bit-identical to the imported reference: same random-init weights from the same seed, same forward
outputs, same loss, same gradients, same parameters after Adam steps. The fixtures are the
reference's outputs; the GPU box (which has no /root/reference) checks the oracle and the CUDA path
against them.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle import rotmv_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
B, V = 8, 2  # BASELINE.json configs[0]
STEP_LR = 1e-3  # CyclicLR max_lr (trainer.py:57); base_lr=1e-6 would move nothing visible in fp32


def tensor_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def assert_same(a, b, what):
    assert a.shape == b.shape and torch.equal(a, b), f"oracle != reference: {what} " \
        f"(max abs diff {(a.float() - b.float()).abs().max().item():.3e})"


def ref_data(ns, images, pose, gt):
    d = {"img_0": images[:, 0].clone(), "img_1": images[:, 1].clone(),
         "rot_0": ns.rotation_matrix_2d(pose[:, 0]), "rot_1": ns.rotation_matrix_2d(pose[:, 1]),
         "gt_gaze": gt[:, 0].clone(), "gt_gaze_1": gt[:, 1].clone()}
    return d


def main():
    torch.set_num_threads(os.cpu_count())
    ns = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)

    # ---- weights: same seed -> same init ------------------------------------------------------
    O.seed_all(0)
    ref = ns.FeatRotationSymm(backbone_depth=50, num_iter=3, share_weights=False,
                              encode_rotmat=False, share_feature=False, ignore_rotmat=False)
    ora = O.build_model(num_iter=3, depth=50, seed=0)
    rsd, osd = ref.state_dict(), ora.state_dict()
    assert list(rsd.keys()) == list(osd.keys()), "state_dict keys/order differ"
    for k in rsd:
        assert_same(rsd[k], osd[k], f"init {k}")
    init_state = {k: v.clone() for k, v in rsd.items()}
    meta = {"n_keys": len(rsd), "n_params": sum(p.numel() for p in ref.parameters()),
            "weights_sha256": tensor_digest(rsd)}
    print("weights identical:", meta)

    images, pose, gt = O.synthetic_batch(B, V, seed=1)
    rotations = O.pairwise_rotations(pose)
    gold = {}

    # ---- KATs of the math functions -----------------------------------------------
    kp = torch.tensor([[0.1, 0.2], [-0.3, 0.5]])
    gold["kat_pose"] = kp.numpy()
    gold["kat_rotmat"] = ns.rotation_matrix_2d(kp).numpy()
    assert_same(ns.rotation_matrix_2d(kp), O.rotation_matrix_2d(kp), "rotation_matrix_2d")
    gold["kat_vector"] = ns.pitchyaw_to_vector(kp).numpy()
    assert_same(ns.pitchyaw_to_vector(kp), O.pitchyaw_to_vector(kp), "pitchyaw_to_vector")
    kpred = torch.tensor([[0.12, 0.18], [-0.25, 0.55]], requires_grad=True)
    kl = ns.gaze_angular_loss(kpred, kp)
    kl.backward()
    gold["kat_pred"] = kpred.detach().numpy()
    gold["kat_loss"] = np.float32(kl.item())
    gold["kat_loss_grad"] = kpred.grad.numpy()
    kpred2 = kpred.detach().clone().requires_grad_(True)
    kl2 = O.angular_loss_deg(kpred2, kp)
    kl2.backward()
    assert_same(kl.detach(), kl2.detach(), "gaze_angular_loss")
    assert_same(kpred.grad, kpred2.grad, "gaze_angular_loss grad")
    gold["kat_angular_error"] = ns.angular_error(kpred.detach(), kp).numpy()
    assert_same(ns.angular_error(kpred.detach(), kp), O.angular_error_deg(kpred.detach(), kp), "angular_error")
    gold["rotations"] = rotations.numpy()
    r0, r1 = ns.rotation_matrix_2d(pose[:, 0]), ns.rotation_matrix_2d(pose[:, 1])
    assert_same(r0 @ r1.transpose(-1, -2), rotations[:, 0, 1], "pairwise rotations")

    # ---- eval forward, random init (BASELINE config 1) ----------------------------------------
    ref.eval(); ora.eval()
    with torch.no_grad():
        rd = ref(ref_data(ns, images, pose, gt))
        od = ora.forward_views(images, rotations)
    for v in range(V):
        assert_same(rd[f"img_feat_{v}"], od[f"img_feat_{v}"], f"eval img_feat_{v}")
        gold[f"eval_img_feat_{v}"] = rd[f"img_feat_{v}"].numpy()
        gold[f"eval_initial_rot_feat_{v}"] = rd[f"initial_rot_feat_{v}"].numpy()
        for i in range(3):
            assert_same(rd[f"iter_{i}"][f"pred_gaze_{v}"], od[f"iter_{i}"][f"pred_gaze_{v}"], f"eval pred {i} {v}")
            assert_same(rd[f"iter_{i}"][f"feat_{v}"], od[f"iter_{i}"][f"feat_{v}"], f"eval feat {i} {v}")
            gold[f"eval_iter{i}_pred_gaze_{v}"] = rd[f"iter_{i}"][f"pred_gaze_{v}"].numpy()
        gold[f"eval_iter2_feat_{v}"] = rd["iter_2"][f"feat_{v}"].numpy()
    gold["eval_pred_gaze"] = rd["pred_gaze"].numpy()
    # dict API of the oracle (two views) == reference dict API
    with torch.no_grad():
        od2 = ora(ref_data(ns, images, pose, gt))
    assert_same(rd["pred_gaze"], od2["pred_gaze"], "dict-API pred_gaze")
    print("eval forward identical; pred_gaze[0] =", rd["pred_gaze"][0].tolist())

    # ---- train-mode step x2 (trainer.py:119-123,141-143; Adam coupled L2) ----------------------
    metrics = ref_loader.make_loss(ns)
    ropt = torch.optim.Adam(ref.parameters(), lr=STEP_LR, weight_decay=1e-6)
    oopt = O.make_adam(ora, lr=STEP_LR)
    gold["step_lr"] = np.float32(STEP_LR)
    for step in range(2):
        ref.train()
        rdat = ref(ref_data(ns, images, pose, gt))
        rloss = metrics(rdat)
        ropt.zero_grad(); rloss.backward()
        ora.train()
        odat = ora.forward_views(images, rotations)
        oloss = O.iteration_loss(odat, [gt[:, 0], gt[:, 1]])
        oopt.zero_grad(); oloss.backward()
        assert_same(rloss.detach(), oloss.detach(), f"train loss step {step}")
        rg = dict(ref.named_parameters()); og = dict(ora.named_parameters())
        n_grad = 0
        for k in rg:
            if rg[k].grad is None:
                assert og[k].grad is None, k
                continue
            n_grad += 1
            assert_same(rg[k].grad, og[k].grad, f"grad {k} step {step}")
        if step == 0:
            gold["train_loss_0"] = np.float32(rloss.item())
            gold["n_params_with_grad"] = np.int64(n_grad)
            for v in range(V):
                gold[f"train_img_feat_{v}"] = rdat[f"img_feat_{v}"].detach().numpy()
                for i in range(3):
                    gold[f"train_iter{i}_pred_gaze_{v}"] = rdat[f"iter_{i}"][f"pred_gaze_{v}"].detach().numpy()
            for k in ["_feat_extractor.0.conv1.weight", "_feat_extractor.0.bn1.weight",
                      "_feat_extractor.0.layer1.0.conv1.weight", "_feat_extractor.0.layer4.2.bn3.bias",
                      "_feat_extractor.0.layer4.2.conv3.weight",
                      "_lifter._lifter.blocks.1.0.bias", "_img_fusers.2._fuser.blocks.1.0.bias",
                      "_gaze_estimators.2.blocks.1.0.weight", "_gaze_estimators.0.blocks.1.0.bias"]:
                g = rg[k].grad
                gold["grad0::" + k] = (g if g.numel() <= 70000 else g.flatten()[:4096]).detach().clone().numpy()
            gold["grad0_norms"] = np.array([rg[k].grad.norm().item() if rg[k].grad is not None else -1.0
                                            for k in rg], dtype=np.float64)
        else:
            gold["train_loss_1"] = np.float32(rloss.item())
        ropt.step(); oopt.step()
    rsd, osd = ref.state_dict(), ora.state_dict()
    for k in rsd:
        assert_same(rsd[k], osd[k], f"after 2 steps {k}")
    gold["after2_bn1_running_mean"] = rsd["_feat_extractor.0.bn1.running_mean"].detach().clone().numpy()
    gold["after2_bn1_running_var"] = rsd["_feat_extractor.0.bn1.running_var"].detach().clone().numpy()
    gold["after2_num_batches_tracked"] = rsd["_feat_extractor.0.bn1.num_batches_tracked"].detach().clone().numpy()
    gold["after2_head2_w"] = rsd["_gaze_estimators.2.blocks.1.0.weight"].detach().clone().numpy()
    gold["after2_bn1_weight"] = rsd["_feat_extractor.0.bn1.weight"].detach().clone().numpy()
    gold["after2_delta_norms"] = np.array(
        [(rsd[k].double() - init_state[k].double()).norm().item() for k in rsd], dtype=np.float64)
    meta["after2_sha256"] = tensor_digest(rsd)
    print("2 train steps identical; losses", gold["train_loss_0"], gold["train_loss_1"])

    # ---- BN-calibrated eval: O(1) activations -------------------------------
    ref.load_state_dict(init_state); ora.load_state_dict(init_state)
    O.calibrate_bn(ref, images); O.calibrate_bn(ora, images)
    with torch.no_grad():
        rd = ref(ref_data(ns, images, pose, gt))
        od = ora.forward_views(images, rotations)
    for v in range(V):
        assert_same(rd[f"img_feat_{v}"], od[f"img_feat_{v}"], f"calib img_feat_{v}")
        gold[f"calib_img_feat_{v}"] = rd[f"img_feat_{v}"].numpy()
        for i in range(3):
            assert_same(rd[f"iter_{i}"][f"pred_gaze_{v}"], od[f"iter_{i}"][f"pred_gaze_{v}"], "calib pred")
            gold[f"calib_iter{i}_pred_gaze_{v}"] = rd[f"iter_{i}"][f"pred_gaze_{v}"].numpy()
    meta["calib_sha256"] = tensor_digest(ref.state_dict())
    print("calibrated eval identical; pred_gaze[0] =", rd["pred_gaze"][0].tolist())

    # ---- V = 4: no reference semantics; composition of the REFERENCE's sub-modules --
    ref.load_state_dict(init_state); ora.load_state_dict(init_state)
    ref.eval(); ora.eval()
    im4, pose4, gt4 = O.synthetic_batch(4, 4, seed=2)
    rot4 = O.pairwise_rotations(pose4)
    with torch.no_grad():
        feats = [ref._feat_extractor(im4[:, v]) for v in range(4)]
        rf = [ref._lifter(f) for f in feats]
        for i in range(3):
            new = []
            for v in range(4):
                agg = sum(rot4[:, v, u] @ rf[u] for u in range(4) if u != v) / 3
                new.append(ref._img_fusers[i](feats[v], agg).reshape(-1, 3, 512))
            rf = new
        pred4 = [ref._gaze_estimators[2](torch.cat([feats[v], rf[v].flatten(1, -1)], -1)) for v in range(4)]
        od4 = ora.forward_views(im4, rot4)
    for v in range(4):
        assert_same(pred4[v], od4["iter_2"][f"pred_gaze_{v}"], f"V=4 pred view {v}")
        gold[f"v4_iter2_pred_gaze_{v}"] = pred4[v].numpy()
    print("V=4 composition identical")

    gold["meta"] = np.array(repr(meta))
    path = os.path.join(OUT, "rotmv_r50_b8v2.npz")
    np.savez(path, **gold)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", len(gold), "arrays")


if __name__ == "__main__":
    main()
