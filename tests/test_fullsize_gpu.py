"""Parity at BASELINE.json's FULL sizes through size-independent properties (the CPU oracle cannot
run 512-2048 images in seconds, so the whole batch is checked by properties and a scattered subset
of its samples against the oracle):

* eval mode: every sample is independent of the rest of the batch (models/rot_mv.py:187-269 has no
  cross-sample coupling in eval), so scattered samples of the full batch must match the CPU oracle
  run on those samples alone (fp32 engine rtol 1e-4; bf16 engine within the stated angular delta);
* permuting the samples permutes the outputs; permuting the VIEWS (images and both view axes of
  `rotations`) permutes the per-view predictions (the fusion is symmetric in the views, :234-239);
* identical views with identity relative rotation give identical per-view predictions;
* training (configs[3], B=128, V=2): the train-mode loss and the BatchNorm running statistics of one
  step match the CPU oracle's train-mode forward (no backward needed for either), a step with
  lr = 0 leaves every parameter bit-identical (Adam, trainer.py:54), `num_batches_tracked` advances
  by V, `fc.*` is untouched (Q4), and the loss / head gradients do not depend on the
  order of the samples (batch statistics and the batch mean are symmetric).

configs[1]: B=256, V=2; configs[2]: B=512, V=4 (all 12 ordered pairs of rotations); configs[3]:
B=128, V=2 training step.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def close(a, b, rtol=1e-4):
    a, b = a.detach().float().cpu(), torch.as_tensor(b).float()
    atol = rtol * b.pow(2).mean().sqrt().item()
    return torch.allclose(a, b, rtol=rtol, atol=atol), ((a - b).abs().max().item(), atol)


def _ang(a, b):
    def vec(p):
        return torch.stack([torch.cos(p[:, 0]) * torch.sin(p[:, 1]), torch.sin(p[:, 0]),
                            torch.cos(p[:, 0]) * torch.cos(p[:, 1])], 1)
    s = (vec(a.double()) * vec(b.double())).sum(1).clamp(-1, 1)
    return torch.acos(s) * 180 / math.pi


def _device_batch(b, v, seed):
    """Synthetic inputs of SURVEY 8d generated on the device (host generation of 2048 images takes
    longer than the test)."""
    from rotmv_b200 import functional as RF

    g = torch.Generator(device="cuda").manual_seed(seed)
    images = torch.randn((b, v, 3, 224, 224), device="cuda", generator=g)
    pose = torch.rand((b, v, 2), device="cuda", generator=g) - 0.5
    gt = torch.rand((b, v, 2), device="cuda", generator=g) - 0.5
    return images, pose, RF.pose_to_rotations(pose), gt


@pytest.fixture(scope="module")
def calibrated():
    """BN-calibrated weights shared by the oracle and the CUDA module."""
    from oracle import rotmv_oracle as O
    from rotmv_b200.module import FeatRotationSymm

    ora = O.build_model(num_iter=3, depth=50, seed=0)
    images8, _, _ = O.synthetic_batch(8, 2, seed=1)
    O.calibrate_bn(ora, images8)
    model = FeatRotationSymm(50, 3)
    model.load_state_dict(ora.state_dict(), strict=True)
    model = model.cuda().eval()
    model.auto_graph = False
    return O, ora, model


def _oracle_on_subset(O, ora, images, pose, idx):
    """The CPU oracle on the selected samples alone (a handful of samples: well under a second)."""
    with torch.no_grad():
        return ora.forward_views(images[idx].cpu(), O.pairwise_rotations(pose[idx].cpu()))


def _dist(d):
    q = torch.quantile(d.double(), torch.tensor([0.5, 0.9, 0.99], dtype=torch.float64))
    return {"mean": d.mean().item(), "p50": q[0].item(), "p90": q[1].item(), "p99": q[2].item(),
            "max": d.max().item(), "n": d.numel()}


def _fmt(st):
    return (f"n={st['n']} mean {st['mean']:.3f} p50 {st['p50']:.3f} p90 {st['p90']:.3f} "
            f"p99 {st['p99']:.3f} max {st['max']:.3f} deg")


def test_bf16_delta_no_worse_than_reference_autocast(calibrated):
    """THE bf16 tolerance statement: the angular delta of the tcgen05 bf16 engine
    against the reference's fp32 predictions must be no worse than the delta of the REFERENCE'S OWN
    bf16 path (`torch.autocast("cpu", bfloat16)` over the unmodified models/rot_mv.py) on the same 64
    two-view samples and the same BN-calibrated weights. Both fp32 and autocast predictions of the
    reference are the committed fixture tests/golden/rotmv_r50_bf16_autocast_b64.npz
    (oracle/make_golden_bf16.py; reference distribution over its 384 predictions: mean 1.58, p50
    1.47, p90 2.68, p99 3.88, max 4.35 deg -- the 0.46-2.98 deg of BASELINE.md section 2 is the same
    statistic on 8 samples). Bounds: mean, p50, p90 <= 1.10x the reference's, p99 <= 1.15x, max <=
    1.25x (the max of 384 draws of a chaotic random-init network is the noisiest statistic)."""
    import os

    import numpy as np

    O, ora, model = calibrated
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                             "rotmv_r50_bf16_autocast_b64.npz"))
    p32, p16 = torch.tensor(g["pred_fp32"]), torch.tensor(g["pred_bf16_autocast"])   # [3, 2, 64, 2]
    b = int(g["batch"])
    images, pose, _ = O.synthetic_batch(b, 2, seed=int(g["seed"]))
    rot = O.pairwise_rotations(pose)
    with torch.no_grad():
        o32 = model.forward_views(images.cuda(), rot.cuda(), precision="fp32", want_all=False)
        o16 = model.forward_views(images.cuda(), rot.cuda(), precision="bf16", want_all=True)
    ours16 = torch.stack([torch.stack([o16[f"iter_{i}"][f"pred_gaze_{v}"].cpu() for v in range(2)]) for i in range(3)])
    # the fixture's fp32 predictions are the same function the fp32 engine computes (rtol 1e-4)
    ok, info = close(o32["iter_2"]["pred_gaze_0"], p32[2, 0])
    assert ok, info
    ref = _dist(_ang(p16.reshape(-1, 2), p32.reshape(-1, 2)))
    ours = _dist(_ang(ours16.reshape(-1, 2), p32.reshape(-1, 2)))
    print(f"bf16 angular delta vs reference fp32, 64 samples:\n  reference autocast: {_fmt(ref)}\n"
          f"  tcgen05 bf16 engine: {_fmt(ours)}")
    for key, slack in (("mean", 1.10), ("p50", 1.10), ("p90", 1.10), ("p99", 1.15), ("max", 1.25)):
        assert ours[key] <= slack * ref[key], (key, ours[key], ref[key])


def test_config1_scattered_samples_match_oracle(calibrated):
    """configs[1] (B=256, V=2): 8 samples scattered over the batch (first/last rows of tiles and of
    the batch) against the CPU oracle on those samples alone."""
    O, ora, model = calibrated
    images, pose, rot, _ = _device_batch(256, 2, seed=7)
    idx = torch.tensor([0, 31, 77, 127, 128, 200, 254, 255], device="cuda")
    ref = _oracle_on_subset(O, ora, images, pose, idx)
    with torch.no_grad():
        out32 = model.forward_views(images, rot, precision="fp32")
        out16 = model.forward_views(images, rot, precision="bf16")
    deltas = []
    for v in range(2):
        ok, info = close(out32[f"img_feat_{v}"][idx], ref[f"img_feat_{v}"])
        assert ok, ("img_feat", v, info)
        for i in range(3):
            ok, info = close(out32[f"iter_{i}"][f"pred_gaze_{v}"][idx], ref[f"iter_{i}"][f"pred_gaze_{v}"])
            assert ok, ("pred", i, v, info)
            deltas.append(_ang(out16[f"iter_{i}"][f"pred_gaze_{v}"][idx].cpu(),
                               ref[f"iter_{i}"][f"pred_gaze_{v}"]))
    d = torch.cat(deltas)
    print(f"full-size bf16 angular delta vs the oracle on 8 scattered samples: {_fmt(_dist(d))}")
    # whole batch: the bf16 engine against the fp32 engine (which the lines above pin to the oracle at
    # rtol 1e-4) over all 256 samples x 2 views x 3 iterations = 1536 predictions. Bounds = the
    # reference's own bf16-autocast distribution (test_bf16_delta_no_worse_than_reference_autocast:
    # mean 1.58, p99 3.88, max 4.35 deg over 384 predictions) with the same slack; the max of a
    # 4x larger draw gets the 6.4 deg of the uncalibrated regime only as a sanity ceiling.
    full = torch.cat([_ang(out16[f"iter_{i}"][f"pred_gaze_{v}"].cpu(), out32[f"iter_{i}"][f"pred_gaze_{v}"].cpu())
                      for v in range(2) for i in range(3)])
    st = _dist(full)
    print(f"full-size bf16 angular delta vs the fp32 engine, all 256 samples: {_fmt(st)}")
    assert d.max().item() <= 3.0, d.max().item()      # the scattered subset (48 values)
    assert st["mean"] <= 1.10 * 1.58 and st["p99"] <= 1.15 * 3.88 and st["max"] <= 6.4, st
    assert torch.isfinite(out16["pred_gaze"]).all() and tuple(out16["pred_gaze"].shape) == (256, 2)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_config1_sample_and_view_permutations(calibrated, precision):
    """configs[1]: outputs follow a permutation of the samples and a swap of the two views; equal
    views with identity relative rotations give equal per-view predictions. Every output row is
    computed from its own rows only and in the same accumulation order wherever the row sits in a
    tile, so the bound is rounding-level (1e-5 rad), not the bf16 tolerance."""
    _, _, model = calibrated
    images, _, rot, _ = _device_batch(256, 2, seed=8)
    with torch.no_grad():
        base = model.forward_views(images, rot, precision=precision, want_all=False)["iter_2"]
        p0, p1 = base["pred_gaze_0"].clone(), base["pred_gaze_1"].clone()
        perm = torch.randperm(256, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
        o = model.forward_views(images[perm].contiguous(), rot[perm].contiguous(), precision=precision,
                                want_all=False)["iter_2"]
        assert (o["pred_gaze_0"] - p0[perm]).abs().max().item() <= 1e-5
        assert (o["pred_gaze_1"] - p1[perm]).abs().max().item() <= 1e-5
        sw = [1, 0]
        o = model.forward_views(images[:, sw].contiguous(), rot[:, sw][:, :, sw].contiguous(),
                                precision=precision, want_all=False)["iter_2"]
        assert (o["pred_gaze_0"] - p1).abs().max().item() <= 1e-5
        assert (o["pred_gaze_1"] - p0).abs().max().item() <= 1e-5
        same = images[:, :1].expand(-1, 2, -1, -1, -1).contiguous()
        eye = torch.eye(3, device="cuda").expand(256, 2, 2, 3, 3).contiguous()
        o = model.forward_views(same, eye, precision=precision, want_all=False)["iter_2"]
        assert (o["pred_gaze_0"] - o["pred_gaze_1"]).abs().max().item() <= 1e-5


def test_config2_four_views_full_size(calibrated):
    """configs[2] (B=512, V=4 = 2048 images, trunk in 512-image chunks): scattered samples against
    the oracle's V=4 composition of the reference sub-modules, and equivariance under a
    permutation of the four views."""
    O, ora, model = calibrated
    images, pose, rot, _ = _device_batch(512, 4, seed=9)
    idx = torch.tensor([0, 127, 128, 300, 511], device="cuda")
    ref = _oracle_on_subset(O, ora, images, pose, idx)
    with torch.no_grad():
        out = model.forward_views(images, rot, precision="bf16", want_all=False)["iter_2"]
        preds = [out[f"pred_gaze_{k}"].clone() for k in range(4)]
    deltas = [_ang(preds[k][idx].cpu(), ref["iter_2"][f"pred_gaze_{k}"]) for k in range(4)]
    d = torch.cat(deltas)
    print(f"V=4 full-size bf16 angular delta (5 scattered samples): {_fmt(_dist(d))}")
    # same statement as test_bf16_delta_no_worse_than_reference_autocast (reference autocast: mean 1.58,
    # max 4.35 deg); V=4 has no reference bf16 path of its own
    assert d.mean().item() <= 1.10 * 1.58 and d.max().item() <= 1.25 * 4.35, (d.mean().item(), d.max().item())
    sigma = [2, 0, 3, 1]
    with torch.no_grad():
        o = model.forward_views(images[:, sigma].contiguous(), rot[:, sigma][:, :, sigma].contiguous(),
                                precision="bf16", want_all=False)["iter_2"]
    for k in range(4):
        # the partner mean of SURVEY D1 adds the three partners in view order: a permutation changes
        # the order of that fp32 sum (bf16-rounded afterwards), so this bound is the bf16 one
        dk = _ang(o[f"pred_gaze_{k}"].cpu(), preds[sigma[k]].cpu())
        assert dk.mean().item() <= 1.0 and dk.max().item() <= 6.4, (k, dk.mean().item(), dk.max().item())


def test_config2_fp32_subset_of_views(calibrated):
    """V=4, fp32 engine at a batch the FFMA engine runs quickly (B=64): rtol 1e-4 vs the oracle on
    scattered samples."""
    O, ora, model = calibrated
    images, pose, rot, _ = _device_batch(64, 4, seed=10)
    idx = torch.tensor([0, 17, 63], device="cuda")
    ref = _oracle_on_subset(O, ora, images, pose, idx)
    with torch.no_grad():
        out = model.forward_views(images, rot, precision="fp32", want_all=False)["iter_2"]
    for k in range(4):
        ok, info = close(out[f"pred_gaze_{k}"][idx], ref["iter_2"][f"pred_gaze_{k}"])
        assert ok, (k, info)


# -------------------------------------------------------------------------------------------------
# configs[3]: training step at B=128, V=2
# -------------------------------------------------------------------------------------------------
def _train_setup(precision, lr):
    from oracle import rotmv_oracle as O
    from rotmv_b200.module import FeatRotationSymm
    from rotmv_b200.train import TrainEngine

    ora = O.build_model(num_iter=3, depth=50, seed=0)
    model = FeatRotationSymm(50, 3)
    model.load_state_dict(ora.state_dict(), strict=True)
    model = model.cuda().train()
    eng = TrainEngine(model, precision=precision, lr=lr, weight_decay=1e-6)
    return O, ora, model, eng


@pytest.fixture(scope="module")
def train_ref():
    """CPU oracle, train mode, forward only at B=128 (a few seconds): loss + running statistics."""
    from oracle import rotmv_oracle as O

    ora = O.build_model(num_iter=3, depth=50, seed=0).train()
    images, pose, rot, gt = _device_batch(128, 2, seed=11)
    with torch.no_grad():
        out = ora.forward_views(images.cpu(), O.pairwise_rotations(pose.cpu()))
        loss = O.iteration_loss(out, [gt[:, 0].cpu(), gt[:, 1].cpu()]).item()
    stats = {k: v.clone() for k, v in ora.state_dict().items() if "running_" in k or "num_batches" in k}
    return dict(images=images, rot=rot, gt=gt, loss=loss, stats=stats)


def test_config3_fp32_step_loss_and_bn_state_match_oracle(train_ref):
    O, ora, model, eng = _train_setup("fp32", lr=0.0)
    before = eng.flat_p.clone()
    fc_before = model._feat_extractor[0].fc.weight.detach().clone()
    loss = eng.step(train_ref["images"], train_ref["rot"], train_ref["gt"]).item()
    assert abs(loss - train_ref["loss"]) <= 1e-4 * abs(train_ref["loss"]), (loss, train_ref["loss"])
    sd = model.state_dict()
    worst = 0.0
    for k, want in train_ref["stats"].items():
        got = sd[k].cpu()
        if "num_batches" in k:
            assert int(got) == int(want) == 2, k           # +V per step
            continue
        err = ((got - want).abs().max() / want.abs().max().clamp_min(1e-6)).item()
        worst = max(worst, err)
        assert err <= 1e-3, (k, err)
    print(f"full-size fp32 train step: loss {loss:.6f} (oracle {train_ref['loss']:.6f}), worst BN stat rel err {worst:.2e}")
    # lr = 0: Adam moves nothing (bit-exact), and fc.* never takes part (Q4)
    assert torch.equal(eng.flat_p, before)
    assert torch.equal(model._feat_extractor[0].fc.weight.detach(), fc_before)
    assert torch.isfinite(eng.flat_g).all() and eng.flat_g.abs().sum().item() > 0


def test_config3_bf16_step_close_and_order_independent(train_ref):
    """bf16 engine at B=128: loss within 2 % of the fp32 oracle (the stated bf16 training tolerance,
    tests/test_train_gpu.py), and the loss and the last head's gradient do not depend on the order
    of the samples beyond rounding noise."""
    O, ora, model, eng = _train_setup("bf16", lr=0.0)
    images, rot, gt = train_ref["images"], train_ref["rot"], train_ref["gt"]
    loss = eng.step(images, rot, gt).item()
    assert abs(loss - train_ref["loss"]) <= 2e-2 * abs(train_ref["loss"]), (loss, train_ref["loss"])
    head = model._gaze_estimators[2].blocks[1][0].weight
    g_head = eng.grads[id(head)].clone()
    lift = model._lifter._lifter.blocks[1][0].bias
    g_lift = eng.grads[id(lift)].clone()
    perm = torch.randperm(128, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    loss_p = eng.step(images[perm].contiguous(), rot[perm].contiguous(), gt[perm].contiguous()).item()
    cos = torch.nn.functional.cosine_similarity
    c_head = cos(eng.grads[id(head)].flatten(), g_head.flatten(), dim=0).item()
    c_lift = cos(eng.grads[id(lift)].flatten(), g_lift.flatten(), dim=0).item()
    print(f"full-size bf16 train step: loss {loss:.6f} / permuted {loss_p:.6f} (oracle "
          f"{train_ref['loss']:.6f}); grad cosine head {c_head:.5f} lifter {c_lift:.5f}")
    # bf16 rounding noise through 53 layers: measured 1.0e-3 (materialised BatchNorm) / 1.5e-3
    # (recomputed BatchNorm) relative under a permutation, against 3e-4 / 8e-4 to the fp32 oracle
    assert abs(loss_p - loss) <= 4e-3 * abs(loss), (loss, loss_p)
    assert c_head >= 0.99 and c_lift >= 0.9, (c_head, c_lift)
