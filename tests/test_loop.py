"""Scheduler / evaluator / checkpoint helpers around the hot path."""
import os

import numpy as np
import pytest
import torch


def test_cyclic_lr_matches_torch_scheduler():
    """trainer.py:56-62: CyclicLR(base 1e-6, max 1e-3, triangular2, cycle_momentum=False)."""
    from rotmv_b200.loop import cyclic_lr

    for up, down in ((5, 7), (40, 41), (1, 1)):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.Adam([p], lr=0, weight_decay=1e-6)
        sched = torch.optim.lr_scheduler.CyclicLR(opt, base_lr=1e-6, max_lr=1e-3, step_size_up=up,
                                                  step_size_down=down, mode="triangular2",
                                                  cycle_momentum=False)
        for step in range(3 * (up + down) + 2):
            assert abs(opt.param_groups[0]["lr"] - cyclic_lr(step, up, down)) <= 1e-12, (up, down, step)
            opt.step()
            sched.step()


@pytest.mark.gpu
def test_evaluator_matches_reference_metric(tmp_path):
    from oracle import rotmv_oracle as O
    from rotmv_b200.loop import Evaluator, load_checkpoint, save_checkpoint
    from rotmv_b200.module import FeatRotationSymm

    ora = O.build_model(num_iter=2, depth=18, seed=0)
    images, pose, gt = O.synthetic_batch(6, 2, seed=4)
    O.calibrate_bn(ora, images, passes=2)
    model = FeatRotationSymm(18, 2, precision="fp32")
    model.load_state_dict(ora.state_dict(), strict=True)
    model = model.cuda()
    batches = [{"images": images[:4], "head_pose": pose[:4], "gt_gaze": gt[:4, 0]},
               {"images": images[4:], "head_pose": pose[4:], "gt_gaze": gt[4:, 0]}]
    err = Evaluator(model).run(batches)
    with torch.no_grad():
        ref_pred = ora.forward_views(images, O.pairwise_rotations(pose))["pred_gaze"]
    ref = O.angular_error_deg(ref_pred, gt[:, 0]).mean().item()   # trainer.py:192 np.mean(angular_error)
    assert abs(err - ref) <= 1e-3 * max(ref, 1.0), (err, ref)
    # checkpoint round trip in the reference's format (bare state_dict, strict load)
    path = os.path.join(tmp_path, "epoch_10_error=1.23.pth.tar")
    save_checkpoint(path, model)
    sd = torch.load(path)
    assert list(sd.keys()) == list(ora.state_dict().keys())
    other = FeatRotationSymm(18, 2, precision="fp32").cuda()
    load_checkpoint(path, other)
    assert abs(Evaluator(other).run(batches) - err) <= 1e-6
    ora2 = O.build_model(num_iter=2, depth=18, seed=123)
    ora2.load_state_dict(sd, strict=True)   # and the reference-side module loads our file


@pytest.mark.gpu
def test_trainer_loop_schedule_eval_checkpoint(tmp_path):
    """loop.Trainer: the reference's epoch loop (trainer.py:84-96,116-147): CyclicLR stepped once per
    epoch, graph-captured steps from host batches in the reference loader's dict format, evaluation
    and reference-format checkpoints."""
    import glob
    from rotmv_b200.loop import Trainer, cyclic_lr, load_checkpoint
    from rotmv_b200.module import FeatRotationSymm

    torch.manual_seed(0)
    model = FeatRotationSymm(50, 2).cuda()
    B, steps = 4, 3
    g = torch.Generator().manual_seed(1)

    def batches():
        for i in range(steps):
            b = B if i < steps - 1 else B - 1      # a short last batch, as a loader with drop_last=False gives
            yield {"img_0": torch.randn((b, 3, 224, 224), generator=g), "img_1": torch.randn((b, 3, 224, 224), generator=g),
                   "head_pose_0": torch.rand((b, 2), generator=g) - 0.5, "head_pose_1": torch.rand((b, 2), generator=g) - 0.5,
                   "gt_gaze": torch.rand((b, 2), generator=g) - 0.5, "gt_gaze_1": torch.rand((b, 2), generator=g) - 0.5}

    tr = Trainer(model, steps_per_epoch=steps, batch=B, views=2)
    assert abs(float(tr.engine.hyper[0]) - 1e-6) < 1e-12               # CyclicLR step 0 = base_lr
    p_before = tr.engine.flat_p.clone()
    hist = tr.fit(batches, batches, epochs=2, save_epoch=2, ckpt_dir=str(tmp_path))
    assert len(hist) == 3 and all(np.isfinite(h[2]) for h in hist) and all(np.isfinite(h[1]) for h in hist[1:])
    assert tr.train_iter == 2 * steps and tr.sched_steps == 2
    want_lr = cyclic_lr(2, max(steps // 2, 1), steps - steps // 2)
    assert abs(float(tr.engine.hyper[0]) - want_lr) <= 1e-12            # scheduler stepped once per EPOCH
    assert not torch.equal(tr.engine.flat_p, p_before)                  # the steps moved the weights
    files = glob.glob(os.path.join(tmp_path, "epoch_02_error=*.pth.tar"))
    assert len(files) == 1
    # the file Trainer.fit wrote is a BARE state_dict: the reference's strict load accepts it
    # (trainer.py:45-48); the optimizer state lives in the sidecar
    from oracle import rotmv_oracle as O
    sd = torch.load(files[0])
    assert "__optimizer__" not in sd
    O.build_model(num_iter=2, depth=50, seed=5).load_state_dict(sd, strict=True)
    assert os.path.exists(files[0] + ".optim.pt")
    other = FeatRotationSymm(50, 2).cuda()
    load_checkpoint(files[0], other)
    for (k, a), (_, b) in zip(model.state_dict().items(), other.state_dict().items()):
        assert torch.equal(a.cpu(), b.cpu()), k


def test_checkpoint_with_engine_is_bare_state_dict(tmp_path):
    """a checkpoint saved WITH optimizer state must still load into the reference with
    strict=True (trainer.py:45-48). CPU: a stub engine stands in for TrainEngine."""
    import types
    from oracle import rotmv_oracle as O
    from rotmv_b200.loop import load_checkpoint, save_checkpoint
    from rotmv_b200.module import FeatRotationSymm

    model = FeatRotationSymm(18, 2)
    eng = types.SimpleNamespace(flat_m=torch.ones(8), flat_v=torch.full((8,), 2.0),
                                hyper=torch.tensor([1e-3, 0.9, 0.999, 1e-8, 1e-6, 7.0], dtype=torch.float64),
                                names=["a", "b"], sync_replicas=lambda: None)
    path = os.path.join(tmp_path, "epoch_01_error=9.99.pth.tar")
    save_checkpoint(path, model, eng)
    sd = torch.load(path)
    assert all(isinstance(v, torch.Tensor) for v in sd.values())
    assert list(sd.keys()) == list(model.state_dict().keys())
    O.build_model(num_iter=2, depth=18, seed=1).load_state_dict(sd, strict=True)
    from oracle import ref_loader
    if ref_loader.available():          # the unmodified reference module, when present (build container)
        ref = ref_loader.load().FeatRotationSymm(backbone_depth=18, num_iter=2)
        ref.load_state_dict(torch.load(path), strict=True)   # trainer.py:45-48
    eng2 = types.SimpleNamespace(flat_m=torch.zeros(8), flat_v=torch.zeros(8),
                                 hyper=torch.zeros(6, dtype=torch.float64), names=["a", "b"],
                                 sync_replicas=lambda: None)
    other = FeatRotationSymm(18, 2)
    load_checkpoint(path, other, eng2)
    assert torch.equal(eng2.flat_m, eng.flat_m) and torch.equal(eng2.flat_v, eng.flat_v)
    assert float(eng2.hyper[5]) == 7.0
    for (k, a), (_, b) in zip(model.state_dict().items(), other.state_dict().items()):
        assert torch.equal(a, b), k
