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
