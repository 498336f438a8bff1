"""Import the UNMODIFIED reference from /root/reference (build container only) -- TEST
INFRASTRUCTURE. Used by oracle/make_golden.py and the `-m "not gpu"` tests to pin the oracle.

Two non-invasive patches, neither touching a reference file:
  1. `models.rot_mv.resnet50/resnet18` are rebound to random-init constructors: the reference asks
     for ImageNet weights (models/rot_mv.py:120 -> models/resnet.py:281), which needs the network.
  2. empty stub modules for `h5py`, `albumentations`, `omegaconf`, imported but unused by
     utils/math.py:6,12,15.
`trainer.py` cannot be imported as shipped (trainer.py:25 imports a name utils/helper.py lacks), so
the step is the oracle's restatement of trainer.py:54,141-143 driving the imported model + losses.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ROTMV_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "rot_mv.py"))


def load():
    """Returns a namespace with FeatRotationSymm, StereoL1Loss, IterationLoss and the math fns."""
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for name in ("h5py", "albumentations", "omegaconf"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            if name == "omegaconf":
                m.OmegaConf = type("OmegaConf", (), {})
            sys.modules[name] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import models.resnet as ref_resnet
    import models.rot_mv as ref_rot_mv
    import losses.stereo_loss as ref_stereo
    import utils.math as ref_math

    ref_rot_mv.resnet50 = lambda pretrained=True, **kw: ref_resnet.resnet50(pretrained=False, **kw)
    ref_rot_mv.resnet18 = lambda pretrained=True, **kw: ref_resnet.resnet18(pretrained=False, **kw)
    ns = types.SimpleNamespace()
    ns.FeatRotationSymm = ref_rot_mv.FeatRotationSymm
    ns.StereoL1Loss = ref_stereo.StereoL1Loss
    ns.IterationLoss = ref_stereo.IterationLoss
    ns.rotation_matrix_2d = ref_math.rotation_matrix_2d
    ns.pitchyaw_to_vector = ref_math.pitchyaw_to_vector
    ns.angular_error = ref_math.angular_error
    import losses.gaze_loss as ref_gaze

    ns.gaze_angular_loss = ref_gaze.gaze_angular_loss
    return ns


def make_loss(ns):
    """main.py:239-240."""
    return ns.IterationLoss(
        loss=ns.StereoL1Loss(rel_weight=0.01, reference_decay=1.0,
                             distance_metric="angular_error", pred_gaze_key="pred_gaze"),
        iter_decay=0.5)
