"""Callers immediately either side of the hot path, re-done for the GPU:

* `cyclic_lr`      -- the reference's `CyclicLR(base 1e-6, max 1e-3, triangular2,
                      cycle_momentum=False)` (trainer.py:56-62) as a pure function of the scheduler
                      step count; the reference steps it once per EPOCH (trainer.py:147, quirk Q2).
* `Evaluator`      -- `Trainer.test` (trainer.py:164-199) without the per-batch D2H copies: the
                      angular error is accumulated on the device (`rmv_angular_error_accum`) and
                      read back once.
* `save_checkpoint` / `load_checkpoint` -- same file content as the reference
                      (`torch.save(model.state_dict())`, trainer.py:150-160; strict load :45-48),
                      optionally with the optimizer state the reference forgets.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Optional

import torch

from . import functional as RF


def cyclic_lr(step: int, step_size_up: int, step_size_down: int, base_lr: float = 1e-6,
              max_lr: float = 1e-3, mode: str = "triangular2") -> float:
    """Learning rate after `step` scheduler steps (torch.optim.lr_scheduler.CyclicLR semantics:
    the constructor performs step 0, every `scheduler.step()` advances by one)."""
    total = float(step_size_up + step_size_down)
    ratio = step_size_up / total
    cycle = math.floor(1 + step / total)
    x = 1.0 + step / total - cycle
    scale = x / ratio if x <= ratio else (x - 1) / (ratio - 1)
    height = (max_lr - base_lr) * scale
    if mode == "triangular":
        factor = 1.0
    elif mode == "triangular2":
        factor = 1.0 / (2.0 ** (cycle - 1))
    else:
        raise ValueError(f"unsupported CyclicLR mode {mode!r}")
    return base_lr + height * factor


class Evaluator:
    """Mean angular error (degrees) of `pred_gaze` against `gt_gaze` over a loader of batches.

    Each batch is a dict with `images [B,V,3,H,W]`, `head_pose [B,V,2]` (or `rotations`) and
    `gt_gaze [B,2]` (view 0 label), on the host or the device. Mirrors trainer.py:164-199; the
    metric is utils/math.py:96-137 with the cosine clamped (the reference can return NaN).
    """

    def __init__(self, model, precision: Optional[str] = None):
        self.model = model
        self.precision = precision

    @torch.no_grad()
    def run(self, batches: Iterable[Dict[str, torch.Tensor]]) -> float:
        self.model.eval()
        dev = next(self.model.parameters()).device
        acc = torch.zeros((2,), device=dev, dtype=torch.float32)  # [sum of errors, count]
        for batch in batches:
            images = batch["images"].to(dev, non_blocking=True).float()
            if "rotations" in batch:
                rot = batch["rotations"].to(dev, non_blocking=True).float()
            else:
                rot = RF.pose_to_rotations(batch["head_pose"].to(dev, non_blocking=True).float().contiguous())
            pred = self.model(images, rot, precision=self.precision)
            gt = batch["gt_gaze"].to(dev, non_blocking=True).float().contiguous()
            RF.angular_error_accum(pred, gt, acc)
        s, n = acc.tolist()  # the only device->host read of the whole evaluation
        return s / max(n, 1.0)


def save_checkpoint(path: str, model, engine=None) -> None:
    """Reference format: the bare state_dict (trainer.py:150-160). With `engine`, the Adam moments,
    hyper-parameters and step count are stored next to it under '__optimizer__'."""
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    if engine is not None:
        sd["__optimizer__"] = {"exp_avg": engine.flat_m.cpu(), "exp_avg_sq": engine.flat_v.cpu(),
                               "hyper": engine.hyper.cpu(), "names": list(engine.names)}
    torch.save(sd, path)


def load_checkpoint(path: str, model, engine=None, strict: bool = True) -> None:
    sd = torch.load(path, map_location="cpu")
    opt = sd.pop("__optimizer__", None)
    model.load_state_dict(sd, strict=strict)
    if hasattr(model, "invalidate"):
        model.invalidate()
    if engine is not None and opt is not None:
        if list(opt["names"]) != list(engine.names):
            raise ValueError("optimizer state does not match this model's parameter list")
        engine.flat_m.copy_(opt["exp_avg"])
        engine.flat_v.copy_(opt["exp_avg_sq"])
        engine.hyper.copy_(opt["hyper"])
