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
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(acc)   # every rank evaluated its own shard: combine [sum, count]
        s, n = acc.tolist()  # the only device->host read of the whole evaluation
        return s / max(n, 1.0)


def _optim_path(path: str) -> str:
    return path + ".optim.pt"


def _is_rank0() -> bool:
    import torch.distributed as dist

    return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0


def save_checkpoint(path: str, model, engine=None) -> None:
    """Reference format: `path` holds the BARE state_dict (trainer.py:150-160), so the reference's
    `model.load_state_dict(torch.load(path), strict=True)` (trainer.py:45-48) accepts it. With
    `engine`, the Adam moments, hyper-parameters and step count -- which the reference forgets -- go
    to the sidecar file `<path>.optim.pt`. Under torch.distributed only rank 0 writes."""
    if not _is_rank0():
        return
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    torch.save(sd, path)
    if engine is not None:
        torch.save({"exp_avg": engine.flat_m.cpu(), "exp_avg_sq": engine.flat_v.cpu(),
                    "hyper": engine.hyper.cpu(), "names": list(engine.names)}, _optim_path(path))


def load_checkpoint(path: str, model, engine=None, strict: bool = True) -> None:
    """Loads a reference-format checkpoint (bare state_dict; a legacy '__optimizer__' entry is
    accepted) and, with `engine`, the optimizer sidecar if it exists. The engine's replicas are
    re-synchronised from rank 0 afterwards."""
    import os

    sd = torch.load(path, map_location="cpu")
    opt = sd.pop("__optimizer__", None)
    if opt is None and os.path.exists(_optim_path(path)):
        opt = torch.load(_optim_path(path), map_location="cpu")
    model.load_state_dict(sd, strict=strict)
    if hasattr(model, "invalidate"):
        model.invalidate()
    if engine is not None:
        if opt is not None:
            if list(opt["names"]) != list(engine.names):
                raise ValueError("optimizer state does not match this model's parameter list")
            engine.flat_m.copy_(opt["exp_avg"])
            engine.flat_v.copy_(opt["exp_avg_sq"])
            engine.hyper.copy_(opt["hyper"])
        engine.sync_replicas()


def _stack_views(batch: Dict[str, torch.Tensor]):
    """Accepts the reference loader's two-view dict (trainer.py:99-114: img_0/1, head_pose_0/1,
    gt_gaze, gt_gaze_1) or the V-view form (images [B,V,3,H,W], head_pose [B,V,2], gt_gaze [B,V,2])."""
    if "images" in batch:
        return batch["images"], batch["head_pose"], batch["gt_gaze"]
    images = torch.stack([batch["img_0"], batch["img_1"]], dim=1)
    pose = torch.stack([batch["head_pose_0"], batch["head_pose_1"]], dim=1)
    gt = torch.stack([batch["gt_gaze"], batch["gt_gaze_1"]], dim=1)
    return images, pose, gt


class Trainer:
    """The step / epoch loop of the reference `Trainer` (trainer.py:54-62,84-96,116-147,164-199) on
    the GPU engine: `optim.Adam(lr, weight_decay=1e-6)` with `CyclicLR(base 1e-6, max 1e-3,
    triangular2)` stepped once per EPOCH (quirk Q2), CUDA-graph-captured steps fed from pinned host
    batches with two steps in flight, evaluation with the on-device metric, checkpoints in the
    reference's format. TensorBoard / image dumps / config files (trainer.py:66-82,130-139) are
    outside the hot path and not reproduced. One instance per process (= per GPU); under
    `torch.distributed` the gradients are all-reduced by the engine."""

    def __init__(self, model, steps_per_epoch: int, batch: int, views: int = 2, *,
                 precision: str = "bf16", base_lr: float = 1e-6, max_lr: float = 1e-3,
                 weight_decay: float = 1e-6, decoupled: bool = False, process_group=None):
        from .train import GraphedTrainStep, TrainEngine

        self.model = model
        self.step_size_up = int(steps_per_epoch // 2)            # trainer.py:56-58
        self.step_size_down = steps_per_epoch - self.step_size_up
        self.base_lr, self.max_lr = base_lr, max_lr
        self.sched_steps = 0                                     # scheduler.step() calls so far
        model.train()
        self.engine = TrainEngine(model, precision=precision, lr=self._lr(), weight_decay=weight_decay,
                                  decoupled=decoupled, process_group=process_group)
        self.graph = GraphedTrainStep(self.engine, batch, views)
        self.precision = precision
        self.train_iter = 0

    def _lr(self) -> float:
        return cyclic_lr(self.sched_steps, max(self.step_size_up, 1), max(self.step_size_down, 1),
                         self.base_lr, self.max_lr)

    def train_one_epoch(self, batches: Iterable[Dict[str, torch.Tensor]]) -> float:
        """trainer.py:116-147. Returns the mean training loss of the epoch (read back once)."""
        self.model.train()
        losses, prev = [], None
        for batch in batches:
            images, pose, gt = _stack_views(batch)
            if tuple(images.shape[:2]) != tuple(self.graph.images.shape[:2]):
                # a short last batch (drop_last=False): the captured graphs are for the full batch
                # size, this one runs as plain launches
                if prev is not None:
                    losses.append(float(self.graph.result(prev)))
                    prev = None
                dev = self.engine.device
                rot = RF.pose_to_rotations(pose.float().to(dev).contiguous())
                losses.append(float(self.engine.step(images.float().to(dev), rot, gt.float().to(dev))))
                self.train_iter += 1
                continue
            ticket = self.graph.submit(_pinned(images.float()), _pinned(pose.float()), _pinned(gt.float()))
            if prev is not None:
                losses.append(float(self.graph.result(prev)))
            prev = ticket
            self.train_iter += 1
        if prev is not None:
            losses.append(float(self.graph.result(prev)))
        self.sched_steps += 1                                    # scheduler.step(), once per epoch (:147)
        self.engine.set_lr(self._lr())
        return sum(losses) / max(len(losses), 1)

    def test(self, batches: Iterable[Dict[str, torch.Tensor]]) -> float:
        """trainer.py:164-199: mean angular error (degrees) of view 0 over the loader."""
        def adapt(b):
            images, pose, gt = _stack_views(b)
            return {"images": images, "head_pose": pose, "gt_gaze": gt[:, 0]}
        err = Evaluator(self.model, precision=self.precision).run(adapt(b) for b in batches)
        return err

    def fit(self, train_batches, test_batches, epochs: int = 15, save_epoch: int = 0,
            ckpt_dir: Optional[str] = None):
        """trainer.py:84-96. `train_batches` / `test_batches` are callables returning a fresh
        iterable per epoch. Returns [(epoch, train_loss, test_error)]."""
        import os

        history = [(-1, float("nan"), self.test(test_batches()))]
        for epoch in range(epochs):
            loss = self.train_one_epoch(train_batches())
            error = self.test(test_batches())
            history.append((epoch, loss, error))
            if save_epoch and ckpt_dir and (epoch + 1) % save_epoch == 0:
                name = "epoch_" + str(epoch + 1).zfill(2) + "_error=" + str(round(error, 2)) + ".pth.tar"
                save_checkpoint(os.path.join(ckpt_dir, name), self.model, self.engine)
        return history


def _pinned(t: torch.Tensor) -> torch.Tensor:
    t = t.contiguous()
    return t if (t.is_cuda or t.is_pinned()) else t.pin_memory()
