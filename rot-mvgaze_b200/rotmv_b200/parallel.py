"""Data parallelism for the Rot-MV path: one process per GPU, the batch of multi-view
samples is sharded across ranks (all V views of a sample stay on one GPU), weights / BatchNorm
buffers / optimizer state are replicated. Inference needs no collective; training needs ONE
gradient all-reduce of the flat fp32 gradient buffer per step (NCCL over NVLink on GPUs; the same
code runs on `gloo` for the CPU tests). BatchNorm statistics stay per rank, as they would under DDP
of the reference (it has no SyncBN).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of samples owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(global_batch, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place average of a flat gradient buffer over the ranks (sum, then scale)."""
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            flat.mul_(1.0 / world)
    return flat


def flat_layout(named_params, align: int = 64, skip_prefix: str = "_feat_extractor.0.fc."):
    """Offsets of every trained parameter inside the flat fp32 buffers (parameters, gradients, Adam
    moments), in `named_parameters()` order, each padded to `align` elements. `fc.*` never receives a
    gradient and is left out. Returns (names, offsets, total)."""
    names, offs, total = [], [], 0
    for n, p in named_params:
        if n.startswith(skip_prefix):
            continue
        names.append(n)
        offs.append(total)
        total += (p.numel() + align - 1) // align * align
    return names, offs, total


def gradient_buckets(names, offs, total, trunk_prefix: str = "_feat_extractor.0."):
    """Contiguous slices of the flat gradient buffer in the order the backward pass completes them:
    the fusion stage (lifter / fusers / heads: everything behind the trunk in parameter order, 74 % of
    the bytes, final before the trunk backward starts), then layer4, layer3, and layer2 + layer1 + stem
    together (1.4 M parameters). Returns [(name, begin, end)]; the slices partition [0, total)."""
    def first(prefix, default):
        return next((o for n, o in zip(names, offs) if n.startswith(prefix)), default)
    split = next((o for n, o in zip(names, offs) if not n.startswith(trunk_prefix)), total)
    o4 = first(trunk_prefix + "layer4.", split)
    o3 = first(trunk_prefix + "layer3.", o4)
    return [("fusion", split, total), ("layer4", o4, split), ("layer3", o3, o4), ("layer2-stem", 0, o3)]


def broadcast_state_(tensors, module=None, src: int = 0, group=None) -> None:
    """Make rank `src`'s optimizer/parameter tensors (and the module's buffers) every replica's."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in tensors:
            dist.broadcast(t, src=src, group=group)
        if module is not None:
            broadcast_buffers_(module, src=src, group=group)


class OverlappedAllReduce:
    """Sum-all-reduce of a flat gradient buffer in the order its slices become final, each slice as
    an asynchronous collective (NCCL runs it on its own stream, ordered after the work already
    queued on the caller's stream), so that communication of the early slices overlaps the compute
    that is still producing the late ones. `finish()` makes the caller's stream wait for all of
    them (the host does not block on NCCL). With no process group (single process) it is a no-op.

    The Rot-MV step uses the four slices of `gradient_buckets` (fusion stage, layer4, layer3,
    layer2..stem), each started right after the backward kernels that write it have been queued."""

    def __init__(self, flat: torch.Tensor, group=None):
        self.flat, self.group, self.works = flat, group, []
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1

    def start(self, begin: int, end: int) -> None:
        if self.active and end > begin:
            self.works.append(dist.all_reduce(self.flat[begin:end], op=dist.ReduceOp.SUM,
                                              group=self.group, async_op=True))

    def finish(self) -> None:
        for w in self.works:
            w.wait()
        self.works = []


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Max of a host scalar over the ranks (multi-GPU timings are reported as the slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def broadcast_buffers_(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Checkpoint-time helper: make rank `src`'s BatchNorm running statistics the saved ones."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for b in module.buffers():
            dist.broadcast(b, src=src, group=group)
