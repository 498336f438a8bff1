"""CPU oracle for the Rot-MVGaze multi-view hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this file; the product path (rot-mvgaze_b200/) never does and has no CPU fallback.

What it is: an independent restatement, in plain PyTorch CPU ops, of the reference algorithm
(ut-vision/Rot-MVGaze). All reference arithmetic lives in PyTorch itself (un-pinned in
requirements.txt:3; torch 2.11.0+cu128 here), so the restatement calls the same torch.nn ops in the
same order and is bit-identical to the imported reference on CPU -- `oracle/make_golden.py` checks
that in this container (where /root/reference exists) and writes the committed fixtures under
tests/golden/. PARITY PINNING: the reference ships no tests/golden vectors of its own, so
the oracle is pinned by (a) bit-equality with the imported reference module on seeded inputs
(make_golden.py, asserted at generation time) and (b) the committed outputs of that reference run.

Every function cites the reference lines it restates (paths relative to the reference repo).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

NUM_FEAT_VEC = 512  # models/rot_mv.py:117


# ------------------------------------------------------------------------------------------------
# Trunk: torchvision-style ResNet (models/resnet.py:99-275)
# ------------------------------------------------------------------------------------------------
class _Bottleneck(nn.Module):
    """models/resnet.py:99-148 (expansion 4; stride lives on the 3x3 conv)."""

    expansion = 4

    def __init__(self, c_in: int, width: int, stride: int, downsample: Optional[nn.Module]):
        super().__init__()
        self.conv1 = nn.Conv2d(c_in, width, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, width * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(width * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        skip = x if self.downsample is None else self.downsample(x)
        y += skip
        return self.relu(y)


class _BasicBlock(nn.Module):
    """models/resnet.py:50-96 (expansion 1)."""

    expansion = 1

    def __init__(self, c_in: int, width: int, stride: int, downsample: Optional[nn.Module]):
        super().__init__()
        self.conv1 = nn.Conv2d(c_in, width, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(width, width, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.downsample = downsample

    def forward(self, x):
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        skip = x if self.downsample is None else self.downsample(x)
        y += skip
        return self.relu(y)


class TrunkOracle(nn.Module):
    """models/resnet.py:151-275: stem, four stages, global average pool; `fc` is registered (and
    saved in checkpoints) but never used by forward (:201, :261-275)."""

    def __init__(self, depth: int):
        super().__init__()
        block, counts = {50: (_Bottleneck, [3, 4, 6, 3]), 18: (_BasicBlock, [2, 2, 2, 2])}[depth]
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        c_in = 64
        for idx, (width, n_blocks) in enumerate(zip([64, 128, 256, 512], counts)):
            stride = 1 if idx == 0 else 2
            blocks: List[nn.Module] = []
            for b in range(n_blocks):
                s = stride if b == 0 else 1
                ds = None
                if b == 0 and (s != 1 or c_in != width * block.expansion):
                    # created BEFORE the block, as in _make_layer (:226-231): RNG order matters
                    ds = nn.Sequential(
                        nn.Conv2d(c_in, width * block.expansion, 1, stride=s, bias=False),
                        nn.BatchNorm2d(width * block.expansion),
                    )
                blocks.append(block(c_in, width, s, ds))
                c_in = width * block.expansion
            setattr(self, f"layer{idx + 1}", nn.Sequential(*blocks))
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(c_in, 1000)
        self.out_dim = c_in
        # models/resnet.py:203-208
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.avgpool(x)  # take_avg=True default (:162, :272-273)


# ------------------------------------------------------------------------------------------------
# MLPs (models/backbones/blocks.py:7-82): Linear+ReLU blocks, last block bare
# ------------------------------------------------------------------------------------------------
class MlpOracle(nn.Module):
    def __init__(self, c_in: int, widths: List[int]):
        super().__init__()
        dims = [c_in] + list(widths)
        blocks = []
        for i in range(len(widths)):
            layers: List[nn.Module] = [nn.Linear(dims[i], dims[i + 1])]
            if i != len(widths) - 1:
                layers.append(nn.ReLU())
            blocks.append(nn.Sequential(*layers))
        self.blocks = nn.ModuleList(blocks)

    def forward(self, x):
        for b in self.blocks:
            x = b(x)
        return x


class _Lifter(nn.Module):
    """models/rot_mv.py:91-98."""

    def __init__(self, c_in: int):
        super().__init__()
        self._lifter = MlpOracle(c_in, [NUM_FEAT_VEC * 3, NUM_FEAT_VEC * 3])

    def forward(self, x):
        return self._lifter(x).reshape(-1, 3, NUM_FEAT_VEC)


class _ImageFeatFuser(nn.Module):
    """models/rot_mv.py:35-50."""

    def __init__(self, img_dim: int):
        super().__init__()
        c = img_dim + 3 * NUM_FEAT_VEC
        self._fuser = MlpOracle(c, [c, 3 * NUM_FEAT_VEC])

    def forward(self, img_feat, rot_feat):
        return self._fuser(torch.cat([img_feat, rot_feat.flatten(-2, -1)], dim=-1))


class _ImageRotmatFeatFuser(nn.Module):
    """models/rot_mv.py:53-67 (encode_rotmat=True)."""

    def __init__(self, img_dim: int):
        super().__init__()
        c = img_dim + 3 * NUM_FEAT_VEC + 9
        self._fuser = MlpOracle(c, [c, c, 3 * NUM_FEAT_VEC])

    def forward(self, img_feat, rot_feat, rot):
        return self._fuser(
            torch.cat([img_feat, rot_feat.flatten(-2, -1), rot.flatten(-2, -1)], dim=-1))


class _IntensityBatchNorm(nn.Module):
    """models/rot_mv.py:13-32: x / (running_std + eps); the running STD of the per-vector norm lives
    in a buffer called `running_mean` (initialised to 1); eps is used twice (clamp and divide)."""

    def __init__(self, n_channels: int, momentum: float = 0.05, eps: float = 1e-4):
        super().__init__()
        self.register_buffer("running_mean", torch.ones(1, 1, n_channels))
        self._momentum, self._eps = momentum, eps

    def forward(self, x):
        intensity = torch.norm(x, dim=-2, keepdim=True).detach()
        var = torch.var(intensity, unbiased=False, dim=0, keepdim=True)
        std = torch.sqrt(var.clamp_min(self._eps))
        if self.training:
            self.running_mean = self.running_mean * (1 - self._momentum) + std * self._momentum
        return x / (self.running_mean + self._eps)


class _RotFeatFuser(nn.Module):
    """models/rot_mv.py:70-85 (share_feature=True)."""

    def __init__(self):
        super().__init__()
        c = 6 * NUM_FEAT_VEC
        self._fuser = MlpOracle(c, [c, c, 3 * NUM_FEAT_VEC])
        self._batchnorm = _IntensityBatchNorm(NUM_FEAT_VEC)

    def forward(self, feat_0, feat_1):
        x = torch.cat([self._batchnorm(feat_0), self._batchnorm(feat_1)], dim=-1).flatten(-2, -1)
        return self._fuser(x).reshape(-1, 3, NUM_FEAT_VEC)


class RotMVOracle(nn.Module):
    """models/rot_mv.py:102-269 (`FeatRotationSymm`) with the same state_dict keys, all
    constructor flags included (share_feature / encode_rotmat: two views only, like the reference).
    """

    def __init__(self, backbone_depth: int = 50, num_iter: Optional[int] = None,
                 share_weights: bool = False, encode_rotmat: bool = False,
                 share_feature: bool = False, ignore_rotmat: bool = False):
        super().__init__()
        self._num_iter = num_iter
        self._output_index = num_iter - 1  # TypeError for None, like the reference (:115)
        trunk = TrunkOracle(backbone_depth)
        self._feat_extractor = nn.Sequential(trunk, trunk.avgpool, nn.Flatten(-3, -1))  # :124-128
        self._fc_dim = trunk.out_dim
        self._lifter = _Lifter(self._fc_dim)
        assert not (ignore_rotmat and encode_rotmat)
        self._ignore_rotmat, self._encode_rotmat = ignore_rotmat, encode_rotmat
        self._share_feature = share_feature
        fuser_cls = _ImageRotmatFeatFuser if (encode_rotmat and not ignore_rotmat) else _ImageFeatFuser
        head_in = 3 * NUM_FEAT_VEC + self._fc_dim
        if share_weights:  # one module aliased num_iter times (:150-158)
            self._img_fusers = nn.ModuleList([fuser_cls(self._fc_dim)] * num_iter)
            self._gaze_estimators = nn.ModuleList([MlpOracle(head_in, [512, 2])] * num_iter)
        elif share_feature:  # :160-171
            self._img_fusers = nn.ModuleList([_RotFeatFuser() for _ in range(num_iter)])
            self._gaze_estimators = nn.ModuleList(
                [MlpOracle(6 * NUM_FEAT_VEC, [512, 2]) for _ in range(num_iter)])
        else:
            self._img_fusers = nn.ModuleList([fuser_cls(self._fc_dim) for _ in range(num_iter)])
            self._gaze_estimators = nn.ModuleList(
                [MlpOracle(head_in, [512, 2]) for _ in range(num_iter)])

    # -- general V-view forward; identical to the reference at V == 2 ---------------
    def forward_views(self, images: torch.Tensor, rotations: torch.Tensor) -> Dict:
        """images [B,V,3,H,W], rotations [B,V,V,3,3] with rotations[b,i,j] = R_i R_j^T."""
        n_views = images.shape[1]
        # one trunk call PER VIEW: BatchNorm batch statistics are per view (:196-197)
        img_feat = [self._feat_extractor(images[:, v]) for v in range(n_views)]
        rot_feat = [self._lifter(f) for f in img_feat]
        if self._share_feature or self._encode_rotmat:
            assert n_views == 2, "share_feature / encode_rotmat are two-view configurations"
        if self._share_feature:   # :201-203: the lifted feature replaces the image feature
            img_feat = rot_feat
        out: Dict = {"num_iter": self._num_iter}
        for v in range(n_views):
            out[f"img_feat_{v}"] = img_feat[v]
            out[f"initial_rot_feat_{v}"] = rot_feat[v]
        for i, (fuser, head) in enumerate(zip(self._img_fusers, self._gaze_estimators)):
            old = rot_feat  # Jacobi update: every view reads the OLD partner features (:217,234-239)
            new = []
            for v in range(n_views):
                partners = [u for u in range(n_views) if u != v]
                if self._ignore_rotmat:
                    agg = old[partners[0]] if len(partners) == 1 else sum(
                        old[u] for u in partners) / len(partners)
                    f = fuser(img_feat[v], agg)
                else:
                    rot = [rotations[:, v, u] for u in partners]
                    if self._encode_rotmat:
                        # :225-231: the partner feature is NOT rotated, the matrix is an input
                        f = fuser(img_feat[v], old[partners[0]], rot[0])
                    else:
                        if len(partners) == 1:
                            agg = rot[0] @ old[partners[0]]
                        else:
                            agg = sum(r @ old[u] for r, u in zip(rot, partners)) / len(partners)
                        f = fuser(img_feat[v], agg)
                new.append(f.reshape(-1, 3, NUM_FEAT_VEC))
            rot_feat = new
            it = {}
            for v in range(n_views):
                it[f"feat_{v}"] = rot_feat[v]
                if self._share_feature:   # :243-248
                    it[f"pred_gaze_{v}"] = head(torch.cat([img_feat[v], rot_feat[v]], dim=-1).flatten(1, -1))
                else:
                    it[f"pred_gaze_{v}"] = head(torch.cat([img_feat[v], rot_feat[v].flatten(1, -1)], dim=-1))
            out[f"iter_{i}"] = it
        out["pred_gaze"] = out[f"iter_{self._output_index}"]["pred_gaze_0"]  # :265
        return out

    # -- reference dict API (two views) ------------------------------------------------------------
    def forward(self, data: Dict) -> Dict:
        rot_0, rot_1 = data["rot_0"], data["rot_1"]
        eye = torch.eye(3, dtype=rot_0.dtype).expand_as(rot_0)
        r01 = rot_0 @ rot_1.transpose(-1, -2)  # "rot_10" (:193): view-1 frame -> view-0 frame
        r10 = rot_1 @ rot_0.transpose(-1, -2)  # "rot_01" (:194)
        rotations = torch.stack([torch.stack([eye, r01], 1), torch.stack([r10, eye], 1)], 1)
        images = torch.stack([data["img_0"], data["img_1"]], 1)
        data.update(self.forward_views(images, rotations))  # input dict mutated, as in :266-269
        return data


# ------------------------------------------------------------------------------------------------
# Loss / metric / pose math
# ------------------------------------------------------------------------------------------------
def pitchyaw_to_vector(py: torch.Tensor) -> torch.Tensor:
    """utils/math.py:52-60; output is always fp32 (Q7)."""
    s, c = torch.sin(py), torch.cos(py)
    out = torch.empty((py.shape[0], 3), device=py.device)
    out[:, 0] = c[:, 0] * s[:, 1]
    out[:, 1] = s[:, 0]
    out[:, 2] = c[:, 0] * c[:, 1]
    return out


def angular_loss_deg(pred: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """losses/gaze_loss.py:42-52 (hardtanh-clamped cosine, degrees, batch mean)."""
    a, b = pitchyaw_to_vector(label), pitchyaw_to_vector(pred)
    sim = F.hardtanh(F.cosine_similarity(a, b, eps=1e-6), -1.0, 1.0)
    return torch.mean(torch.acos(sim) * (180 / math.pi))


def iteration_loss(out: Dict, gt: List[torch.Tensor], rel_weight: float = 0.01,
                   reference_decay: float = 1.0, iter_decay: float = 0.5) -> torch.Tensor:
    """losses/stereo_loss.py:46-54 (StereoL1Loss) inside :65-84 (IterationLoss), with the
    main.py:239-240 hyper-parameters as defaults; gt[v] is the label of view v. For V > 2 every
    non-zero view is weighted by reference_decay."""
    total = 0
    for i in range(out["num_iter"]):
        it = out[f"iter_{i}"]
        li = angular_loss_deg(it["pred_gaze_0"], gt[0]).mean()
        for v in range(1, len(gt)):
            li = li + angular_loss_deg(it[f"pred_gaze_{v}"], gt[v]).mean() * reference_decay
        total = total * iter_decay + li * rel_weight
    return total


def rotation_matrix_2d(pitch_yaw: torch.Tensor) -> torch.Tensor:
    """utils/math.py:188-219: R = R_y(yaw) @ R_x(-pitch)."""
    py = pitch_yaw.reshape(-1, 2) * torch.tensor([-1.0, 1.0])
    c, s = torch.cos(py), torch.sin(py)
    one, zero = torch.ones_like(c[:, 0]), torch.zeros_like(c[:, 0])
    rx = torch.stack([one, zero, zero, zero, c[:, 0], -s[:, 0], zero, s[:, 0], c[:, 0]], 1).view(-1, 3, 3)
    ry = torch.stack([c[:, 1], zero, s[:, 1], zero, one, zero, -s[:, 1], zero, c[:, 1]], 1).view(-1, 3, 3)
    return ry @ rx


def pairwise_rotations(head_pose: torch.Tensor) -> torch.Tensor:
    """head_pose [B,V,2] -> rotations [B,V,V,3,3], [b,i,j] = R_i R_j^T (models/rot_mv.py:193-194)."""
    b, v, _ = head_pose.shape
    r = rotation_matrix_2d(head_pose.reshape(-1, 2)).view(b, v, 3, 3)
    return r[:, :, None] @ r[:, None].transpose(-1, -2)


def angular_error_deg(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """utils/math.py:122-137 (torch metric, no clamp)."""
    a, b = pitchyaw_to_vector(a), pitchyaw_to_vector(b)
    sim = (a * b).sum(1) / (a.norm(dim=1).clamp(min=1e-7) * b.norm(dim=1).clamp(min=1e-7))
    return torch.acos(sim) * 180.0 / math.pi


# ------------------------------------------------------------------------------------------------
# Step (trainer.py:54,119,121-123,141-143) and synthetic inputs
# ------------------------------------------------------------------------------------------------
def make_adam(model: nn.Module, lr: float = 1e-6, weight_decay: float = 1e-6):
    """trainer.py:54: Adam with COUPLED L2 (not AdamW); CyclicLR base_lr=1e-6 (:57)."""
    return torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay)


def train_step(model: RotMVOracle, opt, images, rotations, gt) -> torch.Tensor:
    model.train()
    out = model.forward_views(images, rotations)
    loss = iteration_loss(out, [gt[:, v] for v in range(gt.shape[1])])
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss.detach()


def seed_all(seed: int = 0) -> None:
    """utils/util.py:7-16 (CPU part)."""
    import random

    import numpy as np

    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def synthetic_batch(batch: int, views: int, seed: int = 1, size: int = 224):
    """images ~ N(0,1); head poses and gaze labels ~ U(-0.5, 0.5) rad."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn((batch, views, 3, size, size), generator=g)
    pose = torch.rand((batch, views, 2), generator=g) - 0.5
    gt = torch.rand((batch, views, 2), generator=g) - 0.5
    return images, pose, gt


def build_model(num_iter: int = 3, depth: int = 50, seed: int = 0, **flags) -> RotMVOracle:
    seed_all(seed)
    return RotMVOracle(backbone_depth=depth, num_iter=num_iter, **flags)


def calibrate_bn(model: nn.Module, images: torch.Tensor, passes: int = 4) -> None:
    """a few train-mode passes with momentum=None (cumulative average) so eval-mode
    activations are O(1) instead of O(100) at random init."""
    bns = [m for m in model.modules() if isinstance(m, nn.BatchNorm2d)]
    saved = [m.momentum for m in bns]
    for m in bns:
        m.momentum = None
        m.reset_running_stats()
    model.train()
    with torch.no_grad():
        for _ in range(passes):
            for v in range(images.shape[1]):
                model._feat_extractor(images[:, v])
    for m, mom in zip(bns, saved):
        m.momentum = mom
    model.eval()
