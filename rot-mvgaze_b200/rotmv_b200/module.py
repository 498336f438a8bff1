"""`FeatRotationSymm` -- host-side mirror of the reference operator (models/rot_mv.py:102-269).

Same constructor, same 348 state_dict keys (so reference checkpoints load with strict=True), same
dict-in/dict-out `forward(data)` contract that `IterationLoss` and `Trainer` consume
(losses/stereo_loss.py:46-50,66-76; trainer.py:122-126,176-181), plus the tensor form
`forward(images[B,V,3,H,W], rotations[B,V,V,3,3]) -> pred_gaze[B,2]` for any V >= 2.

The torch.nn layers built here are PARAMETER CONTAINERS ONLY (they give the reference's key names,
shapes and random-init stream); their `forward` is never called. All compute goes through
librotmv_sm100.so (see engine.py); there is no PyTorch/CPU fallback.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Union

import torch
import torch.nn as nn

from . import engine as E

NUM_FEAT_VEC = 512


def _mlp_params(c_in: int, widths: List[int]) -> nn.Module:
    """Key layout of reference `Mlp` (models/backbones/blocks.py:26-82): blocks.{i}.0.{weight,bias}."""
    holder = nn.Module()
    dims = [c_in] + list(widths)
    holder.blocks = nn.ModuleList(
        [nn.Sequential(nn.Linear(dims[i], dims[i + 1])) for i in range(len(widths))])
    return holder


def _wrap(name: str, child: nn.Module) -> nn.Module:
    holder = nn.Module()
    setattr(holder, name, child)
    return holder


class _BlockParams(nn.Module):
    def __init__(self, kind: str, c_in: int, width: int, stride: int, downsample):
        super().__init__()
        self.kind, self.stride = kind, stride
        if kind == "bottleneck":  # models/resnet.py:99-126
            self.conv1 = nn.Conv2d(c_in, width, 1, bias=False)
            self.bn1 = nn.BatchNorm2d(width)
            self.conv2 = nn.Conv2d(width, width, 3, stride=stride, padding=1, bias=False)
            self.bn2 = nn.BatchNorm2d(width)
            self.conv3 = nn.Conv2d(width, width * 4, 1, bias=False)
            self.bn3 = nn.BatchNorm2d(width * 4)
        else:  # models/resnet.py:50-75
            self.conv1 = nn.Conv2d(c_in, width, 3, stride=stride, padding=1, bias=False)
            self.bn1 = nn.BatchNorm2d(width)
            self.conv2 = nn.Conv2d(width, width, 3, padding=1, bias=False)
            self.bn2 = nn.BatchNorm2d(width)
        self.downsample = downsample


class _TrunkParams(nn.Module):
    """Parameter/buffer tree of the reference ResNet (models/resnet.py:151-218), incl. the unused
    `fc` (Q4) so checkpoints round-trip."""

    def __init__(self, depth: int):
        super().__init__()
        kind, counts, expansion = {50: ("bottleneck", [3, 4, 6, 3], 4),
                                   18: ("basic", [2, 2, 2, 2], 1)}[depth]
        self.kind = kind
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        c_in = 64
        for idx, (width, n_blocks) in enumerate(zip([64, 128, 256, 512], counts)):
            blocks = []
            for b in range(n_blocks):
                s = (1 if idx == 0 else 2) if b == 0 else 1
                ds = None
                if b == 0 and (s != 1 or c_in != width * expansion):
                    ds = nn.Sequential(nn.Conv2d(c_in, width * expansion, 1, stride=s, bias=False),
                                       nn.BatchNorm2d(width * expansion))
                blocks.append(_BlockParams(kind, c_in, width, s, ds))
                c_in = width * expansion
            setattr(self, f"layer{idx + 1}", nn.Sequential(*blocks))
        self.fc = nn.Linear(c_in, 1000)
        self.out_dim = c_in
        for m in self.modules():  # models/resnet.py:203-208
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def blocks(self):
        for li in range(1, 5):
            for blk in getattr(self, f"layer{li}"):
                yield blk


class _TrainStepFunction(torch.autograd.Function):
    """Autograd bridge for train mode: `forward` runs the engine's forward pass (batch-statistic
    BatchNorm, activations kept in the engine's buffers) and returns the per-iteration predictions;
    `backward` receives d(loss)/d(pred) from whatever loss the caller built on them (the reference's
    `IterationLoss`, losses/stereo_loss.py:65-84) and runs the engine's backward kernels, handing the
    parameter gradients to autograd -- so `loss.backward(); optimizer.step()` of trainer.py:141-143
    work unchanged. One backward per forward (no double backward, no retain_graph)."""

    @staticmethod
    def forward(ctx, engine, images, rotations, *params):
        gen = engine._fwd_bwd(images, rotations, None)
        preds = next(gen)
        ctx.gen, ctx.engine = gen, engine
        ctx.pids = [id(p) for p in params]
        ctx.shape = tuple(preds[0].shape)
        return tuple(p.clone() for p in preds)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *dpreds):
        eng, gen = ctx.engine, ctx.gen
        if gen is None:
            raise RuntimeError("FeatRotationSymm (train mode): backward called twice for one forward")
        ctx.gen = None
        ext = [d if d is not None else torch.zeros(ctx.shape, device=eng.device) for d in dpreds]
        try:
            gen.send(ext)
        except StopIteration:
            pass
        grads = tuple(eng.grads[pid].clone() if pid in eng.grads else None for pid in ctx.pids)
        return (None, None, None) + grads


class FeatRotationSymm(nn.Module):
    def __init__(self, backbone_depth: int = 50, num_iter: Optional[int] = None,
                 share_weights: bool = False, encode_rotmat: bool = False,
                 share_feature: bool = False, ignore_rotmat: bool = False, *,
                 precision: str = "bf16", trunk_chunk: int = 512):
        super().__init__()
        self._num_iter = num_iter
        self._output_index = num_iter - 1  # TypeError when num_iter is None, as in the reference
        self._num_feat_vec = NUM_FEAT_VEC
        if backbone_depth not in (18, 50):
            raise ValueError(f"backbone_depth must be 18 or 50, got {backbone_depth}")
        trunk = _TrunkParams(backbone_depth)
        self._feat_extractor = nn.ModuleList([trunk])  # keys "_feat_extractor.0.*" (:124-128)
        self._fc_dim = trunk.out_dim
        self._lifter = _wrap("_lifter", _mlp_params(self._fc_dim, [NUM_FEAT_VEC * 3] * 2))
        assert not (ignore_rotmat and encode_rotmat)
        self._ignore_rotmat, self._encode_rotmat = ignore_rotmat, encode_rotmat
        self._share_feature, self._share_weights = share_feature, share_weights
        if share_feature and share_weights:
            # the reference builds ImageFeatFuser(fc_dim) here (:150-158 wins over :160) and then
            # crashes in torch.cat on the [B,3,512] "image feature" of :201-203
            raise ValueError("share_feature=True with share_weights=True is not a runnable "
                             "configuration of the reference (models/rot_mv.py:150-158,201-203)")
        fuse_in = self._fc_dim + 3 * NUM_FEAT_VEC
        head_in = fuse_in
        if share_feature and not share_weights:   # RotFeatFuser + IntensityBatchNorm (:70-85,13-32)
            fuse_in = head_in = 6 * NUM_FEAT_VEC

        def fuser():
            if share_feature:
                f = _wrap("_fuser", _mlp_params(fuse_in, [fuse_in, fuse_in, 3 * NUM_FEAT_VEC]))
                bn = nn.Module()   # IntensityBatchNorm: running STD kept in a buffer named running_mean
                bn.register_buffer("running_mean", torch.ones(1, 1, NUM_FEAT_VEC))
                bn.momentum, bn.eps = 0.05, 1e-4
                f._batchnorm = bn
                return f
            if encode_rotmat and not ignore_rotmat:   # ImageRotmatFeatFuser (:53-67): + 9 rotation entries
                return _wrap("_fuser", _mlp_params(fuse_in + 9, [fuse_in + 9, fuse_in + 9, 3 * NUM_FEAT_VEC]))
            return _wrap("_fuser", _mlp_params(fuse_in, [fuse_in, 3 * NUM_FEAT_VEC]))

        def head():
            return _mlp_params(head_in, [512, 2])

        if share_weights:  # one module aliased num_iter times (:150-158, Q10)
            self._img_fusers = nn.ModuleList([fuser()] * num_iter)
            self._gaze_estimators = nn.ModuleList([head()] * num_iter)
        else:
            self._img_fusers = nn.ModuleList([fuser() for _ in range(num_iter)])
            self._gaze_estimators = nn.ModuleList([head() for _ in range(num_iter)])
        self.precision = precision
        self.trunk_chunk = trunk_chunk
        # main.py:239-240: StereoL1Loss(rel_weight=0.01, reference_decay=1.0), IterationLoss(0.5)
        self.loss_cfg = {"rel_weight": 0.01, "reference_decay": 1.0, "iter_decay": 0.5}
        self.fuse_loss = False  # dict API: also return data["loss"] from the fused head+loss kernel
        self.auto_graph = True          # tensor API in eval mode: replay CUDA graphs for repeated input shapes
        self.max_graph_sessions = 8
        # uint8 HWC input (images[B,V,H,W,3]): ToTensor + Normalize constants of main.py:38-39
        self.input_mean, self.input_std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
        self._engines: Dict[str, E.InferenceEngine] = {}
        self._train_engine = None        # rotmv_b200.train.TrainEngine behind the train-mode forward

    # -- engine management ---------------------------------------------------------------------
    def engine(self, precision: Optional[str] = None) -> "E.InferenceEngine":
        precision = precision or self.precision
        eng = self._engines.get(precision)
        if eng is None or eng.stale():
            eng = E.InferenceEngine(self, precision)
            self._engines[precision] = eng
        return eng

    def invalidate(self) -> None:
        """Drop cached folded/converted weights (call after changing parameters in place)."""
        self._engines.clear()

    def train(self, mode: bool = True):
        self._engines.clear()
        return super().train(mode)

    # -- forward -------------------------------------------------------------------------------
    def forward(self, data_or_images: Union[Dict[str, Any], torch.Tensor],
                rotations: Optional[torch.Tensor] = None, *, precision: Optional[str] = None):
        if isinstance(data_or_images, dict):
            return self._forward_dict(data_or_images, precision)
        images = data_or_images
        if self.training:
            return self._forward_train(images, rotations, precision)["pred_gaze"]
        if self.auto_graph and not self.training and images.is_cuda and rotations is not None:
            pred = self._forward_graphed(images, rotations, precision)
            if pred is not None:
                return pred
        out = self.forward_views(images, rotations, precision=precision, want_all=False)
        return out["pred_gaze"]

    def _forward_graphed(self, images, rotations, precision):
        """Serving path: the second call with an input signature (batch, views, size, dtype) captures
        the forward as CUDA graphs and every later call replays them -- the eager path spends 1.6 ms of
        host time per call in 72 launches + tensor-map encodes whatever the batch (B = 1: 0.69 ms
        graphed). At most `max_graph_sessions` signatures are kept (least recently used first out)."""
        eng = self.engine(precision)
        if images.dim() != 5 or rotations.dim() != 5 or images.shape[0] == 0:
            return None
        key = (tuple(images.shape), images.dtype, tuple(rotations.shape))
        cache = eng.__dict__.setdefault("_sessions", {})
        seen = eng.__dict__.setdefault("_seen", set())
        sess = cache.get(key)
        if sess is None:
            if key not in seen:          # first sighting: run eagerly, capture only if it comes back
                seen.add(key)
                return None
            b, v = images.shape[0], images.shape[1]
            u8 = images.dtype == torch.uint8
            size = images.shape[2] if u8 else images.shape[3]
            if (images.shape[2] != images.shape[3]) if u8 else (images.shape[3] != images.shape[4]):
                return None              # GraphedForward sessions are square-image only
            if images.dtype not in (torch.uint8, torch.float32):
                return None
            while len(cache) >= self.max_graph_sessions:
                cache.pop(next(iter(cache)))
            sess = E.GraphedForward(self, b, v, precision=precision, size=size,
                                    input_dtype=torch.uint8 if u8 else torch.float32)
            cache[key] = sess
        else:
            cache[key] = cache.pop(key)  # most recently used last
        return sess(images.contiguous(), rotations.float().contiguous()).clone()

    def forward_views(self, images: torch.Tensor, rotations: torch.Tensor, *,
                      precision: Optional[str] = None, want_all: bool = True,
                      gt: Optional[torch.Tensor] = None) -> Dict[str, Any]:
        if self.training:
            raise NotImplementedError(
                "train-mode forward runs through rotmv_b200.train.TrainEngine / GraphedTrainStep "
                "(batch-statistic BatchNorm + backward); module.forward is the inference path")
        if images.dim() != 5 or rotations is None or rotations.dim() != 5:
            raise ValueError("expected images[B,V,3,H,W] (fp32) or [B,V,H,W,3] (uint8) and rotations[B,V,V,3,3]")
        return self.engine(precision).run(images, rotations, want_all=want_all, gt=gt)

    def _forward_dict(self, data: Dict[str, Any], precision) -> Dict[str, Any]:
        """Reference dict contract for two views (models/rot_mv.py:187-269); mutates `data`."""
        from . import functional as RF

        images = torch.stack([data["img_0"], data["img_1"]], dim=1)
        # rot_10 = rot_0 rot_1^T, rot_01 = rot_1 rot_0^T (:193-194) -> rotations[b,i,j] = R_i R_j^T
        rotations = RF.relative_rotations(torch.stack([data["rot_0"], data["rot_1"]], dim=1))
        gt = None
        if "gt_gaze" in data and "gt_gaze_1" in data and self.fuse_loss:
            gt = torch.stack([data["gt_gaze"], data["gt_gaze_1"]], dim=1)
        if self.training:
            out = self._forward_train(images, rotations, precision)
        else:
            out = self.forward_views(images, rotations, precision=precision, want_all=True, gt=gt)
        data.update(out)
        return data

    def _forward_train(self, images: torch.Tensor, rotations: torch.Tensor, precision) -> Dict[str, Any]:
        """Train-mode forward through the training engine, connected to autograd by
        `_TrainStepFunction`: returns the reference's full dict (models/rot_mv.py:205-211,256-266):
        `num_iter`, `img_feat_k`, `initial_rot_feat_k`, `iter_i` -> {`feat_k`, `pred_gaze_k`} and
        `pred_gaze`. The predictions carry the autograd graph (what `IterationLoss`,
        losses/stereo_loss.py:66-76, differentiates); the feature entries are detached fp32 copies of
        the engine's buffers (nothing in the reference's step differentiates through them). The fused,
        CUDA-graph captured step (`rotmv_b200.train.GraphedTrainStep`, `rotmv_b200.loop.Trainer`)
        is the fast path; this is the drop-in one."""
        from .train import TrainEngine

        precision = precision or self.precision
        eng = self._train_engine
        if eng is None or eng.precision != precision:
            eng = self._train_engine = TrainEngine(self, precision=precision, lr=0.0, weight_decay=0.0)
        if images.dim() != 5 or rotations is None or rotations.dim() != 5:
            raise ValueError("expected images[B,V,3,H,W] (fp32) and rotations[B,V,V,3,3]")
        b, v = images.shape[0], images.shape[1]
        preds = _TrainStepFunction.apply(eng, images, rotations, *self.parameters())
        out: Dict[str, Any] = {"num_iter": self._num_iter}
        feats = eng.last_feats
        nv = self._num_feat_vec

        def per_view(t2d, tail):
            t = t2d.detach().float().reshape(b, v, *tail)
            return [t[:, k].contiguous() for k in range(v)]

        img_tail = (3, nv) if self._share_feature else (self._fc_dim,)
        for k, t in enumerate(per_view(feats["img"], img_tail)):
            out[f"img_feat_{k}"] = t
        for k, t in enumerate(per_view(feats["init"], (3, nv))):
            out[f"initial_rot_feat_{k}"] = t
        for i, p in enumerate(preds):
            pv = p.view(b, v, 2)
            it = {f"pred_gaze_{k}": pv[:, k] for k in range(v)}
            for k, t in enumerate(per_view(feats["iters"][i], (3, nv))):
                it[f"feat_{k}"] = t
            out[f"iter_{i}"] = it
        out["pred_gaze"] = out[f"iter_{self._output_index}"]["pred_gaze_0"]
        return out
