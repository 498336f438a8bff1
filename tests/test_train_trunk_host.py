"""Host logic of the WHOLE training step (rotmv_b200/train.py::TrainEngine._fwd_bwd) checked on the
CPU: every kernel-level method of the engine -- convolutions, BatchNorm passes, the recomputed
BatchNorm of the expanding 1x1 convs (`_conv_bn_fwd` / `_conv_bn_bwd`), data / weight gradients with
ReLU-mask pre-multiplication, pools, the fusion stage -- is replaced by its torch formula, so what
runs is the orchestration itself: which tensor feeds which kernel, per-view statistics, the order of
the backward pass, `dyr == d_out` for the recomputed blocks, the masks handed from block to block,
the gradient-bucket hooks. Loss and all 187 parameter gradients are compared with the CPU oracle's
autograd on the same weights and inputs (trainer.py:119-123,141-142). The kernels themselves are
tested against torch on the GPU (tests/test_bnconv_gpu.py, tests/test_train_gpu.py)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import rotmv_oracle as O
from rotmv_b200 import parallel as P
from rotmv_b200 import train as T
from rotmv_b200.module import FeatRotationSymm

from test_train_fusion_host import HostEngine, torch_kernels  # noqa: F401  (fixture)


def nchw(t):
    return t.permute(0, 3, 1, 2)


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


class HostBN:
    def __init__(self, bn, grads):
        self.gamma, self.beta = bn.weight, bn.bias
        self.eps = float(bn.eps)
        self.dgamma, self.dbeta = grads[id(bn.weight)], grads[id(bn.bias)]
        self.nbt = bn.num_batches_tracked
        self.mean = self.invstd = None


class HostTrunkEngine(HostEngine):
    """TrainEngine whose kernel-level methods are torch formulas (fp32, CPU)."""

    def __init__(self, model, recompute=True):
        super().__init__(model)
        self.precision = "bf16"          # selects the bf16 engine's ORCHESTRATION (recompute, masks); math is fp32
        trunk = model._feat_extractor[0]
        self.max_views, self.views = 8, 2
        self.recompute_bn, self.recompute_max_cin, self.fuse_bn_bwd, self.fuse_bn_stats = recompute, 4096, False, True
        self.stem_conv, self.stem_bn = trunk.conv1, HostBN(trunk.bn1, self.grads)
        self.blocks = []
        for blk in trunk.blocks():
            e = {"convs": [blk.conv1, blk.conv2, blk.conv3],
                 "bns": [HostBN(b, self.grads) for b in (blk.bn1, blk.bn2, blk.bn3)]}
            if blk.downsample is not None:
                e["ds_conv"], e["ds_bn"] = blk.downsample[0], HostBN(blk.downsample[1], self.grads)
            self.blocks.append(e)
        names, offs, total = P.flat_layout(model.named_parameters())
        self.buckets = P.gradient_buckets(names, offs, total)
        self._bucket_by_stage = {4: self.buckets[1], 3: self.buckets[2]}
        self._stage_first, bi = {}, 0
        for li in range(1, 5):
            self._stage_first[bi] = li
            bi += len(getattr(trunk, f"layer{li}"))
        self._bits, self._keep = {}, []
        self.calls = []

    # ---- bookkeeping -------------------------------------------------------------------------------
    def _require_device(self, images):
        pass

    def _begin_step(self):
        for g in self.grads.values():
            g.zero_()
        self.loss.zero_()

    def _end_step(self):
        pass

    def _w_fwd(self, conv, tag):
        return conv.weight.detach().permute(0, 2, 3, 1).contiguous()      # KRSC, as the engine's layout

    # ---- BatchNorm (train), statistics per view: image n belongs to view n % views ---------------
    def _bn_apply(self, bn, z, residual, relu):
        v = self.views
        y = torch.empty_like(z)
        bn.mean, bn.invstd = [], []
        for k in range(v):
            zk = z[k::v]
            mean = zk.mean(dim=(0, 1, 2))
            var = zk.var(dim=(0, 1, 2), unbiased=False)
            invstd = torch.rsqrt(var + bn.eps)
            bn.mean.append(mean); bn.invstd.append(invstd)
            y[k::v] = (zk - mean) * invstd * bn.gamma.detach() + bn.beta.detach()
        bn.nbt += v
        if residual is not None:
            y = y + residual
        mask = y > 0
        if relu:
            y = torch.relu(y)
        self._keep.append((y, mask))
        self._bits[id(y)] = mask if relu else None
        return y

    def _bn_grad(self, bn, z, dyr):
        """dz, and dgamma / dbeta (=) of per-view batch-statistic BatchNorm for a masked gradient dyr."""
        v = self.views
        dz = torch.empty_like(z)
        dg = torch.zeros_like(bn.gamma.detach()); db = torch.zeros_like(dg)
        for k in range(v):
            xhat = (z[k::v] - bn.mean[k]) * bn.invstd[k]
            d = dyr[k::v]
            s1, s2 = d.mean(dim=(0, 1, 2)), (d * xhat).mean(dim=(0, 1, 2))
            dz[k::v] = bn.gamma.detach() * bn.invstd[k] * (d - s1 - xhat * s2)
            dg += (d * xhat).sum(dim=(0, 1, 2)); db += d.sum(dim=(0, 1, 2))
        bn.dgamma.copy_(dg); bn.dbeta.copy_(db)
        return dz

    def _conv_stats(self, x, w, *, stride=1, pad=0, out=None, bn=None):
        self.calls.append("conv")
        return nhwc(F.conv2d(nchw(x), w.permute(0, 3, 1, 2), stride=stride, padding=pad)), False

    def _bn_fwd(self, bn, z, residual, relu, tag, stats_done=False):
        self.calls.append("bn_fwd")
        return self._bn_apply(bn, z, residual, relu)

    def _bn_bwd(self, bn, z, dy, mask, tag, want_dyr=False):
        self.calls.append("bn_bwd")
        m = self._bits.get(id(mask)) if mask is not None else None
        dyr = dy if m is None else dy * m
        return self._bn_grad(bn, z, dyr), (dyr if want_dyr else None)

    def _conv_bn_fwd(self, bn, x, w, stride, residual, relu, tag):
        self.calls.append("conv_bn_fwd")
        bn.z_recomputed = nhwc(F.conv2d(nchw(x), w.permute(0, 3, 1, 2), stride=stride))   # NOT handed to the engine
        return self._bn_apply(bn, bn.z_recomputed, residual, relu)

    def _conv_bn_bwd(self, bn, x, w, stride, dy, tag):
        self.calls.append("conv_bn_bwd")
        z = nhwc(F.conv2d(nchw(x), w.permute(0, 3, 1, 2), stride=stride))
        return self._bn_grad(bn, z, dy)          # dy arrives multiplied by the ReLU derivative

    # ---- convolution gradients ------------------------------------------------------------------
    def _dgrad(self, dz, conv, tag, in_shape, residual=None, mask_bits=None, bwd_bn=None):
        self.calls.append("dgrad" + ("+mask" if mask_bits is not None else ""))
        st, pd, k = conv.stride[0], conv.padding[0], conv.kernel_size[0]
        h, w = in_shape[1], in_shape[2]
        op = (h - ((dz.shape[1] - 1) * st - 2 * pd + k), w - ((dz.shape[2] - 1) * st - 2 * pd + k))
        dx = nhwc(F.conv_transpose2d(nchw(dz), conv.weight.detach(), stride=st, padding=pd, output_padding=op))
        if residual is not None:
            dx = dx + residual
        if mask_bits is not None:
            dx = dx * mask_bits
        return dx

    def _wgrad(self, x, dy, conv_or_lin, kh, kw, stride, pad, x_strides=None, grad=None):
        g = grad if grad is not None else self.grads[id(conv_or_lin.weight)]
        if x_strides is not None or x.dim() != 4 or isinstance(conv_or_lin, torch.nn.Linear):
            return super()._wgrad(x, dy, conv_or_lin, kh, kw, stride, pad, x_strides, grad)
        g += torch.nn.grad.conv2d_weight(nchw(x), conv_or_lin.weight.shape, nchw(dy), stride=stride, padding=pad)

    # ---- stem / pools ------------------------------------------------------------------------------
    def _stem_fwd(self, imgs):
        return nhwc(F.conv2d(imgs.to(self.dt), self.stem_conv.weight.detach(), stride=2, padding=3))

    def _stem_wgrad(self, imgs, dz0):
        self.grads[id(self.stem_conv.weight)].add_(
            torch.nn.grad.conv2d_weight(imgs.to(self.dt), self.stem_conv.weight.shape, nchw(dz0), stride=2, padding=3))

    def _maxpool_fwd(self, y0):
        src = nchw(y0).detach().clone().requires_grad_(True)
        out = F.max_pool2d(src, 3, 2, 1)
        return nhwc(out.detach()), (src, out)

    def _maxpool_bwd(self, idx, d_out, y0):
        src, out = idx
        out.backward(nchw(d_out))
        return nhwc(src.grad)

    def _avgpool_bwd(self, dimg, last):
        n, h, w, c = last.shape
        return (dimg[:, None, None, :] / (h * w)).expand(n, h, w, c).contiguous()

    def _mask_grad(self, d, bits):
        d.mul_(bits)


@pytest.mark.parametrize("recompute", [True, False], ids=["recomputed_bn", "materialised_bn"])
def test_whole_step_orchestration_matches_oracle_autograd(torch_kernels, monkeypatch, recompute):  # noqa: F811
    """Both sides run in FLOAT64: at random init with a tiny batch the network is chaotic (single ReLU
    flips move trunk gradients by 1e-2 in fp32 -- the oracle's own noise floor, tests/test_train_gpu.py),
    which would hide an orchestration error; in fp64 the comparison is exact to ~1e-6 (the loss tail is
    fp32 on both sides, utils/math.py:52-60 Q7) and a wrong tensor, a missing mask or a swapped view is O(1)."""
    from rotmv_b200 import functional as RF

    gather = RF.rotate_gather          # the fusion stand-in; the engine hands it fp32 rotations
    monkeypatch.setattr(RF, "rotate_gather",
                        lambda feat, rot, dst, *a, **k: gather(feat, rot.to(feat.dtype), dst, *a, **k))
    torch.manual_seed(0)
    B, V = 2, 2
    ora = O.build_model(num_iter=2, depth=50, seed=0).train().double()
    images, pose, gt = O.synthetic_batch(B, V, seed=5, size=64)
    rot = O.pairwise_rotations(pose)
    images, rot, gt = images.double(), rot.double(), gt.double()
    loss_ref = O.iteration_loss(ora.forward_views(images, rot), [gt[:, 0], gt[:, 1]])
    loss_ref.backward()
    ref = {n: p.grad.clone() for n, p in ora.named_parameters() if p.grad is not None}

    model = FeatRotationSymm(50, 2)
    model.load_state_dict(O.build_model(num_iter=2, depth=50, seed=0).state_dict(), strict=True)
    model.train().double()
    eng = HostTrunkEngine(model, recompute=recompute)
    eng.dt = torch.float64
    seen = []
    out = eng.forward_backward(images, rot, gt, hook=lambda name, b, e: seen.append((name, len(eng.calls))))
    assert [s[0] for s in seen] == ["fusion", "layer4", "layer3", "layer2-stem"]      # bucket order of the backward
    assert seen[0][1] < seen[1][1] < seen[2][1] < seen[3][1]
    if recompute:   # conv3 + downsample of all 16 blocks: recomputed statistics/apply, masked data gradients
        assert eng.calls.count("conv_bn_fwd") == eng.calls.count("conv_bn_bwd") == 16 + 4
        assert eng.calls.count("dgrad+mask") == 15          # every block but the first hands a masked gradient down
    else:
        assert "conv_bn_fwd" not in eng.calls and "dgrad+mask" not in eng.calls
    assert abs(float(out["loss"]) - loss_ref.item()) <= 1e-5 * abs(loss_ref.item()), (float(out["loss"]), loss_ref.item())
    named = dict(model.named_parameters())
    rels = sorted((((eng.grads[id(named[n])] - g_ref).norm() / g_ref.norm().clamp_min(1e-30)).item(), n)
                  for n, g_ref in ref.items())
    print(f"{len(rels)} gradients: median relL2 {rels[len(rels) // 2][0]:.2e}, worst {rels[-1][0]:.2e} ({rels[-1][1]})")
    assert len(rels) > 150                                  # num_iter=2: one fuser/head pair less than the default model
    assert rels[-1][0] <= 1e-4, rels[-3:]
    assert int(model._feat_extractor[0].bn1.num_batches_tracked) == V
    assert not any(n.startswith("_feat_extractor.0.fc.") for n in ref)               # Q4
