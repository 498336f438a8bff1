"""Host logic of the training fusion stage (rotmv_b200/train.py: `_fusion_default`,
`_fusion_encode_rotmat`, `_fusion_share_feature`) checked on the CPU: the C-ABI kernel entry points
the stage calls (Linear, weight gradient, column sum, ReLU mask/add, rotate-gather, pool, head+loss
forward/backward) are replaced by their torch formulas, so that what is tested is the orchestration
itself -- buffer layouts, zero padding of the 3593-wide layers, the [3][2][512] interleave of
RotFeatFuser, the train-mode IntensityBatchNorm sequence, gradient accumulation across iterations
and across aliased (share_weights) modules -- against the CPU oracle's autograd on the same
weights. The kernels themselves are tested against torch in tests/test_ops_gpu.py and
tests/test_train_gpu.py; the same configurations run end to end on the GPU in
tests/test_train_gpu.py::test_depth18_step_fp32_vs_live_oracle_and_bf16.
"""
import math

import pytest
import torch

from oracle import rotmv_oracle as O
from rotmv_b200 import functional as RF
from rotmv_b200 import train as T
from rotmv_b200.module import FeatRotationSymm


class HostEngine(T.TrainEngine):
    """TrainEngine with torch stand-ins for the kernels (fp32, CPU)."""

    def __init__(self, model):
        self.model, self.precision, self.dt, self.dtc = model, "fp32", torch.float32, 0
        self.device = torch.device("cpu")
        self.num_iter, self.fc_dim, self.nvec = model._num_iter, model._fc_dim, model._num_feat_vec
        self.apply_rot = not model._ignore_rotmat
        self.encode_rot = bool(model._encode_rotmat) and not model._ignore_rotmat
        self.share_feat = bool(model._share_feature)
        self.grads = {id(p): torch.zeros_like(p) for p in model.parameters()}
        lif = model._lifter._lifter.blocks
        self.lift = [lif[0][0], lif[1][0]]
        self.fusers = [[blk[0] for blk in m._fuser.blocks] for m in model._img_fusers]
        self.heads = [[m.blocks[0][0], m.blocks[1][0]] for m in model._gaze_estimators]
        self._bufs = {}
        self.loss = torch.zeros((1,))
        self.use_tc_wgrad = False

    def _lin_fwd(self, lin, tag):
        return lin.weight.detach()

    def _lin_t(self, lin, tag):
        return lin.weight.detach().t().contiguous()

    def _wgrad(self, x, dy, lin, kh, kw, stride, pad, x_strides=None, grad=None):
        g = grad if grad is not None else self.grads[id(lin.weight)]
        g += dy.reshape(-1, dy.shape[-1]).t() @ x.reshape(-1, x.shape[-1])

    def _colsum(self, dy, n, dst):
        dst[:n] += dy[:, :n].sum(0)

    def _add(self, src, dst, mask=None, add=None):
        out = src if mask is None else torch.where(mask > 0, src, torch.zeros_like(src))
        dst.copy_(out if add is None else out + add)

    def _head_bwd(self, pred, gt_flat, g, h2, scale, views, cfg, dg, dpred):
        p = pred.detach().clone().requires_grad_(True)
        _head_loss_value(p, gt_flat, scale, views, cfg["reference_decay"]).backward()
        dpred.copy_(p.grad)
        dp = p.grad.to(g.dtype)
        dg.copy_((g > 0).to(g.dtype) * (dp @ h2.weight.detach()))
        self.grads[id(h2.weight)] += dp.t() @ g
        self.grads[id(h2.bias)] += dp.sum(0)


    # the re-layout kernels of the constructor variants (csrc/variant_glue.cu) as torch formulas
    def _zero(self, t):
        t.zero_()

    def _scopy(self, src, dst, scale=None, accumulate=False):
        """rmv_strided_copy: dst (+)= src * scale[last dim] over equally shaped strided views."""
        assert tuple(src.shape) == tuple(dst.shape) and src.dim() <= 3
        val = src.to(torch.float64) if scale is None else src.to(torch.float64) * scale.to(torch.float64)
        if accumulate:
            val = val + dst.to(torch.float64)
        dst.copy_(val.to(dst.dtype))

    def _intensity_scale(self, bn, feat_bv, out):
        """rmv_intensity_bn_train = train-mode IntensityBatchNorm of models/rot_mv.py:13-32."""
        intensity = feat_bv.float().norm(dim=-2, keepdim=True)                      # [B, 1, 512]
        var = intensity.var(dim=0, unbiased=False, keepdim=True)
        std = var.clamp_min(bn.eps).sqrt()
        bn.running_mean.copy_(bn.running_mean * (1 - bn.momentum) + std * bn.momentum)
        out.copy_(1.0 / (bn.running_mean.reshape(-1) + bn.eps))
        return out

    def _head_bwd_ext(self, dpred, g, h2, dg):
        """rmv_head_loss_bwd with gt == NULL (external d(loss)/d(pred)): Linear(512,2) + ReLU backward."""
        dp = dpred.detach().to(g.dtype).reshape(g.shape[0], 2)
        dg.copy_((g > 0).to(g.dtype) * (dp @ h2.weight.detach()))
        self.grads[id(h2.weight)] += dp.t() @ g
        self.grads[id(h2.bias)] += dp.sum(0)


def _head_loss_value(pred, gt, scale, views, aux_decay):
    """scale * sum over rows of w_row * angular(pred, gt) in degrees (losses/gaze_loss.py:42-52)."""
    a, b = O.pitchyaw_to_vector(gt), O.pitchyaw_to_vector(pred)
    sim = torch.nn.functional.hardtanh(torch.nn.functional.cosine_similarity(a, b, eps=1e-6), -1.0, 1.0)
    w = torch.where(torch.arange(pred.shape[0]) % views == 0, 1.0, float(aux_decay))
    return scale * (torch.acos(sim) * (180 / math.pi) * w).sum()


@pytest.fixture()
def torch_kernels(monkeypatch):
    def linear(x, w, bias=None, *, relu=False, out=None, **_):
        y = x @ w.t()
        if bias is not None:
            y = y + bias.detach()
        out.copy_(torch.relu(y) if relu else y)
        return out

    def avgpool(x, out0, out1=None):
        f = x.mean(dim=(1, 2))
        out0[:, :f.shape[1]] = f
        if out1 is not None:
            out1[:, :f.shape[1]] = f

    def rotate_gather(feat, rot, dst, batch, views, nvec=512, apply_rot=True, transpose=False):
        f = feat.reshape(batch, views, 3, nvec)
        out = torch.zeros_like(f)
        eye = torch.eye(3)
        for v in range(views):
            for u in range(views):
                if u == v:
                    continue
                r = (rot[:, u, v].transpose(-1, -2) if transpose else rot[:, v, u]) if apply_rot else eye
                out[:, v] += r @ f[:, u]
        dst.copy_((out / (views - 1)).reshape(batch * views, 3 * nvec))

    def head_loss(hidden, w2, b2, pred, gt=None, loss_scale=0.0, loss_out=None, views=1, aux_decay=1.0):
        pred.copy_(hidden @ w2.t() + b2)
        if gt is not None:
            loss_out += _head_loss_value(pred, gt, loss_scale, views, aux_decay).detach()

    for name, fn in dict(linear=linear, avgpool=avgpool, rotate_gather=rotate_gather,
                         head_loss=head_loss).items():
        monkeypatch.setattr(RF, name, fn)


FLAGS = [dict(), dict(share_weights=True), dict(ignore_rotmat=True), dict(encode_rotmat=True),
         dict(share_feature=True), dict(encode_rotmat=True, share_weights=True)]


IDS = ["default", "share_weights", "ignore_rotmat", "encode_rotmat", "share_feature", "encode_rotmat+share_weights"]
# every flag set with the fused loss; the external-d(pred) path (autograd bridge) once per fusion method
CASES = [pytest.param(f, False, id=i + "-fused_loss") for f, i in zip(FLAGS, IDS)] + \
        [pytest.param(FLAGS[k], True, id=IDS[k] + "-external_dpred") for k in (0, 3, 4)]


@pytest.mark.parametrize("flags,external", CASES)
def test_fusion_stage_host_logic_matches_oracle_autograd(torch_kernels, flags, external):
    b, v, n_it = 5, 2, 2
    ora = O.build_model(num_iter=n_it, depth=18, seed=0, **flags).train()
    model = FeatRotationSymm(18, n_it, **flags)
    model.load_state_dict(ora.state_dict(), strict=True)
    model.train()
    g = torch.Generator().manual_seed(5)
    trunk_out = torch.rand((b * v, 3, 3, 512), generator=g)               # NHWC output of the last block
    pose = torch.rand((b, v, 2), generator=g) - 0.5
    gt = torch.rand((b, v, 2), generator=g) - 0.5
    rot = O.pairwise_rotations(pose)
    # oracle: the same pooled features enter in place of the trunk
    feats = trunk_out.mean(dim=(1, 2)).reshape(b, v, 512).clone().requires_grad_(True)
    ora._feat_extractor = torch.nn.Identity()
    out = ora.forward_views(feats, rot)
    loss_ref = O.iteration_loss(out, [gt[:, k] for k in range(v)])
    loss_ref.backward()

    eng = HostEngine(model)
    cfg = model.loss_cfg
    gt_flat = gt.reshape(b * v, 2).contiguous()

    def stage(gt_arg):
        if eng.encode_rot:
            return eng._fusion_encode_rotmat(trunk_out, rot, gt_arg, b, cfg)
        if eng.share_feat:
            return eng._fusion_share_feature(trunk_out, rot, gt_arg, b, cfg)
        return eng._fusion_default(trunk_out, rot, gt_arg, b, v, cfg)

    gen = stage(None if external else gt_flat)
    fwd_preds = next(gen)                       # forward (+ fused loss when the labels are given)
    ext = None
    if external:
        # the autograd bridge: the caller's own loss objects (here the oracle's IterationLoss
        # restatement) turn the predictions into d(loss)/d(pred) per iteration
        leaves = [p.detach().clone().requires_grad_(True) for p in fwd_preds]
        fake = {"num_iter": n_it}
        for i, p in enumerate(leaves):
            fake[f"iter_{i}"] = {f"pred_gaze_{k}": p.view(b, v, 2)[:, k] for k in range(v)}
        loss_ext = O.iteration_loss(fake, [gt[:, k] for k in range(v)])
        loss_ext.backward()
        ext = [p.grad for p in leaves]
        assert abs(loss_ext.item() - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())
    try:
        gen.send(ext)
        raise AssertionError("the fusion stage must finish after the backward")
    except StopIteration as done:
        dimg, preds = done.value
    if not external:
        assert abs(eng.loss.item() - loss_ref.item()) <= 1e-5 * abs(loss_ref.item()), (eng.loss.item(), loss_ref.item())
    for k in range(v):
        want = out[f"iter_{n_it - 1}"][f"pred_gaze_{k}"].detach()
        assert torch.allclose(preds[-1].view(b, v, 2)[:, k], want, rtol=1e-4, atol=1e-5)
    want = feats.grad.reshape(b * v, 512)
    assert torch.allclose(dimg, want, rtol=1e-3, atol=1e-6 * want.abs().max().item() + 1e-9), \
        (dimg - want).abs().max().item()
    named = dict(model.named_parameters())
    checked = 0
    for n, p in ora.named_parameters():
        if p.grad is None or n.startswith("_feat_extractor"):
            continue
        got = eng.grads[id(named[n])]
        tol = 1e-3 * p.grad.abs().max().item() + 1e-9
        assert (got - p.grad).abs().max().item() <= tol, (n, (got - p.grad).abs().max().item(), tol)
        checked += 1
    assert checked >= 8
    if flags.get("share_feature"):
        for i in range(n_it):
            got = model._img_fusers[i]._batchnorm.running_mean
            want = ora._img_fusers[i]._batchnorm.running_mean
            assert torch.allclose(got, want, rtol=1e-5, atol=1e-7), i


def test_autograd_bridge_plumbing_with_a_stub_engine():
    """`module._TrainStepFunction` alone (CPU, stub engine): predictions come back connected to
    autograd, the caller's d(loss)/d(pred) reaches the engine's backward half, parameter gradients
    are returned in parameter order (None for tensors the engine has no gradient for, e.g. fc.*),
    an output that does not take part in the loss arrives as zeros, and a second backward is refused."""
    from rotmv_b200.module import _TrainStepFunction

    w0 = torch.nn.Parameter(torch.tensor([[1.0, 2.0]]))
    w1 = torch.nn.Parameter(torch.tensor([[-1.0, 0.5]]))
    unused = torch.nn.Parameter(torch.ones(3))
    x = torch.arange(4.0).view(4, 1)                      # stands for the images

    class StubEngine:
        device = torch.device("cpu")

        def __init__(self):
            self.grads = {id(w0): torch.zeros_like(w0), id(w1): torch.zeros_like(w1)}
            self.seen = None

        def _fwd_bwd(self, images, rotations, gt, hook=None):
            assert gt is None
            preds = [images @ w0.detach(), images @ w1.detach()]          # two "iterations", [4, 2] each
            ext = yield preds
            self.seen = [e.clone() for e in ext]
            for w, e in zip((w0, w1), ext):
                self.grads[id(w)].copy_(images.t() @ e)
            return {}

    eng = StubEngine()
    p0, p1 = _TrainStepFunction.apply(eng, x, None, w0, unused, w1)
    assert p0.requires_grad and p1.requires_grad and torch.equal(p0, x @ w0.detach())
    loss = (p0 * torch.tensor([1.0, -2.0])).sum()         # p1 does not take part
    loss.backward()
    assert torch.equal(eng.seen[0], torch.tensor([1.0, -2.0]).expand(4, 2)) and eng.seen[1].abs().sum() == 0
    assert torch.equal(w0.grad, x.t() @ eng.seen[0]) and torch.equal(w1.grad, torch.zeros_like(w1))
    assert unused.grad is None
    q0, _ = _TrainStepFunction.apply(eng, x, None, w0, unused, w1)
    q0.sum().backward(retain_graph=True)
    with pytest.raises(RuntimeError):
        q0.sum().backward()
