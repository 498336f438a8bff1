"""Parity of the RECOMPUTED-BatchNorm kernels of the training step (rmv_conv_bn_stats,
rmv_conv_bn_bwd_reduce, the bn_mode 1 / 2 epilogues of rmv_conv2d_fwd, mask_bits, rmv_mask_bits)
against their torch formulas on the bf16-rounded operands, through the C ABI. These kernels replace
nn.BatchNorm2d (train mode) behind the expanding 1x1 convolutions of the bottlenecks
(models/resnet.py:122-123,139-146,229-230) without ever writing the conv output to HBM."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


# n_img, h, w, c_in, c_out, stride
CASES = [
    (6, 14, 14, 64, 256, 1),       # flattened, several images per 256-pixel tile
    (4, 56, 56, 64, 256, 1),       # layer1 shape: one view per tile almost everywhere
    (5, 7, 7, 512, 2048, 1),       # odd image count, ragged last tile, 16 channel tiles, K = 512
    (3, 9, 11, 128, 128, 1),       # ragged everything, a single channel tile
    (4, 28, 28, 256, 512, 2),      # stride-2 downsample: boxed tiles over the strided pixels
    (6, 14, 14, 1024, 2048, 2),
    (2, 13, 9, 64, 384, 2),        # odd extents with stride 2
]


def _inputs(case, seed=0):
    n, h, w, ci, co, s = case
    g = torch.Generator(device="cuda").manual_seed(1000 * seed + sum(case))
    x = torch.randn((n, h, w, ci), device="cuda", generator=g).bfloat16()
    wt = (torch.randn((co, 1, 1, ci), device="cuda", generator=g) / math.sqrt(ci)).bfloat16()
    z = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), stride=s)
    return g, x, wt, z.permute(0, 2, 3, 1).contiguous()        # z: [n, oh, ow, co] fp32


def _view_sums(t):
    """[2, C]: sum over the images of each view (image n -> view n & 1) and all pixels."""
    return torch.stack([t[v::2].double().sum(dim=(0, 1, 2)) for v in range(2)])


@pytest.mark.parametrize("case", CASES)
def test_conv_bn_stats_matches_torch(case):
    from rotmv_b200 import functional as RF

    _, x, wt, z = _inputs(case)
    co = case[4]
    acc = torch.zeros((2, co, 2), device="cuda", dtype=torch.float64)
    RF.conv_bn_stats(x, wt, acc, stride=case[5])
    RF.conv_bn_stats(x, wt, acc, stride=case[5])     # accumulates
    ref1, ref2 = _view_sums(z), _view_sums(z * z)
    scale = z.double().abs().sum(dim=(0, 1, 2)).max().item()
    assert (acc[:, :, 0] - 2 * ref1).abs().max().item() <= 2e-5 * scale
    assert (acc[:, :, 1] - 2 * ref2).abs().max().item() <= 2e-5 * ref2.max().item()


@pytest.mark.parametrize("case", CASES)
def test_conv_bn_bwd_reduce_matches_torch(case):
    from rotmv_b200 import functional as RF

    g, x, wt, z = _inputs(case, seed=1)
    co = case[4]
    dy = torch.randn(z.shape, device="cuda", generator=g).bfloat16()
    dy = dy * (torch.rand(z.shape, device="cuda", generator=g) > 0.4)      # a masked gradient
    mean = torch.randn((2, co), device="cuda", generator=g) * 0.1
    invstd = torch.rand((2, co), device="cuda", generator=g) + 0.5
    acc = torch.zeros((2, co, 2), device="cuda", dtype=torch.float64)
    RF.conv_bn_bwd_reduce(x, wt, dy, mean, invstd, acc, stride=case[5])
    d = dy.double()
    ref1 = _view_sums(d)
    xhat = torch.empty_like(z, dtype=torch.float64)
    for v in range(2):
        xhat[v::2] = (z[v::2].double() - mean[v].double()) * invstd[v].double()
    ref2 = _view_sums(d * xhat)
    s1 = d.abs().sum(dim=(0, 1, 2)).max().item()
    s2 = (d * xhat).abs().sum(dim=(0, 1, 2)).max().item()
    assert (acc[:, :, 0] - ref1).abs().max().item() <= 2e-5 * s1
    assert (acc[:, :, 1] - ref2).abs().max().item() <= 5e-5 * s2


def _bn_params(c, momentum=0.1, eps=1e-5):
    from rotmv_b200 import _lib as L

    t = {"ticket": torch.zeros((1,), device="cuda", dtype=torch.int32),
         "gamma": torch.rand((c,), device="cuda") + 0.5, "beta": torch.randn((c,), device="cuda"),
         "running_mean": torch.randn((c,), device="cuda") * 0.1, "running_var": torch.rand((c,), device="cuda") + 0.5,
         "num_batches": torch.zeros((), device="cuda", dtype=torch.int64)}
    for k in ("mean", "invstd", "a", "b", "k0", "k1", "k2"):
        t[k] = torch.full((2, c), float("nan"), device="cuda")
    for k in ("dgamma", "dbeta"):
        t[k] = torch.full((c,), float("nan"), device="cuda")
    p = L.BnParams()
    for k, v in t.items():
        setattr(p, k, v.data_ptr())
    p.eps, p.momentum = eps, momentum
    return p, t


@pytest.mark.parametrize("case", [CASES[0], CASES[1], CASES[4], (6, 7, 7, 512, 2048, 1)])
@pytest.mark.parametrize("via", ["recompute", "epilogue"])
def test_statistics_finalized_inside_the_launch(case, via):
    """The last thread block of the statistics launch produces the BatchNorm coefficients and the
    running statistics (two sequential view updates, unbiased variance, num_batches += 2: SURVEY Q1)
    -- against nn.BatchNorm2d semantics applied view by view; the accumulator is left zeroed.
    `recompute` = rmv_conv_bn_stats, `epilogue` = the STATS epilogue of rmv_conv2d_fwd."""
    from rotmv_b200 import functional as RF

    n, h, w, ci, co, s_ = case
    _, x, wt, z = _inputs(case, seed=4)
    torch.manual_seed(3)
    p, t = _bn_params(co)
    rm0, rv0 = t["running_mean"].clone(), t["running_var"].clone()
    acc = torch.zeros((2, co, 2), device="cuda", dtype=torch.float64)
    if via == "recompute":
        RF.conv_bn_stats(x, wt, acc, stride=s_, finalize=p)
        zz = z
    else:
        y = RF.conv2d(x, wt, stride=s_, stat_acc=acc, stat_views=2, stat_finalize=p)
        zz = y.float()                      # the epilogue statistics are those of the bf16 values stored
    torch.cuda.synchronize()
    assert int(t["ticket"]) == 0 and int(t["num_batches"]) == 2 and float(acc.abs().max()) == 0.0
    rm, rv = rm0.clone(), rv0.clone()
    for v in range(2):
        zv = zz[v::2].double().reshape(-1, co)
        mean, var = zv.mean(0), zv.var(0, unbiased=False)
        invstd = 1.0 / torch.sqrt(var + 1e-5)
        assert (t["mean"][v].double() - mean).abs().max().item() <= 1e-4 * max(mean.abs().max().item(), 1e-2)
        assert ((t["invstd"][v].double() - invstd) / invstd).abs().max().item() <= 1e-3
        a = t["gamma"].double() * invstd
        assert ((t["a"][v].double() - a) / a).abs().max().item() <= 1e-3
        assert (t["b"][v].double() - (t["beta"].double() - mean * a)).abs().max().item() <= 1e-3
        cnt = zv.shape[0]
        rm = 0.9 * rm + 0.1 * mean.float()
        rv = 0.9 * rv + 0.1 * (var * cnt / (cnt - 1)).float()
    assert (t["running_mean"] - rm).abs().max().item() <= 1e-4
    assert ((t["running_var"] - rv) / rv).abs().max().item() <= 1e-3


def test_bwd_reduce_finalized_inside_the_launch():
    """rmv_conv_bn_bwd_reduce with finalize: dgamma, dbeta and the (k0, k1, k2) of
    dz = k0*dy + k1*z + k2 against autograd through F.batch_norm applied per view."""
    from rotmv_b200 import functional as RF

    case = (6, 14, 14, 64, 256, 1)
    g, x, wt, z = _inputs(case, seed=5)
    co = case[4]
    torch.manual_seed(4)
    p, t = _bn_params(co)
    dy = torch.randn(z.shape, device="cuda", generator=g).bfloat16()
    zr = z.clone().requires_grad_(True)
    gamma = t["gamma"].clone().requires_grad_(True)
    beta = t["beta"].clone().requires_grad_(True)
    outs = [F.batch_norm(zr[v::2].permute(0, 3, 1, 2), None, None, gamma, beta, True, 0.1, 1e-5) for v in range(2)]
    loss = sum((o.permute(0, 2, 3, 1) * dy[v::2].float()).sum() for v, o in enumerate(outs))
    loss.backward()
    for v in range(2):
        zv = z[v::2].double().reshape(-1, co)
        t["mean"][v] = zv.mean(0).float()
        t["invstd"][v] = (1.0 / torch.sqrt(zv.var(0, unbiased=False) + 1e-5)).float()
    acc = torch.zeros((2, co, 2), device="cuda", dtype=torch.float64)
    RF.conv_bn_bwd_reduce(x, wt, dy, t["mean"], t["invstd"], acc, finalize=p)
    torch.cuda.synchronize()
    assert float(acc.abs().max()) == 0.0 and int(t["ticket"]) == 0
    assert (t["dbeta"] - beta.grad).abs().max().item() <= 2e-3 * beta.grad.abs().max().item()
    assert (t["dgamma"] - gamma.grad).abs().max().item() <= 2e-3 * gamma.grad.abs().max().item()
    dz = torch.empty_like(z)
    for v in range(2):
        dz[v::2] = t["k0"][v] * dy[v::2].float() + t["k1"][v] * z[v::2] + t["k2"][v]
    assert (dz - zr.grad).abs().max().item() <= 2e-3 * zr.grad.abs().max().item()


def _unpack_bits(bits, shape):
    b = bits.view(-1, 1).int()
    return ((b >> torch.arange(8, device=bits.device).view(1, 8)) & 1).bool().view(shape)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("use_res", [True, False])
def test_bn_mode1_forward_apply(case, use_res):
    """y = relu(a[v]*z + b[v] + residual) and the packed sign mask of the pre-ReLU value."""
    from rotmv_b200 import functional as RF

    g, x, wt, z = _inputs(case, seed=2)
    co = case[4]
    a = torch.rand((2, co), device="cuda", generator=g) + 0.5
    b = torch.randn((2, co), device="cuda", generator=g)
    res = torch.randn(z.shape, device="cuda", generator=g).bfloat16() if use_res else None
    bits = torch.zeros((z.numel() // 8,), device="cuda", dtype=torch.uint8)
    y = torch.full(z.shape, float("nan"), device="cuda", dtype=torch.bfloat16)
    RF.conv2d(x, wt, stride=case[5], residual=res, relu=True, out=y, bn_mode=1, bn_a=a, bn_b=b,
              bn_bits=bits)
    pre = torch.empty_like(z)
    for v in range(2):
        pre[v::2] = z[v::2] * a[v] + b[v]
    if use_res:
        pre = pre + res.float()
    ref = torch.relu(pre)
    err = (y.float() - ref).abs().max().item()
    assert err <= 1.2e-2 * ref.abs().max().item(), err
    got = _unpack_bits(bits, z.shape)
    want = pre > 0
    # the sign can only differ where the fp32 pre-activation is at rounding distance from zero
    bad = (got != want) & (pre.abs() > 1e-3 * pre.abs().max())
    assert not bad.any(), int(bad.sum())
    assert (got == want).float().mean().item() > 0.9999


@pytest.mark.parametrize("case", CASES)
def test_bn_mode2_backward_apply(case):
    """dz = k0[v]*dy + k1[v]*z + k2[v] with z recomputed in the kernel."""
    from rotmv_b200 import functional as RF

    g, x, wt, z = _inputs(case, seed=3)
    co = case[4]
    k0 = torch.rand((2, co), device="cuda", generator=g) + 0.5
    k1 = torch.randn((2, co), device="cuda", generator=g) * 0.3
    k2 = torch.randn((2, co), device="cuda", generator=g) * 0.1
    dy = torch.randn(z.shape, device="cuda", generator=g).bfloat16()
    dz = torch.full(z.shape, float("nan"), device="cuda", dtype=torch.bfloat16)
    RF.conv2d(x, wt, stride=case[5], residual=dy, out=dz, bn_mode=2, bn_a=k0, bn_b=k1, bn_c=k2)
    ref = torch.empty_like(z)
    for v in range(2):
        ref[v::2] = k0[v] * dy[v::2].float() + k1[v] * z[v::2] + k2[v]
    err = (dz.float() - ref).abs().max().item()
    assert err <= 1.2e-2 * ref.abs().max().item(), err


@pytest.mark.parametrize("case", [
    # n, h, ci (channels of x = dx), co (channels of y = dy), k, stride, pad
    (4, 14, 1024, 256, 1, 1, 0),     # data gradient of a reducing conv1: expanding, HBM-bound class
    (3, 28, 512, 128, 1, 1, 0),
    (4, 14, 128, 128, 3, 2, 1),      # stride 2: four parity classes with doubled strides / mask offsets
    (2, 15, 128, 64, 3, 2, 1),       # odd extent
    (3, 14, 256, 128, 3, 1, 1),
    (5, 7, 2048, 512, 1, 1, 0),      # 7x7: ragged last tile, 16 N tiles
])
def test_dgrad_mask_bits(case):
    """rmv_conv2d_dgrad with mask_bits: dx = (conv_transpose(dy, w) + residual) * mask, against
    autograd on the bf16-rounded operands."""
    from rotmv_b200 import functional as RF

    n, hh, ci, co, k, s_, p_ = case
    torch.manual_seed(sum(case))
    x = torch.randn((n, ci, hh, hh), device="cuda", requires_grad=True)
    w = (torch.randn((co, ci, k, k), device="cuda") / (k * k * co) ** 0.5).bfloat16().float().requires_grad_(True)
    y = F.conv2d(x, w, stride=s_, padding=p_)
    dy = torch.randn_like(y).bfloat16().float()
    y.backward(dy)
    res = torch.randn((n, hh, hh, ci), device="cuda").bfloat16()
    keep = torch.rand((n, hh, hh, ci), device="cuda") > 0.5
    ref = (x.grad.permute(0, 2, 3, 1) + res.float()) * keep
    bits = (keep.reshape(-1, 8).int() * (1 << torch.arange(8, device="cuda")).view(1, 8)).sum(1).to(torch.uint8)
    wt = w.detach().permute(1, 2, 3, 0).flip(1, 2).contiguous().bfloat16()   # [C, kh, kw, K], taps reversed
    dx = RF.conv2d_dgrad(dy.permute(0, 2, 3, 1).contiguous().bfloat16(), wt, stride=s_, pad=p_,
                         in_hw=(hh, hh), residual=res, mask_bits=bits)
    err = (dx.float() - ref).abs().max().item()
    assert err <= 1e-2 * ref.abs().max().item() + 1e-3, (case, err)
    assert (dx.float()[~keep] == 0).all()


@pytest.mark.parametrize("case", [
    # n, h, ci (channels of dx = the mid tensor), co (channels of dy_out), k, pad
    (4, 56, 64, 256, 1, 0),        # dgrad of layer1's conv3: 64-wide tiles
    (6, 14, 256, 1024, 1, 0),      # dgrad of layer3's conv3
    (4, 28, 128, 128, 3, 1),       # dgrad of a 3x3 conv2 (boxed tiles, ragged rows masked by validity)
    (4, 7, 512, 512, 3, 1),        # 7x7: several images per tile
])
def test_dgrad_bn_mode4_masked_gradient_and_backward_reduction(case):
    """bn_mode 4 of rmv_conv2d_dgrad: dx = conv_transpose(dy, w) * mask is stored and, in the same
    launch, reduced for the BatchNorm backward of the layer below (sum dx, sum dx*xhat per view and
    channel) and finalized (dgamma, dbeta, k0, k1, k2) -- against autograd + the torch formulas."""
    from rotmv_b200 import functional as RF

    n, hh, ci, co, k, p_ = case
    torch.manual_seed(sum(case) + 5)
    x = torch.randn((n, ci, hh, hh), device="cuda", requires_grad=True)
    w = (torch.randn((co, ci, k, k), device="cuda") / (k * k * co) ** 0.5).bfloat16().float().requires_grad_(True)
    y = F.conv2d(x, w, stride=1, padding=p_)
    dy = torch.randn_like(y).bfloat16().float()
    y.backward(dy)
    keep = torch.rand((n, hh, hh, ci), device="cuda") > 0.4
    ref_dx = x.grad.permute(0, 2, 3, 1) * keep                     # fp32, masked
    bits = (keep.reshape(-1, 8).int() * (1 << torch.arange(8, device="cuda")).view(1, 8)).sum(1).to(torch.uint8)
    z = torch.randn((n, hh, hh, ci), device="cuda").bfloat16()     # pre-BatchNorm output of the layer below
    pr, t = _bn_params(ci)
    for v in range(2):
        zv = z[v::2].double().reshape(-1, ci)
        t["mean"][v] = zv.mean(0).float()
        t["invstd"][v] = (1.0 / torch.sqrt(zv.var(0, unbiased=False) + 1e-5)).float()
    acc = torch.zeros((2, ci, 2), device="cuda", dtype=torch.float64)
    wt = w.detach().permute(1, 2, 3, 0).flip(1, 2).contiguous().bfloat16()
    dx = RF.conv2d_dgrad(dy.permute(0, 2, 3, 1).contiguous().bfloat16(), wt, stride=1, pad=p_, in_hw=(hh, hh),
                         mask_bits=bits, bwd_bn={"z": z, "mean": t["mean"], "invstd": t["invstd"], "acc": acc,
                                                 "finalize": pr})
    torch.cuda.synchronize()
    err = (dx.float() - ref_dx).abs().max().item()
    assert err <= 1e-2 * ref_dx.abs().max().item() + 1e-3, (case, err)
    assert (dx.float()[~keep] == 0).all()
    assert float(acc.abs().max()) == 0.0 and int(t["ticket"]) == 0          # finalized and reset
    # reductions of the STORED (bf16) gradient, as the apply pass will read it
    d = dx.double()
    s1 = torch.stack([d[v::2].sum(dim=(0, 1, 2)) for v in range(2)])
    xhat = torch.empty_like(d)
    for v in range(2):
        xhat[v::2] = (z[v::2].double() - t["mean"][v].double()) * t["invstd"][v].double()
    s2 = torch.stack([(d * xhat)[v::2].sum(dim=(0, 1, 2)) for v in range(2)])
    scale1 = d.abs().sum(dim=(0, 1, 2)).max().item()
    assert (t["dbeta"].double() - s1.sum(0)).abs().max().item() <= 1e-4 * scale1
    assert (t["dgamma"].double() - s2.sum(0)).abs().max().item() <= 2e-4 * (d * xhat).abs().sum(dim=(0, 1, 2)).max().item()
    count = (n // 2) * hh * hh                                             # the library's count per view (even n)
    if n % 2 == 0:
        k0 = t["gamma"].double() * t["invstd"].double()
        k1 = -k0 * t["invstd"].double() * s2 / count
        k2 = -k0 * s1 / count - k1 * t["mean"].double()
        assert ((t["k0"].double() - k0) / k0).abs().max().item() <= 1e-5
        assert (t["k1"].double() - k1).abs().max().item() <= 2e-4 * k1.abs().max().item() + 1e-9
        assert (t["k2"].double() - k2).abs().max().item() <= 2e-4 * k2.abs().max().item() + 1e-9


def test_mask_bits_kernel():
    from rotmv_b200 import _lib as L
    from rotmv_b200 import functional as RF

    g = torch.Generator(device="cuda").manual_seed(5)
    src = torch.randn((3, 7, 7, 2048), device="cuda", generator=g).bfloat16()
    keep = torch.rand(src.shape, device="cuda", generator=g) > 0.3
    bits = (keep.view(-1, 8).int() * (1 << torch.arange(8, device="cuda")).view(1, 8)).sum(1).to(torch.uint8)
    dst = src.clone()
    RF._call("rmv_mask_bits", {}, L.load().rmv_mask_bits, dst.data_ptr(), bits.data_ptr(), dst.data_ptr(),
             dst.numel(), L.BF16, L.stream_ptr())
    assert torch.equal(dst, src * keep)


def test_recompute_step_matches_materialised_step():
    """One bf16 training step (B=6, V=2, ResNet-50) with the recomputed BatchNorm against the same
    step with RMV_BN_RECOMPUTE=0 (conv output written, HBM-bound passes), both judged against the
    fp32 (FFMA) engine's gradients. The random-init small-batch network is chaotic in its inputs
    (tests/test_train_gpu.py header: the reference's own bf16-autocast gradients have cosine 0.12 at
    the stem against its fp32 gradients), so two bf16 paths that round z at different places
    decorrelate in the early layers; what must hold is that the recomputed path is AS CLOSE to fp32
    as the materialised one, tensor by tensor on aggregate, that both agree near the loss, and that
    the integer bookkeeping is identical."""
    from oracle import rotmv_oracle as O
    from rotmv_b200.module import FeatRotationSymm
    from rotmv_b200.train import TrainEngine

    ora = O.build_model(num_iter=3, depth=50, seed=0)
    sd0 = {k: v.clone() for k, v in ora.state_dict().items()}
    images, pose, gt = O.synthetic_batch(6, 2, seed=11)
    rot = O.pairwise_rotations(pose)
    out = {}
    for mode in ("recompute", "materialised", "fp32"):
        model = FeatRotationSymm(50, 3)
        model.load_state_dict(sd0, strict=True)
        model = model.cuda().train()
        eng = TrainEngine(model, precision="fp32" if mode == "fp32" else "bf16", lr=1e-3, weight_decay=1e-6)
        eng.recompute_bn = mode == "recompute"
        eng.recompute_max_cin = 4096      # every bottleneck's conv3 / downsample, not only layer1-3
        eng.fuse_bn_bwd = mode == "recompute"   # and the mid-layer backward reductions inside the data gradients
        eng.forward_backward(images.cuda(), rot.cuda(), gt.cuda())
        torch.cuda.synchronize()
        named = dict(model.named_parameters())
        out[mode] = (eng.loss.item(), {n: eng.grads[id(p)].clone() for n, p in named.items() if id(p) in eng.grads},
                     {n: b.clone() for n, b in model.named_buffers()})
    (l1, g1, b1), (l0, g0, b0), (lf, gf, _) = out["recompute"], out["materialised"], out["fp32"]

    def cos(a, b):
        return F.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0).item()

    c1 = {n: cos(g1[n], gf[n]) for n in gf}
    c0 = {n: cos(g0[n], gf[n]) for n in gf}
    groups = ["conv1", "bn1", "layer1", "layer2", "layer3", "layer4", "_lifter", "_img_fusers", "_gaze_estimators"]
    print(f"loss: recompute {l1:.5f} materialised {l0:.5f} fp32 {lf:.5f}")
    for gname in groups:
        ns = [n for n in gf if (n.startswith("_feat_extractor.0." + gname) or n.startswith(gname))]
        if ns:
            m1 = sum(c1[n] for n in ns) / len(ns)
            m0 = sum(c0[n] for n in ns) / len(ns)
            print(f"  mean cosine vs fp32, {gname:18s} ({len(ns):3d} tensors): recompute {m1:.4f} materialised {m0:.4f}")
    assert abs(l1 - lf) <= 3e-2 * abs(lf) and abs(l0 - lf) <= 3e-2 * abs(lf), (l1, l0, lf)
    mean1, mean0 = sum(c1.values()) / len(c1), sum(c0.values()) / len(c0)
    assert mean1 >= mean0 - 0.05, (mean1, mean0)
    for gname in ("layer4", "_lifter", "_img_fusers", "_gaze_estimators"):
        ns = [n for n in gf if (n.startswith("_feat_extractor.0." + gname) or n.startswith(gname))]
        m1 = sum(c1[n] for n in ns) / len(ns)
        m0 = sum(c0[n] for n in ns) / len(ns)
        assert m1 >= m0 - 0.05, (gname, m1, m0)
    heads = [n for n in c1 if n.startswith("_gaze_estimators.2")]
    # near the loss both bf16 paths agree with fp32 (measured 0.975-0.995 for either; run-to-run noise of
    # the fp32 atomics is +-0.005): the recomputed path within 0.02 of the materialised one, tensor by tensor
    assert all(c1[n] > 0.95 and c1[n] >= c0[n] - 0.02 for n in heads), [(n, c1[n], c0[n]) for n in heads]
    for n in b0:
        if n.endswith("num_batches_tracked"):
            assert torch.equal(b0[n], b1[n]), n
        elif n.endswith("running_mean") and "_feat_extractor" in n and ("bn3" in n or "downsample" in n):
            # running statistics of the recomputed layers: same batch means up to the bf16 rounding of z
            assert (b0[n] - b1[n]).abs().max().item() <= 5e-2 * max(b0[n].abs().max().item(), 1e-2), n
