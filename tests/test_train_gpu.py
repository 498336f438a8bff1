"""Step parity on the GPU (SURVEY 4 "Step parity"): loss, gradients, parameters after Adam steps,
BatchNorm running statistics and num_batches_tracked of rotmv_b200.train.TrainEngine against the
golden outputs of the UNMODIFIED reference driven by torch.optim.Adam(weight_decay=1e-6)
(tests/golden, oracle/make_golden.py: B=8, V=2, lr=1e-3, two steps).

fp32 engine tolerances: loss rtol 1e-4; gradient NORMS of all 187 trained tensors within 2e-2
(median within 1e-3). Element-wise gradient tolerances are set by the reference's own noise floor:
at random init with B=8 the network is chaotic, and the CPU reference's gradients move by
relL2 = 9.8e-3 (conv1.weight), 1.2e-2 (bn1.weight), 4.9e-3 (layer4.2.conv3.weight),
3.5e-3 (lifter bias), 2.6e-6 (last head weight) when nothing but its thread count changes (1 vs 8
threads, measured with the oracle in the build container). We allow 5x that floor for trunk
tensors and 1e-3 for the heads. For bf16 the reference's OWN torch.autocast(bfloat16) gradients
have cosine 0.12 (conv1.weight) ... 0.56 (layer4.2.conv3.weight) ... 0.89 (lifter bias) ... 0.998
(last head weight) against its fp32 gradients on these inputs, so the bf16 engine is only required
to match that profile (see test_bf16_step_close_to_fp32_reference).
"""

# relL2 of the reference's fp32 gradient between 1 and 8 CPU threads (noise floor), and cosine of
# the reference's bf16-autocast gradient vs its fp32 gradient, per checked tensor
REF_NOISE = {
    "_feat_extractor.0.conv1.weight": (9.8e-3, 0.1246),
    "_feat_extractor.0.bn1.weight": (1.2e-2, 0.1210),
    "_feat_extractor.0.layer1.0.conv1.weight": (1.0e-2, 0.1077),
    "_feat_extractor.0.layer4.2.bn3.bias": (5e-3, 0.5),
    "_feat_extractor.0.layer4.2.conv3.weight": (4.9e-3, 0.5551),
    "_lifter._lifter.blocks.1.0.bias": (3.5e-3, 0.8906),
    "_img_fusers.2._fuser.blocks.1.0.bias": (2e-3, 0.9),
    "_gaze_estimators.2.blocks.1.0.weight": (2.6e-6, 0.9983),
    "_gaze_estimators.0.blocks.1.0.bias": (1e-4, 0.99),
}
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "rotmv_r50_b8v2.npz")


def _setup(precision):
    from oracle import rotmv_oracle as O
    from rotmv_b200.module import FeatRotationSymm
    from rotmv_b200.train import TrainEngine

    ora = O.build_model(num_iter=3, depth=50, seed=0)
    model = FeatRotationSymm(50, 3)
    model.load_state_dict(ora.state_dict(), strict=True)
    model = model.cuda().train()
    gold = np.load(GOLD)
    eng = TrainEngine(model, precision=precision, lr=float(gold["step_lr"]), weight_decay=1e-6)
    images, pose, gt = O.synthetic_batch(8, 2, seed=1)
    rot = O.pairwise_rotations(pose)
    return O, ora, model, eng, gold, images.cuda(), rot.cuda(), gt.cuda()


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_fp32_step_parity_vs_reference_golden():
    O, ora, model, eng, gold, images, rot, gt = _setup("fp32")
    init = {k: v.clone() for k, v in model.state_dict().items()}
    out = eng.forward_backward(images, rot, gt)
    loss0 = out["loss"].item()
    assert abs(loss0 - float(gold["train_loss_0"])) <= 1e-4 * float(gold["train_loss_0"]), loss0
    # predictions of the train-mode forward
    for i in range(3):
        p = out["preds"][i].view(8, 2, 2).cpu()
        for v in range(2):
            ref = torch.tensor(gold[f"train_iter{i}_pred_gaze_{v}"])
            assert torch.allclose(p[:, v], ref, rtol=1e-3, atol=1e-3 * ref.abs().max().item()), (i, v)
    named = dict(model.named_parameters())
    # every trained tensor: gradient norm against the reference's
    norms = gold["grad0_norms"]
    ref_names = [n for n, _ in ora.named_parameters()]
    errs = []
    for n, ref_norm in zip(ref_names, norms):
        if ref_norm < 0:  # fc.*: no gradient in the reference either
            assert n.startswith("_feat_extractor.0.fc.")
            continue
        g = eng.grads[id(named[n])]
        got = g.double().norm().item()
        errs.append((abs(got - ref_norm) / max(ref_norm, 1e-12), n, got, ref_norm))
    errs.sort(reverse=True)
    print("worst gradient-norm mismatches:", [(f"{e:.2e}", n) for e, n, _, _ in errs[:5]],
          "median %.2e" % errs[len(errs) // 2][0])
    assert len(errs) == 187
    assert errs[0][0] <= 2e-2, errs[:3]            # every one of the 187 trained tensors
    assert errs[len(errs) // 10][0] <= 5e-3, errs[len(errs) // 10]   # 90 % of them within 0.5 %
    assert errs[len(errs) // 2][0] <= 1e-3, errs[len(errs) // 2]
    # selected tensors element-wise
    for key in [k for k in gold.files if k.startswith("grad0::")]:
        n = key[len("grad0::"):]
        g = eng.grads[id(named[n])].cpu()
        ref = torch.tensor(gold[key])
        got = g if g.numel() == ref.numel() else g.flatten()[:ref.numel()]
        err, tol = rel_l2(got.reshape(ref.shape), ref), max(5 * REF_NOISE[n][0], 1e-3)
        print(f"grad {n}: relL2 {err:.2e} (tol {tol:.1e})")
        assert err <= tol, (n, err, tol)
    # two optimisation steps (forward_backward above did not touch the parameters)
    eng.flat_m.zero_(); eng.flat_v.zero_()
    for bn in [m for m in model.modules() if isinstance(m, torch.nn.BatchNorm2d)]:
        bn.running_mean.copy_(init_state(bn, model, init, "running_mean"))
        bn.running_var.copy_(init_state(bn, model, init, "running_var"))
        bn.num_batches_tracked.zero_()
    l0 = eng.step(images, rot, gt).item()
    l1 = eng.step(images, rot, gt).item()
    assert abs(l0 - float(gold["train_loss_0"])) <= 1e-4 * float(gold["train_loss_0"])
    assert abs(l1 - float(gold["train_loss_1"])) <= 2e-2 * float(gold["train_loss_1"]), (l1, float(gold["train_loss_1"]))
    sd = model.state_dict()
    assert int(sd["_feat_extractor.0.bn1.num_batches_tracked"]) == int(gold["after2_num_batches_tracked"]) == 4
    assert np.allclose(sd["_feat_extractor.0.bn1.running_mean"].cpu().numpy(), gold["after2_bn1_running_mean"],
                       rtol=1e-3, atol=1e-5)
    assert np.allclose(sd["_feat_extractor.0.bn1.running_var"].cpu().numpy(), gold["after2_bn1_running_var"],
                       rtol=1e-3, atol=1e-5)
    # parameter movement: per-tensor norm of (param - init) vs the reference's
    keys = list(ora.state_dict().keys())
    deltas = gold["after2_delta_norms"]
    bad = []
    for k, ref_d in zip(keys, deltas):
        if "num_batches_tracked" in k or "running_" in k:
            continue
        got = (sd[k].double() - init[k].double().cuda()).norm().item()
        # the second step's gradient is taken at a point where the loss has jumped 0.97 -> 2.3 and
        # inherits the chaotic trunk noise documented above; small BN tensors move by up to ~7 %
        if abs(got - ref_d) > 0.15 * max(ref_d, 1e-9) + 1e-7:
            bad.append((k, got, ref_d))
    assert not bad, bad[:5]
    # element-wise: error of the updated head weight small against the distance it moved
    hk = "_gaze_estimators.2.blocks.1.0.weight"
    ref_w = torch.tensor(gold["after2_head2_w"]).double()
    moved = (ref_w - init[hk].double().cpu()).norm().item()
    assert (sd[hk].double().cpu() - ref_w).norm().item() <= 0.2 * moved
    # fc.* untouched (no gradient, no weight decay applied by Adam when grad is None)
    assert torch.equal(sd["_feat_extractor.0.fc.weight"].cpu(), init["_feat_extractor.0.fc.weight"].cpu())


def init_state(bn, model, init, name):
    for k, m in model.named_modules():
        if m is bn:
            return init[f"{k}.{name}"]
    raise KeyError


def test_bf16_step_close_to_fp32_reference():
    """bf16 engine (tcgen05 forward + data gradients): loss within 2 % of the fp32 reference (the
    reference's own bf16 autocast: 0.65 %), and the cosine of every checked gradient against the fp32
    reference no worse than 0.6x the cosine the reference's own bf16-autocast gradient reaches where
    that cosine is meaningful (>= 0.5: layer4, lifter, fusers, heads); in the chaotic trunk (reference
    bf16 cosine ~0.1) only the gradient magnitude is checked (norm ratio within [0.5, 2])."""
    O, ora, model, eng, gold, images, rot, gt = _setup("bf16")
    out = eng.forward_backward(images, rot, gt)
    loss0 = out["loss"].item()
    assert abs(loss0 - float(gold["train_loss_0"])) <= 2e-2 * float(gold["train_loss_0"]), loss0
    named = dict(model.named_parameters())
    for key in [k for k in gold.files if k.startswith("grad0::")]:
        n = key[len("grad0::"):]
        g = eng.grads[id(named[n])].cpu().flatten().double()
        ref = torch.tensor(gold[key]).flatten().double()
        g = g[:ref.numel()]
        cos = (g @ ref / (g.norm() * ref.norm())).item()
        ratio = (g.norm() / ref.norm()).item()
        print(f"bf16 grad {n}: cos {cos:.4f} (reference's own bf16 autocast: {REF_NOISE[n][1]:.4f}), "
              f"norm ratio {ratio:.3f}")
        if REF_NOISE[n][1] >= 0.5:
            assert cos >= 0.6 * REF_NOISE[n][1], (n, cos)
        else:
            # chaotic regime (the reference's own bf16 gradient has cosine ~0.1 here and the value
            # moves from run to run): require the right magnitude and a finite, non-zero gradient
            assert 0.5 <= ratio <= 2.0 and cos == cos, (n, cos, ratio)
    l = eng.step(images, rot, gt).item()
    assert l == l


def test_adam_kernel_matches_torch_adam():
    from rotmv_b200 import _lib as L
    from rotmv_b200.train import _ck

    torch.manual_seed(0)
    n = 100003
    p = torch.randn(n, device="cuda"); g = torch.randn(n, device="cuda") * 0.1
    for decoupled, opt_cls in ((0, torch.optim.Adam), (1, torch.optim.AdamW)):
        ref_p = p.clone().requires_grad_(True)
        opt = opt_cls([ref_p], lr=1e-3, weight_decay=1e-2)
        mine = p.clone(); m = torch.zeros_like(p); v = torch.zeros_like(p)
        hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 1e-2, 0.0], device="cuda", dtype=torch.float64)
        for _ in range(3):
            ref_p.grad = g.clone()
            opt.step()
            _ck("rmv_adam_step", mine.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(),
                hyper.data_ptr(), n, decoupled, 1.0)
        assert hyper[5].item() == 3.0
        assert torch.allclose(mine, ref_p.detach(), rtol=1e-5, atol=1e-7), (mine - ref_p).abs().max()


def test_bn_train_kernels_match_torch():
    from rotmv_b200 import _lib as L
    from rotmv_b200.train import _ck

    torch.manual_seed(1)
    n, h, w, c, views = 6, 9, 7, 64, 2
    z = torch.randn((n, h, w, c), device="cuda") * 2 + 0.5
    gamma = torch.rand(c, device="cuda") + 0.5; beta = torch.randn(c, device="cuda")
    rm = torch.zeros(c, device="cuda"); rv = torch.ones(c, device="cuda")
    nbt = torch.zeros((), device="cuda", dtype=torch.long)
    acc = torch.zeros((views, c, 2), device="cuda", dtype=torch.float64)
    mean, invstd, a, b, k0, k1, k2 = (torch.empty((views, c), device="cuda") for _ in range(7))
    y = torch.empty_like(z)
    _ck("rmv_bn_stats", z.data_ptr(), 0, n, h * w, c, views, acc.data_ptr())
    _ck("rmv_bn_finalize", acc.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(), rv.data_ptr(),
        nbt.data_ptr(), mean.data_ptr(), invstd.data_ptr(), a.data_ptr(), b.data_ptr(), c, views,
        (n // views) * h * w, 1e-5, 0.1)
    bits = torch.zeros((n * h * w * c // 8,), device="cuda", dtype=torch.uint8)
    _ck("rmv_bn_apply", z.data_ptr(), a.data_ptr(), b.data_ptr(), None, y.data_ptr(), bits.data_ptr(), 0,
        n, h * w, c, views, 1)
    # packed ReLU mask: bit e of byte i <=> element 8*i+e passed the ReLU
    want_bits = ((y.reshape(-1, 8) > 0).to(torch.int32) << torch.arange(8, device="cuda", dtype=torch.int32)).sum(1)
    assert torch.equal(bits.to(torch.int32), want_bits)
    # reference: one nn.BatchNorm2d called once per view, in view order
    bn = torch.nn.BatchNorm2d(c).cuda().train()
    bn.weight.data.copy_(gamma); bn.bias.data.copy_(beta)
    zz = z.clone().requires_grad_(True)
    outs = {}
    for v in range(views):
        outs[v] = torch.relu(bn(zz[v::views].permute(0, 3, 1, 2))).permute(0, 2, 3, 1)
    ref = torch.empty_like(z)
    for v in range(views):
        ref[v::views] = outs[v].detach()
    assert torch.allclose(y, ref, rtol=1e-4, atol=1e-5)
    assert torch.allclose(rm, bn.running_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(rv, bn.running_var, rtol=1e-5, atol=1e-6)
    assert int(nbt) == views and acc.abs().max().item() == 0.0
    # one-launch variant (the last block finalizes): same coefficients, running stats, counters
    rm2 = torch.zeros(c, device="cuda"); rv2 = torch.ones(c, device="cuda")
    nbt2 = torch.zeros((), device="cuda", dtype=torch.long)
    ticket = torch.zeros((1,), device="cuda", dtype=torch.int32)
    mean2, invstd2, a2, b2, k02, k12, k22 = (torch.empty((views, c), device="cuda") for _ in range(7))
    for _ in range(2):   # twice: the accumulator/ticket reset must leave a clean state
        rm2.zero_(); rv2.fill_(1.0); nbt2.zero_()
        _ck("rmv_bn_stats_finalize", z.data_ptr(), 0, n, h * w, c, views, acc.data_ptr(), ticket.data_ptr(),
            gamma.data_ptr(), beta.data_ptr(), rm2.data_ptr(), rv2.data_ptr(), nbt2.data_ptr(),
            mean2.data_ptr(), invstd2.data_ptr(), a2.data_ptr(), b2.data_ptr(), 1e-5, 0.1)
        for got, want in ((mean2, mean), (invstd2, invstd), (a2, a), (b2, b), (rm2, rm), (rv2, rv)):
            assert torch.allclose(got, want, rtol=1e-6, atol=1e-7)
        assert int(nbt2) == views and int(ticket) == 0 and acc.abs().max().item() == 0.0
    dy = torch.randn_like(z)
    loss = sum((outs[v] * dy[v::views]).sum() for v in range(views))
    loss.backward()
    for mask, is_bits in ((y, 0), (bits, 1)):   # ReLU mask from the output tensor / from the packed bits
        dgamma = torch.empty(c, device="cuda"); dbeta = torch.empty(c, device="cuda")
        dz = torch.empty_like(z)
        dyr = torch.empty_like(z)
        _ck("rmv_bn_bwd_reduce", z.data_ptr(), dy.data_ptr(), mask.data_ptr(), is_bits, mean.data_ptr(),
            invstd.data_ptr(), 0, n, h * w, c, views, acc.data_ptr())
        _ck("rmv_bn_bwd_finalize", acc.data_ptr(), gamma.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
            dgamma.data_ptr(), dbeta.data_ptr(), k0.data_ptr(), k1.data_ptr(), k2.data_ptr(), c, views,
            (n // views) * h * w)
        _ck("rmv_bn_bwd_apply", z.data_ptr(), dy.data_ptr(), mask.data_ptr(), is_bits, k0.data_ptr(),
            k1.data_ptr(), k2.data_ptr(), dz.data_ptr(), dyr.data_ptr(), 0, n, h * w, c, views)
        assert torch.allclose(dz, zz.grad, rtol=1e-3, atol=1e-5), (dz - zz.grad).abs().max()
        assert torch.equal(dyr, dy * (y > 0))
        assert torch.allclose(dgamma, bn.weight.grad, rtol=1e-4, atol=1e-4)
        assert torch.allclose(dbeta, bn.bias.grad, rtol=1e-4, atol=1e-4)
        dg2 = torch.empty(c, device="cuda"); db2 = torch.empty(c, device="cuda")
        _ck("rmv_bn_bwd_reduce_finalize", z.data_ptr(), dy.data_ptr(), mask.data_ptr(), is_bits,
            mean.data_ptr(), invstd.data_ptr(), 0, n, h * w, c, views, acc.data_ptr(), ticket.data_ptr(),
            gamma.data_ptr(), dg2.data_ptr(), db2.data_ptr(), k02.data_ptr(), k12.data_ptr(), k22.data_ptr())
        for got, want in ((dg2, dgamma), (db2, dbeta), (k02, k0), (k12, k1), (k22, k2)):
            assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
        assert int(ticket) == 0 and acc.abs().max().item() == 0.0


def test_wgrad_dgrad_pool_bwd_match_torch():
    import torch.nn.functional as F
    from rotmv_b200 import _lib as L
    from rotmv_b200 import functional as RF
    from rotmv_b200.train import _ck
    import ctypes as C

    torch.manual_seed(2)
    for (n, hh, ci, co, k, s, p) in [(3, 14, 64, 128, 3, 1, 1), (2, 28, 64, 64, 3, 2, 1), (2, 14, 128, 256, 1, 2, 0)]:
        x = torch.randn((n, ci, hh, hh), device="cuda", requires_grad=True)
        wt = (torch.randn((co, ci, k, k), device="cuda") / (k * k * ci) ** 0.5).requires_grad_(True)
        yref = F.conv2d(x, wt, stride=s, padding=p)
        dy = torch.randn_like(yref)
        yref.backward(dy)
        xn = x.detach().permute(0, 2, 3, 1).contiguous()
        dyn = dy.permute(0, 2, 3, 1).contiguous()
        dw = torch.zeros_like(wt)
        a = L.ConvArgs()
        a.x_dtype = 0; a.x = xn.data_ptr()
        a.x_sn, a.x_sh, a.x_sw, a.x_sc = xn.stride(0), xn.stride(1), xn.stride(2), 1
        a.n_img, a.in_h, a.in_w, a.c_in = n, hh, hh, ci
        a.c_out, a.kh, a.kw, a.stride, a.pad = co, k, k, s, p
        a.y_sn, a.y_sh, a.y_sw = dyn.stride(0), dyn.stride(1), dyn.stride(2)
        a.out_h, a.out_w = dyn.shape[1], dyn.shape[2]
        L.check(L.load().rmv_conv2d_wgrad(C.byref(a), dyn.data_ptr(), dw.data_ptr(), L.stream_ptr()), "wgrad")
        assert torch.allclose(dw, wt.grad, rtol=1e-3, atol=1e-3 * wt.grad.abs().max().item())
        # dgrad through the forward kernel with flipped/transposed filters (+ dilation for stride 2)
        w_t = torch.empty((ci, k, k, co), device="cuda")
        _ck("rmv_permute_cast", wt.data_ptr(), w_t.data_ptr(), ci, k, k, co, k * k, k, 1, ci * k * k, 1, 1, 0)
        src = dyn
        if s == 2:
            src = torch.empty((n, 2 * dyn.shape[1], 2 * dyn.shape[2], co), device="cuda")
            _ck("rmv_dilate2", dyn.data_ptr(), src.data_ptr(), n, dyn.shape[1], dyn.shape[2], co, 0)
        dx = RF.conv2d(src, w_t, stride=1, pad=k - 1 - p)
        ref = x.grad.permute(0, 2, 3, 1)
        assert torch.allclose(dx, ref, rtol=1e-3, atol=1e-3 * ref.abs().max().item()), (k, s)
    # max-pool / avg-pool backward
    xp = torch.relu(torch.randn((2, 64, 12, 12), device="cuda")).requires_grad_(True)
    yp = F.max_pool2d(xp, 3, 2, 1)
    dyp = torch.randn_like(yp)
    yp.backward(dyp)
    xn = xp.detach().permute(0, 2, 3, 1).contiguous(); dn = dyp.permute(0, 2, 3, 1).contiguous()
    dxp = torch.empty_like(xn)
    _ck("rmv_maxpool3x3s2_bwd", xn.data_ptr(), dn.data_ptr(), dxp.data_ptr(), 2, 12, 12, 64, 0)
    assert torch.allclose(dxp, xp.grad.permute(0, 2, 3, 1), atol=1e-6)


def test_head_loss_bwd_matches_autograd():
    from rotmv_b200 import functional as RF
    from rotmv_b200.train import _ck
    from oracle import rotmv_oracle as O

    torch.manual_seed(3)
    m, hid, views = 24, 512, 2
    g_in = torch.relu(torch.randn((m, hid), device="cuda")).requires_grad_(True)
    w2 = (torch.randn((2, hid), device="cuda") * 0.02).requires_grad_(True)
    b2 = torch.zeros(2, device="cuda", requires_grad=True)
    gt = (torch.rand((m, 2), device="cuda") - 0.5)
    pred_ref = g_in @ w2.t() + b2
    scale = 0.01 / (m // views)
    loss_ref = sum(O.angular_loss_deg(pred_ref[v::views].cpu(), gt[v::views].cpu()) for v in range(views)) * 0.01
    loss_ref.backward()
    pred = torch.empty((m, 2), device="cuda"); loss = torch.zeros(1, device="cuda")
    RF.head_loss(g_in.detach(), w2.detach(), b2.detach(), pred, gt, scale, loss, views=views)
    assert abs(loss.item() - loss_ref.item()) <= 1e-4 * abs(loss_ref.item())
    dg = torch.empty_like(g_in); dpred = torch.empty((m, 2), device="cuda")
    dw2 = torch.zeros_like(w2); db2 = torch.zeros_like(b2)
    _ck("rmv_head_loss_bwd", pred.data_ptr(), gt.data_ptr(), g_in.data_ptr(), hid, 0, w2.data_ptr(), m, hid,
        scale, views, 1.0, dg.data_ptr(), hid, dpred.data_ptr(), dw2.data_ptr(), db2.data_ptr())
    mask = (g_in.detach() > 0).float()
    assert torch.allclose(dg, g_in.grad * mask, rtol=1e-3, atol=1e-6 + 1e-3 * g_in.grad.abs().max().item())
    assert torch.allclose(dw2, w2.grad, rtol=1e-3, atol=1e-3 * w2.grad.abs().max().item())
    assert torch.allclose(db2, b2.grad, rtol=1e-3, atol=1e-5)


WGRAD_TC_CASES = [
    # n, h, c_in, c_out, k, stride, pad
    (4, 56, 64, 64, 3, 1, 1), (4, 56, 64, 256, 1, 1, 0), (3, 28, 128, 128, 3, 1, 1),
    (4, 56, 128, 128, 3, 2, 1), (5, 14, 256, 256, 3, 1, 1), (6, 28, 512, 1024, 1, 2, 0),
    (7, 7, 512, 512, 3, 1, 1), (9, 7, 2048, 512, 1, 1, 0), (2, 14, 1024, 256, 1, 1, 0),
    (3, 23, 64, 64, 3, 1, 1), (2, 40, 64, 64, 3, 1, 1),   # tap-row kernel, ragged tiles
]


@pytest.mark.parametrize("case", WGRAD_TC_CASES)
def test_wgrad_tcgen05_matches_autograd(case):
    """rmv_conv2d_wgrad_tc (MN-major tcgen05 + TMA reduce-add split over pixels) vs torch autograd
    on the same bf16-rounded operands."""
    import ctypes as C
    import torch.nn.functional as F
    from rotmv_b200 import _lib as L

    n, hh, ci, co, k, s, p = case
    g = torch.Generator(device="cuda").manual_seed(sum(case))
    x = torch.randn((n, hh, hh, ci), device="cuda", generator=g).bfloat16()
    oh = (hh + 2 * p - k) // s + 1
    dy = torch.randn((n, oh, oh, co), device="cuda", generator=g).bfloat16()
    wt = torch.zeros((co, ci, k, k), device="cuda", requires_grad=True)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wt, stride=s, padding=p)
    y.backward(dy.float().permute(0, 3, 1, 2))
    ref = wt.grad.permute(0, 2, 3, 1).contiguous()  # KRSC
    dw = torch.zeros((co, k, k, ci), device="cuda")
    a = L.ConvArgs()
    a.x_dtype = 1; a.x = x.data_ptr()
    a.x_sn, a.x_sh, a.x_sw, a.x_sc = x.stride(0), x.stride(1), x.stride(2), 1
    a.n_img, a.in_h, a.in_w, a.c_in = n, hh, hh, ci
    a.c_out, a.kh, a.kw, a.stride, a.pad = co, k, k, s, p
    a.y_sn, a.y_sh, a.y_sw = dy.stride(0), dy.stride(1), dy.stride(2)
    a.out_h, a.out_w = oh, oh
    L.check(L.load().rmv_conv2d_wgrad_tc(C.byref(a), dy.data_ptr(), dw.data_ptr(), L.stream_ptr()), "wgrad_tc")
    err = (dw - ref).abs().max().item()
    assert err <= 2e-4 * ref.abs().max().item() + 1e-4, (case, err, ref.abs().max().item())
    # accumulation semantics (+=)
    L.check(L.load().rmv_conv2d_wgrad_tc(C.byref(a), dy.data_ptr(), dw.data_ptr(), L.stream_ptr()), "wgrad_tc")
    assert (dw - 2 * ref).abs().max().item() <= 4e-4 * ref.abs().max().item() + 2e-4


def test_permute_batch_kinds_match_generic():
    """rmv_permute_cast_batch: the tiled access patterns (kind 1/2/3) give exactly what the generic
    4-D permute (kind 0 = rmv_permute_cast) gives, for bf16 and fp32 destinations."""
    import ctypes as C
    from rotmv_b200 import _lib as L
    from rotmv_b200.train import _ck

    torch.manual_seed(3)
    cases = []  # (weight, dims, strides, flip, kind)
    for (k, c, r) in ((64, 64, 3), (128, 64, 1), (72, 200, 3), (40, 24, 1)):
        w = torch.randn((k, c, r, r), device="cuda")
        cases.append((w, (k, r, r, c), (c * r * r, r, 1, r * r), 0, 1 if r == 1 else 3))   # forward KRSC
        cases.append((w, (c, r, r, k), (r * r, r, 1, c * r * r), 1, 2))                    # dgrad layout
    lw = torch.randn((136, 520), device="cuda")
    cases.append((lw, (1, 1, 136, 520), (0, 0, 520, 1), 0, 1))
    cases.append((lw, (520, 1, 1, 136), (1, 0, 0, 520), 0, 2))
    for dt_code, dt in ((L.BF16, torch.bfloat16), (L.F32, torch.float32)):
        jobs = (L.PermuteJob * len(cases))()
        outs, refs, block = [], [], 0
        for i, (w, dims, strides, flip, kind) in enumerate(cases):
            out = torch.zeros(dims, device="cuda", dtype=dt)
            ref = torch.zeros(dims, device="cuda", dtype=dt)
            _ck("rmv_permute_cast", w.data_ptr(), ref.data_ptr(), *dims, *strides, flip, flip, dt_code)
            j = jobs[i]
            j.src, j.dst = w.data_ptr(), out.data_ptr()
            j.d0, j.d1, j.d2, j.d3 = dims
            j.s0, j.s1, j.s2, j.s3 = strides
            j.flip1, j.flip2, j.dst_dtype, j.first_block, j.kind = flip, flip, dt_code, block, kind
            if kind == 2:
                block += ((dims[0] * dims[1] * dims[2] + 63) // 64) * ((dims[3] + 63) // 64)
            elif kind == 3:
                block += dims[0]
            else:
                block += (out.numel() + 1023) // 1024
            outs.append(out); refs.append(ref)
        table = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).cuda()
        L.check(L.load().rmv_permute_cast_batch(table.data_ptr(), len(cases), block, L.stream_ptr()),
                "rmv_permute_cast_batch")
        torch.cuda.synchronize()
        for (w, dims, strides, flip, kind), out, ref in zip(cases, outs, refs):
            assert torch.equal(out, ref), (dims, kind, dt)
    # and the generic kernel itself against torch on one case
    w = cases[1][0]
    assert torch.equal(refs[1], w.permute(1, 2, 3, 0).flip(1, 2).contiguous())


def test_conv2d_dgrad_tc_matches_torch():
    """rmv_conv2d_dgrad (tcgen05): stride 1 and the parity-class stride-2 path (no dilated copy),
    3x3 and 1x1, odd and even input sizes, with and without a residual, against autograd on the
    bf16-rounded operands (fp32 accumulation: only the bf16 rounding of dx remains)."""
    import torch.nn.functional as F
    from rotmv_b200 import functional as RF

    torch.manual_seed(5)
    cases = [(3, 14, 64, 128, 3, 1, 1, True), (2, 28, 64, 64, 3, 2, 1, False), (2, 28, 128, 64, 3, 2, 1, True),
             (2, 14, 128, 256, 1, 2, 0, False), (3, 15, 64, 64, 3, 2, 1, True), (2, 9, 64, 128, 1, 2, 0, False),
             (4, 56, 64, 64, 1, 1, 0, True)]
    for (n, hh, ci, co, k, s, p, with_res) in cases:
        x = torch.randn((n, ci, hh, hh), device="cuda", requires_grad=True)
        w = (torch.randn((co, ci, k, k), device="cuda") / (k * k * co) ** 0.5).bfloat16().float().requires_grad_(True)
        y = F.conv2d(x, w, stride=s, padding=p)
        dy = torch.randn_like(y).bfloat16().float()
        y.backward(dy)
        ref = x.grad.permute(0, 2, 3, 1)
        res = torch.randn((n, hh, hh, ci), device="cuda").bfloat16() if with_res else None
        if with_res:
            ref = ref + res.float()
        wt = w.detach().permute(1, 2, 3, 0).flip(1, 2).contiguous().bfloat16()   # [C, kh, kw, K], taps reversed
        dx = RF.conv2d_dgrad(dy.permute(0, 2, 3, 1).contiguous().bfloat16(), wt, stride=s, pad=p,
                             in_hw=(hh, hh), residual=res)
        err = (dx.float() - ref).abs().max().item()
        assert err <= 1e-2 * ref.abs().max().item() + 1e-3, (n, hh, ci, co, k, s, p, err)


@pytest.mark.parametrize("case", [
    # n, h, w, ci, co, k, stride, pad
    (6, 56, 56, 64, 256, 1, 1, 0),    # flattened pointwise, image boundaries warp aligned
    (6, 28, 28, 128, 512, 1, 1, 0),   # flattened, 784 px per image (16-row aligned)
    (10, 14, 14, 256, 1024, 1, 1, 0), # flattened, 196 px per image: views change inside 16-row groups
    (14, 7, 7, 512, 2048, 1, 1, 0),   # flattened, 49 px per image
    (4, 56, 56, 64, 64, 3, 1, 1),     # halo-patch kernel
    (6, 28, 28, 128, 128, 3, 1, 1),   # boxed 4x4x8 tiles
    (10, 14, 14, 256, 256, 3, 1, 1),  # boxed 2x2x32 tiles
    (6, 7, 7, 512, 512, 3, 1, 1),     # boxed 1x1x128 tiles
    (4, 56, 56, 128, 128, 3, 2, 1),   # stride 2 (parity planes)
    (6, 28, 28, 256, 512, 1, 2, 0),   # stride-2 1x1 (downsample)
    (4, 23, 19, 64, 64, 3, 1, 1),     # ragged tiles
    (2, 30, 30, 64, 128, 1, 1, 0),    # ragged flattened tail
])
def test_conv_epilogue_bn_statistics(case):
    """rmv_conv_args.stat_acc: the per-(view, channel) sum / sum of squares the tcgen05 conv
    epilogue accumulates equal those of the bf16 tensor it wrote (fp64 reference), for every tile
    geometry; the output itself is unchanged by the option."""
    from rotmv_b200 import functional as RF, _lib as L

    n, h, w, ci, co, k, stride, pad = case
    g = torch.Generator(device="cuda").manual_seed(sum(case))
    x = torch.randn((n, h, w, ci), device="cuda", generator=g).bfloat16()
    wt = (torch.randn((co, k, k, ci), device="cuda", generator=g) / (k * k * ci) ** 0.5).bfloat16()
    acc = torch.zeros((2, co, 2), device="cuda", dtype=torch.float64)
    y = RF.conv2d(x, wt, stride=stride, pad=pad, engine=L.ENGINE_TC, stat_acc=acc, stat_views=2)
    y0 = RF.conv2d(x, wt, stride=stride, pad=pad, engine=L.ENGINE_TC)
    assert torch.equal(y, y0)
    yd = y.double()
    for v in range(2):
        s1 = yd[v::2].sum(dim=(0, 1, 2))
        s2 = (yd[v::2] ** 2).sum(dim=(0, 1, 2))
        assert torch.allclose(acc[v, :, 0], s1, rtol=1e-5, atol=1e-4 * s2.sqrt().max().item()), (case, v)
        assert torch.allclose(acc[v, :, 1], s2, rtol=1e-5, atol=1e-6 * s2.max().item()), (case, v)


@pytest.mark.parametrize("n,h,w", [(2, 12, 12), (3, 13, 11), (1, 112, 112), (2, 7, 9)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_maxpool_index_forward_backward(n, h, w, dt):
    """Training max-pool: the forward records the winning window position, the backward (one thread
    per 2x2 input block) routes every gradient through it. Against autograd, with ties (ReLU zeros):
    the first maximum in scan order wins, as in ATen."""
    import torch.nn.functional as F
    from rotmv_b200 import _lib as L
    from rotmv_b200.train import _ck

    torch.manual_seed(n * 100 + h + w)
    x = torch.relu(torch.randn((n, 64, h, w), device="cuda")).to(dt).float().requires_grad_(True)
    y = F.max_pool2d(x, 3, 2, 1)
    dy = torch.randn_like(y).to(dt).float()
    y.backward(dy)
    xn = x.detach().permute(0, 2, 3, 1).contiguous().to(dt)
    dn = dy.permute(0, 2, 3, 1).contiguous().to(dt)
    oh, ow = y.shape[2], y.shape[3]
    yn = torch.empty((n, oh, ow, 64), device="cuda", dtype=dt)
    idx = torch.empty((n, oh, ow, 64), device="cuda", dtype=torch.uint8)
    dx = torch.full((n, h, w, 64), 7.0, device="cuda", dtype=dt)
    code = L.dtype_code(dt)
    _ck("rmv_maxpool3x3s2_fwd_idx", xn.data_ptr(), yn.data_ptr(), idx.data_ptr(), n, h, w, 64, code)
    _ck("rmv_maxpool3x3s2_bwd_idx", idx.data_ptr(), dn.data_ptr(), dx.data_ptr(), n, h, w, 64, code)
    assert torch.equal(yn.float(), y.detach().permute(0, 2, 3, 1))
    ref = x.grad.permute(0, 2, 3, 1)
    tol = 0.0 if dt == torch.float32 else 2e-2 * ref.abs().max().item()   # bf16: sums of up to 4 terms rounded
    assert (dx.float() - ref).abs().max().item() <= tol + 1e-6


@pytest.mark.parametrize("flags", [dict(), dict(share_weights=True), dict(ignore_rotmat=True),
                                   dict(encode_rotmat=True), dict(share_feature=True),
                                   dict(encode_rotmat=True, share_weights=True)],
                         ids=["default", "share_weights", "ignore_rotmat", "encode_rotmat",
                              "share_feature", "encode_rotmat+share_weights"])
def test_depth18_step_fp32_vs_live_oracle_and_bf16(flags):
    """backbone_depth=18 (BasicBlock trunk, models/resnet.py:50-96): one training step of the fp32
    engine against the CPU oracle's autograd step on the same batch/weights -- loss, every gradient
    norm, BatchNorm running statistics, parameters after Adam; the bf16 (tcgen05) engine within the
    usual bf16 distance."""
    from oracle import rotmv_oracle as O
    from rotmv_b200.module import FeatRotationSymm
    from rotmv_b200.train import TrainEngine

    B, V, lr = 6, 2, 1e-3
    ora = O.build_model(num_iter=2, depth=18, seed=0, **flags)
    sd0 = {k: v.clone() for k, v in ora.state_dict().items()}
    images, pose, gt = O.synthetic_batch(B, V, seed=2)
    rot = O.pairwise_rotations(pose)
    ora.train()
    out = ora.forward_views(images, rot)
    loss_ref = O.iteration_loss(out, [gt[:, 0], gt[:, 1]])
    loss_ref.backward()
    ref_grads = {n: p.grad.clone() for n, p in ora.named_parameters() if p.grad is not None}

    for precision in ("fp32", "bf16"):
        model = FeatRotationSymm(18, 2, **flags)
        model.load_state_dict(sd0, strict=True)
        model = model.cuda().train()
        eng = TrainEngine(model, precision=precision, lr=lr, weight_decay=1e-6)
        res = eng.forward_backward(images.cuda(), rot.cuda(), gt.cuda())
        loss = res["loss"].item()
        named = dict(model.named_parameters())
        errs = []
        for n, g_ref in ref_grads.items():
            g = eng.grads[id(named[n])].cpu().double()
            errs.append((abs(g.norm().item() - g_ref.double().norm().item()) / max(g_ref.double().norm().item(), 1e-12), n))
        errs.sort(reverse=True)
        if precision == "fp32":
            assert abs(loss - loss_ref.item()) <= 1e-4 * abs(loss_ref.item()), (loss, loss_ref.item())
            assert len(errs) == len(ref_grads) and errs[0][0] <= 2e-2, errs[:3]
            assert errs[len(errs) // 2][0] <= 2e-3, errs[len(errs) // 2]
            bn = model._feat_extractor[0].bn1
            obn = ora._feat_extractor[0].bn1
            assert torch.allclose(bn.running_mean.cpu(), obn.running_mean, rtol=1e-3, atol=1e-5)
            assert torch.allclose(bn.running_var.cpu(), obn.running_var, rtol=1e-3, atol=1e-5)
            assert int(bn.num_batches_tracked) == V
            if flags.get("share_feature"):
                # IntensityBatchNorm (models/rot_mv.py:13-32): running STD after 4 train-mode calls
                for i in range(2):
                    got = model._img_fusers[i]._batchnorm.running_mean.cpu()
                    want = ora._img_fusers[i]._batchnorm.running_mean
                    assert torch.allclose(got, want, rtol=1e-4, atol=1e-6), i
        else:
            assert abs(loss - loss_ref.item()) <= 3e-2 * abs(loss_ref.item()), (loss, loss_ref.item())
            assert errs[len(errs) // 2][0] <= 0.1, errs[len(errs) // 2]   # median gradient norm within 10 %
        # the optimiser step runs and moves every trained tensor
        p0 = eng.flat_p.clone()
        eng.step(images.cuda(), rot.cuda(), gt.cuda())
        assert torch.isfinite(eng.flat_p).all() and not torch.equal(p0, eng.flat_p)


def test_three_view_step_fp32_vs_live_oracle():
    """V = 3 training step (SURVEY D1 generalisation; BatchNorm statistics per view, partner mean
    in the fusion and its transposed gather in the backward pass): fp32 engine against the CPU
    oracle's autograd on the same batch and weights."""
    from oracle import rotmv_oracle as O
    from rotmv_b200.module import FeatRotationSymm
    from rotmv_b200.train import TrainEngine

    B, V = 4, 3
    ora = O.build_model(num_iter=2, depth=18, seed=0)
    sd0 = {k: v.clone() for k, v in ora.state_dict().items()}
    images, pose, gt = O.synthetic_batch(B, V, seed=4)
    rot = O.pairwise_rotations(pose)
    ora.train()
    out = ora.forward_views(images, rot)
    loss_ref = O.iteration_loss(out, [gt[:, v] for v in range(V)])
    loss_ref.backward()
    ref_grads = {n: p.grad.clone() for n, p in ora.named_parameters() if p.grad is not None}
    model = FeatRotationSymm(18, 2)
    model.load_state_dict(sd0, strict=True)
    model = model.cuda().train()
    eng = TrainEngine(model, precision="fp32", lr=1e-3, weight_decay=1e-6)
    res = eng.forward_backward(images.cuda(), rot.cuda(), gt.cuda())
    loss = res["loss"].item()
    assert abs(loss - loss_ref.item()) <= 1e-4 * abs(loss_ref.item()), (loss, loss_ref.item())
    named = dict(model.named_parameters())
    errs = []
    for n, g_ref in ref_grads.items():
        g = eng.grads[id(named[n])].cpu().double()
        errs.append((abs(g.norm().item() - g_ref.double().norm().item()) / max(g_ref.double().norm().item(), 1e-12), n))
    errs.sort(reverse=True)
    assert len(errs) == len(ref_grads) and errs[0][0] <= 2e-2, errs[:3]
    assert errs[len(errs) // 2][0] <= 2e-3, errs[len(errs) // 2]
    # lifter gradient element-wise (it collects every path through the rotated gathers)
    n = "_lifter._lifter.blocks.1.0.bias"
    assert rel_l2(eng.grads[id(named[n])].cpu(), ref_grads[n]) <= 2e-2
    assert int(model._feat_extractor[0].bn1.num_batches_tracked) == V


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_autograd_bridge_reference_style_step(precision):
    """The reference's own step, unchanged (trainer.py:119-123,141-143): `data = model(data)` in
    train mode, the loss built OUTSIDE the model on the returned predictions (IterationLoss
    restatement, torch ops), `zero_grad(); loss.backward(); torch.optim.Adam.step()`. The module's
    autograd bridge runs the engine's forward/backward kernels underneath; loss and gradients are
    checked against the CPU oracle's autograd, the update against torch's Adam on the oracle."""
    from oracle import rotmv_oracle as O
    from rotmv_b200.module import FeatRotationSymm

    B, V, lr = 6, 2, 1e-3
    ora = O.build_model(num_iter=2, depth=18, seed=0)
    sd0 = {k: v.clone() for k, v in ora.state_dict().items()}
    images, pose, gt = O.synthetic_batch(B, V, seed=2)
    rot = O.pairwise_rotations(pose)
    ora.train()
    oopt = O.make_adam(ora, lr=lr)
    ref_out = ora.forward_views(images, rot)
    loss_ref = O.iteration_loss(ref_out, [gt[:, 0], gt[:, 1]])
    oopt.zero_grad(); loss_ref.backward()
    ref_grads = {n: p.grad.clone() for n, p in ora.named_parameters() if p.grad is not None}
    oopt.step()

    model = FeatRotationSymm(18, 2, precision=precision)
    model.load_state_dict(sd0, strict=True)
    model = model.cuda().train()
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=1e-6)      # trainer.py:54
    r = O.rotation_matrix_2d(pose.reshape(-1, 2)).view(B, V, 3, 3)
    data = {"img_0": images[:, 0].cuda(), "img_1": images[:, 1].cuda(), "rot_0": r[:, 0].cuda(),
            "rot_1": r[:, 1].cuda(), "gt_gaze": gt[:, 0].cuda(), "gt_gaze_1": gt[:, 1].cuda()}
    out = model(data)
    assert out is data and out["num_iter"] == 2 and out["pred_gaze"].requires_grad
    # the train-mode dict carries every key of the reference's (models/rot_mv.py:205-211,256-266)
    tol_feat = 1e-3 if precision == "fp32" else 0.1
    for k in range(V):
        for key, shape in ((f"img_feat_{k}", (B, 512)), (f"initial_rot_feat_{k}", (B, 3, 512))):
            assert tuple(out[key].shape) == shape, (key, tuple(out[key].shape))
            assert rel_l2(out[key].cpu(), ref_out[key].detach()) <= tol_feat, key
        for i in range(2):
            f = out[f"iter_{i}"][f"feat_{k}"]
            assert tuple(f.shape) == (B, 3, 512)
            assert rel_l2(f.cpu(), ref_out[f"iter_{i}"][f"feat_{k}"].detach()) <= tol_feat, (i, k)
            assert tuple(out[f"iter_{i}"][f"pred_gaze_{k}"].shape) == (B, 2)
    loss = O.iteration_loss(out, [data["gt_gaze"], data["gt_gaze_1"]])
    opt.zero_grad(); loss.backward()
    tol_loss, tol_med = (1e-4, 2e-3) if precision == "fp32" else (3e-2, 0.1)
    assert abs(loss.item() - loss_ref.item()) <= tol_loss * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    errs = []
    for n, p in model.named_parameters():
        if n.startswith("_feat_extractor.0.fc."):
            assert p.grad is None, n                                        # Q4
            continue
        g_ref = ref_grads[n].double()
        errs.append(abs(p.grad.double().cpu().norm().item() - g_ref.norm().item()) / max(g_ref.norm().item(), 1e-12))
    errs.sort()
    assert errs[len(errs) // 2] <= tol_med, errs[len(errs) // 2]
    if precision == "fp32":
        assert errs[-1] <= 2e-2, errs[-1]
    opt.step()
    w = "_gaze_estimators.1.blocks.1.0.weight"
    moved = (model.state_dict()[w].cpu() - sd0[w]).abs().max().item()
    assert 0.5 * lr <= moved <= 1.5 * lr, moved                            # first Adam step: ~lr per element
    if precision == "fp32":
        assert (model.state_dict()[w].cpu() - ora.state_dict()[w]).abs().max().item() <= 2 * lr
    bn = model._feat_extractor[0].bn1
    assert int(bn.num_batches_tracked) == V
    # the next step starts from the updated weights; the tensor form returns pred_gaze [B, 2]
    images_d = torch.stack([data["img_0"], data["img_1"]], 1)
    pred = model(images_d, rot.cuda())
    assert tuple(pred.shape) == (B, 2) and pred.requires_grad
    pred.sum().backward()
    assert torch.isfinite(model._lifter._lifter.blocks[0][0].weight.grad).all()
    # and eval mode serves inference from the trained weights
    model.eval()
    with torch.no_grad():
        p_eval = model(images_d, rot.cuda())
    assert torch.isfinite(p_eval).all()
