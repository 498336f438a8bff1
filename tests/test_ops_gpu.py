"""Per-kernel parity on the GPU: each sm_100a kernel against the single torch op it replaces
(fp32, TF32 disabled), on the same seeded inputs. Calls go through the C ABI (ctypes)."""
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


def _ref_conv(x_nhwc, w_krsc, stride, pad, scale, shift, residual, relu):
    x = x_nhwc.float().permute(0, 3, 1, 2)
    w = w_krsc.float().permute(0, 3, 1, 2)
    y = F.conv2d(x, w, stride=stride, padding=pad)
    if scale is not None:
        y = y * scale.view(1, -1, 1, 1)
    if shift is not None:
        y = y + shift.view(1, -1, 1, 1)
    y = y.permute(0, 2, 3, 1)
    if residual is not None:
        y = y + residual.float()
    if relu:
        y = torch.relu(y)
    return y.contiguous()


CONV_CASES = [
    # n, h, w, c_in, c_out, k, stride, pad, residual, relu
    (4, 56, 56, 64, 64, 1, 1, 0, False, True),
    (4, 56, 56, 64, 256, 1, 1, 0, True, True),
    (3, 56, 56, 64, 64, 3, 1, 1, False, True),
    (5, 28, 28, 128, 128, 3, 1, 1, False, True),
    (7, 14, 14, 256, 256, 3, 1, 1, False, True),
    (3, 7, 7, 512, 512, 3, 1, 1, False, False),
    (4, 56, 56, 128, 128, 3, 2, 1, False, True),
    (2, 28, 28, 256, 256, 3, 2, 1, False, True),
    (6, 14, 14, 512, 512, 3, 2, 1, False, True),
    (4, 56, 56, 256, 512, 1, 2, 0, False, False),
    (9, 14, 14, 1024, 2048, 1, 2, 0, False, False),
    (16, 7, 7, 2048, 512, 1, 1, 0, False, True),
    (33, 7, 7, 512, 2048, 1, 1, 0, True, True),
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("engine", ["tc", "simt_bf16", "simt_f32"])
def test_conv_parity(case, engine):
    from rotmv_b200 import functional as RF, _lib as L

    n, h, w, ci, co, k, stride, pad, use_res, relu = case
    g = torch.Generator(device="cuda").manual_seed(sum(case) * 7 + 1)
    dt = torch.float32 if engine == "simt_f32" else torch.bfloat16
    x = torch.randn((n, h, w, ci), device="cuda", generator=g).to(dt)
    wt = (torch.randn((co, k, k, ci), device="cuda", generator=g) / math.sqrt(k * k * ci)).to(dt)
    scale = torch.rand((co,), device="cuda", generator=g) + 0.5
    shift = torch.randn((co,), device="cuda", generator=g)
    oh = (h + 2 * pad - k) // stride + 1
    res = torch.randn((n, oh, oh, co), device="cuda", generator=g).to(dt) if use_res else None
    eng = L.ENGINE_TC if engine == "tc" else L.ENGINE_SIMT
    y = RF.conv2d(x, wt, stride=stride, pad=pad, scale=scale, shift=shift, residual=res,
                  relu=relu, engine=eng)
    ref = _ref_conv(x, wt, stride, pad, scale, shift, res, relu)
    torch.cuda.synchronize()
    err = (y.float() - ref).abs().max().item()
    tol = (2e-5 if dt == torch.float32 else 1.2e-2) * ref.abs().max().item()
    assert err <= tol, f"{engine} {case}: max err {err} > {tol}"


@pytest.mark.parametrize("block_n", [64, 128, 256])
def test_conv_tc_fp32_out_tight(block_n):
    """fp32 output removes the bf16 rounding of y: tcgen05 accumulation must agree with the fp32
    reference to accumulation-order noise."""
    from rotmv_b200 import functional as RF, _lib as L

    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn((6, 14, 14, 256), device="cuda", generator=g).bfloat16()
    wt = (torch.randn((512, 3, 3, 256), device="cuda", generator=g) / 48).bfloat16()
    y = RF.conv2d(x, wt, stride=1, pad=1, out_dtype=torch.float32, engine=L.ENGINE_TC,
                  block_n=block_n)
    ref = _ref_conv(x, wt, 1, 1, None, None, None, False)
    err = (y - ref).abs().max().item()
    assert err <= 2e-5 * ref.abs().max().item() + 1e-5, err


LINEAR_CASES = [(300, 2048, 1536), (16, 3584, 3584), (513, 3584, 512), (128, 1536, 1536),
                (1024, 3584, 1536),
                # the fusion-stage shapes at M = B*V rows (models/rot_mv.py:35-50,91-98,179-184): split-K
                (2, 3584, 512), (16, 2048, 1536), (512, 3584, 3584), (512, 3584, 1536), (512, 3584, 512),
                (512, 2048, 1536), (512, 1536, 1536), (2048, 3584, 512), (256, 3584, 3584)]


@pytest.mark.parametrize("m,k,n", LINEAR_CASES)
@pytest.mark.parametrize("engine", ["tc", "simt_f32"])
def test_linear_parity(m, k, n, engine):
    from rotmv_b200 import functional as RF, _lib as L

    g = torch.Generator(device="cuda").manual_seed(m + k + n)
    dt = torch.float32 if engine == "simt_f32" else torch.bfloat16
    # strided input and output views, as the fusion block uses them (concat-free buffers)
    xbuf = torch.randn((m, k + 64), device="cuda", generator=g).to(dt)
    x = xbuf[:, 64:]
    wt = (torch.randn((n, k), device="cuda", generator=g) / math.sqrt(k)).to(dt)
    b = torch.randn((n,), device="cuda", generator=g)
    obuf = torch.zeros((m, n + 128), device="cuda", dtype=dt)
    out = obuf[:, 128:]
    RF.linear(x, wt, b, relu=True, out=out,
              engine=L.ENGINE_TC if engine == "tc" else L.ENGINE_SIMT)
    ref = torch.relu(x.float() @ wt.float().t() + b)
    err = (out.float() - ref).abs().max().item()
    tol = (2e-5 if dt == torch.float32 else 1.2e-2) * ref.abs().max().item()
    assert err <= tol, (err, tol)
    assert obuf[:, :128].abs().max().item() == 0.0  # nothing written outside the view


@pytest.mark.parametrize("m,k,n", [(512, 3584, 512), (256, 3584, 1536), (16, 2048, 1536), (512, 1536, 1536)])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_linear_splitk_is_deterministic_and_matches_unsplit(m, k, n, out_dtype):
    """Split-K of the small-M GEMMs: the fp32 partial tiles are added in a fixed order, so repeated
    launches are BIT-identical; against the unsplit kernel (RMV_SPLITK=0) only the fp32 summation
    order differs."""
    from rotmv_b200 import functional as RF, _lib as L

    g = torch.Generator(device="cuda").manual_seed(m * 3 + k + n)
    x = torch.randn((m, k), device="cuda", generator=g).bfloat16()
    wt = (torch.randn((n, k), device="cuda", generator=g) / math.sqrt(k)).bfloat16()
    b = torch.randn((n,), device="cuda", generator=g)
    outs = [RF.linear(x, wt, b, relu=True, out_dtype=out_dtype, engine=L.ENGINE_TC).clone() for _ in range(3)]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    L.check(L.load().rmv_set_tuning(b"SPLITK", 0), "rmv_set_tuning")
    try:
        plain = RF.linear(x, wt, b, relu=True, out_dtype=out_dtype, engine=L.ENGINE_TC)
    finally:
        L.check(L.load().rmv_set_tuning(b"SPLITK", 1), "rmv_set_tuning")
    ref = torch.relu(x.float() @ wt.float().t() + b)
    tol = (2e-5 if out_dtype == torch.float32 else 1.2e-2) * ref.abs().max().item() + 1e-5
    assert (outs[0].float() - ref).abs().max().item() <= tol
    assert (outs[0].float() - plain.float()).abs().max().item() <= tol


def test_maxpool_avgpool():
    from rotmv_b200 import functional as RF

    g = torch.Generator(device="cuda").manual_seed(3)
    for dt in (torch.float32, torch.bfloat16):
        x = torch.randn((3, 112, 112, 64), device="cuda", generator=g).to(dt)
        y = RF.maxpool3x3s2(x)
        ref = F.max_pool2d(x.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
        assert torch.equal(y.float(), ref)
        z = torch.randn((5, 7, 7, 2048), device="cuda", generator=g).to(dt)
        o0 = torch.zeros((5, 3584), device="cuda", dtype=dt)
        o1 = torch.zeros((5, 3584), device="cuda", dtype=dt)
        RF.avgpool(z, o0, o1)
        refm = z.float().mean(dim=(1, 2))
        tol = 1e-6 if dt == torch.float32 else 1e-2
        assert (o0[:, :2048].float() - refm).abs().max().item() <= tol
        assert torch.equal(o0, o1) and o0[:, 2048:].abs().max().item() == 0


@pytest.mark.parametrize("views", [2, 3, 4, 8, 20])
def test_rotate_gather(views):
    """models/rot_mv.py:234,238 generalised to V views: the pair kernel (V = 2), the
    shared-memory staged kernel (V = 3, 4) and the general kernel (V = 8, 20) against the
    formula, fp32 and bf16; the transposed mode (backward of the gather) through the adjoint identity
    <A x, y> = <x, A^T y>."""
    from rotmv_b200 import functional as RF

    g = torch.Generator(device="cuda").manual_seed(11)
    b = 9
    pose = (torch.rand((b, views, 2), device="cuda", generator=g) - 0.5)
    rot = RF.pose_to_rotations(pose)
    feat = torch.randn((b * views, 1536), device="cuda", generator=g)
    dst = torch.zeros((b * views, 3584), device="cuda")
    RF.rotate_gather(feat, rot, dst[:, 2048:], b, views)
    f3 = feat.view(b, views, 3, 512)
    ref = torch.zeros_like(f3)
    for v in range(views):
        for u in range(views):
            if u != v:
                ref[:, v] += rot[:, v, u] @ f3[:, u]
    ref /= (views - 1)
    assert (dst[:, 2048:].reshape(b, views, 3, 512) - ref).abs().max().item() < 1e-5
    assert dst[:, :2048].abs().max().item() == 0
    # bf16 storage: same arithmetic on the rounded inputs, one rounding of the result
    f16 = feat.bfloat16()
    d16 = torch.zeros((b * views, 1536), device="cuda", dtype=torch.bfloat16)
    RF.rotate_gather(f16, rot, d16, b, views)
    ref16 = torch.zeros_like(f3)
    f3r = f16.float().view(b, views, 3, 512)
    for v in range(views):
        for u in range(views):
            if u != v:
                ref16[:, v] += rot[:, v, u] @ f3r[:, u]
    ref16 /= (views - 1)
    assert (d16.float().view(b, views, 3, 512) - ref16).abs().max().item() <= 2e-2 * ref16.abs().max().item()
    # identity mode (ignore_rotmat): plain partner mean
    RF.rotate_gather(feat, rot, dst[:, 2048:], b, views, 512, False)
    mean_ref = (f3.sum(1, keepdim=True) - f3) / (views - 1)
    assert (dst[:, 2048:].reshape(b, views, 3, 512) - mean_ref).abs().max().item() < 1e-5
    # transposed mode = adjoint of the forward gather
    y = torch.randn((b * views, 1536), device="cuda", generator=g)
    aty = torch.empty_like(y)
    RF.rotate_gather(y, rot, aty, b, views, 512, True, transpose=True)
    ax = torch.empty_like(feat)
    RF.rotate_gather(feat, rot, ax, b, views, 512, True)
    lhs, rhs = (ax.double() * y.double()).sum().item(), (feat.double() * aty.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs)), (lhs, rhs)


def test_stem_im2col_matches_conv():
    from rotmv_b200 import functional as RF, _lib as L

    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((2, 3, 224, 224), device="cuda", generator=g)
    w = torch.randn((64, 3, 7, 7), device="cuda", generator=g) / 12
    a, oh, ow = RF.stem_im2col(x)
    wk = torch.zeros((64, 192), device="cuda")
    wk[:, :147] = w.permute(0, 2, 3, 1).reshape(64, 147)
    y = RF.linear(a, wk.bfloat16(), None, engine=L.ENGINE_TC, out_dtype=torch.float32)
    ref = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), stride=2, padding=3)
    ref = ref.permute(0, 2, 3, 1).reshape(-1, 64)
    assert (y - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 1e-5
    # fp32 engine straight from NCHW strides
    y32 = RF.conv2d_nchw_input(x, w.permute(0, 2, 3, 1).contiguous(), stride=2, pad=3)
    ref32 = F.conv2d(x, w, stride=2, padding=3).permute(0, 2, 3, 1)
    assert (y32 - ref32).abs().max().item() <= 2e-5 * ref32.abs().max().item()


@pytest.mark.parametrize("n,h,w", [(3, 224, 224), (2, 96, 96), (1, 70, 106)])
def test_fused_stem_matches_conv_bn_relu(n, h, w):
    """rmv_stem_conv_fwd (in-smem im2col + tcgen05 + TMA store) vs conv7x7/s2 + affine + ReLU."""
    from rotmv_b200 import functional as RF

    g = torch.Generator(device="cuda").manual_seed(n * 1000 + h)
    x = torch.randn((n, 3, h, w), device="cuda", generator=g)
    wt = torch.randn((64, 3, 7, 7), device="cuda", generator=g) / 12
    scale = torch.rand((64,), device="cuda", generator=g) + 0.5
    shift = torch.randn((64,), device="cuda", generator=g)
    y = RF.stem_conv(x, RF.stem_pack_weights(wt), scale, shift)
    ref = F.conv2d(x.bfloat16().float(), wt.bfloat16().float(), stride=2, padding=3)
    ref = torch.relu(ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)).permute(0, 2, 3, 1)
    assert tuple(y.shape) == tuple(ref.shape)
    err = (y.float() - ref).abs().max().item()
    assert err <= 1.2e-2 * ref.abs().max().item(), err


@pytest.mark.parametrize("n,h,w", [(3, 56, 56), (2, 20, 23), (5, 16, 8), (1, 33, 40)])
def test_conv3x3_halo_kernel(n, h, w):
    """3x3/s1/p1 64->64 runs on the halo-patch kernel (one 16x18 input patch per 8x16 output tile,
    nine row-shifted UMMA descriptors, resident filters): full tiles, ragged right/bottom edges,
    several images; fp32 output isolates the accumulation, bf16 output is the production path."""
    from rotmv_b200 import functional as RF, _lib as L

    g = torch.Generator(device="cuda").manual_seed(n * 1000 + h * 10 + w)
    x = torch.randn((n, h, w, 64), device="cuda", generator=g).bfloat16()
    wt = (torch.randn((64, 3, 3, 64), device="cuda", generator=g) / 24).bfloat16()
    scale = torch.rand((64,), device="cuda", generator=g) + 0.5
    shift = torch.randn((64,), device="cuda", generator=g)
    y = RF.conv2d(x, wt, stride=1, pad=1, scale=scale, shift=shift, relu=True, engine=L.ENGINE_TC)
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), padding=1)
    ref = torch.relu(ref.permute(0, 2, 3, 1) * scale + shift)
    err = (y.float() - ref).abs().max().item()
    assert err <= 1.2e-2 * ref.abs().max().item(), (err, ref.abs().max().item())
    # the generic tap-by-tap kernel (block_n given explicitly) must agree to bf16 rounding
    y2 = RF.conv2d(x, wt, stride=1, pad=1, scale=scale, shift=shift, relu=True, engine=L.ENGINE_TC, block_n=64)
    assert (y.float() - y2.float()).abs().max().item() <= 4e-2 * ref.abs().max().item() / 8


PAIR_CASES = [
    # n, h, w, ci, co, k, stride, pad
    (6, 14, 14, 256, 256, 3, 1, 1),     # boxed 2x2x32 tiles, odd number of M tiles (10 -> tail pair)
    (7, 14, 14, 1024, 256, 1, 1, 0),    # flattened pointwise, K = 1024
    (5, 7, 7, 512, 2048, 1, 1, 0),      # 8 N tiles
    (4, 28, 28, 256, 256, 3, 2, 1),     # stride 2 (parity planes)
    (3, 28, 28, 512, 1024, 1, 2, 0),    # stride-2 pointwise (downsample)
    (1, 16, 16, 64, 256, 1, 1, 0),      # exactly one pair
    (9, 7, 7, 512, 512, 3, 1, 1),       # 1x1x128 tiles
    (6, 28, 28, 128, 128, 3, 1, 1),     # 128-wide pair tiles (each CTA loads 64 filter rows)
    (3, 56, 56, 128, 128, 3, 2, 1),     # 128-wide, stride 2
    (5, 14, 14, 512, 384, 1, 1, 0),     # 3 N tiles of 128
    (1, 1, 512, 3584, 1536, 1, 1, 0),   # the fuser Linear shape (M = 512 rows)
]


@pytest.mark.parametrize("case", PAIR_CASES)
def test_conv_cta_pair_kernel(case):
    """cta_group::2 pair kernel (RMV_CTA2 / rmv_set_tuning("CTA2", 2) forces it for every
    c_out % 256 == 0 layer): same results as torch on the bf16-rounded operands, scale/shift/ReLU
    epilogue included, and bit-identical to the single-CTA kernel's output."""
    from rotmv_b200 import functional as RF, _lib as L

    n, h, w, ci, co, k, stride, pad = case
    g = torch.Generator(device="cuda").manual_seed(sum(case) + 11)
    x = torch.randn((n, h, w, ci), device="cuda", generator=g).bfloat16()
    wt = (torch.randn((co, k, k, ci), device="cuda", generator=g) / math.sqrt(k * k * ci)).bfloat16()
    scale = torch.rand((co,), device="cuda", generator=g) + 0.5
    shift = torch.randn((co,), device="cuda", generator=g)
    lib = L.load()
    try:
        L.check(lib.rmv_set_tuning(b"CTA2", 0), "set_tuning")
        y1 = RF.conv2d(x, wt, stride=stride, pad=pad, scale=scale, shift=shift, relu=True, engine=L.ENGINE_TC)
        L.check(lib.rmv_set_tuning(b"CTA2", 2), "set_tuning")
        y2 = RF.conv2d(x, wt, stride=stride, pad=pad, scale=scale, shift=shift, relu=True, engine=L.ENGINE_TC)
        torch.cuda.synchronize()
    finally:
        L.check(lib.rmv_set_tuning(b"CTA2", 1), "set_tuning")   # back to the default
    ref = _ref_conv(x, wt, stride, pad, scale, shift, None, True)
    err = (y2.float() - ref).abs().max().item()
    assert err <= 1.2e-2 * ref.abs().max().item(), (case, err)
    assert (y1.float() - y2.float()).abs().max().item() <= 8e-3 * ref.abs().max().item()
