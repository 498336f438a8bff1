"""Parity of the re-layout kernels of the constructor variants (csrc/variant_glue.cu:
rmv_strided_copy, rmv_intensity_bn_train, rmv_fill_zero) against their torch formulas, through the
C ABI. They replace the tensor glue of ImageRotmatFeatFuser / RotFeatFuser / IntensityBatchNorm
(models/rot_mv.py:13-32,53-85,225-248): zero-padded 3593-wide layers and their transposes, the 9
rotation entries per row, the [3][2][512] interleave, the train-mode running-std update. Copies
are bit-exact (one fp32 -> bf16 rounding at most); the statistics are compared at fp32 tolerance.
The same kernels run end to end in tests/test_variants.py and the depth-18 training test."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, dtype, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, device="cuda", generator=g).to(dtype)


def _check_copy(src, dst, scale=None, accumulate=False):
    from rotmv_b200 import functional as RF

    base = dst.clone()
    want = src.float() if scale is None else src.float() * scale
    if accumulate:
        want = want + base.float()
    want = want.to(dst.dtype)
    RF.strided_copy(src, dst, scale, accumulate)
    torch.cuda.synchronize()
    assert torch.equal(dst, want), (dst.float() - want.float()).abs().max().item()


@pytest.mark.parametrize("sd,dd", [(torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16),
                                   (torch.bfloat16, torch.float32), (torch.float32, torch.float32)])
def test_strided_copy_padded_corner_and_its_transpose(sd, dd):
    """w[:n,:k] = W and wt[:k,:n] = W^T into zero-padded buffers (ImageRotmatFeatFuser's 3593-wide
    Linear layers run at 3648): the padding must stay untouched, the transposing form (32x32
    shared-memory tiles) must agree with the element-wise one."""
    n, k, pn, pk = 203, 77, 256, 128
    w = _rand((n, k), sd, 1)
    wp = torch.zeros((pn, pk), device="cuda", dtype=dd)
    wt = torch.zeros((pk, pn), device="cuda", dtype=dd)
    _check_copy(w, wp[:n, :k])
    _check_copy(w.t(), wt[:k, :n])
    assert wp[n:].abs().sum() == 0 and wp[:, k:].abs().sum() == 0
    assert wt[k:].abs().sum() == 0 and wt[:, n:].abs().sum() == 0
    assert torch.equal(wt[:k, :n], wp[:n, :k].t())
    # accumulate the padded corner of a scratch gradient into a contiguous parameter gradient
    g = _rand((n, k), torch.float32, 2)
    scratch = _rand((pn, pk), torch.float32, 3)
    _check_copy(scratch[:n, :k], g, accumulate=True)


def test_strided_copy_full_size_transpose_3593():
    n = k = 3593
    p = 3648
    w = _rand((n, k), torch.float32, 4)
    wt = torch.zeros((p, p), device="cuda", dtype=torch.bfloat16)
    _check_copy(w.t(), wt[:k, :n])
    bias = torch.zeros((p,), device="cuda")
    _check_copy(_rand((n,), torch.float32, 5), bias[:n])


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
def test_strided_copy_interleave_scale_and_scatter(dt):
    """cat([a, b], -1).flatten(-2, -1) of two [3, nv] features = the [3][2][nv] interleave, with
    IntensityBatchNorm's per-vector factor; and its backward (strided gather, scaled accumulate)."""
    b, v, nv = 5, 2, 512
    m = b * v
    f_init = _rand((m, 3 * nv), dt, 6)
    rotf = _rand((m, 3 * nv), dt, 7)
    sc = _rand((v, 2, nv), torch.float32, 8).abs() + 0.5
    x = torch.full((m, 6 * nv), 7.0, device="cuda", dtype=dt)
    xv = x.view(b, v, 3, 2, nv)
    for kk in range(v):
        _check_copy(f_init.view(b, v, 3, nv)[:, kk], xv[:, kk, :, 0], scale=sc[kk, 0])
        _check_copy(rotf.view(b, v, 3, nv)[:, kk], xv[:, kk, :, 1], scale=sc[kk, 1])
    want = torch.cat([(f_init.float().view(b, v, 3, nv) * sc[:, 0].view(1, v, 1, nv)).to(dt),
                      (rotf.float().view(b, v, 3, nv) * sc[:, 1].view(1, v, 1, nv)).to(dt)], dim=-1)
    assert torch.equal(x, want.reshape(m, 6 * nv))        # every element written exactly once
    # backward: d_init (fp32) += d_x[..., 0, :] * scale
    d_x = _rand((m, 6 * nv), dt, 9).view(b, v, 3, 2, nv)
    d_init = _rand((m, 3 * nv), torch.float32, 10)
    for kk in range(v):
        _check_copy(d_x[:, kk, :, 0], d_init.view(b, v, 3, nv)[:, kk], scale=sc[kk, 0], accumulate=True)
    # unscaled head-input interleave over all rows at once
    y = torch.zeros((m, 6 * nv), device="cuda", dtype=dt)
    _check_copy(f_init.view(m, 3, nv), y.view(m, 3, 2, nv)[:, :, 0])
    _check_copy(rotf.view(m, 3, nv), y.view(m, 3, 2, nv)[:, :, 1])


def test_strided_copy_pair_rotations_and_per_view_outputs():
    """Row (b, view) of the fuser input gets the 9 entries of R_{view <- partner}
    (models/rot_mv.py:193-194,225-231); output assembly gathers one view's rows as fp32."""
    from rotmv_b200 import functional as RF

    b, v, p, off = 6, 2, 3648, 3584
    rot = _rand((b, v, v, 3, 3), torch.float32, 11)
    xin = torch.zeros((b * v, p), device="cuda", dtype=torch.bfloat16)
    RF.strided_copy(rot.view(b, v * v, 9)[:, 1:3], xin.view(b, v, p)[:, :, off:off + 9])
    want = torch.stack([rot[:, 0, 1], rot[:, 1, 0]], dim=1).reshape(b * v, 9).bfloat16()
    assert torch.equal(xin[:, off:off + 9], want)
    assert xin[:, :off].abs().sum() == 0 and xin[:, off + 9:].abs().sum() == 0
    # per-view gather of a row-strided column block (engine._per_view)
    buf = _rand((b * v, 3584), torch.bfloat16, 12)
    t2d = buf[:, 2048:]
    src = t2d.as_strided((b, v, 1536), (v * t2d.stride(0), t2d.stride(0), 1), t2d.storage_offset())
    for kk in range(v):
        o = torch.empty((b, 3, 512), device="cuda")
        RF.strided_copy(src[:, kk], o.view(b, 1536))
        assert torch.equal(o, t2d.float().reshape(b, v, 3, 512)[:, kk])


def test_strided_copy_empty_and_argument_errors():
    from rotmv_b200 import _lib as L
    from rotmv_b200 import functional as RF

    e = torch.empty((0, 8), device="cuda")
    RF.strided_copy(e, torch.empty((0, 8), device="cuda", dtype=torch.bfloat16))      # no launch, no error
    lib = L.load()
    a = torch.zeros((4,), device="cuda")
    rc = lib.rmv_strided_copy(a.data_ptr(), 7, 0, 0, 1, a.data_ptr(), 0, 0, 0, 1, 1, 1, 4, None, 0, None)
    assert rc < 0 and b"dtype" in lib.rmv_last_error()
    rc = lib.rmv_strided_copy(None, 0, 0, 0, 1, a.data_ptr(), 0, 0, 0, 1, 1, 1, 4, None, 0, None)
    assert rc < 0


@pytest.mark.parametrize("dt,rows", [(torch.float32, 7), (torch.bfloat16, 64), (torch.float32, 300)])
def test_intensity_bn_train_matches_reference_formula(dt, rows):
    """models/rot_mv.py:13-32 in train mode, called twice in a row (the running STD of the first call
    feeds the second), on one view's rows of a [b, v, 3, nv] feature."""
    from rotmv_b200 import functional as RF

    v, nv, mom, eps = 2, 512, 0.05, 1e-4
    feat = (_rand((rows, v, 3, nv), torch.float32, 13) * 3.0).to(dt)
    running = torch.ones((1, 1, nv), device="cuda")
    ref_run = running.clone()
    for kk in range(v):
        x = feat[:, kk]                                                  # [rows, 3, nv], row stride v*3*nv
        out = torch.empty((nv,), device="cuda")
        RF.intensity_bn_train(x, running.view(-1), mom, eps, out)
        intensity = torch.norm(x.float(), dim=-2, keepdim=True)
        var = torch.var(intensity, unbiased=False, dim=0, keepdim=True)
        std = torch.sqrt(var.clamp_min(eps))
        ref_run = ref_run * (1 - mom) + std * mom
        torch.cuda.synchronize()
        assert torch.allclose(running, ref_run, rtol=2e-6, atol=1e-7), (running - ref_run).abs().max().item()
        assert torch.allclose(out, 1.0 / (ref_run.view(-1) + eps), rtol=3e-6, atol=1e-7)
    # constant intensity: variance 0 -> clamped at eps (no NaN)
    ones = torch.ones((rows, 3, nv), device="cuda", dtype=dt)
    r2 = torch.ones((nv,), device="cuda")
    o2 = torch.empty((nv,), device="cuda")
    RF.intensity_bn_train(ones, r2, mom, eps, o2)
    want = 1.0 * (1 - mom) + (eps ** 0.5) * mom
    assert torch.allclose(r2, torch.full_like(r2, want), rtol=1e-6) and torch.isfinite(o2).all()


def test_fill_zero_any_alignment_and_size():
    from rotmv_b200 import functional as RF

    raw = torch.full((4096 + 64,), 0xAB, device="cuda", dtype=torch.uint8)
    for off, n in [(0, 4096), (1, 1), (3, 30), (16, 15), (5, 4000), (17, 16), (0, 0)]:
        raw.fill_(0xAB)
        RF.fill_zero(raw[off:off + n])
        torch.cuda.synchronize()
        assert raw[off:off + n].sum().item() == 0
        assert (raw[:off] == 0xAB).all() and (raw[off + n:] == 0xAB).all(), (off, n)
    big = torch.ones((89_600_000 // 8 + 3,), device="cuda")             # an odd number of floats, 44.8 MB
    RF.fill_zero(big)
    assert big.abs().sum().item() == 0
    one = torch.ones((1,), device="cuda")
    RF.fill_zero(one)
    assert one.item() == 0.0
