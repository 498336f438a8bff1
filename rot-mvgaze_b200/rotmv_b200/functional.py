"""Tensor-level wrappers over the C ABI. Activations are NHWC; every function launches on the
current CUDA stream and returns without synchronising."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L


# When set to a list, conv2d/linear append (engine, flops, start_event, end_event) per launch
# (CUDA events on the launching stream) -- used by bench.py for the live roofline measurement.
PROFILE = None


def _call(what, meta, fn, *args):
    """Launch through the C ABI; with PROFILE on, bracket the launch with CUDA events."""
    if PROFILE is None:
        L.check(fn(*args), what)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(fn(*args), what)
    e1.record()
    PROFILE.append((meta.get("engine", what), meta.get("flops", 0.0), e0, e1, what, meta))


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.RotmvError("rotmv_b200 ops need CUDA tensors (there is no CPU path)")


_WORKSPACE = {}


def splitk_workspace(device):
    """Per-device scratch of rmv_splitk_workspace_bytes() for the split-K GEMMs (allocated on first
    use -- the engines' warm-up passes run before any CUDA-graph capture; launches on one stream
    use it one after the other)."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    t = _WORKSPACE.get(key)
    if t is None:
        t = torch.empty((L.load().rmv_splitk_workspace_bytes(),), dtype=torch.uint8, device=device)
        _WORKSPACE[key] = t
    return t


def _conv_args(x, w, out, stride, pad):
    """rmv_conv_args of y = conv(x, w) for NHWC x [N,H,W,C], KRSC w, NHWC out."""
    n, h, wd, c = x.shape
    k, kh, kw, _ = w.shape
    a = L.ConvArgs()
    a.x_dtype = L.dtype_code(x.dtype)
    a.y_dtype = L.dtype_code(out.dtype if out is not None else x.dtype)
    a.x = x.data_ptr()
    a.x_sn, a.x_sh, a.x_sw, a.x_sc = x.stride(0), x.stride(1), x.stride(2), 1
    a.n_img, a.in_h, a.in_w, a.c_in = n, h, wd, c
    a.w = w.data_ptr()
    a.c_out, a.kh, a.kw, a.stride, a.pad = k, kh, kw, stride, pad
    a.out_h = (h + 2 * pad - kh) // stride + 1
    a.out_w = (wd + 2 * pad - kw) // stride + 1
    if out is not None:
        a.y = out.data_ptr()
        a.y_sn, a.y_sh, a.y_sw = out.stride(0), out.stride(1), out.stride(2)
    return a


def conv_bn_stats(x, w, acc, *, stride=1, finalize=None):
    """acc[v][k] += (sum z, sum z^2) of z = conv1x1(x, w), recomputed on the tensor cores (z is never
    written; rmv_conv_bn_stats). x [N,H,W,C] bf16, w [K,1,1,C] bf16, acc fp64 [>=2, K, 2]."""
    _need_cuda(x, w, acc)
    assert w.shape[1] == w.shape[2] == 1 and w.is_contiguous() and w.dtype == x.dtype == torch.bfloat16
    a = _conv_args(x, w, None, stride, 0)
    meta = {}
    if PROFILE is not None:
        n, h, wd, c = x.shape
        meta = {"engine": "tcgen05-bnstat", "flops": 2.0 * n * a.out_h * a.out_w * w.shape[0] * c,
                "bytes": float(n * a.out_h * a.out_w * c * 2),
                "desc": f"conv-bn stats 1x1s{stride} [{n},{h},{wd},{c}]->{w.shape[0]}"}
    _call("rmv_conv_bn_stats", meta, L.load().rmv_conv_bn_stats, C.byref(a), acc.data_ptr(),
          None if finalize is None else C.byref(finalize), L.stream_ptr())


def conv_bn_bwd_reduce(x, w, dy, mean, invstd, acc, *, stride=1, finalize=None):
    """acc[v][k] += (sum dy, sum dy*xhat) with xhat from the recomputed z = conv1x1(x, w)
    (rmv_conv_bn_bwd_reduce); dy [N,OH,OW,K] bf16, already ReLU-masked."""
    _need_cuda(x, w, dy, mean, invstd, acc)
    assert w.shape[1] == w.shape[2] == 1 and w.is_contiguous() and w.dtype == x.dtype == dy.dtype == torch.bfloat16
    a = _conv_args(x, w, None, stride, 0)
    assert tuple(dy.shape) == (x.shape[0], a.out_h, a.out_w, w.shape[0]) and dy.stride(3) == 1
    a.y_sn, a.y_sh, a.y_sw = dy.stride(0), dy.stride(1), dy.stride(2)
    meta = {}
    if PROFILE is not None:
        n, h, wd, c = x.shape
        meta = {"engine": "tcgen05-bnstat", "flops": 2.0 * dy.numel() * c,
                "bytes": float(n * a.out_h * a.out_w * c * 2 + dy.numel() * 2),
                "desc": f"conv-bn bwd reduce 1x1s{stride} [{n},{h},{wd},{c}]->{w.shape[0]}"}
    _call("rmv_conv_bn_bwd_reduce", meta, L.load().rmv_conv_bn_bwd_reduce, C.byref(a), dy.data_ptr(),
          mean.data_ptr(), invstd.data_ptr(), acc.data_ptr(),
          None if finalize is None else C.byref(finalize), L.stream_ptr())


def conv2d(x, w, *, stride=1, pad=0, scale=None, shift=None, residual=None, relu=False,
           out=None, out_dtype=None, engine=L.ENGINE_AUTO, block_n=0, stat_acc=None, stat_views=0,
           bn_mode=0, bn_a=None, bn_b=None, bn_c=None, bn_bits=None, mask_bits=None, stat_finalize=None):
    """y = act(scale * conv(x, w) + shift + residual).

    x: [N, H, W, C] (any pixel strides, channel stride 1); w: [K, kh, kw, C] contiguous, same
    dtype as x; scale/shift: fp32 [K]; residual/out: [N, OH, OW, K].
    Mirrors nn.Conv2d+BatchNorm2d(eval)+ReLU of reference models/resnet.py:128-148.
    """
    _need_cuda(x, w, scale, shift, residual, out)
    n, h, wd, c = x.shape
    k, kh, kw, c2 = w.shape
    assert c == c2 and w.is_contiguous() and w.dtype == x.dtype and x.stride(3) == 1
    oh = (h + 2 * pad - kh) // stride + 1
    ow = (wd + 2 * pad - kw) // stride + 1
    if out is None:
        out = torch.empty((n, oh, ow, k), dtype=out_dtype or x.dtype, device=x.device)
    assert tuple(out.shape) == (n, oh, ow, k) and out.stride(3) == 1
    a = L.ConvArgs()
    a.x_dtype = L.dtype_code(x.dtype)
    a.y_dtype = L.dtype_code(out.dtype)
    a.engine = engine
    a.block_n = block_n
    a.x = x.data_ptr()
    a.x_sn, a.x_sh, a.x_sw, a.x_sc = x.stride(0), x.stride(1), x.stride(2), 1
    a.n_img, a.in_h, a.in_w, a.c_in = n, h, wd, c
    a.w = w.data_ptr()
    a.c_out, a.kh, a.kw, a.stride, a.pad = k, kh, kw, stride, pad
    a.y = out.data_ptr()
    a.y_sn, a.y_sh, a.y_sw = out.stride(0), out.stride(1), out.stride(2)
    a.out_h, a.out_w = oh, ow
    for t in (scale, shift):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.numel() == k)
    a.scale = L.ptr(scale)
    a.shift = L.ptr(shift)
    if residual is not None:
        assert tuple(residual.shape) == tuple(out.shape) and residual.dtype == out.dtype
        assert residual.stride(3) == 1
        a.residual = residual.data_ptr()
        a.r_sn, a.r_sh, a.r_sw = residual.stride(0), residual.stride(1), residual.stride(2)
    a.relu = int(relu)
    if stat_acc is not None:
        # training: per-(view, channel) sum / sum of squares of `out` fused into the conv epilogue
        assert stat_acc.dtype == torch.float64 and stat_acc.is_contiguous() and stat_acc.numel() >= stat_views * k * 2
        a.stat_acc = stat_acc.data_ptr()
        a.stat_views = stat_views
        if stat_finalize is not None:   # L.BnParams: coefficients by the last CTA of this launch
            a.stat_finalize = C.addressof(stat_finalize)
    if bn_mode:
        # recomputed-BatchNorm epilogues (rmv_conv_args.bn_mode): per-(view, channel) tables [2][K]
        for t in (bn_a, bn_b) + ((bn_c,) if bn_mode == 2 else ()):
            assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() >= 2 * k
        a.bn_mode = bn_mode
        a.bn_a, a.bn_b, a.bn_c = bn_a.data_ptr(), bn_b.data_ptr(), L.ptr(bn_c)
        a.bn_bits = L.ptr(bn_bits)
    a.mask_bits = L.ptr(mask_bits)
    if kh == 1 and kw == 1 and x.dtype == torch.bfloat16 and n * oh * ow <= 4096:
        ws = splitk_workspace(x.device)       # small-M pointwise GEMM: the library may split K
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    meta = {}
    if PROFILE is not None:
        tc = x.dtype == torch.bfloat16 and engine != L.ENGINE_SIMT and c % 64 == 0
        es, eo = x.element_size(), out.element_size()
        meta = {"engine": "tcgen05" if tc else "ffma",
                "flops": 2.0 * n * oh * ow * k * kh * kw * c,
                "bytes": float(n * h * wd * c * es / (stride * stride if kh == 1 else 1)
                               + w.numel() * es + n * oh * ow * k * eo * (2 if residual is not None else 1)),
                "desc": f"conv {kh}x{kw}s{stride} [{n},{h},{wd},{c}]->{k}"}
    _call("rmv_conv2d_fwd", meta, L.load().rmv_conv2d_fwd, C.byref(a), L.stream_ptr())
    return out


def conv2d_dgrad(dy, wt, *, stride, pad, in_hw, residual=None, out=None, mask_bits=None, bwd_bn=None):
    """dx (+ residual) of y = conv(x, w, stride, pad) on the tcgen05 engine.

    dy: [N, OH, OW, K] bf16; wt: [C, kh, kw, K] = w reversed and transposed
    (wt[c, kh-1-r, kw-1-s, k] = w[k, c, r, s]); in_hw = (H, W) of x; out/residual: [N, H, W, C].
    stride 2 needs no zero-dilated copy of dy (four parity-class convolutions inside the library).
    The backward of reference models/resnet.py:31-47 (autograd/cuDNN at trainer.py:142)."""
    _need_cuda(dy, wt, residual, out)
    n, oh, ow, k = dy.shape
    c, kh, kw, k2 = wt.shape
    h, w = in_hw
    assert k == k2 and wt.is_contiguous() and wt.dtype == dy.dtype and dy.stride(3) == 1
    if out is None:
        out = torch.empty((n, h, w, c), dtype=dy.dtype, device=dy.device)
    assert tuple(out.shape) == (n, h, w, c) and out.stride(3) == 1
    a = L.ConvArgs()
    a.x_dtype = a.y_dtype = L.dtype_code(dy.dtype)
    a.engine = L.ENGINE_TC
    a.x = dy.data_ptr()
    a.x_sn, a.x_sh, a.x_sw, a.x_sc = dy.stride(0), dy.stride(1), dy.stride(2), 1
    a.n_img, a.in_h, a.in_w, a.c_in = n, oh, ow, k
    a.w = wt.data_ptr()
    a.c_out, a.kh, a.kw, a.stride, a.pad = c, kh, kw, stride, pad
    a.y = out.data_ptr()
    a.y_sn, a.y_sh, a.y_sw = out.stride(0), out.stride(1), out.stride(2)
    a.out_h, a.out_w = h, w
    if residual is not None:
        assert tuple(residual.shape) == tuple(out.shape) and residual.dtype == out.dtype
        a.residual = residual.data_ptr()
        a.r_sn, a.r_sh, a.r_sw = residual.stride(0), residual.stride(1), residual.stride(2)
    a.mask_bits = L.ptr(mask_bits)   # dx is zeroed where the packed ReLU mask of that tensor is 0
    if bwd_bn is not None:
        # bn_mode 4: dx (masked) is also reduced for the BatchNorm backward of the layer that produced
        # the tensor: bwd_bn = dict(z, mean, invstd, acc, finalize) -- z (same geometry as dx) travels as
        # `residual` and is NOT added
        assert residual is None and mask_bits is not None and stride == 1
        z = bwd_bn["z"]
        assert tuple(z.shape) == tuple(out.shape) and z.dtype == out.dtype and z.stride(3) == 1
        a.bn_mode = 4
        a.residual = z.data_ptr()
        a.r_sn, a.r_sh, a.r_sw = z.stride(0), z.stride(1), z.stride(2)
        a.bn_a, a.bn_b = bwd_bn["mean"].data_ptr(), bwd_bn["invstd"].data_ptr()
        a.stat_acc, a.stat_views = bwd_bn["acc"].data_ptr(), 2
        a.stat_finalize = C.addressof(bwd_bn["finalize"])
    meta = {}
    if PROFILE is not None:
        meta = {"engine": "tcgen05", "flops": 2.0 * n * oh * ow * k * kh * kw * c,
                "bytes": float((dy.numel() + wt.numel() + out.numel() * (2 if residual is not None else 1)) * 2),
                "desc": f"dgrad {kh}x{kw}s{stride} [{n},{oh},{ow},{k}]->{c}"}
    _call("rmv_conv2d_dgrad", meta, L.load().rmv_conv2d_dgrad, C.byref(a), L.stream_ptr())
    return out


def conv2d_nchw_input(x_nchw, w, *, stride, pad, scale=None, shift=None, relu=False,
                      out_dtype=None):
    """FFMA conv that consumes the caller's NCHW fp32 tensor directly through strides
    (fp32 parity mode stem; reference models/resnet.py:184-188,262-264)."""
    _need_cuda(x_nchw, w)
    xv = x_nchw.permute(0, 2, 3, 1)  # logical NHWC view of NCHW storage
    n, h, wd, c = xv.shape
    k, kh, kw, _ = w.shape
    oh = (h + 2 * pad - kh) // stride + 1
    ow = (wd + 2 * pad - kw) // stride + 1
    out = torch.empty((n, oh, ow, k), dtype=out_dtype or x_nchw.dtype, device=x_nchw.device)
    a = L.ConvArgs()
    a.x_dtype = L.dtype_code(x_nchw.dtype)
    a.y_dtype = L.dtype_code(out.dtype)
    a.engine = L.ENGINE_SIMT
    a.x = xv.data_ptr()
    a.x_sn, a.x_sh, a.x_sw, a.x_sc = xv.stride(0), xv.stride(1), xv.stride(2), xv.stride(3)
    a.n_img, a.in_h, a.in_w, a.c_in = n, h, wd, c
    a.w = w.data_ptr()
    a.c_out, a.kh, a.kw, a.stride, a.pad = k, kh, kw, stride, pad
    a.y = out.data_ptr()
    a.y_sn, a.y_sh, a.y_sw = out.stride(0), out.stride(1), out.stride(2)
    a.out_h, a.out_w = oh, ow
    a.scale = L.ptr(scale)
    a.shift = L.ptr(shift)
    a.relu = int(relu)
    _call("rmv_conv2d_fwd", {"desc": "rmv_conv2d_fwd"}, L.load().rmv_conv2d_fwd, C.byref(a), L.stream_ptr())
    return out


def linear(x, w, bias=None, *, relu=False, out=None, out_dtype=None, engine=L.ENGINE_AUTO,
           block_n=0):
    """y = act(x @ w.T + bias); x [M, K] (row stride free), w [N, K] contiguous, out [M, N] (row
    stride free). Mirrors nn.Linear(+ReLU) of reference models/backbones/blocks.py:41-60."""
    m, kdim = x.shape
    n = w.shape[0]
    if out is None:
        out = torch.empty((m, n), dtype=out_dtype or x.dtype, device=x.device)
    x4 = x.as_strided((1, 1, m, kdim), (x.stride(0) * max(m, 1), x.stride(0) * max(m, 1), x.stride(0), 1),
                      x.storage_offset())
    o4 = out.as_strided((1, 1, m, n), (out.stride(0) * max(m, 1), out.stride(0) * max(m, 1), out.stride(0), 1),
                        out.storage_offset())
    conv2d(x4, w.view(n, 1, 1, kdim), shift=bias, relu=relu, out=o4, engine=engine, block_n=block_n)
    return out


def stem_im2col(x_nchw, k_pad=192, dtype=torch.bfloat16, kh=7, kw=7, stride=2, pad=3):
    _need_cuda(x_nchw)
    assert x_nchw.dtype == torch.float32 and x_nchw.is_contiguous()
    n, c, h, w = x_nchw.shape
    oh = (h + 2 * pad - kh) // stride + 1
    ow = (w + 2 * pad - kw) // stride + 1
    a = torch.empty((n * oh * ow, k_pad), dtype=dtype, device=x_nchw.device)
    _call("rmv_stem_im2col", {"desc": "rmv_stem_im2col"}, L.load().rmv_stem_im2col, x_nchw.data_ptr(), a.data_ptr(), n, c, h, w, kh, kw, stride,
                                     pad, oh, ow, k_pad, L.dtype_code(dtype), L.stream_ptr())
    return a, oh, ow


def stem_pack_weights(w_oihw):
    """fp32 [64,3,7,7] -> bf16 [64,192] in the k = c*56 + kh*8 + kw order of the fused stem."""
    _need_cuda(w_oihw)
    assert tuple(w_oihw.shape) == (64, 3, 7, 7) and w_oihw.dtype == torch.float32
    w_oihw = w_oihw.contiguous()
    packed = torch.empty((64, 192), dtype=torch.bfloat16, device=w_oihw.device)
    _call("rmv_stem_pack_weights", {"desc": "rmv_stem_pack_weights"},
          L.load().rmv_stem_pack_weights, w_oihw.data_ptr(), packed.data_ptr(), L.stream_ptr())
    return packed


def stem_conv(x_nchw, w_packed, scale, shift, out=None, relu=True):
    """Fused conv7x7/s2 + BN(eval) + ReLU, fp32 NCHW in -> bf16 NHWC out (tcgen05)."""
    _need_cuda(x_nchw, w_packed, scale, shift, out)
    assert x_nchw.dtype == torch.float32 and x_nchw.is_contiguous() and x_nchw.shape[1] == 3
    n, _, h, w = x_nchw.shape
    oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    if out is None:
        out = torch.empty((n, oh, ow, 64), dtype=torch.bfloat16, device=x_nchw.device)
    assert out.is_contiguous() and tuple(out.shape) == (n, oh, ow, 64)
    meta = {"desc": f"stem conv7x7s2+bn+relu [{n},3,{h},{w}]", "engine": "tcgen05-stem",
            "flops": 2.0 * n * oh * ow * 64 * 147,
            "bytes": float(x_nchw.numel() * 4 + out.numel() * 2)}
    _call("rmv_stem_conv_fwd", meta, L.load().rmv_stem_conv_fwd, x_nchw.data_ptr(),
          w_packed.data_ptr(), L.ptr(scale), L.ptr(shift), out.data_ptr(), n, h, w, int(relu),
          L.stream_ptr())
    return out


def stem_conv_u8(x_nhwc_u8, mean, std, w_packed, scale, shift, out=None, relu=True):
    """Fused ToTensor + Normalize + conv7x7/s2 + BN(eval) + ReLU from uint8 HWC images
    [n, H, W, 3] (main.py:38-56 + models/resnet.py:184-188,262-264) -> bf16 NHWC (tcgen05)."""
    _need_cuda(x_nhwc_u8, w_packed, scale, shift, out)
    assert x_nhwc_u8.dtype == torch.uint8 and x_nhwc_u8.is_contiguous() and x_nhwc_u8.shape[3] == 3
    n, h, w, _ = x_nhwc_u8.shape
    oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    if out is None:
        out = torch.empty((n, oh, ow, 64), dtype=torch.bfloat16, device=x_nhwc_u8.device)
    assert out.is_contiguous() and tuple(out.shape) == (n, oh, ow, 64)
    m3 = (C.c_float * 3)(*[float(v) for v in mean])
    s3 = (C.c_float * 3)(*[float(v) for v in std])
    meta = {"desc": f"stem(u8) conv7x7s2+bn+relu [{n},{h},{w},3]", "engine": "tcgen05-stem",
            "flops": 2.0 * n * oh * ow * 64 * 147, "bytes": float(x_nhwc_u8.numel() + out.numel() * 2)}
    _call("rmv_stem_conv_fwd_u8", meta, L.load().rmv_stem_conv_fwd_u8, x_nhwc_u8.data_ptr(), m3, s3,
          w_packed.data_ptr(), L.ptr(scale), L.ptr(shift), out.data_ptr(), n, h, w, int(relu),
          L.stream_ptr())
    return out


def nchw_to_nhwc(x_nchw, dtype):
    _need_cuda(x_nchw)
    assert x_nchw.dtype == torch.float32 and x_nchw.is_contiguous()
    n, c, h, w = x_nchw.shape
    y = torch.empty((n, h, w, c), dtype=dtype, device=x_nchw.device)
    _call("rmv_nchw_to_nhwc", {"desc": "rmv_nchw_to_nhwc"}, L.load().rmv_nchw_to_nhwc, x_nchw.data_ptr(), y.data_ptr(), n, c, h, w,
                                      L.dtype_code(dtype), L.stream_ptr())
    return y


def maxpool3x3s2(x, out=None):
    _need_cuda(x, out)
    assert x.is_contiguous()
    n, h, w, c = x.shape
    oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    y = out if out is not None else torch.empty((n, oh, ow, c), dtype=x.dtype, device=x.device)
    assert tuple(y.shape) == (n, oh, ow, c) and y.dtype == x.dtype and y.is_contiguous()
    _call("rmv_maxpool3x3s2_fwd", {"desc": "rmv_maxpool3x3s2_fwd"}, L.load().rmv_maxpool3x3s2_fwd, x.data_ptr(), y.data_ptr(), n, h, w, c,
                                          L.dtype_code(x.dtype), L.stream_ptr())
    return y


def avgpool(x, out0, out1=None):
    """x [N, H, W, C] contiguous -> out0[:, :C] (and out1[:, :C]); out* are [N, >=C] row-strided."""
    _need_cuda(x, out0, out1)
    assert x.is_contiguous()
    n, h, w, c = x.shape
    assert out0.dtype == x.dtype and (out1 is None or out1.dtype == x.dtype)
    _call("rmv_avgpool_fwd", {"desc": "rmv_avgpool_fwd"}, L.load().rmv_avgpool_fwd, x.data_ptr(), n, h * w, c, L.dtype_code(x.dtype),
                                     out0.data_ptr(), out0.stride(0), L.ptr(out1),
                                     0 if out1 is None else out1.stride(0), L.stream_ptr())


def rotate_gather(feat, rot, dst, batch, views, nvec=512, apply_rot=True, transpose=False):
    """dst[b*V+v, r*nvec+k] = 1/(V-1) sum_{u!=v} sum_c rot[b,v,u,r,c] feat[b*V+u, c*nvec+k].
    feat/dst: [B*V, 3*nvec] row-strided views. Reference models/rot_mv.py:234,238."""
    _need_cuda(feat, rot, dst)
    assert feat.dtype == dst.dtype and feat.stride(1) == 1 and dst.stride(1) == 1
    assert rot.dtype == torch.float32 and rot.is_contiguous()
    assert tuple(rot.shape) == (batch, views, views, 3, 3)
    _call("rmv_rotate_gather_fwd", {"desc": "rmv_rotate_gather_fwd"}, L.load().rmv_rotate_gather_fwd, feat.data_ptr(), feat.stride(0), rot.data_ptr(),
                                           dst.data_ptr(), dst.stride(0), batch, views, nvec,
                                           L.dtype_code(feat.dtype), (3 if transpose else 1) if apply_rot else 0,
                                           L.stream_ptr())


def head_loss(hidden, w2, b2, pred, gt=None, loss_scale=0.0, loss_out=None, views=1,
              aux_decay=1.0):
    _need_cuda(hidden, w2, b2, pred, gt, loss_out)
    rows, hid = hidden.shape
    assert w2.dtype == torch.float32 and w2.is_contiguous() and tuple(w2.shape) == (2, hid)
    assert pred.dtype == torch.float32 and pred.is_contiguous()
    assert gt is None or (gt.dtype == torch.float32 and gt.is_contiguous())
    _call("rmv_head_loss_fwd", {"desc": "rmv_head_loss_fwd"}, L.load().rmv_head_loss_fwd, hidden.data_ptr(), hidden.stride(0),
                                       L.dtype_code(hidden.dtype), w2.data_ptr(), b2.data_ptr(),
                                       rows, hid, pred.data_ptr(), L.ptr(gt), float(loss_scale),
                                       int(views), float(aux_decay), L.ptr(loss_out),
                                       L.stream_ptr())


def angular_error_accum(pred, gt, err_sum):
    _need_cuda(pred, gt, err_sum)
    assert pred.dtype == gt.dtype == err_sum.dtype == torch.float32
    _call("rmv_angular_error_accum", {"desc": "rmv_angular_error_accum"}, L.load().rmv_angular_error_accum, pred.data_ptr(), pred.stride(0), gt.data_ptr(),
                                             gt.stride(0), pred.shape[0], err_sum.data_ptr(),
                                             L.stream_ptr())


def pose_to_rotations(head_pose, out=None):
    """[B, V, 2] (pitch, yaw) -> [B, V, V, 3, 3]; reference utils/math.py:188-219 + rot_mv.py:193-194."""
    _need_cuda(head_pose, out)
    assert head_pose.dtype == torch.float32 and head_pose.is_contiguous()
    b, v, _ = head_pose.shape
    rot = out if out is not None else torch.empty((b, v, v, 3, 3), dtype=torch.float32, device=head_pose.device)
    assert tuple(rot.shape) == (b, v, v, 3, 3) and rot.dtype == torch.float32 and rot.is_contiguous()
    _call("rmv_pose_to_rotations", {"desc": "rmv_pose_to_rotations"}, L.load().rmv_pose_to_rotations, head_pose.data_ptr(), rot.data_ptr(), b, v,
                                           L.stream_ptr())
    return rot


def relative_rotations(rot):
    """[B, V, 3, 3] per-view rotations -> [B, V, V, 3, 3] with [b,i,j] = R_i R_j^T
    (reference models/rot_mv.py:193-194)."""
    _need_cuda(rot)
    rot = rot.float().contiguous()
    b, v = rot.shape[0], rot.shape[1]
    out = torch.empty((b, v, v, 3, 3), dtype=torch.float32, device=rot.device)
    _call("rmv_relative_rotations", {"desc": "rmv_relative_rotations"}, L.load().rmv_relative_rotations, rot.data_ptr(), out.data_ptr(), b, v, L.stream_ptr())
    return out


# ---- re-layout kernels of the constructor variants (encode_rotmat / share_feature) ---------------
def strided_copy(src, dst, scale=None, accumulate=False):
    """dst (+)= src * scale over two equally shaped strided VIEWS of up to three dimensions (fp32 or
    bf16 each, any strides, non-overlapping); `scale` is an fp32 vector over the last dimension.
    One launch of rmv_strided_copy: the padded-corner, interleave and scatter copies of
    ImageRotmatFeatFuser / RotFeatFuser (models/rot_mv.py:53-85,225-248)."""
    _need_cuda(src, dst, scale)
    assert tuple(src.shape) == tuple(dst.shape) and 1 <= src.dim() <= 3, (src.shape, dst.shape)
    pad = 3 - src.dim()
    dims = [1] * pad + [int(d) for d in src.shape]
    ss = [0] * pad + [int(s) for s in src.stride()]
    ds = [0] * pad + [int(s) for s in dst.stride()]
    if scale is not None:
        assert scale.dtype == torch.float32 and scale.is_contiguous() and scale.numel() == dims[2]
    _call("rmv_strided_copy", {"desc": "rmv_strided_copy"}, L.load().rmv_strided_copy,
          src.data_ptr(), L.dtype_code(src.dtype), ss[0], ss[1], ss[2],
          dst.data_ptr(), L.dtype_code(dst.dtype), ds[0], ds[1], ds[2], dims[0], dims[1], dims[2],
          L.ptr(scale), int(bool(accumulate)), L.stream_ptr())
    return dst


def intensity_bn_train(feat, running, momentum, eps, scale_out):
    """Train-mode IntensityBatchNorm statistics of one call (models/rot_mv.py:13-32): feat is a
    [rows, 3, nvec] view (row stride free, the [3, nvec] block contiguous); updates `running` (the
    reference's `running_mean` buffer, fp32, nvec elements) in place and writes the factor
    1 / (running + eps) this call applies into `scale_out` (fp32 [nvec])."""
    _need_cuda(feat, running, scale_out)
    rows, three, nvec = feat.shape
    assert three == 3 and feat.stride(2) == 1 and feat.stride(1) == nvec
    assert running.dtype == scale_out.dtype == torch.float32 and running.is_contiguous() and scale_out.is_contiguous()
    assert running.numel() == nvec and scale_out.numel() == nvec
    _call("rmv_intensity_bn_train", {"desc": "rmv_intensity_bn_train"}, L.load().rmv_intensity_bn_train,
          feat.data_ptr(), feat.stride(0), L.dtype_code(feat.dtype), rows, nvec, running.data_ptr(),
          float(momentum), float(eps), scale_out.data_ptr(), L.stream_ptr())
    return scale_out


def fill_zero(t):
    """t[...] = 0 for a contiguous tensor (rmv_fill_zero; `optimizer.zero_grad()` of trainer.py:141 for
    the flat gradient buffer, padding / accumulator buffers elsewhere)."""
    _need_cuda(t)
    assert t.is_contiguous()
    _call("rmv_fill_zero", {"desc": "rmv_fill_zero"}, L.load().rmv_fill_zero, t.data_ptr(),
          t.numel() * t.element_size(), L.stream_ptr())
    return t
