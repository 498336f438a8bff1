"""Run the recomputed-BatchNorm kernels alone on one layer shape (development aid / ncu target):
rmv_conv_bn_stats, rmv_conv_bn_bwd_reduce, bn_mode 1 and bn_mode 2 convolutions, CUDA-event times.
    N=256 HW=56 C=64 K=256 python scripts/bnconv_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rot-mvgaze_b200"))
import torch
from rotmv_b200 import functional as RF

n, hw, c, k = (int(os.environ.get(x, d)) for x, d in (("N", 256), ("HW", 56), ("C", 64), ("K", 256)))
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((n, hw, hw, c), device="cuda", generator=g).bfloat16()
w = (torch.randn((k, 1, 1, c), device="cuda", generator=g) / c ** 0.5).bfloat16()
dy = torch.randn((n, hw, hw, k), device="cuda", generator=g).bfloat16()
res = torch.randn((n, hw, hw, k), device="cuda", generator=g).bfloat16()
y = torch.empty_like(dy)
bits = torch.zeros((dy.numel() // 8,), device="cuda", dtype=torch.uint8)
acc = torch.zeros((2, k, 2), device="cuda", dtype=torch.float64)
co = [torch.rand((2, k), device="cuda") + 0.5 for _ in range(3)]
flush = torch.empty((256 << 20,), device="cuda", dtype=torch.uint8)


def timed(name, fn, nbytes):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(5):
        flush.zero_()                       # evict the operands from the 126 MB L2
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:34s} {t * 1e3:8.1f} us  {nbytes / t / 1e6:7.0f} GB/s (algorithmic bytes)")


sx, sy = x.numel() * 2, dy.numel() * 2
print(f"[{n},{hw},{hw},{c}] -> {k}")
timed("conv_bn_stats", lambda: RF.conv_bn_stats(x, w, acc), sx)
timed("conv_bn_bwd_reduce", lambda: RF.conv_bn_bwd_reduce(x, w, dy, co[0], co[1], acc), sx + sy)
timed("conv bn_mode 1 (+res, relu, bits)", lambda: RF.conv2d(x, w, residual=res, relu=True, out=y, bn_mode=1, bn_a=co[0], bn_b=co[1], bn_bits=bits), sx + 2 * sy)
timed("conv bn_mode 2 (dz)", lambda: RF.conv2d(x, w, residual=dy, out=y, bn_mode=2, bn_a=co[0], bn_b=co[1], bn_c=co[2]), sx + 2 * sy)
sc = torch.rand((k,), device="cuda") + 0.5
timed("conv scale/shift + res + relu (eval)", lambda: RF.conv2d(x, w, scale=sc, shift=sc, residual=res, relu=True, out=y), sx + 2 * sy)
timed("conv dgrad-like + res + mask", lambda: RF.conv2d(x, w, residual=res, out=y, mask_bits=bits), sx + 2 * sy)
