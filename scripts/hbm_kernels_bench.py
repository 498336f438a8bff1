#!/usr/bin/env python
"""Achieved HBM bandwidth of the memory-bound kernels of the path (north_star: "achieved HBM GB/s for
the fusion and head kernels against B200 peak").

At the bench batch (512 rows) the fusion/head kernels move a few MB and take ~5 us each -- launch
latency, not bandwidth. This script runs each HBM-bound kernel alone on operands larger than the
126 MB L2 (the sizes of a B=2048..65536 sweep point, BASELINE configs[4]) and reports
ALGORITHMIC bytes / CUDA-event time against the measured copy bandwidth of MEASURED_PEAKS.json.

    python scripts/hbm_kernels_bench.py [--reps 20] > profiles/<round>_hbm_kernels.txt
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rot-mvgaze_b200"))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    from rotmv_b200 import _lib as L
    from rotmv_b200 import functional as RF

    L.load()
    dev = torch.device("cuda", 0)
    peak = 6650.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    bf = torch.bfloat16
    rows = []

    def timed(name, nbytes, fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append((name, nbytes / 1e6, ms * 1e3, gbs, gbs / peak))

    g = torch.Generator(device="cuda").manual_seed(0)
    # ---- rotate + gather (models/rot_mv.py:234,238): read (V-1) partner rows + 36 B, write one row
    for b, v in ((131072, 2), (32768, 4), (16384, 8)):
        m = b * v
        feat = torch.randn((m, 1536), device=dev, generator=g).to(bf)
        dst = torch.empty_like(feat)
        rot = RF.pose_to_rotations(torch.rand((b, v, 2), device=dev, generator=g) - 0.5)
        nbytes = m * ((v - 1) * 1536 * 2 + (v - 1) * 36 + 1536 * 2)
        timed(f"rotate_gather bf16 B={b} V={v} (algorithmic: (V-1) partner reads)", nbytes,
              lambda: RF.rotate_gather(feat, rot, dst, b, v, 512, True))
        if v > 2:   # every feature row is read by V-1 rows of the same sample: HBM sees it once
            timed(f"rotate_gather bf16 B={b} V={v} (unique bytes: one read + one write)",
                  m * (1536 * 2 * 2 + (v - 1) * 36),
                  lambda: RF.rotate_gather(feat, rot, dst, b, v, 512, True))
        del feat, dst, rot
    # ---- gaze head tail + pitch-yaw -> vector + angular loss (blocks.py:41-60, gaze_loss.py:42-52)
    m = 1 << 20
    hid = torch.randn((m, 512), device=dev, generator=g).to(bf)
    w2 = torch.randn((2, 512), device=dev, generator=g) * 0.02
    b2 = torch.zeros((2,), device=dev)
    pred = torch.empty((m, 2), device=dev)
    gt = torch.rand((m, 2), device=dev, generator=g) - 0.5
    loss = torch.zeros((1,), device=dev)
    timed(f"head_loss fwd bf16 rows={m} (Linear(512,2) + vector + acos + weighted mean)",
          m * (512 * 2 + 8 + 8), lambda: RF.head_loss(hid, w2, b2, pred, gt, 1e-3, loss, views=2))
    dg = torch.empty_like(hid)
    dpred = torch.empty((m, 2), device=dev)
    gw, gb = torch.zeros((2, 512), device=dev), torch.zeros((2,), device=dev)
    lib = L.load()

    def head_bwd():
        rc = lib.rmv_head_loss_bwd(pred.data_ptr(), gt.data_ptr(), hid.data_ptr(), hid.stride(0),
                                   L.dtype_code(bf), w2.data_ptr(), m, 512, 1e-3, 2, 1.0, dg.data_ptr(),
                                   dg.stride(0), dpred.data_ptr(), gw.data_ptr(), gb.data_ptr(),
                                   L.stream_ptr())
        L.check(rc, "rmv_head_loss_bwd")

    timed(f"head_loss bwd bf16 rows={m} (read hidden, write d hidden, dW2/db2)", m * (512 * 2 * 2 + 16),
          head_bwd)
    del hid, dg
    # ---- pools (models/resnet.py:265,272) at the bench batch
    n = 512
    y = torch.randn((n, 112, 112, 64), device=dev, generator=g).to(bf)
    timed(f"maxpool3x3s2 bf16 [{n},112,112,64]", n * (112 * 112 + 56 * 56) * 64 * 2, lambda: RF.maxpool3x3s2(y))
    del y
    x = torch.randn((4096, 7, 7, 2048), device=dev, generator=g).to(bf)
    o0 = torch.empty((4096, 3584), device=dev, dtype=bf)
    o1 = torch.empty((4096, 3584), device=dev, dtype=bf)
    timed("avgpool bf16 [4096,7,7,2048] -> two fusion buffers", 4096 * (49 + 2) * 2048 * 2,
          lambda: RF.avgpool(x, o0, o1))
    del x, o0, o1

    # ---- round 2: zero fill (zero_grad of the flat gradient), the strided copy of the variants /
    # output assembly, Adam (trainer.py:141,143; models/rot_mv.py:53-85,205-211)
    flat = torch.empty((89_591_366,), device=dev)
    timed("fill_zero fp32 flat gradient (89.6 M elements)", flat.numel() * 4, lambda: RF.fill_zero(flat))
    del flat
    m = 65536
    f0 = torch.randn((m, 1536), device=dev, generator=g).to(bf)
    xi = torch.empty((m, 3072), device=dev, dtype=bf)
    sc = torch.rand((512,), device=dev, generator=g) + 0.5
    timed(f"strided_copy bf16 rows={m}: [3][512] -> slot 0 of the [3][2][512] interleave, scaled",
          m * 1536 * 2 * 2, lambda: RF.strided_copy(f0.view(m, 3, 512), xi.view(m, 3, 2, 512)[:, :, 0], scale=sc))
    o32 = torch.empty((m // 2, 1536), device=dev)
    src = f0.as_strided((m // 2, 2, 1536), (2 * 1536, 1536, 1))
    timed(f"strided_copy bf16 -> fp32 rows={m // 2}: one view's rows (output assembly)",
          (m // 2) * 1536 * (2 + 4), lambda: RF.strided_copy(src[:, 0], o32))
    del f0, xi, o32
    w = torch.randn((3593 * 4, 3593), device=dev, generator=g)
    wt = torch.empty((3648, 3648 * 4), device=dev, dtype=bf)
    timed("strided_copy fp32 -> bf16 transposing [14372,3593] -> padded [3648,14592] (32x32 tiles)",
          w.numel() * (4 + 2), lambda: RF.strided_copy(w.t(), wt[:3593, :3593 * 4]))
    del w, wt

    print(f"# HBM-bound kernels alone, operands > L2, {args.reps} launches each, CUDA events; peak = "
          f"{peak:.0f} GB/s (MEASURED_PEAKS.json copy bandwidth)")
    print(f"{'kernel':88s} {'MB':>9s} {'us':>9s} {'GB/s':>8s} {'frac':>6s}")
    for name, mb, us, gbs, frac in rows:
        print(f"{name:88s} {mb:9.1f} {us:9.1f} {gbs:8.0f} {frac:6.2f}")


if __name__ == "__main__":
    main()
