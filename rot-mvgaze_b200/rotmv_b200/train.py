"""Training step of the Rot-MV path on the GPU: forward with batch-statistic BatchNorm (per-view
statistics), fused angular loss, hand-written backward, fused Adam, optional data-parallel gradient
all-reduce. Mirrors the step body of the reference trainer (trainer.py:119-123,141-143) with the
optimizer of trainer.py:54 (`optim.Adam(lr, weight_decay=1e-6)` = COUPLED L2; `decoupled=True`
gives AdamW as BASELINE config 4 words it).

No autograd: the backward pass is an explicit sequence of librotmv_sm100 launches (dgrad through
the same tcgen05/FFMA convolution kernels with flipped/transposed filters, FFMA weight gradients,
BatchNorm backward reductions, ...). Parameters are re-pointed at one flat fp32 buffer (as are the
gradients and the Adam moments) so the optimizer is one kernel and the DP all-reduce one NCCL call.

precision "fp32": FFMA engine end to end (step parity with torch.optim.Adam on the CPU oracle).
precision "bf16": bf16 activations/filters, tcgen05 forward + data gradients, fp32 master weights.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any, Dict, List, Optional

import torch

from . import _lib as L
from . import functional as RF
from . import parallel as P

_DT = {"bf16": torch.bfloat16, "fp32": torch.float32}


def _ck(name, *args, desc=None):
    RF._call(name, {"desc": desc or name}, getattr(L.load(), name), *args, L.stream_ptr())


def dims4(dims):
    """Pad a shape to four dims on the left (the permute kernels are 4-D)."""
    return (1,) * (4 - len(dims)) + tuple(dims)


class _BN:
    """Per-layer BatchNorm state: parameter views + saved statistics of the last forward."""

    def __init__(self, eng, bn, gamma_g, beta_g):
        c = bn.num_features
        self.c, self.eps, self.momentum = c, float(bn.eps), float(bn.momentum)
        self.gamma, self.beta = bn.weight, bn.bias
        self.rm, self.rv, self.nbt = bn.running_mean, bn.running_var, bn.num_batches_tracked
        self.dgamma, self.dbeta = gamma_g, beta_g
        v, dev = eng.max_views, eng.device
        f = lambda: torch.empty((v, c), device=dev, dtype=torch.float32)  # noqa: E731
        self.mean, self.invstd, self.a, self.b, self.k0, self.k1, self.k2 = (f() for _ in range(7))
        # rmv_bn_params: lets the last thread block of a statistics launch finalize the coefficients
        p = self.params = L.BnParams()
        p.ticket = eng.ticket.data_ptr()
        p.gamma, p.beta = self.gamma.data_ptr(), self.beta.data_ptr()
        p.running_mean, p.running_var, p.num_batches = self.rm.data_ptr(), self.rv.data_ptr(), self.nbt.data_ptr()
        p.mean, p.invstd, p.a, p.b = (t.data_ptr() for t in (self.mean, self.invstd, self.a, self.b))
        p.dgamma, p.dbeta = self.dgamma.data_ptr(), self.dbeta.data_ptr()
        p.k0, p.k1, p.k2 = self.k0.data_ptr(), self.k1.data_ptr(), self.k2.data_ptr()
        p.eps, p.momentum = self.eps, self.momentum


class TrainEngine:
    def __init__(self, model, precision: str = "bf16", lr: float = 1e-6, betas=(0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 1e-6, decoupled: bool = False,
                 process_group=None, max_views: int = 8):
        if precision not in _DT:
            raise ValueError("precision must be 'bf16' or 'fp32'")
        L.load()
        self.model, self.precision, self.dt = model, precision, _DT[precision]
        self.dtc = L.dtype_code(self.dt)
        trunk = model._feat_extractor[0]
        # constructor variants, two views only like the reference: encode_rotmat
        # (ImageRotmatFeatFuser, models/rot_mv.py:53-67,225-231) and share_feature (RotFeatFuser +
        # IntensityBatchNorm, :13-32,70-85,201-203,243-248)
        self.encode_rot = bool(getattr(model, "_encode_rotmat", False)) and not model._ignore_rotmat
        self.share_feat = bool(getattr(model, "_share_feature", False))
        self.device = trunk.conv1.weight.device
        if self.device.type != "cuda":
            raise L.RotmvError("TrainEngine needs the model on a CUDA device; there is no CPU path")
        self.max_views = max_views
        self.use_tc_wgrad = True  # bf16: weight gradients on tcgen05 (False -> FFMA kernel)
        # bf16, V == 2: BatchNorm statistics in the conv epilogue (RMV_FUSE_BN_STATS=0: separate pass)
        self.fuse_bn_stats = os.environ.get("RMV_FUSE_BN_STATS", "1") != "0"
        # bf16, V == 2, bottleneck blocks: the BatchNorm of the expanding 1x1 convolutions (conv3,
        # downsample) never materialises the conv output -- statistics and apply passes recompute it on
        # the tensor cores (RMV_BN_RECOMPUTE=0: the conv writes z and the HBM-bound passes read it)
        self.recompute_bn = os.environ.get("RMV_BN_RECOMPUTE", "1") != "0" and trunk.kind == "bottleneck"
        # ... where the recomputed GEMM is cheap next to the tensor it avoids: reduction depth c_in <= 256
        # (layer1-3). Measured per block (four passes, B=128, kernels alone, cold L2): 56^2 (K=64) 586 us
        # recomputed vs 777 us materialised; 14^2 (K=256) 219 vs ~237; 7^2 (K=512) 203 vs ~133 -- at 7^2
        # the passes are bound by re-streaming the 2 MB of filters per tile, not by the 51 MB tensor.
        # Whole step, same box: threshold 512 -> 20.10 ms, 128 -> 20.16, 256 -> 19.97.
        # RMV_BN_RECOMPUTE_MAXC overrides the threshold.
        self.recompute_max_cin = int(os.environ.get("RMV_BN_RECOMPUTE_MAXC", "256"))
        # bf16, V == 2: the backward reduction of the mid-layer BatchNorms (bn1, bn2) runs inside the
        # epilogue of the data-gradient kernel that produces their dy (bn_mode 4), except behind the
        # halo 3x3 kernel (no shared memory left) and stride-2 data gradients (RMV_BN_BWD_FUSE=1: on).
        # Measured (B=128, one lease, interleaved): 19.80 / 19.75 ms without, 19.85 / 19.68 ms with -- 26
        # launches less but no time: the fused form only saves ONE read of dy (S/4 per layer), the z read
        # and the mask read move into the data-gradient epilogue. Kept as a tested variant, off by default.
        self.fuse_bn_bwd = os.environ.get("RMV_BN_BWD_FUSE", "0") != "0"
        # data parallel: all-reduce the fusion-stage gradients while the trunk backward runs
        # (RMV_DP_OVERLAP=0: one all-reduce of the whole buffer after the backward pass)
        self.dp_overlap = os.environ.get("RMV_DP_OVERLAP", "1") != "0"
        self.num_iter, self.fc_dim, self.nvec = model._num_iter, model._fc_dim, model._num_feat_vec
        self.apply_rot = not model._ignore_rotmat
        self.decoupled = bool(decoupled)
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        # ---- flat parameter / gradient / moment buffers (fc.* never gets a gradient: SURVEY Q4) ----
        names, offs, total = P.flat_layout(model.named_parameters())
        params = dict(model.named_parameters())
        named = [(n, params[n]) for n in names]
        self.flat_p = torch.zeros(total, device=self.device, dtype=torch.float32)
        self.flat_g = torch.zeros_like(self.flat_p)
        self.flat_m = torch.zeros_like(self.flat_p)
        self.flat_v = torch.zeros_like(self.flat_p)
        self.grads: Dict[int, torch.Tensor] = {}
        self.names: List[str] = []
        for (n, p), o in zip(named, offs):
            view = self.flat_p[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            self.grads[id(p)] = self.flat_g[o:o + p.numel()].view(p.shape)
            self.names.append(n)
        self.n_trained = sum(p.numel() for _, p in named)
        # data-parallel overlap: flat_g[:grad_split] = trunk gradients (written last, by the trunk
        # backward), flat_g[grad_split:] = lifter/fuser/head gradients (74 % of the bytes, complete
        # before the trunk backward starts) -- the second slice is all-reduced while the trunk
        # backward runs
        self.grad_split = next((o for (n, _), o in zip(named, offs) if not n.startswith("_feat_extractor.")), total)
        # gradient buckets in the order the backward pass completes them (parallel.gradient_buckets):
        # `forward_backward` reports each through its hook as soon as the last kernel writing into it
        # has been launched
        self.buckets = P.gradient_buckets(names, offs, total)
        self._bucket_by_stage = {4: self.buckets[1], 3: self.buckets[2]}
        self.hyper = torch.tensor([lr, betas[0], betas[1], eps, weight_decay, 0.0],
                                  device=self.device, dtype=torch.float64)
        self.ticket = torch.zeros((1,), device=self.device, dtype=torch.int32)
        # ---- layer table ----
        g = self.grads
        self.stem_bn = _BN(self, trunk.bn1, g[id(trunk.bn1.weight)], g[id(trunk.bn1.bias)])
        self.stem_conv = trunk.conv1
        self.blocks = []
        for blk in trunk.blocks():
            # Bottleneck (models/resnet.py:99-148): 1x1, 3x3(stride), 1x1; BasicBlock (:50-96): 3x3(stride), 3x3
            convs = [blk.conv1, blk.conv2, blk.conv3] if trunk.kind == "bottleneck" else [blk.conv1, blk.conv2]
            bns = [blk.bn1, blk.bn2, blk.bn3] if trunk.kind == "bottleneck" else [blk.bn1, blk.bn2]
            e = {"convs": convs,
                 "bns": [_BN(self, b, g[id(b.weight)], g[id(b.bias)]) for b in bns]}
            if blk.downsample is not None:
                e["ds_conv"] = blk.downsample[0]
                d = blk.downsample[1]
                e["ds_bn"] = _BN(self, d, g[id(d.weight)], g[id(d.bias)])
            self.blocks.append(e)
        lif = model._lifter._lifter.blocks
        self.lift = [lif[0][0], lif[1][0]]
        self.fusers = [[blk[0] for blk in m._fuser.blocks] for m in model._img_fusers]
        self.heads = [[m.blocks[0][0], m.blocks[1][0]] for m in model._gaze_estimators]
        c_max = max(2048, self.fc_dim)
        assert L.load().rmv_bn_workspace_bytes(c_max, max_views) == max_views * c_max * 2 * 8
        self.acc = torch.zeros((max_views, c_max, 2), device=self.device, dtype=torch.float64)
        self._bufs: Dict[Any, torch.Tensor] = {}
        self._bits: Dict[int, Optional[torch.Tensor]] = {}   # id(ReLU output) -> packed mask
        self.loss = torch.zeros((1,), device=self.device, dtype=torch.float32)
        self.launches_last_step = 0
        self._wjobs, self._wjob_tags, self._wjobs_ready = [], set(), False
        self.last_feats = None   # views of the last forward's features (train-mode dict contract)
        # first block index of every trunk stage (layer1..layer4), for the per-stage gradient buckets
        self._stage_first, bi = {}, 0
        for li in range(1, 5):
            self._stage_first[bi] = li
            bi += len(getattr(trunk, f"layer{li}"))
        self.sync_replicas()

    def sync_replicas(self, src: int = 0) -> None:
        """Data parallel: make rank `src`'s parameters, Adam moments, hyper-parameters (incl. the
        step count) and BatchNorm buffers the state of every replica. Called at construction and
        after loading a checkpoint -- the gradient all-reduce averages gradients, which is only
        meaningful when all replicas hold the same weights."""
        if self.world > 1:
            P.broadcast_state_((self.flat_p, self.flat_m, self.flat_v, self.hyper), self.model,
                               src=src, group=self.pg)

    # ------------------------------------------------------------------------------------------
    def set_lr(self, lr: float) -> None:
        self.hyper[0:1].copy_(torch.tensor([lr], dtype=torch.float64), non_blocking=True)

    def _buf(self, tag, shape, dtype=None):
        key = (tag, tuple(shape), dtype or self.dt)
        t = self._bufs.get(key)
        if t is None:
            t = torch.empty(shape, device=self.device, dtype=dtype or self.dt)
            self._bufs[key] = t
        return t

    # ---- filter layout transforms (fp32 OIHW master -> engine layouts) ------------------------
    # The first step launches one rmv_permute_cast per tensor and records the jobs; from the second
    # step on, ONE batched launch at the top of the step re-derives all of them (the masters only
    # change in the Adam kernel at the end of a step) and these helpers just return the buffers.
    def _transform(self, src, tag, dims, strides, flip, kind=0):
        """`kind` (rmv_permute_job.kind) names the access pattern the batched kernel uses for this
        job; the first step runs the generic kernel, so both are exercised against each other."""
        out = self._buf(tag, dims)
        if self._wjobs_ready:
            return out
        job = (src.data_ptr(), out.data_ptr(), *[int(d) for d in dims4(dims)], *strides, flip, flip,
               self.dtc, kind)
        if tag not in self._wjob_tags:
            self._wjob_tags.add(tag)
            self._wjobs.append(job)
        d = dims4(dims)
        _ck("rmv_permute_cast", src.data_ptr(), out.data_ptr(), d[0], d[1], d[2], d[3], *strides,
            flip, flip, self.dtc)
        return out

    def _w_fwd(self, conv, tag):
        k, c, r, s = conv.weight.shape
        return self._transform(conv.weight, ("wf", tag), (k, r, s, c), (c * r * s, s, 1, r * s), 0,
                               kind=1 if r * s == 1 else 3)

    def _w_dgrad(self, conv, tag):
        """Filters of the data-gradient convolution: wt[c][r'][s'][k] = w[k][c][R-1-r'][S-1-s']."""
        k, c, r, s = conv.weight.shape
        return self._transform(conv.weight, ("wd", tag), (c, r, s, k), (r * s, s, 1, c * r * s), 1, kind=2)

    def _lin_fwd(self, lin, tag):
        n, k = lin.weight.shape
        return self._transform(lin.weight, ("lf", tag), (n, k), (0, 0, k, 1), 0, kind=1)

    def _lin_t(self, lin, tag):
        n, k = lin.weight.shape
        # as a (C,R,S,K) = (k,1,1,n) transpose job; the buffer is the [k, n] matrix
        out = self._transform(lin.weight, ("lt", tag), (k, 1, 1, n), (1, 0, 0, k), 0, kind=2)
        return out.view(k, n)

    def _finish_wjobs(self):
        """Upload the recorded job table (once)."""
        jobs = (L.PermuteJob * len(self._wjobs))()
        block = 0
        for i, (src, dst, d0, d1, d2, d3, s0, s1, s2, s3, f1, f2, dt, kind) in enumerate(self._wjobs):
            j = jobs[i]
            j.src, j.dst, j.d0, j.d1, j.d2, j.d3 = src, dst, d0, d1, d2, d3
            j.s0, j.s1, j.s2, j.s3, j.flip1, j.flip2, j.dst_dtype = s0, s1, s2, s3, f1, f2, dt
            j.first_block, j.kind = block, kind
            if kind == 2:      # (C,R,S,K): 64x64 tiles of the [K][C*R*S] matrix
                block += ((d0 * d1 * d2 + 63) // 64) * ((d3 + 63) // 64)
            elif kind == 3:    # (K,R,S,C): one block per k
                block += d0
            else:
                block += (d0 * d1 * d2 * d3 + 1023) // 1024
        raw = bytes(jobs)
        self._wjob_table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device)
        self._wjob_count, self._wjob_blocks = len(self._wjobs), block
        self._wjobs_ready = True

    def _run_wjobs(self):
        RF._call("rmv_permute_cast_batch", {"desc": "rmv_permute_cast_batch"},
                 L.load().rmv_permute_cast_batch, self._wjob_table.data_ptr(), self._wjob_count,
                 self._wjob_blocks, L.stream_ptr())

    # ---- BatchNorm -----------------------------------------------------------------------------
    def _conv_stats(self, x, w, *, stride=1, pad=0, out=None, bn=None):
        """Forward conv of the training step. bf16 / two views: the BatchNorm batch statistics of
        the output are accumulated by the conv epilogue itself and turned into the coefficients of
        `bn` by the last thread block of the launch (returns stats_done=True)."""
        fused = self.fuse_bn_stats and self.precision == "bf16" and self.views == 2
        # measured (B=128, V=2): free for the tensor-bound 3x3 and the reducing 1x1 convs, but the
        # expanding 1x1 convs are HBM-bound with the epilogue on the critical path (+70 % with the
        # statistics in it) -- those keep the separate one-wave reduction
        if w.shape[1] == 1 and w.shape[0] > w.shape[3]:
            fused = False
        fused = fused and bn is not None and x.shape[0] % 2 == 0
        z = RF.conv2d(x, w, stride=stride, pad=pad, out=out,
                      stat_acc=self.acc if fused else None, stat_views=2 if fused else 0,
                      stat_finalize=bn.params if fused else None)
        return z, fused

    def _bn_fwd(self, bn: _BN, z, residual, relu, tag, stats_done=False):
        n, h, w, c = z.shape
        v = self.views
        shp = f" [{n},{h},{w},{c}]" if RF.PROFILE is not None else ""
        if not stats_done:   # else: statistics AND coefficients came out of the conv launch itself
            _ck("rmv_bn_stats_finalize", z.data_ptr(), self.dtc, n, h * w, c, v, self.acc.data_ptr(),
                self.ticket.data_ptr(), bn.gamma.data_ptr(), bn.beta.data_ptr(), bn.rm.data_ptr(),
                bn.rv.data_ptr(), bn.nbt.data_ptr(), bn.mean.data_ptr(), bn.invstd.data_ptr(),
                bn.a.data_ptr(), bn.b.data_ptr(), bn.eps, bn.momentum, desc="rmv_bn_stats_finalize" + shp)
        y = self._buf(tag, z.shape)
        # packed ReLU mask (1 bit/element) for the backward pass instead of re-reading y there
        bits = self._buf(("bits", tag), (n * h * w * c // 8,), torch.uint8) if relu else None
        _ck("rmv_bn_apply", z.data_ptr(), bn.a.data_ptr(), bn.b.data_ptr(), L.ptr(residual),
            y.data_ptr(), L.ptr(bits), self.dtc, n, h * w, c, v, int(relu), desc="rmv_bn_apply" + shp)
        self._bits[id(y)] = bits
        return y

    def _bn_bwd(self, bn: _BN, z, dy, mask, tag, want_dyr=False):
        """`mask` is the ReLU output tensor `_bn_fwd` returned (its packed bit mask is used) or None."""
        n, h, w, c = z.shape
        v = self.views
        is_bits = 0
        if mask is not None and self._bits.get(id(mask)) is not None:
            mask, is_bits = self._bits[id(mask)], 1
        shp = f" [{n},{h},{w},{c}]" if RF.PROFILE is not None else ""
        _ck("rmv_bn_bwd_reduce_finalize", z.data_ptr(), dy.data_ptr(), L.ptr(mask), is_bits,
            bn.mean.data_ptr(), bn.invstd.data_ptr(), self.dtc, n, h * w, c, v, self.acc.data_ptr(),
            self.ticket.data_ptr(), bn.gamma.data_ptr(), bn.dgamma.data_ptr(), bn.dbeta.data_ptr(),
            bn.k0.data_ptr(), bn.k1.data_ptr(), bn.k2.data_ptr(), desc="rmv_bn_bwd_reduce_finalize" + shp)
        dz = self._buf(("dz", tag), z.shape)
        dyr = self._buf(("dyr", tag), z.shape) if want_dyr else None
        _ck("rmv_bn_bwd_apply", z.data_ptr(), dy.data_ptr(), L.ptr(mask), is_bits, bn.k0.data_ptr(),
            bn.k1.data_ptr(), bn.k2.data_ptr(), dz.data_ptr(), L.ptr(dyr), self.dtc, n, h * w, c, v,
            desc="rmv_bn_bwd_apply" + shp)
        return dz, dyr

    # ---- BatchNorm over a recomputed 1x1 convolution (conv3 / downsample of the bottlenecks) ------
    def _use_recompute(self, conv) -> bool:
        return (self.recompute_bn and self.precision == "bf16" and self.views == 2
                and conv.kernel_size[0] == 1 and conv.out_channels % 128 == 0 and conv.in_channels % 64 == 0
                and conv.in_channels <= self.recompute_max_cin)

    def _conv_bn_fwd(self, bn: _BN, x, w, stride, residual, relu, tag):
        """y = relu?(BN_train(conv1x1(x, w)) + residual) without writing the conv output: statistics
        from the transposed recompute GEMM (rmv_conv_bn_stats, coefficients by its last thread block), then
        the conv again with the BatchNorm-apply epilogue (bn_mode 1), which also packs the ReLU mask."""
        n, h, wd, _ = x.shape
        oh, ow = (h - 1) // stride + 1, (wd - 1) // stride + 1
        c, v = w.shape[0], self.views
        RF.conv_bn_stats(x, w, self.acc, stride=stride, finalize=bn.params)   # sums + coefficients
        y = self._buf(tag, (n, oh, ow, c))
        bits = self._buf(("bits", tag), (n * oh * ow * c // 8,), torch.uint8) if relu else None
        RF.conv2d(x, w, stride=stride, residual=residual, relu=relu, out=y, bn_mode=1, bn_a=bn.a,
                  bn_b=bn.b, bn_bits=bits)
        self._bits[id(y)] = bits
        return y

    def _conv_bn_bwd(self, bn: _BN, x, w, stride, dy, tag):
        """dz of z = conv1x1(x, w) under y = BN_train(z) for a dy that is ALREADY multiplied by the
        derivative of the ReLU behind the BatchNorm: reductions over the recomputed z
        (rmv_conv_bn_bwd_reduce, coefficients by its last thread block), then the conv again with the
        backward-apply epilogue (bn_mode 2: dz = k0*dy + k1*z + k2)."""
        n, oh, ow, c = dy.shape
        v = self.views
        RF.conv_bn_bwd_reduce(x, w, dy, bn.mean, bn.invstd, self.acc, stride=stride, finalize=bn.params)
        dz = self._buf(("dz", tag), dy.shape)
        RF.conv2d(x, w, stride=stride, residual=dy, out=dz, bn_mode=2, bn_a=bn.k0, bn_b=bn.k1, bn_c=bn.k2)
        return dz

    # ---- convolution gradients -----------------------------------------------------------------
    def _wgrad(self, x, dy, conv_or_lin, kh, kw, stride, pad, x_strides=None, grad=None):
        """dW (fp32, parameter layout) += sum_p dy[p,k] x[p+(r,s), c]. `grad` overrides the
        destination (default: the parameter's slice of the flat gradient buffer)."""
        a = L.ConvArgs()
        a.x_dtype = L.dtype_code(x.dtype)
        a.x = x.data_ptr()
        if x_strides is None:
            a.x_sn, a.x_sh, a.x_sw, a.x_sc = x.stride(0), x.stride(1), x.stride(2), 1
            n, h, w, c = x.shape
        else:
            (n, h, w, c), (a.x_sn, a.x_sh, a.x_sw, a.x_sc) = x_strides
        a.n_img, a.in_h, a.in_w, a.c_in = n, h, w, c
        a.c_out, a.kh, a.kw, a.stride, a.pad = dy.shape[3], kh, kw, stride, pad
        a.y_sn, a.y_sh, a.y_sw = dy.stride(0), dy.stride(1), dy.stride(2)
        a.out_h, a.out_w = dy.shape[1], dy.shape[2]
        assert L.dtype_code(dy.dtype) == a.x_dtype
        if grad is None:
            grad = self.grads[id(conv_or_lin.weight)]
        if self.use_tc_wgrad and x.dtype == torch.bfloat16 and c % 64 == 0 and a.x_sc == 1:
            # tcgen05 weight gradient into an fp32 [k][r][s][c] buffer; 1x1 / Linear: that IS the
            # parameter layout, so accumulate straight into the (zeroed) flat gradient slice.
            k = dy.shape[3]
            meta = {"desc": f"wgrad-tc {kh}x{kw}s{stride} [{n},{h},{w},{c}]->{k}", "engine": "tcgen05-wgrad",
                    "flops": 2.0 * n * dy.shape[1] * dy.shape[2] * k * kh * kw * c}
            if kh * kw == 1:
                RF._call("rmv_conv2d_wgrad_tc", meta, L.load().rmv_conv2d_wgrad_tc, C.byref(a),
                         dy.data_ptr(), grad.data_ptr(), L.stream_ptr())
            else:
                assert L.load().rmv_conv2d_wgrad_tc_workspace_bytes(C.byref(a)) == k * kh * kw * c * 4
                scratch = self._buf(("wg", id(conv_or_lin)), (k, kh, kw, c), torch.float32)
                self._zero(scratch)
                RF._call("rmv_conv2d_wgrad_tc", meta, L.load().rmv_conv2d_wgrad_tc, C.byref(a),
                         dy.data_ptr(), scratch.data_ptr(), L.stream_ptr())
                _ck("rmv_permute_cast", scratch.data_ptr(), grad.data_ptr(), k, c, kh, kw,
                    kh * kw * c, 1, kw * c, c, 0, 0, L.F32)
            return
        RF._call("rmv_conv2d_wgrad", {"desc": f"wgrad {kh}x{kw}s{stride} [{n},{h},{w},{c}]->{dy.shape[3]}",
                                      "engine": "ffma-wgrad",
                                      "flops": 2.0 * n * dy.shape[1] * dy.shape[2] * dy.shape[3] * kh * kw * c},
                 L.load().rmv_conv2d_wgrad, C.byref(a), dy.data_ptr(), grad.data_ptr(), L.stream_ptr())

    def _dgrad(self, dz, conv, tag, in_shape, residual=None, mask_bits=None, bwd_bn=None):
        """dx = conv_transpose(dz, w) (+ residual) via the forward kernel with flipped filters;
        `mask_bits` (bf16 engine): packed ReLU mask of the tensor dx is the gradient of -- dx is
        zeroed where that ReLU was inactive."""
        wt = self._w_dgrad(conv, tag)
        r = conv.kernel_size[0]
        stride, pad = conv.stride[0], conv.padding[0]
        if self.precision == "bf16":
            return RF.conv2d_dgrad(dz, wt, stride=stride, pad=pad, in_hw=in_shape[1:3],
                                   residual=residual, out=self._buf(("dx", tag), in_shape),
                                   mask_bits=mask_bits, bwd_bn=bwd_bn)
        assert mask_bits is None and bwd_bn is None
        src = dz
        if stride == 2:
            n, oh, ow, k = dz.shape
            src = self._buf(("dil", tag), (n, 2 * oh, 2 * ow, k))
            _ck("rmv_dilate2", dz.data_ptr(), src.data_ptr(), n, oh, ow, k, self.dtc)
        elif stride != 1:
            raise NotImplementedError("stride > 2")
        out = self._buf(("dx", tag), in_shape)
        assert src.shape[1] == in_shape[1] and src.shape[2] == in_shape[2], (src.shape, in_shape)
        return RF.conv2d(src, wt, stride=1, pad=r - 1 - pad, residual=residual, out=out)

    def _lin(self, x, w, bias, relu, out):
        return RF.linear(x, w, bias, relu=relu, out=out)

    def _lin_bwd(self, lin, tag, x, dy, dx_out, need_dx=True):
        """bias grad, weight grad and (optionally) input grad of y = x W^T + b."""
        m, n = dy.shape
        self._colsum(dy, n, self.grads[id(lin.bias)])
        x4 = x.as_strided((1, 1, m, x.shape[1]), (0, 0, x.stride(0), 1), x.storage_offset())
        d4 = dy.as_strided((1, 1, m, n), (0, 0, dy.stride(0), 1), dy.storage_offset())
        self._wgrad(x4, d4, lin, 1, 1, 1, 0)
        if need_dx:
            RF.linear(dy, self._lin_t(lin, tag), None, out=dx_out)
        return dx_out

    def _add(self, src, dst, mask=None, add=None):
        """dst = (mask > 0 ? src : 0) + add   (2-D, row-strided)."""
        rows, cols = src.shape
        _ck("rmv_relu_bwd", src.data_ptr(), src.stride(0), L.ptr(mask), 0 if mask is None else mask.stride(0),
            L.ptr(add), 0 if add is None else add.stride(0), dst.data_ptr(), dst.stride(0), rows, cols,
            self.dtc)

    # ------------------------------------------------------------------------------------------
    def step(self, images: torch.Tensor, rotations: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
        """One optimisation step; returns the (device) loss tensor of this step's forward pass."""
        before = L.STATS["launches"]
        ar = P.OverlappedAllReduce(self.flat_g, self.pg)
        n = self.flat_g.numel()
        if ar.active and self.dp_overlap:   # every bucket is all-reduced as soon as it is final
            self.forward_backward(images, rotations, gt, hook=lambda name, b, e: ar.start(b, e))
        else:
            self.forward_backward(images, rotations, gt)
            ar.start(0, n)
        ar.finish()   # the current stream waits for NCCL's stream; the host does not block
        _ck("rmv_adam_step", self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.flat_m.data_ptr(),
            self.flat_v.data_ptr(), self.hyper.data_ptr(), self.flat_p.numel(), int(self.decoupled),
            1.0 / self.world)
        self.launches_last_step = L.STATS["launches"] - before
        self.model.invalidate()   # cached inference weights / graph sessions are stale now
        return self.loss

    # ---- fusion stage: forward + loss + backward down to d(loss)/d(pooled image feature) ----------
    def _fusion_default(self, x, rot, gt_flat, b, v, cfg):
        """ImageFeatFuser configuration (the one main.py builds; models/rot_mv.py:35-50,198-254)."""
        m = b * v
        dt, dtc = self.dt, self.dtc
        wide = self.fc_dim + 3 * self.nvec
        fd = self.fc_dim
        xs = [self._buf(("X", i), (m, wide)) for i in range(self.num_iter)]
        ys = [self._buf(("Y", i), (m, wide)) for i in range(self.num_iter)]
        hs = [self._buf(("H", i), (m, wide)) for i in range(self.num_iter)]
        gs = [self._buf(("G", i), (m, 512)) for i in range(self.num_iter)]
        preds = [self._buf(("pred", i), (m, 2), torch.float32) for i in range(self.num_iter)]
        y_init = self._buf("Yinit", (m, wide))
        RF.avgpool(x, y_init, xs[0])
        img = y_init[:, :fd]
        for i in range(self.num_iter):
            if i > 0:
                self._add(img, xs[i][:, :fd])
            self._add(img, ys[i][:, :fd])
        l1 = self._buf("L1", (m, 3 * self.nvec))
        self._lin(img, self._lin_fwd(self.lift[0], "l0"), self.lift[0].bias, True, l1)
        self._lin(l1, self._lin_fwd(self.lift[1], "l1"), self.lift[1].bias, False, y_init[:, fd:])
        scales = []
        for i in range(self.num_iter):
            f_prev = (y_init if i == 0 else ys[i - 1])[:, fd:]
            RF.rotate_gather(f_prev, rot, xs[i][:, fd:], b, v, self.nvec, self.apply_rot)
            f1, f2 = self.fusers[i]
            self._lin(xs[i], self._lin_fwd(f1, ("f1", i)), f1.bias, True, hs[i])
            self._lin(hs[i], self._lin_fwd(f2, ("f2", i)), f2.bias, False, ys[i][:, fd:])
            h1, h2 = self.heads[i]
            self._lin(ys[i], self._lin_fwd(h1, ("h1", i)), h1.bias, True, gs[i])
            scale = (cfg["iter_decay"] ** (self.num_iter - 1 - i)) * cfg["rel_weight"] / b
            scales.append(scale)
            RF.head_loss(gs[i], h2.weight.detach(), h2.bias.detach(), preds[i], gt_flat, scale,
                         self.loss, views=v, aux_decay=cfg["reference_decay"])

        self.last_feats = {"img": img, "init": y_init[:, fd:], "iters": [y[:, fd:] for y in ys]}
        # ================================ backward ================================
        ext = yield preds   # None: gradient of the fused loss; else d(loss)/d(pred) per iteration
        dimg = self._buf("dimg", (m, fd))
        self._zero(dimg)
        d_f_next = None
        for i in reversed(range(self.num_iter)):
            h1, h2 = self.heads[i]
            f1, f2 = self.fusers[i]
            dg = self._buf("dG", (m, 512))
            dpred = self._buf("dpred", (m, 2), torch.float32)
            if ext is None:
                self._head_bwd(preds[i], gt_flat, gs[i], h2, scales[i], v, cfg, dg, dpred)
            else:
                self._head_bwd_ext(ext[i], gs[i], h2, dg)
            d_y = self._lin_bwd(h1, ("h1", i), ys[i], dg, self._buf("dY", (m, wide)))
            self._add(d_y[:, :fd], dimg, add=dimg)
            d_f = self._buf("dF", (m, 3 * self.nvec))
            self._add(d_y[:, fd:], d_f, add=d_f_next)
            d_h = self._lin_bwd(f2, ("f2", i), hs[i], d_f, self._buf("dH", (m, wide)))
            self._add(d_h, d_h, mask=hs[i])
            d_x = self._lin_bwd(f1, ("f1", i), xs[i], d_h, self._buf("dX", (m, wide)))
            self._add(d_x[:, :fd], dimg, add=dimg)
            d_f_next = self._buf("dFn", (m, 3 * self.nvec))
            RF.rotate_gather(d_x[:, fd:], rot, d_f_next, b, v, self.nvec, self.apply_rot, transpose=True)
        d_l1 = self._lin_bwd(self.lift[1], "l1", l1, d_f_next, self._buf("dL1", (m, 3 * self.nvec)))
        self._add(d_l1, d_l1, mask=l1)
        d_img2 = self._lin_bwd(self.lift[0], "l0", img, d_l1, self._buf("dimg2", (m, fd)))
        self._add(d_img2, dimg, add=dimg)
        return dimg, preds

    # ---- constructor variants (two views) --------------------------------------------
    # The GEMMs, weight gradients, rotation gathers, head/loss kernels are the sm_100a kernels of the
    # default configuration. What these variants add -- zero-padding the 3593-wide layers to a
    # multiple of 64, the 9 rotation entries per row, the [3][2][512] interleave of RotFeatFuser and
    # IntensityBatchNorm's 512-element statistics -- runs in the re-layout kernels of
    # csrc/variant_glue.cu (`rmv_strided_copy` over strided views, `rmv_intensity_bn_train`,
    # `rmv_fill_zero`); tests/test_train_fusion_host.py replaces the three methods below by their
    # torch formulas.
    def _zero(self, t) -> None:
        RF.fill_zero(t)

    def _scopy(self, src, dst, scale=None, accumulate=False) -> None:
        """dst (+)= src * scale[last dim] for equally shaped strided views (<= 3-D, fp32 or bf16 each)."""
        RF.strided_copy(src, dst, scale, accumulate)

    def _intensity_scale(self, bn, feat_bv, out):
        """Train-mode IntensityBatchNorm (models/rot_mv.py:13-32) of one call: feat_bv [B, 3, 512].
        Updates the running STD (buffer `running_mean`) from the batch and writes the per-vector
        scale 1 / (running_std + eps) (fp32 [512]) that this call applies into `out`; the norm is
        detached in the reference, so the scale is a constant of the backward pass."""
        return RF.intensity_bn_train(feat_bv, bn.running_mean.view(-1), bn.momentum, bn.eps, out)

    def _colsum(self, dy, n, dst):
        """dst[:n] += column sums of dy[:, :n] (bias gradient)."""
        _ck("rmv_colsum", dy.data_ptr(), dy.stride(0), dy.shape[0], n, self.dtc, dst.data_ptr())

    def _head_bwd(self, pred, gt_flat, g, h2, scale, views, cfg, dg, dpred):
        """Backward of the fused head tail + loss (rmv_head_loss_bwd): d hidden, d pred, dW2, db2."""
        _ck("rmv_head_loss_bwd", pred.data_ptr(), gt_flat.data_ptr(), g.data_ptr(), g.stride(0), self.dtc,
            h2.weight.data_ptr(), g.shape[0], g.shape[1], scale, views, cfg["reference_decay"],
            dg.data_ptr(), dg.stride(0), dpred.data_ptr(), self.grads[id(h2.weight)].data_ptr(),
            self.grads[id(h2.bias)].data_ptr())

    def _head_bwd_ext(self, dpred, g, h2, dg):
        """Backward of the head's last Linear(512,2) and the ReLU before it for an EXTERNAL
        d(loss)/d(pred) [m,2] (the autograd bridge: the caller's own loss objects produced it):
        rmv_head_loss_bwd with gt == NULL reads `dpred` instead of computing the loss gradient."""
        dp = self._buf("dpred_ext", (g.shape[0], 2), torch.float32)
        dp.copy_(dpred.detach().reshape(g.shape[0], 2))
        _ck("rmv_head_loss_bwd", None, None, g.data_ptr(), g.stride(0), self.dtc, h2.weight.data_ptr(),
            g.shape[0], g.shape[1], 0.0, 1, 1.0, dg.data_ptr(), dg.stride(0), dp.data_ptr(),
            self.grads[id(h2.weight)].data_ptr(), self.grads[id(h2.bias)].data_ptr())

    def _padded_linear(self, lin, tag, pk, pn):
        """bf16/fp32 copies of lin.weight zero-padded to [pn, pk] and its transpose [pk, pn], the
        bias zero-padded to [pn] (refreshed every step: the fp32 masters move in the Adam kernel)."""
        n, k = lin.weight.shape
        key = ("padw", tag)
        if key not in self._bufs:
            self._bufs[key] = (torch.empty((pn, pk), device=self.device, dtype=self.dt),
                               torch.empty((pk, pn), device=self.device, dtype=self.dt),
                               torch.empty((pn,), device=self.device, dtype=torch.float32))
            for t in self._bufs[key]:
                self._zero(t)              # the padding stays zero; only the corners are rewritten
        w, wt, bias = self._bufs[key]
        src = lin.weight.detach()
        self._scopy(src, w[:n, :k])
        self._scopy(src.t(), wt[:k, :n])   # 32x32 shared-memory tiles (source contiguous along dim 0)
        self._scopy(lin.bias.detach(), bias[:n])
        return w, wt, bias

    def _padded_lin_bwd(self, lin, wt, x, dy, dx_out):
        """Backward of y = x Wp^T + bp on zero-padded operands: x [m, pk], dy [m, pn] (padding
        columns are zero). The weight gradient is formed at the padded size in an fp32 scratch and its
        [n, k] corner added to the parameter's gradient (+=: share_weights aliases one fuser)."""
        n, k = lin.weight.shape
        m, pn = dy.shape
        pk = x.shape[1]
        self._colsum(dy, n, self.grads[id(lin.bias)])
        scratch = self._buf(("padg", pn, pk), (pn, pk), torch.float32)
        self._zero(scratch)
        x4 = x.as_strided((1, 1, m, pk), (0, 0, x.stride(0), 1), x.storage_offset())
        d4 = dy.as_strided((1, 1, m, pn), (0, 0, dy.stride(0), 1), dy.storage_offset())
        self._wgrad(x4, d4, lin, 1, 1, 1, 0, grad=scratch)
        self._scopy(scratch[:n, :k], self.grads[id(lin.weight)], accumulate=True)
        if dx_out is not None:
            RF.linear(dy, wt, None, out=dx_out)
        return dx_out

    def _fusion_encode_rotmat(self, x, rot, gt_flat, b, cfg):
        """ImageRotmatFeatFuser (models/rot_mv.py:53-67,225-231): the partner feature is NOT rotated;
        the 9 entries of R_{self<-partner} are appended to the fuser input instead. Fuser = three
        Linear layers of width fc+1536+9 (3593 for ResNet-50), run zero-padded to a multiple of 64."""
        v, m = 2, 2 * b
        dt, dtc = self.dt, self.dtc
        fd, nv3, n_it = self.fc_dim, 3 * self.nvec, self.num_iter
        wide = fd + nv3
        w_true = wide + 9
        p = (w_true + 63) // 64 * 64
        xs = [self._buf(("Xe", i), (m, p)) for i in range(n_it)]
        h1s = [self._buf(("He1", i), (m, p)) for i in range(n_it)]
        h2s = [self._buf(("He2", i), (m, p)) for i in range(n_it)]
        ys = [self._buf(("Y", i), (m, wide)) for i in range(n_it)]
        gs = [self._buf(("G", i), (m, 512)) for i in range(n_it)]
        preds = [self._buf(("pred", i), (m, 2), torch.float32) for i in range(n_it)]
        y_init = self._buf("Yinit", (m, wide))
        for t in xs:
            self._zero(t)
        RF.avgpool(x, y_init, xs[0])
        img = y_init[:, :fd]
        # row (b, view) gets R_{view <- partner}: rot[b,0,1] = rot_10, rot[b,1,0] = rot_01 (:193-194)
        pair = rot.view(b, v * v, 9)[:, 1:3]                   # [b, view, 9], strides (36, 9, 1)
        for i in range(n_it):
            if i > 0:
                self._add(img, xs[i][:, :fd])
            self._add(img, ys[i][:, :fd])
            self._scopy(pair, xs[i].view(b, v, p)[:, :, wide:w_true])
        l1 = self._buf("L1", (m, nv3))
        self._lin(img, self._lin_fwd(self.lift[0], "l0"), self.lift[0].bias, True, l1)
        self._lin(l1, self._lin_fwd(self.lift[1], "l1"), self.lift[1].bias, False, y_init[:, fd:])
        scales, padded = [], []
        for i in range(n_it):
            f_prev = (y_init if i == 0 else ys[i - 1])[:, fd:]
            RF.rotate_gather(f_prev, rot, xs[i][:, fd:wide], b, v, self.nvec, False)   # partner, unrotated
            f1, f2, f3 = self.fusers[i]
            pw = [self._padded_linear(f1, ("e1", i), p, p), self._padded_linear(f2, ("e2", i), p, p),
                  self._padded_linear(f3, ("e3", i), p, nv3)]
            padded.append(pw)
            self._lin(xs[i], pw[0][0], pw[0][2], True, h1s[i])
            self._lin(h1s[i], pw[1][0], pw[1][2], True, h2s[i])
            self._lin(h2s[i], pw[2][0], pw[2][2], False, ys[i][:, fd:])
            h1, h2 = self.heads[i]
            self._lin(ys[i], self._lin_fwd(h1, ("h1", i)), h1.bias, True, gs[i])
            scale = (cfg["iter_decay"] ** (n_it - 1 - i)) * cfg["rel_weight"] / b
            scales.append(scale)
            RF.head_loss(gs[i], h2.weight.detach(), h2.bias.detach(), preds[i], gt_flat, scale,
                         self.loss, views=v, aux_decay=cfg["reference_decay"])
        self.last_feats = {"img": img, "init": y_init[:, fd:], "iters": [y[:, fd:] for y in ys]}
        # backward
        ext = yield preds   # None: gradient of the fused loss; else d(loss)/d(pred) per iteration
        dimg = self._buf("dimg", (m, fd))
        self._zero(dimg)
        d_f_next = None
        for i in reversed(range(n_it)):
            h1, h2 = self.heads[i]
            f1, f2, f3 = self.fusers[i]
            pw = padded[i]
            dg = self._buf("dG", (m, 512))
            dpred = self._buf("dpred", (m, 2), torch.float32)
            if ext is None:
                self._head_bwd(preds[i], gt_flat, gs[i], h2, scales[i], v, cfg, dg, dpred)
            else:
                self._head_bwd_ext(ext[i], gs[i], h2, dg)
            d_y = self._lin_bwd(h1, ("h1", i), ys[i], dg, self._buf("dY", (m, wide)))
            self._add(d_y[:, :fd], dimg, add=dimg)
            d_f = self._buf("dF", (m, nv3))
            self._add(d_y[:, fd:], d_f, add=d_f_next)
            d_h2 = self._padded_lin_bwd(f3, pw[2][1], h2s[i], d_f, self._buf("dHe2", (m, p)))
            self._add(d_h2, d_h2, mask=h2s[i])
            d_h1 = self._padded_lin_bwd(f2, pw[1][1], h1s[i], d_h2, self._buf("dHe1", (m, p)))
            self._add(d_h1, d_h1, mask=h1s[i])
            d_x = self._padded_lin_bwd(f1, pw[0][1], xs[i], d_h1, self._buf("dXe", (m, p)))
            self._add(d_x[:, :fd], dimg, add=dimg)
            d_f_next = self._buf("dFn", (m, nv3))
            RF.rotate_gather(d_x[:, fd:wide], rot, d_f_next, b, v, self.nvec, False, transpose=True)
        d_l1 = self._lin_bwd(self.lift[1], "l1", l1, d_f_next, self._buf("dL1", (m, nv3)))
        self._add(d_l1, d_l1, mask=l1)
        d_img2 = self._lin_bwd(self.lift[0], "l0", img, d_l1, self._buf("dimg2", (m, fd)))
        self._add(d_img2, dimg, add=dimg)
        return dimg, preds

    def _fusion_share_feature(self, x, rot, gt_flat, b, cfg):
        """share_feature=True (models/rot_mv.py:70-85,160-171,201-203,243-248): the lifted feature
        replaces the image feature; fuser input = cat(BN(F_init), BN(R F_partner), -1).flatten() of
        two [3,512] features = the [3][2][512] interleave, three Linear layers of width 3072; head
        input = cat(F_init, F_new, -1).flatten(). IntensityBatchNorm runs in train mode: four
        sequential running-std updates per iteration (view 0: own, rotated partner; view 1: same)."""
        v, m = 2, 2 * b
        dt, dtc = self.dt, self.dtc
        fd, nv, n_it = self.fc_dim, self.nvec, self.num_iter
        nv3, w6 = 3 * nv, 6 * nv
        xs = [self._buf(("Xs", i), (m, w6)) for i in range(n_it)]
        h1s = [self._buf(("Hs1", i), (m, w6)) for i in range(n_it)]
        h2s = [self._buf(("Hs2", i), (m, w6)) for i in range(n_it)]
        ys = [self._buf(("Ys", i), (m, w6)) for i in range(n_it)]
        fs = [self._buf(("Fs", i), (m, nv3)) for i in range(n_it)]
        gs = [self._buf(("G", i), (m, 512)) for i in range(n_it)]
        preds = [self._buf(("pred", i), (m, 2), torch.float32) for i in range(n_it)]
        img = self._buf("img", (m, fd))
        RF.avgpool(x, img, None)
        l1 = self._buf("L1", (m, nv3))
        f_init = self._buf("Finit", (m, nv3))
        self._lin(img, self._lin_fwd(self.lift[0], "l0"), self.lift[0].bias, True, l1)
        self._lin(l1, self._lin_fwd(self.lift[1], "l1"), self.lift[1].bias, False, f_init)
        fi4 = f_init.view(b, v, 3, nv)
        rotf = self._buf("rotF", (m, nv3))
        scales, bn_scales = [], []
        for i in range(n_it):
            f_prev = f_init if i == 0 else fs[i - 1]
            RF.rotate_gather(f_prev, rot, rotf, b, v, nv, self.apply_rot)       # R_{self<-partner} F_partner
            r4 = rotf.view(b, v, 3, nv)
            bn = self.model._img_fusers[i]._batchnorm
            xv = xs[i].view(b, v, 3, 2, nv)
            sc = self._buf(("isc", i), (v, 2, nv), torch.float32)
            for k in range(v):                                                   # view order, as the reference
                self._intensity_scale(bn, fi4[:, k], sc[k, 0])
                self._scopy(fi4[:, k], xv[:, k, :, 0], scale=sc[k, 0])
                self._intensity_scale(bn, r4[:, k], sc[k, 1])
                self._scopy(r4[:, k], xv[:, k, :, 1], scale=sc[k, 1])
            bn_scales.append(sc)
            f1, f2, f3 = self.fusers[i]
            self._lin(xs[i], self._lin_fwd(f1, ("f1", i)), f1.bias, True, h1s[i])
            self._lin(h1s[i], self._lin_fwd(f2, ("f2", i)), f2.bias, True, h2s[i])
            self._lin(h2s[i], self._lin_fwd(f3, ("f3", i)), f3.bias, False, fs[i])
            yv = ys[i].view(m, 3, 2, nv)
            self._scopy(f_init.view(m, 3, nv), yv[:, :, 0])
            self._scopy(fs[i].view(m, 3, nv), yv[:, :, 1])
            h1, h2 = self.heads[i]
            self._lin(ys[i], self._lin_fwd(h1, ("h1", i)), h1.bias, True, gs[i])
            scale = (cfg["iter_decay"] ** (n_it - 1 - i)) * cfg["rel_weight"] / b
            scales.append(scale)
            RF.head_loss(gs[i], h2.weight.detach(), h2.bias.detach(), preds[i], gt_flat, scale,
                         self.loss, views=v, aux_decay=cfg["reference_decay"])
        # share_feature: the lifted feature IS the "image feature" of the dict (models/rot_mv.py:201-203)
        self.last_feats = {"img": f_init, "init": f_init, "iters": list(fs)}
        # backward
        ext = yield preds   # None: gradient of the fused loss; else d(loss)/d(pred) per iteration
        d_init = self._buf("dInit", (m, nv3), torch.float32)                       # d loss / d F_init
        self._zero(d_init)
        d_f_next = None
        for i in reversed(range(n_it)):
            h1, h2 = self.heads[i]
            f1, f2, f3 = self.fusers[i]
            dg = self._buf("dG", (m, 512))
            dpred = self._buf("dpred", (m, 2), torch.float32)
            if ext is None:
                self._head_bwd(preds[i], gt_flat, gs[i], h2, scales[i], v, cfg, dg, dpred)
            else:
                self._head_bwd_ext(ext[i], gs[i], h2, dg)
            d_y = self._lin_bwd(h1, ("h1", i), ys[i], dg, self._buf("dYs", (m, w6))).view(m, 3, 2, nv)
            self._scopy(d_y[:, :, 0], d_init.view(m, 3, nv), accumulate=True)
            d_f = self._buf("dF", (m, nv3))
            self._scopy(d_y[:, :, 1], d_f.view(m, 3, nv))
            if d_f_next is not None:
                self._scopy(d_f_next, d_f, accumulate=True)
            d_h2 = self._lin_bwd(f3, ("f3", i), h2s[i], d_f, self._buf("dHs2", (m, w6)))
            self._add(d_h2, d_h2, mask=h2s[i])
            d_h1 = self._lin_bwd(f2, ("f2", i), h1s[i], d_h2, self._buf("dHs1", (m, w6)))
            self._add(d_h1, d_h1, mask=h1s[i])
            d_x = self._lin_bwd(f1, ("f1", i), xs[i], d_h1, self._buf("dXs", (m, w6))).view(b, v, 3, 2, nv)
            sc = bn_scales[i]
            d_rotf = self._buf("dRotF", (m, nv3))
            for k in range(v):
                self._scopy(d_x[:, k, :, 0], d_init.view(b, v, 3, nv)[:, k], scale=sc[k, 0], accumulate=True)
                self._scopy(d_x[:, k, :, 1], d_rotf.view(b, v, 3, nv)[:, k], scale=sc[k, 1])
            d_f_next = self._buf("dFn", (m, nv3))
            RF.rotate_gather(d_rotf, rot, d_f_next, b, v, nv, self.apply_rot, transpose=True)
        d_total = self._buf("dFinit", (m, nv3))
        self._scopy(d_f_next, d_init, accumulate=True)    # iteration 0 gathered F_init itself
        self._scopy(d_init, d_total)
        d_l1 = self._lin_bwd(self.lift[1], "l1", l1, d_total, self._buf("dL1", (m, nv3)))
        self._add(d_l1, d_l1, mask=l1)
        dimg = self._lin_bwd(self.lift[0], "l0", img, d_l1, self._buf("dimg", (m, fd)))
        return dimg, preds

    def forward_backward(self, images, rotations, gt, hook=None) -> Dict[str, Any]:
        """Forward + loss + backward into the flat gradient buffer. `hook(name, begin, end)` is called
        once per gradient bucket (`self.buckets`: fusion stage, layer4, layer3, layer2..stem), right
        after the last kernel writing into flat_g[begin:end] has been launched: the data-parallel step
        starts that bucket's all-reduce there, the graph-captured step splits its graphs there."""
        gen = self._fwd_bwd(images, rotations, gt, hook)
        next(gen)                 # forward + fused loss
        try:
            gen.send(None)        # backward from the fused loss
        except StopIteration as done:
            return done.value
        raise RuntimeError("forward_backward: the step generator yielded twice")

    def _fwd_bwd(self, images, rotations, gt, hook=None):
        """The step as a generator: runs the forward, yields the per-iteration predictions
        (list of [B*V, 2] fp32), and on `send(ext)` runs the backward -- from the fused loss when
        `ext` is None (`gt` required), from the caller's d(loss)/d(pred) list otherwise (`gt` may be
        None: rotmv_b200.module's autograd bridge). Returns {"loss", "preds", "pool_out"}."""
        self._require_device(images)
        b, v = images.shape[0], images.shape[1]
        if v > self.max_views:
            raise ValueError(f"views={v} exceeds max_views={self.max_views}")
        m = b * v
        self.views = v
        dt, dtc = self.dt, self.dtc
        cfg = self.model.loss_cfg
        imgs = images.reshape(m, *images.shape[2:]).float().contiguous()
        rot = rotations.float().contiguous()
        gt_flat = None if gt is None else gt.float().reshape(m, 2).contiguous()
        self._begin_step()

        # ================================ forward ================================
        # stem (models/resnet.py:262-265)
        z0 = self._stem_fwd(imgs)
        y0 = self._bn_fwd(self.stem_bn, z0, None, True, "y_stem")
        x, pool_idx = self._maxpool_fwd(y0)
        pool_out = x
        saved = []
        for bi, e in enumerate(self.blocks):
            convs, bns = e["convs"], e["bns"]
            x_in = x
            # conv -> (statistics in the epilogue) -> finalize -> apply, strictly in this order: the
            # fp64 accumulator `self.acc` is shared and reset by every finalize
            t, zs, ys = x_in, [], []
            for si in range(len(convs) - 1):
                cv = convs[si]
                k, st, pd = cv.kernel_size[0], cv.stride[0], cv.padding[0]
                oh = (t.shape[1] + 2 * pd - k) // st + 1
                z, sd = self._conv_stats(t, self._w_fwd(cv, (bi, si + 1)), stride=st, pad=pd, bn=bns[si],
                                         out=self._buf((f"z{si + 1}", bi), (m, oh, oh, cv.out_channels)))
                t = self._bn_fwd(bns[si], z, None, True, (f"y{si + 1}", bi), stats_done=sd)
                zs.append(z); ys.append(t)
            oh = t.shape[1]
            zd = None
            if "ds_conv" in e:
                dc = e["ds_conv"]
                if self._use_recompute(dc):
                    skip = self._conv_bn_fwd(e["ds_bn"], x_in, self._w_fwd(dc, (bi, "d")), dc.stride[0],
                                             None, False, ("skip", bi))
                else:
                    zd, sd = self._conv_stats(x_in, self._w_fwd(dc, (bi, "d")), stride=dc.stride[0], bn=e["ds_bn"],
                                              out=self._buf(("zd", bi), (m, oh, oh, dc.out_channels)))
                    skip = self._bn_fwd(e["ds_bn"], zd, None, False, ("skip", bi), stats_done=sd)
            else:
                skip = x_in
            cv = convs[-1]
            if self._use_recompute(cv):
                zs.append(None)   # never materialised
                x = self._conv_bn_fwd(bns[-1], t, self._w_fwd(cv, (bi, len(convs))), 1, skip, True, ("out", bi))
            else:
                z, sd = self._conv_stats(t, self._w_fwd(cv, (bi, len(convs))), stride=cv.stride[0],
                                         pad=cv.padding[0], bn=bns[-1],
                                         out=self._buf((f"z{len(convs)}", bi), (m, oh, oh, cv.out_channels)))
                zs.append(z)
                x = self._bn_fwd(bns[-1], z, skip, True, ("out", bi), stats_done=sd)
            saved.append((x_in, zs, ys, zd, x))
        if self.encode_rot or self.share_feat:
            if v != 2:
                raise NotImplementedError("encode_rotmat / share_feature are two-view configurations "
                                          "(the reference defines nothing else)")
            fusion = self._fusion_encode_rotmat if self.encode_rot else self._fusion_share_feature
            dimg, preds = yield from fusion(x, rot, gt_flat, b, cfg)
        else:
            dimg, preds = yield from self._fusion_default(x, rot, gt_flat, b, v, cfg)
        if hook is not None:
            hook(*self.buckets[0])
        # trunk
        last = saved[-1][4]
        d_out = self._avgpool_bwd(dimg, last)
        pre_masked = False   # d_out already multiplied by the ReLU derivative of the block output
        if self.precision == "bf16" and self._bits.get(id(last)) is not None and self._use_recompute(self.blocks[-1]["convs"][-1]):
            self._mask_grad(d_out, self._bits[id(last)])
            pre_masked = True
        for bi in reversed(range(len(self.blocks))):
            e = self.blocks[bi]
            convs, bns = e["convs"], e["bns"]
            x_in, zs, ys, zd, out = saved[bi]
            n_st = len(convs)
            # last conv of the block: its BatchNorm also yields the gradient of the skip branch
            if zs[-1] is None:    # recomputed BatchNorm: d_out arrives masked, dyr IS d_out
                assert pre_masked
                dz = self._conv_bn_bwd(bns[-1], ys[-1], self._w_fwd(convs[-1], (bi, n_st)), 1, d_out, (bi, n_st))
                dyr = d_out
            else:
                dz, dyr = self._bn_bwd(bns[-1], zs[-1], d_out, out, (bi, n_st), want_dyr=True)
            # the data gradient this block hands to the previous one is multiplied by that block's ReLU
            # derivative (its packed mask) when that block's BatchNorm backward expects it that way
            in_bits = None
            if bi > 0 and self._use_recompute(self.blocks[bi - 1]["convs"][-1]):
                in_bits = self._bits.get(id(x_in))     # the previous block runs the recomputed BatchNorm
            pre_masked = in_bits is not None
            for si in reversed(range(n_st)):
                cv = convs[si]
                k, st, pd = cv.kernel_size[0], cv.stride[0], cv.padding[0]
                src = x_in if si == 0 else ys[si - 1]        # input of conv si
                self._wgrad(src, dz, cv, k, k, st, pd)
                if si > 0:
                    bn_lo, z_lo, y_lo = bns[si - 1], zs[si - 1], ys[si - 1]
                    bits_lo = self._bits.get(id(y_lo))
                    halo = k == 3 and st == 1 and cv.in_channels == 64 and cv.out_channels == 64
                    if (self.fuse_bn_bwd and self.precision == "bf16" and self.views == 2 and st == 1
                            and not halo and bits_lo is not None and cv.in_channels % 64 == 0):
                        # the data gradient arrives masked and already reduced (bn_mode 4): apply only
                        dy = self._dgrad(dz, cv, (bi, si + 1), src.shape, mask_bits=bits_lo,
                                         bwd_bn={"z": z_lo, "mean": bn_lo.mean, "invstd": bn_lo.invstd,
                                                 "acc": self.acc, "finalize": bn_lo.params})
                        dz = self._buf(("dz", (bi, si)), z_lo.shape)
                        n_, h_, w_, c_ = z_lo.shape
                        _ck("rmv_bn_bwd_apply", z_lo.data_ptr(), dy.data_ptr(), None, 0, bn_lo.k0.data_ptr(),
                            bn_lo.k1.data_ptr(), bn_lo.k2.data_ptr(), dz.data_ptr(), None, self.dtc, n_,
                            h_ * w_, c_, self.views,
                            desc="rmv_bn_bwd_apply" + (f" [{n_},{h_},{w_},{c_}]" if RF.PROFILE is not None else ""))
                    else:
                        dy = self._dgrad(dz, cv, (bi, si + 1), src.shape)
                        dz, _ = self._bn_bwd(bn_lo, z_lo, dy, y_lo, (bi, si))
                else:
                    if "ds_conv" in e:
                        dc = e["ds_conv"]
                        if zd is None:
                            dzd = self._conv_bn_bwd(e["ds_bn"], x_in, self._w_fwd(dc, (bi, "d")), dc.stride[0],
                                                    dyr, (bi, "d"))
                        else:
                            dzd, _ = self._bn_bwd(e["ds_bn"], zd, dyr, None, (bi, "d"))
                        self._wgrad(x_in, dzd, dc, 1, 1, dc.stride[0], 0)
                        res = self._dgrad(dzd, dc, (bi, "d"), x_in.shape)
                    else:
                        res = dyr
                    d_out = self._dgrad(dz, cv, (bi, 1), x_in.shape, residual=res, mask_bits=in_bits)
            if hook is not None and self._stage_first.get(bi) in self._bucket_by_stage:
                hook(*self._bucket_by_stage[self._stage_first[bi]])   # this stage's gradients are final
        d_y0 = self._maxpool_bwd(pool_idx, d_out, y0)
        dz0, _ = self._bn_bwd(self.stem_bn, z0, d_y0, y0, "stem")
        self._stem_wgrad(imgs, dz0)
        if hook is not None:
            hook(*self.buckets[3])
        self._end_step()
        return {"loss": self.loss, "preds": preds, "pool_out": pool_out}

    # ---- the trunk's non-convolution steps as tensor-level methods (the kernels behind them are raw
    # C-ABI calls; tests/test_train_trunk_host.py replaces these methods by their torch formulas to
    # check the orchestration of `_fwd_bwd` on the CPU) -----------------------------------------------
    def _require_device(self, images) -> None:
        if not images.is_cuda:
            raise L.RotmvError("images must be CUDA tensors (there is no CPU path)")

    def _begin_step(self) -> None:
        self._zero(self.flat_g)     # optimizer.zero_grad() (trainer.py:141)
        self._zero(self.loss)
        if self._wjobs_ready:
            self._run_wjobs()

    def _end_step(self) -> None:
        if not self._wjobs_ready:
            self._finish_wjobs()

    def _stem_fwd(self, imgs):
        """conv7x7/s2 of the fp32 NCHW images -> z0 NHWC (models/resnet.py:184-186,262), no BatchNorm yet."""
        m = imgs.shape[0]
        if self.precision == "bf16":
            wp = RF.stem_pack_weights(self.stem_conv.weight.detach())
            return RF.stem_conv(imgs, wp, None, None, out=self._buf("z_stem", (m, 112, 112, 64))
                                if imgs.shape[2] == 224 and imgs.shape[3] == 224 else None, relu=False)
        return RF.conv2d_nchw_input(imgs, self._w_fwd(self.stem_conv, "stem"), stride=2, pad=3)

    def _stem_wgrad(self, imgs, dz0) -> None:
        m = imgs.shape[0]
        if self.precision == "bf16":   # tcgen05 stem weight gradient straight from the fp32 NCHW images
            RF._call("rmv_stem_wgrad", {"desc": "stem wgrad (tcgen05)", "engine": "tcgen05-wgrad",
                                        "flops": 2.0 * m * dz0.shape[1] * dz0.shape[2] * 64 * 147},
                     L.load().rmv_stem_wgrad, imgs.data_ptr(), dz0.data_ptr(),
                     self._buf("stem_wg", (L.load().rmv_stem_wgrad_workspace_bytes() // 4,),
                               torch.float32).data_ptr(),
                     self.grads[id(self.stem_conv.weight)].data_ptr(), m, imgs.shape[2], imgs.shape[3],
                     L.stream_ptr())
        else:
            xv = imgs.permute(0, 2, 3, 1)
            self._wgrad(imgs, dz0, self.stem_conv, 7, 7, 2, 3,
                        x_strides=(tuple(xv.shape), (xv.stride(0), xv.stride(1), xv.stride(2), xv.stride(3))))

    def _maxpool_fwd(self, y0):
        """MaxPool2d(3, 2, 1) with the winning window position recorded for the backward."""
        m, h, w, c = y0.shape
        ph, pw = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        x = self._buf("pool", (m, ph, pw, c))
        idx = self._buf("pool_idx", (m, ph, pw, c), torch.uint8)
        _ck("rmv_maxpool3x3s2_fwd_idx", y0.data_ptr(), x.data_ptr(), idx.data_ptr(), m, h, w, c, self.dtc)
        return x, idx

    def _maxpool_bwd(self, idx, d_out, y0):
        d_y0 = self._buf("dy_stem", y0.shape)
        _ck("rmv_maxpool3x3s2_bwd_idx", idx.data_ptr(), d_out.data_ptr(), d_y0.data_ptr(), y0.shape[0],
            y0.shape[1], y0.shape[2], y0.shape[3], self.dtc)
        return d_y0

    def _avgpool_bwd(self, dimg, last):
        """d(block output) of the global average pool: dimg [m, C] / (H*W) broadcast over the pixels."""
        d_out = self._buf(("dx", "avg"), last.shape)
        _ck("rmv_avgpool_bwd", dimg.data_ptr(), dimg.stride(0), d_out.data_ptr(), last.shape[0],
            last.shape[1] * last.shape[2], last.shape[3], self.dtc)
        return d_out

    def _mask_grad(self, d, bits) -> None:
        """d *= ReLU derivative recorded as a packed sign mask (in place)."""
        _ck("rmv_mask_bits", d.data_ptr(), bits.data_ptr(), d.data_ptr(), d.numel(), self.dtc)


class GraphedTrainStep:
    """CUDA-graph-captured training step (north_star: "trainer.py step loop (CUDA-graph captured)").

    The step is captured as one graph per gradient bucket (`TrainEngine.buckets`): forward + loss +
    fusion-stage backward, then the trunk backward split after layer4, after layer3 and at the end,
    plus the Adam update. In a data-parallel job every bucket of the flat fp32 gradient is all-reduced
    over NCCL as soon as its graph has been queued, so the collectives run on NCCL's stream while the
    next graph computes; only the last bucket (layer2 + layer1 + stem, 5.7 MB) is exposed. Per step
    the host replays five graphs (and issues four collectives); learning rate and step count live in
    device memory (`TrainEngine.hyper`), so `set_lr` needs no re-capture. The per-step D2H of the
    reference (trainer.py:128) is gone: `loss` stays on the device until the caller reads it.
    """

    def __init__(self, engine: TrainEngine, batch: int, views: int, size: int = 224):
        self.engine = engine
        dev = engine.device
        self.images = torch.zeros((batch, views, 3, size, size), device=dev)
        self.rotations = torch.eye(3, device=dev).expand(batch, views, views, 3, 3).contiguous()
        self.gt = torch.zeros((batch, views, 2), device=dev)
        self.skip_allreduce = False   # bench.py: time the step without its collectives
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        # warm-up on a side stream allocates every cached buffer; everything it touched (parameters,
        # BatchNorm buffers, Adam state) is restored afterwards -- a checkpoint loaded before the
        # capture survives it
        saved = [t.clone() for t in (engine.flat_p, engine.flat_m, engine.flat_v, engine.hyper)]
        bufs = [b.clone() for b in engine.model.buffers()]
        with torch.cuda.stream(side):
            for _ in range(2):
                engine.forward_backward(self.images, self.rotations, self.gt)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        n0 = L.STATS["launches"]
        self.graphs = [torch.cuda.CUDAGraph()]
        self.bucket_of_graph = []

        def split(name, begin, end):
            self.bucket_of_graph.append((name, begin, end))
            self.graphs[-1].capture_end()
            if len(self.bucket_of_graph) < len(engine.buckets):
                g = torch.cuda.CUDAGraph()
                g.capture_begin(pool=self.graphs[0].pool())
                self.graphs.append(g)

        cap = torch.cuda.Stream(device=dev)
        cap.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cap):
            self.graphs[0].capture_begin()
            engine.forward_backward(self.images, self.rotations, self.gt, hook=split)
        torch.cuda.current_stream(dev).wait_stream(cap)
        assert len(self.graphs) == len(self.bucket_of_graph) == len(engine.buckets)
        self.adam = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.adam, pool=self.graphs[0].pool()):
            _ck("rmv_adam_step", engine.flat_p.data_ptr(), engine.flat_g.data_ptr(),
                engine.flat_m.data_ptr(), engine.flat_v.data_ptr(), engine.hyper.data_ptr(),
                engine.flat_p.numel(), int(engine.decoupled), 1.0 / engine.world)
        self.launches_per_step = L.STATS["launches"] - n0
        for t, c in zip((engine.flat_p, engine.flat_m, engine.flat_v, engine.hyper), saved):
            t.copy_(c)
        for b, c in zip(engine.model.buffers(), bufs):
            b.copy_(c)
        torch.cuda.synchronize(dev)
        self._pipe = None
        self._copy_stream = torch.cuda.Stream(device=dev)

    def collective_desc(self) -> str:
        eng = self.engine
        if not eng.dp_overlap:
            return "one all-reduce of the whole flat gradient after the backward pass"
        return "; ".join(f"{n}: {(e - b) * 4 / 1e6:.1f} MB" for n, b, e in self.bucket_of_graph) + \
            " -- each all-reduced (NCCL, async) right after its graph is queued"

    def step(self, images=None, rotations=None, gt=None) -> torch.Tensor:
        if images is not None:
            self.images.copy_(images, non_blocking=True)
        if rotations is not None:
            self.rotations.copy_(rotations, non_blocking=True)
        if gt is not None:
            self.gt.copy_(gt, non_blocking=True)
        eng = self.engine
        for p in (eng.model._feat_extractor[0].conv1.weight,):
            # the parameters must still be views of the flat buffer Adam updates (a later
            # model.to()/.float() would silently detach them)
            if not (eng.flat_p.data_ptr() <= p.data_ptr() < eng.flat_p.data_ptr() + eng.flat_p.numel() * 4):
                raise L.RotmvError("GraphedTrainStep: model parameters were moved after TrainEngine "
                                   "was built (model.to()/.float()?); rebuild the engine")
        ar = P.OverlappedAllReduce(eng.flat_g, eng.pg)
        if self.skip_allreduce:
            ar.active = False
        for g, (_, begin, end) in zip(self.graphs, self.bucket_of_graph):
            g.replay()
            if eng.dp_overlap:
                ar.start(begin, end)   # overlaps the next graph
        if not eng.dp_overlap:
            ar.start(0, eng.flat_g.numel())
        ar.finish()
        self.adam.replay()
        eng.model.invalidate()
        return eng.loss

    # ---- host-buffer entry, two steps in flight -------------------------------------------------
    def submit(self, images_host, head_pose_host, gt_host) -> int:
        """One optimisation step from pinned HOST buffers (images [B,V,3,H,W] fp32, head poses
        [B,V,2], labels [B,V,2] -- what trainer.py:99-123 hands to the model); returns a ticket for
        `result`. The host->HBM copies of step k+1 run on a copy stream into staging buffers while
        step k computes; the loss of each step is copied back to pinned host memory."""
        dev = self.engine.device
        if self._pipe is None:
            b, v = self.gt.shape[0], self.gt.shape[1]
            self._pipe = {
                "img": [torch.empty_like(self.images) for _ in range(2)],
                "pose": [torch.empty((b, v, 2), device=dev) for _ in range(2)],
                "gt": [torch.empty_like(self.gt) for _ in range(2)],
                "loss": [torch.empty((1,), dtype=torch.float32).pin_memory() for _ in range(2)],
                "copied": [torch.cuda.Event() for _ in range(2)],
                "free": [torch.cuda.Event() for _ in range(2)],
                "done": [torch.cuda.Event() for _ in range(2)],
                "count": 0, "taken": [True, True]}
            for ev in self._pipe["free"]:
                ev.record(torch.cuda.current_stream(dev))
        p = self._pipe
        k = p["count"]
        slot = k & 1
        if not p["taken"][slot]:
            raise RuntimeError("GraphedTrainStep.submit: two steps are already outstanding")
        main = torch.cuda.current_stream(dev)
        cs = self._copy_stream
        cs.wait_event(p["free"][slot])
        with torch.cuda.stream(cs):
            p["img"][slot].copy_(images_host, non_blocking=True)
            p["pose"][slot].copy_(head_pose_host, non_blocking=True)
            p["gt"][slot].copy_(gt_host, non_blocking=True)
            p["copied"][slot].record(cs)
        main.wait_event(p["copied"][slot])
        self.images.copy_(p["img"][slot], non_blocking=True)
        self.gt.copy_(p["gt"][slot], non_blocking=True)
        RF.pose_to_rotations(p["pose"][slot], out=self.rotations)
        p["free"][slot].record(main)
        loss = self.step()
        p["loss"][slot].copy_(loss, non_blocking=True)
        p["done"][slot].record(main)
        p["taken"][slot] = False
        p["count"] = k + 1
        return k

    def result(self, ticket: int) -> torch.Tensor:
        p = self._pipe
        if p is None or not (p["count"] - 2 <= ticket < p["count"]):
            raise RuntimeError(f"GraphedTrainStep.result: ticket {ticket} is not outstanding")
        slot = ticket & 1
        p["done"][slot].synchronize()
        p["taken"][slot] = True
        return p["loss"][slot]
