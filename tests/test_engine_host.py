"""Host logic of the INFERENCE path (rotmv_b200/module.py + rotmv_b200/engine.py) checked on the CPU:
the tensor-level kernel wrappers of rotmv_b200.functional (conv + folded BatchNorm + residual +
ReLU, pools, Linear, rotation gather, head, strided copy, zero fill, relative rotations) are
replaced by their torch formulas and the engine's device check is lifted, so what runs is the
orchestration itself -- BatchNorm folding, the block wiring of both ResNet kinds, the concat-free
X/Y buffers, the zero-padded 3593-wide fuser of `encode_rotmat`, the [3][2][512] interleave and the
IntensityBatchNorm factor of `share_feature`, the V > 2 gather, per-view output assembly, the
reference dict adapter -- compared with the CPU oracle's eval forward on the same weights and
inputs (models/rot_mv.py:187-269). The kernels themselves are tested against torch on the GPU
(tests/test_ops_gpu.py, tests/test_variant_glue_gpu.py), the same configurations end to end in
tests/test_variants.py / tests/test_module_gpu.py."""
import pytest
import torch
import torch.nn.functional as F

from oracle import rotmv_oracle as O
from rotmv_b200 import engine as E
from rotmv_b200 import functional as RF
from rotmv_b200.module import FeatRotationSymm

from test_train_fusion_host import torch_kernels  # noqa: F401  (fixture: linear, avgpool, rotate_gather, head_loss)


@pytest.fixture()
def host_engine(torch_kernels, monkeypatch):  # noqa: F811
    def conv2d(x, w, *, stride=1, pad=0, scale=None, shift=None, residual=None, relu=False, out=None, **_):
        y = F.conv2d(x.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), stride=stride, padding=pad).permute(0, 2, 3, 1)
        if scale is not None:
            y = y * scale
        if shift is not None:
            y = y + shift
        if residual is not None:
            y = y + residual
        if relu:
            y = torch.relu(y)
        if out is None:
            return y.contiguous()
        out.copy_(y)
        return out

    def conv2d_nchw_input(x_nchw, w, *, stride, pad, scale=None, shift=None, relu=False, out_dtype=None):
        return conv2d(x_nchw.permute(0, 2, 3, 1), w, stride=stride, pad=pad, scale=scale, shift=shift, relu=relu)

    def maxpool3x3s2(x, out=None):
        y = F.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
        if out is None:
            return y.contiguous()
        out.copy_(y)
        return out

    def strided_copy(src, dst, scale=None, accumulate=False):
        assert tuple(src.shape) == tuple(dst.shape) and src.dim() <= 3
        val = src.float() if scale is None else src.float() * scale
        dst.copy_((val + dst.float() if accumulate else val).to(dst.dtype))
        return dst

    def fill_zero(t):
        return t.zero_()

    def relative_rotations(rot):
        return rot[:, :, None] @ rot[:, None].transpose(-1, -2)

    for name, fn in dict(conv2d=conv2d, conv2d_nchw_input=conv2d_nchw_input, maxpool3x3s2=maxpool3x3s2,
                         strided_copy=strided_copy, fill_zero=fill_zero,
                         relative_rotations=relative_rotations).items():
        monkeypatch.setattr(RF, name, fn)
    monkeypatch.setattr(E.InferenceEngine, "_require_device", lambda self, dev, what: None)


def _pair(depth, n_it, flags, b, v, size=64, seed=3):
    ora = O.build_model(num_iter=n_it, depth=depth, seed=0, **flags)
    images, pose, gt = O.synthetic_batch(b, v, seed=seed, size=size)
    rot = O.pairwise_rotations(pose)
    O.calibrate_bn(ora, images, passes=2)
    if flags.get("share_feature"):      # a running std that is not the initial 1
        for i, fuser in enumerate(ora._img_fusers):
            fuser._batchnorm.running_mean.copy_(0.5 + torch.rand(1, 1, 512, generator=torch.Generator().manual_seed(i)))
    ora.eval()
    model = FeatRotationSymm(depth, n_it, **flags)
    model.load_state_dict(ora.state_dict(), strict=True)
    model.eval()
    model.auto_graph = False
    return ora, model, images, rot, pose, gt


def _close(got, want, what, rtol=1e-4):
    tol = rtol * want.abs().max().item() + 1e-6
    err = (got - want).abs().max().item()
    assert tuple(got.shape) == tuple(want.shape) and got.dtype == torch.float32, (what, got.shape, want.shape)
    assert err <= tol, (what, err, tol)


FLAGS = [dict(), dict(share_weights=True), dict(ignore_rotmat=True), dict(encode_rotmat=True),
         dict(share_feature=True), dict(encode_rotmat=True, share_weights=True)]
IDS = ["default", "share_weights", "ignore_rotmat", "encode_rotmat", "share_feature", "encode_rotmat+share_weights"]


@pytest.mark.parametrize("flags", FLAGS, ids=IDS)
def test_inference_orchestration_matches_oracle_depth18(host_engine, flags):
    ora, model, images, rot, _, gt = _pair(18, 2, flags, b=3, v=2)
    with torch.no_grad():
        want = ora.forward_views(images, rot)
        got = model.forward_views(images, rot, precision="fp32", want_all=True, gt=gt)
    assert got["num_iter"] == want["num_iter"] == 2
    for k in range(2):
        _close(got[f"img_feat_{k}"], want[f"img_feat_{k}"], f"img_feat_{k}")
        _close(got[f"initial_rot_feat_{k}"], want[f"initial_rot_feat_{k}"], f"initial_rot_feat_{k}")
        for i in range(2):
            _close(got[f"iter_{i}"][f"feat_{k}"], want[f"iter_{i}"][f"feat_{k}"], f"iter_{i}/feat_{k}")
            _close(got[f"iter_{i}"][f"pred_gaze_{k}"], want[f"iter_{i}"][f"pred_gaze_{k}"], f"iter_{i}/pred_gaze_{k}")
    _close(got["pred_gaze"], want["pred_gaze"], "pred_gaze")
    loss_ref = O.iteration_loss(want, [gt[:, 0], gt[:, 1]]).item()
    assert abs(got["loss"].item() - loss_ref) <= 1e-4 * abs(loss_ref), (got["loss"].item(), loss_ref)
    # the tensor API returns the last iteration's view-0 prediction (north_star signature)
    with torch.no_grad():
        _close(model(images, rot, precision="fp32"), want["pred_gaze"], "forward(images, rotations)")


def test_bottleneck_trunk_three_views_and_dict_adapter(host_engine):
    """ResNet-50 wiring (conv3 + residual, downsample branches), the V = 3 gather, an
    empty batch, and the reference's dict contract (models/rot_mv.py:187-269: the input dict is
    mutated and returned)."""
    ora, model, images, rot, pose, _ = _pair(50, 1, {}, b=2, v=3, size=64)
    rt = 1e-3   # fp32 with folded vs unfolded BatchNorm through 53 layers calibrated on six images
    with torch.no_grad():
        want = ora.forward_views(images, rot)
        got = model.forward_views(images, rot, precision="fp32")
    for k in range(3):
        _close(got[f"img_feat_{k}"], want[f"img_feat_{k}"], f"img_feat_{k}", rt)
        _close(got["iter_0"][f"feat_{k}"], want["iter_0"][f"feat_{k}"], f"feat_{k}", rt)
        _close(got["iter_0"][f"pred_gaze_{k}"], want["iter_0"][f"pred_gaze_{k}"], f"pred_gaze_{k}", rt)
    empty = model.forward_views(images[:0], rot[:0], precision="fp32")
    assert tuple(empty["pred_gaze"].shape) == (0, 2) and tuple(empty["img_feat_2"].shape) == (0, 2048)
    # dict adapter, two views
    r = O.rotation_matrix_2d(pose.reshape(-1, 2)).reshape(2, 3, 3, 3)
    data = {"img_0": images[:, 0], "img_1": images[:, 1], "rot_0": r[:, 0], "rot_1": r[:, 1]}
    ref = ora.forward({k: v.clone() for k, v in data.items()})
    with torch.no_grad():
        out = model(data, precision="fp32")
    assert out is data
    for key in ("img_feat_0", "img_feat_1", "initial_rot_feat_0", "initial_rot_feat_1", "pred_gaze"):
        _close(out[key], ref[key].detach(), key, rt)
    _close(out["iter_0"]["pred_gaze_1"], ref["iter_0"]["pred_gaze_1"].detach(), "iter_0/pred_gaze_1", rt)


def test_device_check_is_what_keeps_the_cpu_out():
    """Without the test's override the same call raises: there is no CPU path in the product."""
    from rotmv_b200 import _lib as L

    model = FeatRotationSymm(18, 1).eval()
    with pytest.raises(L.RotmvError):
        model.forward_views(torch.zeros((1, 2, 3, 32, 32)), torch.eye(3).expand(1, 2, 2, 3, 3).contiguous())
