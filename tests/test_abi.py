"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/rotmv_sm100.h declares, the ctypes table covers the header exactly, argument validation
returns error codes (no compute is launched without a GPU), and the product package refuses to run
without CUDA instead of falling back."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rotmv_sm100.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rmv_[a-z0-9_]+)\s*\(", src)))


def test_header_table_and_library_agree():
    from rotmv_b200 import _lib as L

    syms = header_symbols()
    assert len(syms) >= 10
    assert sorted(L.SIGNATURES.keys()) == syms
    lib = L.load()
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert lib.rmv_version() == 100


def test_header_arg_counts_match_ctypes():
    from rotmv_b200 import _lib as L

    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, argtypes) in L.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, src, flags=re.S)
        assert m, name
        args = m.group(1).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        assert n == len(argtypes), (name, n, len(argtypes))


def test_conv_args_struct_layout():
    from rotmv_b200 import _lib as L

    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    body = re.search(r"typedef struct rmv_conv_args \{(.*?)\} rmv_conv_args;", src, flags=re.S).group(1)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(const\s+)?(void|float|double|int|long long|size_t)\s*\*?\s*", "", decl)
        names += [n.strip().lstrip("*") for n in decl.split(",")]
    assert names == [f[0] for f in L.ConvArgs._fields_]


def test_invalid_arguments_return_codes_not_crashes():
    from rotmv_b200 import _lib as L

    lib = L.load()
    assert lib.rmv_conv2d_fwd(None, None) < 0
    assert b"null" in lib.rmv_last_error()
    a = L.ConvArgs()
    assert lib.rmv_conv2d_fwd(C.byref(a), None) < 0
    assert lib.rmv_stem_im2col(None, None, 1, 3, 224, 224, 7, 7, 2, 3, 112, 112, 100, 1, None) < 0
    assert lib.rmv_rotate_gather_fwd(None, 8, None, None, 8, 1, 1, 512, 1, 1, None) < 0
    # empty inputs are a no-op, not an error
    assert lib.rmv_maxpool3x3s2_fwd(None, None, 0, 112, 112, 64, 1, None) == 0
    assert lib.rmv_pose_to_rotations(None, None, 0, 2, None) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_path_fails_loudly_without_cuda():
    from rotmv_b200 import _lib as L
    from rotmv_b200.module import FeatRotationSymm

    m = FeatRotationSymm(18, 1).eval()
    x = torch.zeros((1, 2, 3, 224, 224))
    r = torch.eye(3).expand(1, 2, 2, 3, 3).contiguous()
    with pytest.raises(L.RotmvError):
        m(x, r)


def test_constructor_surface_matches_reference():
    """Same ctor kwargs and error behaviour as models/rot_mv.py:103-184."""
    from rotmv_b200.module import FeatRotationSymm

    with pytest.raises(TypeError):
        FeatRotationSymm(backbone_depth=18)  # num_iter=None -> TypeError, like the reference
    with pytest.raises(AssertionError):
        FeatRotationSymm(18, 1, encode_rotmat=True, ignore_rotmat=True)
    m = FeatRotationSymm(backbone_depth=18, num_iter=2, share_weights=True)
    sd = m.state_dict()
    assert "_img_fusers.1._fuser.blocks.0.0.weight" in sd  # aliased copies appear per iteration
    assert sd["_img_fusers.0._fuser.blocks.0.0.weight"].data_ptr() == \
        sd["_img_fusers.1._fuser.blocks.0.0.weight"].data_ptr()


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints ONE JSON line
    with the contract's keys; it runs the oracle port on the host cores and needs no GPU."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["config"]["workload"].startswith("configs[1]")
    # the default run carries the training half of the metric (configs[3]) in the `train` block
    t = d["train"]
    assert t["impl"] == "reference" and t["value"] > 0 and t["config"]["workload"].startswith("configs[3]")
    assert "fwd+bwd" in t["metric"] and t["cpu_baseline"]["kind"] == "port"
