"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/gm3d.h declares,
validates arguments before touching the device, and the Python package refuses to run without it."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gm3d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gm3d_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_all_exported_and_bound():
    from gm3d_b200 import _lib
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in gm3d.h but not exported by libgm3d_sm100.so"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in gm3d_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(syms)
    # nothing but the ABI is exported (visibility=hidden elsewhere)
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.SO_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert {e for e in exported if e.startswith("gm3d_")} == set(syms)


def test_version_errors_and_workspace_are_host_only():
    from gm3d_b200 import _lib
    lib = _lib.load()
    assert lib.gm3d_abi_version() == _lib.GM3D_ABI_VERSION == 5
    assert b"invalid" in lib.gm3d_strerror(_lib.GM3D_EINVAL)
    assert b"not supported" in lib.gm3d_strerror(_lib.GM3D_ENOSUP)
    assert lib.gm3d_strerror(0) == b"success"
    assert lib.gm3d_workspace_bytes(_lib.OP_FPS, 4, 1024, 64, 0) == 0          # register-resident kernel
    assert lib.gm3d_workspace_bytes(_lib.OP_FPS, 4, 10000, 64, 0) == 4 * 10000 * 4  # running-min array
    assert lib.gm3d_workspace_bytes(_lib.OP_CHAMFER_FWD, 4992, 32, 32, 0) == 16 + 4096 * 32 + 4992 * 4  # ticket, CTA partials, per-patch scratch
    assert lib.gm3d_workspace_bytes(_lib.OP_KNN, 4, 1024, 64, 32) == 0


def test_argument_validation_returns_einval_without_a_device():
    """Bad shapes / NULL pointers are rejected before any CUDA call (so this runs without a GPU)."""
    from gm3d_b200 import _lib
    lib = _lib.load()
    E, U = _lib.GM3D_EINVAL, _lib.GM3D_ENOSUP
    p = ctypes.c_void_p(16)  # never dereferenced: validation fails first
    assert lib.gm3d_fps_f32(None, 1, 8, 2, p, None, None, None) == E
    assert lib.gm3d_fps_f32(p, 0, 8, 2, p, None, None, None) == E
    assert lib.gm3d_knn_f32(p, p, 1, 8, 2, 9, None, p, None, None) == E      # k > N
    assert lib.gm3d_knn_f32(p, p, 1, 80, 2, 33, None, p, None, None) == U    # k > GM3D_KNN_MAX_K
    assert lib.gm3d_group_f32(p, 1, 8, 9, 2, p, p, None, p, None, None, None) == E  # G > N
    assert lib.gm3d_knn_group_f32(p, p, 1, 8, 2, 4, None, None, None, None) == E    # nbhd required
    assert lib.gm3d_chamfer_fwd_f32(p, p, None, 4, 8, 8, p, p, p, p, None, None, None, 3, None, None) == E  # norm
    assert lib.gm3d_chamfer_fwd_f32(p, p, None, 4, 8, 8, p, p, p, p, None, p, None, 2, None, None) == E     # total needs ws
    assert lib.gm3d_chamfer_fused_f32(p, p, None, 4, 40, 8, 1.0, 1.0, None, None, None, None, None, None, None, 2, p, None, None, 0, None, None) == U
    assert lib.gm3d_chamfer_bwd_f32(p, p, None, None, p, None, None, 1.0, 1.0, 4, 8, 8, p, None, None) == E
    assert lib.gm3d_hard_mask_f32(None, 2, 64, 25, 15, None, 0, 0, p, None, 0, None) == E  # len_loss > 0 needs loss_pred
    assert lib.gm3d_hard_mask_f32(p, 2, 64, 25, 40, None, 0, 0, p, None, 0, None) == E     # len_loss > L - len_keep
    assert lib.gm3d_hard_mask_f32(p, 2, 5000, 25, 4, None, 0, 0, p, None, 0, None) == U
    assert lib.gm3d_select_patches_f32(None, p, 2, 64, 96, 65, 0, None, p, None, None) == E
    assert lib.gm3d_gather_f32(p, p, 1, 0, 8, 2, p, None) == E
    assert lib.gm3d_loss_stats_f32(None, 4, p, None) == E
    # the per-step peer all-reduce descriptor: more ranks than inboxes, or a rank outside the world
    red = _lib.StepReduce()
    red.world, red.rank, red.epoch = _lib.MAX_PEERS + 1, 0, 16
    assert lib.gm3d_chamfer_fused_f32(p, p, None, 4, 8, 8, 1.0, 1.0, None, None, None, None, None, None, None, 2, p, None,
                                      ctypes.byref(red), 0, p, None) == E
    red.world, red.rank = 2, 2
    assert lib.gm3d_chamfer_fused_f32(p, p, None, 4, 8, 8, 1.0, 1.0, None, None, None, None, None, None, None, 2, p, None,
                                      ctypes.byref(red), 0, p, None) == E
    assert ctypes.sizeof(_lib.StepReduce) == 112 and _lib.INBOX_BYTES == 1024
    assert lib.gm3d_peer_alloc(0, None, None) == E and lib.gm3d_peer_open(None, None) == E
    with pytest.raises(ValueError):
        _lib.check("x", E)
    with pytest.raises(NotImplementedError):
        _lib.check("x", U)
    with pytest.raises(_lib.Gm3dError):
        _lib.check("x", 700)  # a cudaError_t


def test_step_reduce_struct_matches_the_header(tmp_path):
    """gm3d_step_reduce_t is passed by HOST pointer from ctypes: every field offset and the size of _lib.StepReduce must
    equal what a C compiler makes of include/gm3d.h (compiled here with gcc), and the header must be plain C."""
    from gm3d_b200 import _lib
    src = tmp_path / "off.c"
    fields = ["head", "world", "rank", "inbox", "epoch", "timeout_us", "defer", "status", "collected"]
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "gm3d.h"\nint main(void) {\n'
                   + "".join(f'  printf("{f} %zu\\n", offsetof(gm3d_step_reduce_t, {f}));\n' for f in fields)
                   + '  printf("sizeof %zu\\n", sizeof(gm3d_step_reduce_t));\n'
                   + '  printf("inbox_bytes %d\\n", GM3D_INBOX_BYTES);\n  printf("abi %d\\n", GM3D_ABI_VERSION);\n  return 0;\n}\n')
    exe = tmp_path / "off"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    for f in fields:
        assert int(got[f]) == getattr(_lib.StepReduce, f).offset, f
    assert int(got["sizeof"]) == ctypes.sizeof(_lib.StepReduce)
    assert int(got["inbox_bytes"]) == _lib.INBOX_BYTES and int(got["abi"]) == _lib.GM3D_ABI_VERSION


def test_missing_library_fails_loudly(tmp_path):
    """No silent fallback: with the .so absent the operators cannot even be loaded."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from gm3d_b200 import _lib\n"
        "_lib.SO_PATH = %r\n"
        "try:\n"
        "    _lib.load()\n"
        "except ImportError as e:\n"
        "    assert 'no CPU or PyTorch fallback' in str(e); print('raised')\n"
    ) % (ROOT, str(tmp_path / "nope.so"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.stdout.strip() == "raised", r.stderr


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gm3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
                assert "libgm3d_oracle" not in src
