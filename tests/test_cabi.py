"""The C-ABI library loads and exports every symbol include/vsseg_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

from vs_seg_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "vsseg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vsseg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(lib.LIB_PATH):
        lib.build()
    handle = ctypes.CDLL(lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 9
    for n in names:
        assert hasattr(handle, n), f"{n} declared in vsseg_b200.h but not exported"
    assert sorted(lib.exported_symbols()) == names, "ctypes binding and header disagree"
    assert lib.load().vsseg_abi_version() == 2


def test_bad_arguments_return_einval_without_gpu():
    h = lib.load()
    code = h.vsseg_sw_finalize(None, None, None, 2, 10, None, None, 0, None, None)
    assert code == 100001
    assert b"sw_finalize" in h.vsseg_last_error()


def test_inconsistent_window_set_is_rejected_without_gpu():
    """vsseg_f32view.n_windows: a window set must hold out->B single-window records of one volume; everything that
    does not take window sets rejects them (validated before any launch, so no GPU is needed)."""
    import ctypes as C
    h = lib.load()
    n = 8 * 8 * 8
    views = (lib.F32View * 2)()
    for i in range(2):
        views[i] = lib.F32View(4096 + 64 * i, n, n, 64, 8, 1, 1, 1, 8, 8, 8, 0, None)
    out = lib.Act8(4096, 2 * 16 * n, 16 * n, 2, 16, 8, 8, 8)
    g = lib.ConvGeom(3, 3, 1, 1, 1, 1, 0)
    ep = lib.Epilogue(4096, 4096, 0.25, 0)
    views[0].n_windows = 3                      # three records announced for a batch of two
    assert h.vsseg_conv3d_cin1(views, C.byref(out), C.byref(g), 4096, C.byref(ep), None) == 100001
    assert b"shape mismatch" in h.vsseg_last_error()
    views[0].n_windows = 2
    views[1].sy = 16                            # the records do not describe windows of one volume
    assert h.vsseg_conv3d_cin1(views, C.byref(out), C.byref(g), 4096, C.byref(ep), None) == 100001
    assert b"window set" in h.vsseg_last_error()
    views[1].sy = 8
    att = lib.Act8(4096, 2 * 8 * n, 8 * n, 2, 8, 8, 8, 8)
    assert h.vsseg_att_gate(C.byref(att), views, C.byref(att), None) == 100001   # plain views only


def test_product_path_fails_loudly_without_the_library_or_a_gpu(tmp_path):
    """No CPU or eager fallback behind the native path: a missing library and a CPU device both raise."""
    import subprocess
    import sys

    import pytest
    env = dict(os.environ, VSSEG_LIB_PATH=str(tmp_path / "missing.so"))
    code = ("from vs_seg_b200 import lib\n"
            "try:\n    lib.load()\nexcept lib.NativeLibraryError as e:\n    print('RAISED', e)\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, env=env, timeout=300)
    assert "RAISED" in out.stdout and "no CPU fallback" in out.stdout, out.stdout + out.stderr
    from oracle import unet_oracle
    from vs_seg_b200.engine import UNetEvalPlan, conv_block_ncdhw
    import torch
    with pytest.raises(lib.NativeLibraryError):
        UNetEvalPlan(unet_oracle.seeded_state_dict(0), (64, 64, 16), device="cpu")
    with pytest.raises(lib.NativeLibraryError):
        conv_block_ncdhw(torch.zeros(1, 16, 4, 4, 8), {"conv.weight": torch.zeros(16, 16, 3, 3, 1)}, (3, 3, 1), (1, 1, 1),
                         False, False, "none")


def test_header_is_plain_c():
    """include/vsseg_b200.h is the C-ABI contract: it must compile as C99 (plain pointers and sizes, no C++/torch types)."""
    import shutil
    import subprocess

    import pytest
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc on this box")
    r = subprocess.run([gcc, "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Werror", os.path.join(ROOT, "include", "vsseg_b200.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
