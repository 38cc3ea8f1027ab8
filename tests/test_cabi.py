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
