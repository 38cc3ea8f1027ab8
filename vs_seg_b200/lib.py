"""ctypes binding of libvsseg_b200.so (the C ABI declared in include/vsseg_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VSSEG_LIB_PATH") or os.path.join(_HERE, "libvsseg_b200.so")  # override: kernel-variant experiments
CSRC_DIR = os.path.join(_HERE, "csrc")


class NativeLibraryError(RuntimeError):
    pass


class Act8(C.Structure):
    _fields_ = [("hi", C.c_void_p), ("lo_offset", C.c_int64), ("batch_stride", C.c_int64),
                ("B", C.c_int32), ("C", C.c_int32), ("X", C.c_int32), ("Y", C.c_int32), ("Z", C.c_int32)]


class F32View(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("sb", C.c_int64), ("sc", C.c_int64), ("sx", C.c_int64),
                ("sy", C.c_int64), ("sz", C.c_int64),
                ("B", C.c_int32), ("C", C.c_int32), ("X", C.c_int32), ("Y", C.c_int32), ("Z", C.c_int32),
                ("n_windows", C.c_int32), ("indirect", C.c_void_p)]   # indirect: device cell holding the base address


MAX_WINDOWS = 16   # VSSEG_MAX_WINDOWS: records of a window set (F32View.n_windows, include/vsseg_b200.h)


class Epilogue(C.Structure):
    _fields_ = [("scale", C.c_void_p), ("shift", C.c_void_p), ("slope", C.c_float), ("act", C.c_int32)]


class ConvGeom(C.Structure):
    _fields_ = [("kx", C.c_int32), ("ky", C.c_int32), ("kz", C.c_int32),
                ("sx", C.c_int32), ("sy", C.c_int32), ("sz", C.c_int32), ("transposed", C.c_int32)]


_P = C.POINTER
ABI_VERSION = 2   # VSSEG_ABI_VERSION of include/vsseg_b200.h
_SIGNATURES = {
    "vsseg_abi_version": (C.c_int, []),
    "vsseg_last_error": (C.c_char_p, []),
    "vsseg_device_sm_count": (C.c_int, [C.c_int, _P(C.c_int)]),
    "vsseg_pack_act8": (C.c_int, [_P(F32View), _P(Act8), C.c_void_p]),
    "vsseg_unpack_act8": (C.c_int, [_P(Act8), _P(F32View), C.c_void_p]),
    "vsseg_conv3d_act8": (C.c_int, [_P(Act8), _P(Act8), _P(ConvGeom), C.c_void_p, C.c_int32, _P(Epilogue),
                                    _P(Act8), _P(F32View), C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsseg_conv3d_tc_supported": (C.c_int, [_P(Act8), _P(Act8), _P(ConvGeom), C.c_int32, _P(Act8)]),
    "vsseg_conv3d_tc_suggest_split": (C.c_int, [_P(Act8), _P(Act8), _P(ConvGeom), _P(Act8)]),
    "vsseg_conv3d_tc_f32out": (C.c_int, [_P(Act8), _P(F32View), _P(ConvGeom), C.c_void_p, _P(Epilogue), C.c_void_p,
                                         C.c_void_p]),
    "vsseg_conv3d_tc_f32out_2p": (C.c_int, [_P(Act8), _P(F32View), _P(ConvGeom), C.c_void_p, _P(Epilogue), C.c_void_p,
                                            C.c_void_p]),
    "vsseg_conv3d_tc_attgate": (C.c_int, [_P(Act8), _P(F32View), _P(ConvGeom), C.c_void_p, _P(Epilogue), _P(Act8), C.c_void_p]),
    "vsseg_conv3d_tc_describe": (C.c_int, [_P(Act8), _P(Act8), _P(ConvGeom), C.c_int32, _P(Act8), C.c_char_p, C.c_int32]),
    "vsseg_conv3d_tc": (C.c_int, [_P(Act8), _P(Act8), _P(ConvGeom), C.c_void_p, C.c_int32, _P(Epilogue),
                                  _P(Act8), _P(F32View), C.c_void_p, C.c_void_p,
                                  _P(Act8), C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsseg_conv3d_cin1": (C.c_int, [_P(F32View), _P(Act8), _P(ConvGeom), C.c_void_p, _P(Epilogue), C.c_void_p]),
    "vsseg_conv3d_smallcout": (C.c_int, [_P(Act8), _P(F32View), _P(ConvGeom), C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_float, C.c_void_p, C.c_void_p]),
    "vsseg_conv3d_gate_logits": (C.c_int, [_P(Act8), _P(F32View), C.c_void_p, C.c_void_p, C.c_int32, _P(F32View), C.c_int32,
                                           C.c_void_p, C.c_int32, C.c_void_p]),
    "vsseg_att_gate": (C.c_int, [_P(Act8), _P(F32View), _P(Act8), C.c_void_p]),
    "vsseg_peer_alloc": (C.c_int, [C.c_int64, _P(C.c_void_p), C.c_void_p]),
    "vsseg_peer_open": (C.c_int, [C.c_void_p, _P(C.c_void_p)]),
    "vsseg_peer_close": (C.c_int, [C.c_void_p]),
    "vsseg_peer_free": (C.c_int, [C.c_void_p]),
    "vsseg_flag_set": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "vsseg_flag_wait": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "vsseg_maxpool3d": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 7 + [C.c_void_p]),
    "vsseg_dice_sums": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_float, C.c_void_p,
                                  C.c_void_p]),
    "vsseg_dice_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsseg_dice_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_float, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsseg_dice_general_sums": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32,
                                          C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "vsseg_dice_general_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32,
                                              C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsseg_bn_stats": (C.c_int, [_P(Act8), C.c_void_p, C.c_void_p]),
    "vsseg_bn_finalize": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsseg_bn_act_fwd": (C.c_int, [_P(Act8), _P(Act8), C.c_void_p, C.c_void_p, C.c_float, C.c_uint64, C.c_void_p, _P(Act8),
                                   C.c_void_p]),
    "vsseg_bn_act_bwd_reduce": (C.c_int, [_P(Act8), _P(Act8), C.c_void_p, C.c_void_p, C.c_float, C.c_uint64, C.c_void_p,
                                          C.c_void_p, C.c_void_p]),
    "vsseg_bn_act_bwd_apply": (C.c_int, [_P(Act8), _P(Act8), C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_uint64,
                                         C.c_void_p, _P(Act8), C.c_void_p]),
    "vsseg_act_bwd": (C.c_int, [_P(Act8), _P(Act8), C.c_float, _P(Act8), C.c_void_p]),
    "vsseg_act8_add": (C.c_int, [_P(Act8), _P(Act8), _P(Act8), C.c_void_p]),
    "vsseg_conv3d_wgrad": (C.c_int, [_P(Act8), _P(Act8), _P(ConvGeom), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "vsseg_conv3d_wgrad_tc_supported": (C.c_int, [_P(Act8), _P(Act8), _P(ConvGeom)]),
    "vsseg_conv3d_wgrad_tc": (C.c_int, [_P(Act8), _P(Act8), _P(ConvGeom), C.c_void_p, C.c_int32, C.c_void_p]),
    "vsseg_conv3d_cin1_wgrad": (C.c_int, [_P(F32View), _P(Act8), _P(ConvGeom), C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsseg_conv3d_smallcout_bwd": (C.c_int, [_P(Act8), _P(F32View), _P(F32View), _P(ConvGeom), C.c_void_p, C.c_int32,
                                             _P(Act8), C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsseg_att_gate_bwd": (C.c_int, [_P(Act8), _P(F32View), _P(Act8), _P(Act8), _P(F32View), C.c_int32, C.c_void_p]),
    "vsseg_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_float,
                                  C.c_float, C.c_float, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vsseg_pack_conv_weight_tc": (C.c_int, [C.c_void_p] + [C.c_int32] * 11 + [C.c_void_p, C.c_void_p]),
    "vsseg_sw_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p,
                                    C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC_DIR, "-j4"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise NativeLibraryError("building libvsseg_b200.so failed")
    return LIB_PATH


def load():
    """Load the native library (once) and type its entry points."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the CUDA path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.vsseg_abi_version() != ABI_VERSION:
        raise NativeLibraryError("libvsseg_b200.so ABI version mismatch")
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(code: int, what: str = ""):
    if code != 0:
        msg = load().vsseg_last_error().decode(errors="replace")
        raise NativeLibraryError(f"{what or 'vsseg call'} failed (code {code}): {msg}")


_launch_count = 0


def count_launch(n: int = 1):
    global _launch_count
    _launch_count += n


def launches() -> int:
    return _launch_count
