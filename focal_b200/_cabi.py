"""ctypes binding of libfocal_b200.so (the C ABI declared in include/focal_b200.h).

There is no fallback: if the shared library is missing or does not load, importing the loss fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# FOCAL_B200_LIB: an experiment build of the same library (tools/variant_bench.py); never a different implementation
LIB_PATH = os.environ.get("FOCAL_B200_LIB") or os.path.join(_PKG, "libfocal_b200.so")

FOCAL_TERM_NCE, FOCAL_TERM_ORTH, FOCAL_TERM_TEMPORAL, FOCAL_TERM_ALL = 1, 2, 4, 7
FOCAL_PREC_BF16, FOCAL_PREC_FP32 = 0, 1
FOCAL_MAX_MODALITIES = 8
ABI_VERSION = 5
FOCAL_OK, FOCAL_EINVAL, FOCAL_ESHAPE, FOCAL_ECUDA, FOCAL_EWORKSPACE = 0, -1, -2, -3, -4

EXPORTS = (
    "focal_b200_abi_version", "focal_b200_strerror", "focal_b200_workspace_info", "focal_b200_prologue",
    "focal_b200_nce_rowsum", "focal_b200_nce_lse", "focal_b200_nce_grad", "focal_b200_temporal",
    "focal_b200_finalize", "focal_b200_loss", "focal_b200_set_ptrs",
    "focal_b200_peer_alloc", "focal_b200_peer_open", "focal_b200_peer_close", "focal_b200_peer_free",
    "focal_b200_loss_sharded", "focal_b200_spectrum_rotate", "focal_b200_knn_predict", "focal_b200_debug_stage_times",
)
FOCAL_MAX_PEERS = 8


class FocalCfg(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("S", C.c_int32), ("M", C.c_int32), ("D", C.c_int32),
        ("temperature", C.c_float), ("margin", C.c_float),
        ("w_shared", C.c_float), ("w_private", C.c_float), ("w_orth", C.c_float), ("w_rank", C.c_float),
        ("no_private", C.c_int32), ("need_grad", C.c_int32), ("terms", C.c_int32), ("precision", C.c_int32),
        ("seq_begin", C.c_int32), ("seq_end", C.c_int32), ("num_sms", C.c_int32),
        ("in_block_rows", C.c_int32), ("in_block_stride", C.c_int32),
        ("local_rows", C.c_int32), ("indirect_ptrs", C.c_int32),
    ]


class FocalPeers(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("ws", C.c_void_p * 8), ("mc", C.c_void_p)]


class FocalWsInfo(C.Structure):
    _fields_ = [
        ("total_bytes", C.c_size_t),
        ("b", C.c_int32), ("bpad", C.c_int32), ("Bpad", C.c_int32), ("n_problems", C.c_int32),
        ("n_ops", C.c_int32), ("kb_full", C.c_int32),
        ("rowsum_off", C.c_size_t), ("rowsum_bytes", C.c_size_t),
        ("cnt_off", C.c_size_t), ("cnt_bytes", C.c_size_t),
        ("mintra_off", C.c_size_t), ("lossparts_off", C.c_size_t),
        ("dz_off", C.c_size_t), ("dz_bytes", C.c_size_t),
        ("dx_off", C.c_size_t), ("dx_bytes", C.c_size_t),
        ("cnt_piece_stride", C.c_size_t), ("n_pieces_nce", C.c_int32), ("n_pieces_tmp", C.c_int32),
    ]


class FocalError(RuntimeError):
    def __init__(self, code: int, what: str):
        self.code = code
        super().__init__(f"{what}: {strerror(code)} (code {code})")


_lib = None


def load(path: str | None = None) -> C.CDLL:
    """Load the shared library once.  Raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -m focal_b200.build` (nvcc, sm_100a). "
            "focal_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(path)
    vp, cfgp = C.c_void_p, C.POINTER(FocalCfg)
    lib.focal_b200_abi_version.restype = C.c_int
    lib.focal_b200_strerror.restype = C.c_char_p
    lib.focal_b200_strerror.argtypes = [C.c_int]
    lib.focal_b200_workspace_info.argtypes = [cfgp, C.POINTER(FocalWsInfo)]
    lib.focal_b200_prologue.argtypes = [cfgp, C.POINTER(vp), vp, C.c_size_t, vp]
    lib.focal_b200_nce_rowsum.argtypes = [cfgp, vp, C.c_size_t, vp]
    lib.focal_b200_nce_lse.argtypes = [cfgp, vp, C.c_size_t, C.c_int, vp]
    lib.focal_b200_nce_grad.argtypes = [cfgp, vp, C.c_size_t, vp]
    lib.focal_b200_temporal.argtypes = [cfgp, vp, C.c_size_t, vp]
    lib.focal_b200_finalize.argtypes = [cfgp, C.POINTER(vp), vp, C.c_size_t, vp, C.POINTER(vp), vp]
    lib.focal_b200_loss.argtypes = [cfgp, C.POINTER(vp), vp, C.c_size_t, vp, C.POINTER(vp), vp]
    old = bool(os.environ.get("FOCAL_B200_LIB_BISECT"))    # bisecting with a build of an older commit (tools only)
    if not old:
        lib.focal_b200_spectrum_rotate.argtypes = [vp, vp, C.c_longlong, C.c_int, C.c_int, C.c_float, C.c_float, vp]
        lib.focal_b200_knn_predict.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
    lib.focal_b200_set_ptrs.argtypes = [cfgp, vp, C.c_size_t, C.POINTER(vp), vp, C.POINTER(vp), vp]
    lib.focal_b200_peer_alloc.argtypes = [C.c_size_t, C.POINTER(vp), C.c_char_p]
    lib.focal_b200_peer_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    lib.focal_b200_peer_close.argtypes = [vp]
    lib.focal_b200_peer_free.argtypes = [vp]
    lib.focal_b200_loss_sharded.argtypes = [cfgp, C.POINTER(vp), C.POINTER(FocalPeers), C.c_size_t, vp, C.POINTER(vp), vp]
    for name in EXPORTS:
        if name != "focal_b200_strerror" and not (old and not hasattr(lib, name)):
            getattr(lib, name).restype = C.c_int
    if lib.focal_b200_abi_version() != ABI_VERSION and not old:
        raise ImportError(f"{path}: ABI version {lib.focal_b200_abi_version()} != expected {ABI_VERSION}")
    if path == LIB_PATH:
        _lib = lib
    return lib


BRINGUP_PATH = os.path.join(_PKG, "libfocal_bringup.so")
_bringup = None


def load_bringup() -> C.CDLL:
    """Hardware probes / micro-benchmarks (focal_b200/csrc/bringup.h): tests and tools only, never the product path."""
    global _bringup
    if _bringup is None:
        if not os.path.exists(BRINGUP_PATH):
            raise ImportError(f"{BRINGUP_PATH} not found: build it with `python -m focal_b200.build`")
        lib = C.CDLL(BRINGUP_PATH)
        vp = C.c_void_p
        lib.focal_b200_debug_umma.argtypes = [vp, C.c_uint32, vp, C.c_uint32] + [C.c_uint32] * 10 + [vp, vp]
        lib.focal_b200_debug_umma_rate.argtypes = [C.c_uint32] * 5 + [vp, vp]
        lib.focal_b200_debug_tma_rate.argtypes = [vp] + [C.c_uint32] * 6 + [vp, vp]
        for name in ("focal_b200_debug_umma", "focal_b200_debug_umma_rate", "focal_b200_debug_tma_rate"):
            getattr(lib, name).restype = C.c_int
        _bringup = lib
    return _bringup


def strerror(code: int) -> str:
    return load().focal_b200_strerror(code).decode()


def check(code: int, what: str) -> None:
    if code != 0:
        raise FocalError(code, what)


def ptr_array(ptrs):
    arr = (C.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr
