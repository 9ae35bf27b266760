"""ctypes binding of libpbllm.so (include/pbllm.h). Fails loudly if the library is missing or
cannot run: there is no Python/CPU fallback for any compute entry point."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

PBL_F16, PBL_BF16, PBL_F32 = 0, 1, 2
TILE_ROWS, TILE_COLS, RG_ROWS = 128, 64, 32


class PblSizes(C.Structure):
    _fields_ = [("n_pad", C.c_int64), ("k_pad", C.c_int64), ("tiles_r", C.c_int64), ("tiles_c", C.c_int64),
                ("groups", C.c_int64), ("planes_bytes", C.c_size_t), ("vptr_bytes", C.c_size_t),
                ("affine_bytes", C.c_size_t), ("vals_elem_bytes", C.c_size_t)]


class PblLayerDesc(C.Structure):
    _fields_ = [("N", C.c_int64), ("K", C.c_int64), ("groupsize", C.c_int64), ("dtype", C.c_int32),
                ("flags", C.c_uint32), ("affine", C.c_void_p), ("bias", C.c_void_p),
                ("planes", C.c_void_p), ("vptr", C.c_void_p), ("vals", C.c_void_p), ("sign_planes", C.c_void_p),
                ("fsign", C.c_void_p), ("eptr", C.c_void_p), ("ent", C.c_void_p), ("exc", C.c_void_p), ("n_exc", C.c_int64)]


MAX_PEERS = 8


class PblPeerPush(C.Structure):
    _fields_ = [("y", C.c_void_p * MAX_PEERS), ("flags", C.c_void_p * MAX_PEERS), ("sync_ctr", C.c_void_p),
                ("n_ranks", C.c_int32), ("rank", C.c_int32), ("wait_prev", C.c_int32), ("reserved", C.c_int32)]


class PblStreamSizes(C.Structure):
    _fields_ = [("blocks", C.c_int64), ("fsign_bytes", C.c_size_t), ("eptr_bytes", C.c_size_t)]


# name -> (restype, argtypes): every symbol include/pbllm.h declares
SYMBOLS = {
    "pbl_pack_sizes": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int, C.POINTER(PblSizes)]),
    "pbl_pack_affine": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                  C.c_void_p, C.c_void_p]),
    "pbl_pack_planes": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                  C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pbl_pack_vals": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int,
                                C.c_void_p, C.c_void_p]),
    "pbl_layer_create": (C.c_int, [C.POINTER(PblLayerDesc), C.POINTER(C.c_void_p)]),
    "pbl_layer_destroy": (None, [C.c_void_p]),
    "pbl_unpack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pbl_linear_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                     C.c_void_p]),
    "pbl_linear_forward_ws": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                        C.c_void_p, C.c_size_t, C.c_void_p]),
    "pbl_stream_layout": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int, C.POINTER(PblStreamSizes)]),
    "pbl_stream_count": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "pbl_stream_fill": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pbl_stream_position": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_uint32)]),
    "pbl_linear_forward_push": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(PblPeerPush), C.c_int64, C.c_int64,
                                          C.c_void_p, C.c_size_t, C.c_void_p]),
    "pbl_peer_wait": (C.c_int, [C.POINTER(PblPeerPush), C.c_void_p]),
    "pbl_kth_workspace": (C.c_size_t, []),
    "pbl_kth_value": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pbl_gptq_block": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "pbl_decode_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64]),
    "pbl_decode_set_trace": (None, [C.c_void_p, C.c_size_t]),
    "pbl_decode_plan": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_uint32)]),
    "pbl_forward_host_workspace": (C.c_size_t, [C.c_void_p, C.c_int64]),
    "pbl_linear_forward_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pbl_bireal_workspace": (C.c_size_t, [C.c_void_p, C.c_int64]),
    "pbl_bireal_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                     C.c_void_p]),
    "pbl_bireal_fixup_workspace": (C.c_size_t, [C.c_void_p, C.c_int64]),
    "pbl_bireal_forward_ws": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                        C.c_void_p, C.c_size_t, C.c_void_p]),
    "pbl_select_kernel": (C.c_int, [C.c_void_p, C.c_int64]),
    "pbl_decode_variant": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]),
    "pbl_launch_count": (C.c_int64, []),
    "pbl_last_error": (C.c_char_p, []),
    "pbl_abi_version": (C.c_int, []),
    "pbl_device_check": (C.c_int, []),
}

_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Load libpbllm.so (building it with nvcc if the sources are newer). Raises RuntimeError
    if it cannot be had -- the package is unusable without its CUDA library."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if _build.needs_build():
        try:
            path = _build.build()
        except Exception as e:  # noqa: BLE001
            # never load a binary that was built from other sources than the ones in the tree
            raise RuntimeError(f"libpbllm.so is missing or stale (source hash mismatch) and could not be rebuilt: {e}") from e
    try:
        lib = C.CDLL(path)
    except OSError as e:
        raise RuntimeError(f"cannot load {path}: {e}; pb-llm_b200 has no CPU fallback") from e
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.pbl_abi_version() != 4:
        raise RuntimeError("libpbllm.so ABI version mismatch")
    _lib = lib
    return lib


def last_error() -> str:
    return load().pbl_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "libpbllm"):
    if rc != 0:
        raise RuntimeError(f"{what} failed (status {rc}): {last_error()}")
