"""The C-ABI library loads here (no GPU) and exports every symbol include/pbllm.h declares;
compute entry points refuse to run without a device instead of falling back."""
import ctypes as C
import os
import re

import pytest
import torch

import pbllm_b200 as pb
from pbllm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "pbllm.h")).read()
    return re.findall(r"PBL_API\s+[\w\s\*]+?\b(pbl_\w+)\s*\(", src)


def test_every_declared_symbol_is_exported_and_bound():
    names = header_functions()
    assert len(names) >= 15 and len(set(names)) == len(names)
    lib = C.CDLL(_lib.lib_path())
    for n in names:
        assert hasattr(lib, n), f"{n} declared in pbllm.h but not exported by libpbllm.so"
    assert set(names) == set(_lib.SYMBOLS), "ctypes binding table and header disagree"


def test_no_oracle_or_reference_in_product_path():
    # product package must not import/execute anything under oracle/ (nor the reference)
    pkg = os.path.join(ROOT, "pb-llm_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.lower().replace("# oracle", ""), f"{f} mentions oracle"
                assert "/root/reference" not in txt, f"{f} reads the reference tree"


def test_pack_sizes_host_only():
    sz = pb.pack_sizes(4096, 11008, -1, torch.float16)
    assert (sz.n_pad, sz.k_pad, sz.tiles_r, sz.tiles_c, sz.groups) == (4096, 11008, 32, 172, 1)
    assert sz.planes_bytes == 4096 * 11008 // 4          # 2 bit/weight: sign plane + salient bitmap
    assert sz.vptr_bytes == (32 * 172 * 4 + 1) * 4
    sz = pb.pack_sizes(100, 70, -1, torch.float32)       # ragged -> padded to whole tiles
    assert (sz.n_pad, sz.k_pad, sz.vals_elem_bytes) == (128, 128, 4)
    sz = pb.pack_sizes(256, 512, 128, torch.bfloat16)
    assert sz.groups == 4
    with pytest.raises(RuntimeError, match="multiple of 64"):
        pb.pack_sizes(256, 512, 100, torch.float16)
    with pytest.raises(RuntimeError, match="positive"):
        pb.pack_sizes(0, 512, -1, torch.float16)


def test_error_codes_without_compute():
    lib = _lib.load()
    assert lib.pbl_abi_version() == 4
    assert lib.pbl_layer_create(None, None) == -1 and "null" in _lib.last_error()
    assert lib.pbl_linear_forward(None, None, 0, None, 0, 1, None) == -1
    assert lib.pbl_forward_host_workspace(None, 4) == 0
    assert lib.pbl_linear_forward_ws(None, None, 0, None, 0, 1, None, 0, None) == -1
    assert lib.pbl_decode_workspace_bytes(None, 8) == 0
    assert lib.pbl_decode_variant(None, None, 0, 8) == -1
    assert lib.pbl_stream_layout(4096, 4096, -1, 0, None) == -1 and lib.pbl_stream_position(0, 0, None) == -1
    ss = _lib.PblStreamSizes()
    assert lib.pbl_stream_layout(4096, 11008, -1, 0, C.byref(ss)) == 0
    assert (ss.blocks, ss.fsign_bytes, ss.eptr_bytes) == (128 * 172, 4096 * 11008 // 8, (128 * 172 + 1) * 4)   # 1 bit / weight
    assert lib.pbl_stream_layout(64, 64, -1, 2, C.byref(ss)) == -2 and "fp16" in _lib.last_error()             # fp32: planes layout
    out4 = (C.c_uint32 * 4)()
    assert lib.pbl_stream_position(32, 0, out4) == -3
    assert lib.pbl_launch_count() >= 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_fails_loudly_without_gpu():
    lib = _lib.load()
    assert lib.pbl_device_check() == -4 and "no CPU fallback" in _lib.last_error()
    w = torch.randn(64, 64)
    with pytest.raises(RuntimeError, match="CUDA"):
        pb.PackedLinear.from_dense(w)
    m = pb.XnorBinaryLinear(w, None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(2, 64))
    # a raw ABI compute call on host pointers is refused, not emulated
    buf = (C.c_float * 16)()
    rc = lib.pbl_pack_affine(C.cast(buf, C.c_void_p), 4, None, 4, 4, -1, 2, C.cast(buf, C.c_void_p), None)
    assert rc == -4
