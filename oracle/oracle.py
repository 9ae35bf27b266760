"""numpy front-end to oracle/libpbllm_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (and only as the checker / the timed CPU baseline).  The product package
pb-llm_b200/ never does.  Parity pinned by tests/test_oracle_golden.py against fixtures made
by executing the unmodified reference (oracle/gen_golden.py).

Every function cites the reference file:line (relative to /root/reference) it follows; the
arithmetic itself lives in pbllm_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpbllm_oracle.so")


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (building the checker is not using it)."""
    src = os.path.join(_HERE, "pbllm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpbllm_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        f32p, u8p, i64 = C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.c_int64
        _lib.orc_sign.argtypes = [f32p, f32p, i64]
        _lib.orc_binary_wsim.argtypes = [f32p, i64, i64, f32p]
        _lib.orc_xnor_wsim.argtypes = [f32p, i64, i64, f32p, f32p, f32p]
        _lib.orc_weight_quant_8bit.argtypes = [f32p, i64, i64, f32p, u8p, C.c_int]
        _lib.orc_outlier_gen_mask.argtypes = [f32p, i64, i64, C.c_double, u8p, f32p, f32p, f32p, C.c_int]
        _lib.orc_outlier_gen_mask.restype = i64
        _lib.orc_outlier_wsim.argtypes = [f32p, u8p, C.c_float, C.c_float, i64, i64, f32p, C.c_int]
        _lib.orc_outlier_train_scale.argtypes = [f32p, u8p, i64, C.c_int]
        _lib.orc_outlier_train_scale.restype = C.c_float
        _lib.orc_outlier_nbits.argtypes = [f32p, u8p, i64, i64, C.c_int]
        _lib.orc_outlier_nbits.restype = C.c_double
        _lib.orc_linear.argtypes = [f32p, f32p, f32p, i64, i64, i64, f32p]
        _lib.orc_bireal_forward.argtypes = [f32p, f32p, i64, i64, i64, f32p]
        _lib.orc_low_xnor_calibrate.argtypes = [f32p, u8p, i64, i64, i64, i64, f32p, f32p]
        _lib.orc_high_calibrate.argtypes = [f32p, i64, i64, C.c_int, f32p, f32p]
        _lib.orc_gptqpb_rtn.argtypes = [f32p, u8p, i64, i64, i64, C.c_int, f32p, f32p, f32p, C.c_int]
    return _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a, t=C.c_float):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def linear(x, w, bias=None):
    """F.linear (quant/quantizer.py:86,193; quant/outlier_quantizer.py:105), double accumulate."""
    x, w = _f32(x), _f32(w)
    lead = x.shape[:-1]
    K = x.shape[-1]
    N = w.shape[0]
    assert w.shape[1] == K
    x2 = x.reshape(-1, K)
    b = None if bias is None else _f32(bias)
    y = np.empty((x2.shape[0], N), np.float32)
    lib().orc_linear(_p(x2), _p(w), _p(b), x2.shape[0], N, K, _p(y))
    return y.reshape(*lead, N)


def binary_wsim(W):
    """BinaryLinear / FdaBinaryLinear effective weight (quant/quantizer.py:84-85,123-128)."""
    W = _f32(W)
    out = np.empty_like(W)
    lib().orc_binary_wsim(_p(W), W.shape[0], W.shape[1], _p(out))
    return out


def xnor_wsim(W):
    """XnorBinaryLinear / IrBinaryLinear effective weight (quant/quantizer.py:181-189,98-105).
    Returns (w_sim, mu[N], alpha[N])."""
    W = _f32(W)
    N, K = W.shape
    out, mu, al = np.empty_like(W), np.empty(N, np.float32), np.empty(N, np.float32)
    lib().orc_xnor_wsim(_p(W), N, K, _p(out), _p(mu), _p(al))
    return out, mu, al


def weight_quant_8bit(W, half_mode=False):
    """weight_quant_8bit (quant/outlier_quantizer.py:10-29). Returns (simulated, codes)."""
    W = _f32(W)
    sim, codes = np.empty_like(W), np.empty(W.shape, np.uint8)
    lib().orc_weight_quant_8bit(_p(W), W.shape[0], W.shape[1], _p(sim), _p(codes, C.c_uint8), int(half_mode))
    return sim, codes


def outlier_state(W, outlier_fraction, half_mode=False):
    """gen_outlier_mask (quant/outlier_quantizer.py:54-81).
    Returns dict(mask bool[N,K] (True = salient), binary_scale, w8, thr=(lo,hi), count, nbits)."""
    W = _f32(W)
    N, K = W.shape
    mask = np.empty((N, K), np.uint8)
    w8 = np.empty_like(W)
    bs = np.zeros(1, np.float32)
    thr = np.zeros(2, np.float32)
    cnt = lib().orc_outlier_gen_mask(_p(W), N, K, float(outlier_fraction), _p(mask, C.c_uint8), _p(bs), _p(w8),
                                     _p(thr), int(half_mode))
    if cnt < 0:
        raise ValueError("kthvalue rank out of range (outlier_fraction too small for this layer)")
    nbits = lib().orc_outlier_nbits(_p(w8), _p(mask, C.c_uint8), N, K, int(half_mode))
    return dict(mask=mask.astype(bool), binary_scale=float(bs[0]), w8=w8, thr=(float(thr[0]), float(thr[1])),
                count=int(cnt), nbits=float(nbits))


def outlier_wsim(state, outlier_scale=1.0, training=False, half_mode=False):
    """binarize_except_outliers (quant/outlier_quantizer.py:83-99)."""
    w8 = _f32(state["w8"])
    mask = np.ascontiguousarray(state["mask"].astype(np.uint8))
    N, K = w8.shape
    bs = state["binary_scale"]
    if training:  # :90-93
        bs = float(lib().orc_outlier_train_scale(_p(w8), _p(mask, C.c_uint8), N * K, int(half_mode)))
    out = np.empty_like(w8)
    lib().orc_outlier_wsim(_p(w8), _p(mask, C.c_uint8), bs, float(outlier_scale), N, K, _p(out), int(half_mode))
    return out


def bireal_forward(x, W):
    """BiRealLinear.forward value (quant/quantizer.py:151-169); bias is dropped (:168)."""
    x, W = _f32(x), _f32(W)
    lead, K = x.shape[:-1], x.shape[-1]
    x2 = x.reshape(-1, K)
    y = np.empty((x2.shape[0], W.shape[0]), np.float32)
    lib().orc_bireal_forward(_p(x2), _p(W), x2.shape[0], W.shape[0], K, _p(y))
    return y.reshape(*lead, W.shape[0])


def gptqpb_rtn(W, low_mask, groupsize=-1, bits=8, to_half=True):
    """GPTQ-PB output weight with disable_gptq=True (gptq_pb/gptq.py:116-128,180-184;
    low_quant.py:25-32,75-82; high_quant.py:6-8,29-67). low_mask True = binarized.
    Returns (w_out[N,K], lo[N,G], hi[N,G])."""
    W = _f32(W)
    N, K = W.shape
    m = np.ascontiguousarray(np.asarray(low_mask).astype(np.uint8))
    gs = K if groupsize <= 0 else groupsize
    G = (K + gs - 1) // gs
    out, lo, hi = np.empty_like(W), np.empty((N, G), np.float32), np.empty((N, G), np.float32)
    lib().orc_gptqpb_rtn(_p(W), _p(m, C.c_uint8), N, K, int(groupsize), int(bits), _p(out), _p(lo), _p(hi),
                         int(to_half))
    return out, lo, hi


# ---- whole-module forwards (what the parity tests compare the CUDA path against) ---------

def forward_binary(x, W, bias=None):
    """BinaryLinear.forward (quant/quantizer.py:84-86)."""
    return linear(x, binary_wsim(W), bias)


def forward_xnor(x, W, bias=None):
    """XnorBinaryLinear.forward (quant/quantizer.py:191-193)."""
    return linear(x, xnor_wsim(W)[0], bias)


def forward_outlier(x, W, bias, outlier_fraction, outlier_scale=1.0, half_mode=False):
    """BinaryXnorExceptOutliersLinear.forward, eval mode (quant/outlier_quantizer.py:101-106)."""
    st = outlier_state(W, outlier_fraction, half_mode)
    return linear(x, outlier_wsim(st, outlier_scale, False, half_mode), bias)
