"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (/root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python oracle/gen_golden.py

The reference modules are imported as they are; the only harness-side shims are the ones
SURVEY.md 8c lists (no reference file is edited):
  * torch.Tensor.cuda -> identity   (quant/quantizer.py:33-34,116 call .cuda() at import time)
  * torch.cuda.synchronize / empty_cache -> no-ops  (gptq_pb/gptq.py:176,194)
Inputs come from numpy's legacy RandomState (stream-stable across numpy versions) so large
inputs can be regenerated from (seed, shape) instead of being stored; `make_weight` /
`make_x` below are the generators the tests re-use.
"""
from __future__ import annotations

import hashlib
import io
import contextlib
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def make_weight(seed, N, K, kind="normal", dtype=np.float32):
    """Deterministic synthetic latent weight. kind: normal (N(0,0.02^2)) | heavy (student-t nu=3)."""
    rs = np.random.RandomState(seed)
    if kind == "normal":
        w = rs.standard_normal((N, K)) * 0.02
    elif kind == "heavy":
        w = rs.standard_t(3, (N, K)) * 0.02
    else:
        raise ValueError(kind)
    return w.astype(np.float32).astype(dtype)


def make_x(seed, shape, dtype=np.float32):
    return np.random.RandomState(seed).standard_normal(shape).astype(np.float32).astype(dtype)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def import_reference():
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.empty_cache = lambda *a, **k: None
    sys.path.insert(0, "/root/reference")
    sys.path.insert(0, "/root/reference/gptq_pb")
    import quant  # noqa
    return quant


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def main():
    quant = import_reference()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(1)
    meta = {}

    # ---- 1. quantizer.py classes, small, everything stored ---------------------------------
    N, K = 96, 160
    W = make_weight(11, N, K)
    b = make_x(12, (N,)) * 0.1
    x = make_x(13, (2, 3, K))  # 3-D activations (any leading dims)
    out = dict(W=W, b=b, x=x)
    with torch.no_grad():
        for name in ["BinaryLinear", "XnorBinaryLinear", "IrBinaryLinear", "FdaBinaryLinear", "BiRealLinear"]:
            mod = getattr(quant, name)(t(W), t(b))
            out["y_" + name] = mod(t(x)).numpy()
        out["y_Xnor_nobias"] = quant.XnorBinaryLinear(t(W), None)(t(x)).numpy()
        xm = quant.XnorBinaryLinear(t(W), t(b))
        out["wsim_Xnor"] = xm.quant_weight().numpy()
        out["wsim_Binary"] = t(W).sign().numpy()
        sd = xm.state_dict()
        assert sorted(sd.keys()) == ["bias", "weight"]
        swd = xm.get_save_weight_dict()
        assert swd["weight"].dtype == torch.float16
    np.savez_compressed(os.path.join(GOLD, "quantizer_small.npz"), **out)

    # ---- 2. BASELINE config 1: OPT-125m-shaped 768x768 XnorBinaryLinear on CPU -------------
    N = K = 768
    for seed in range(100, 200):
        W = make_weight(seed, N, K)
        mu = t(W).mean(-1, keepdim=True)
        margin = ((t(W) - mu).abs() / mu.abs().clamp_min(1e-12)).min().item()
        if margin > 1e-3:  # no element within 1e-3*|mu| of its row mean: sign bits are device-proof
            break
    b = make_x(seed + 1, (N,)) * 0.1
    x = make_x(seed + 2, (4, K))
    with torch.no_grad():
        m = quant.XnorBinaryLinear(t(W), t(b))
        y = m(t(x)).numpy()
        ws = m.quant_weight()
        alpha = ws.abs().amax(-1).numpy()
        bits = np.packbits((ws > 0).numpy(), axis=-1, bitorder="little")
        assert (ws != 0).all()
        bm = quant.BinaryLinear(t(W), t(b))
        yb = bm(t(x)).numpy()
    np.savez_compressed(os.path.join(GOLD, "cfg1_xnor_768.npz"), seed=seed, y=y, y_binary=yb, alpha=alpha,
                        sign_bits=bits, w_sha=sha(W), x_sha=sha(x), b_sha=sha(b), wsim_sha=sha(ws.numpy()),
                        margin=margin)

    # ---- 3. outlier class: small fp32 / fp16, and the 768^2 known answers ------------------
    for tag, N, K, frac, dt, kind, seed in [
        ("outlier_f32_small", 128, 192, 0.1, np.float32, "normal", 21),
        ("outlier_f16_small", 128, 192, 0.1, np.float16, "normal", 22),
        ("outlier_f32_heavy", 64, 320, 0.2, np.float32, "heavy", 23),
        ("outlier_f16_heavy", 64, 320, 0.05, np.float16, "heavy", 24),
    ]:
        W = make_weight(seed, N, K, kind, dt)
        b = (make_x(seed + 1, (N,)) * 0.1).astype(dt)
        x = make_x(seed + 2, (5, K), dt)
        with torch.no_grad():
            m = quant.BinaryXnorExceptOutliersLinear(t(W).clone(), t(b), frac)
            m.eval()
            y = quiet(m, t(x)).float().numpy()
            ws = m.binarize_except_outliers().float().numpy()
            reg = m.to_regular_linear()
            y_reg = reg(t(x)).float().numpy()
            assert np.array_equal(y, y_reg)
            m.train()
            ws_train = m.binarize_except_outliers().float().numpy()
            scale_train = float(m.binary_scale.float())
            m.eval()
        np.savez_compressed(os.path.join(GOLD, tag + ".npz"), W=W, b=b, x=x, frac=frac, y=y, wsim=ws,
                            mask=m.outlier_mask.numpy(), w8=m.weight.data.float().numpy(),
                            wsim_train=ws_train, scale_train=scale_train,
                            binary_scale_shape=np.array(m.binary_scale.shape), nbits=m.outlier_nbits)

    N = K = 768
    W = make_weight(31, N, K)
    x = make_x(33, (4, K))
    with torch.no_grad():
        m = quant.BinaryXnorExceptOutliersLinear(t(W).clone(), None, 0.1)
        m.eval()
        y = quiet(m, t(x)).numpy()
        ws = m.binarize_except_outliers().numpy()
        nm = ws[~m.outlier_mask.numpy()]
        lev = np.unique(nm)
    np.savez_compressed(os.path.join(GOLD, "outlier_768_kat.npz"), seed=31, y=y, count=int(m.outlier_mask.sum()),
                        binary_scale=float(m.binary_scale), nbits=m.outlier_nbits, nonsalient_levels=lev,
                        n_zero=int((ws == 0).sum()), n_neg=int((ws < 0).sum()), w_sha=sha(W),
                        mask_bits=np.packbits(m.outlier_mask.numpy(), axis=-1, bitorder="little"),
                        wsim_sha=sha(ws))

    # ---- 4. Hessian subclass: mask file present / absent ---------------
    N, K = 64, 128
    W = make_weight(41, N, K)
    x = make_x(43, (3, K))
    low_mask = np.random.RandomState(44).rand(N, K) < 0.9  # True = binarized (file convention)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td, torch.no_grad():
        os.chdir(td)
        try:
            os.makedirs("gptq_pb/outputs/mask")
            m = quant.BinaryXnorExceptOutliersLinearHessian(t(W).clone(), None, 0.1)
            m.global_name = "synthetic/model.layers.0.q_proj"
            torch.save(t(low_mask), f"gptq_pb/outputs/mask/mask_0.9_{m.global_name.replace('/', '_')}.pkl")
            m.eval()
            quiet(m.gen_outlier_mask)
            assert m.binary_scale is None
            raised = False
            try:
                m(t(x))
            except TypeError:
                raised = True
            assert raised  # eval forward with binary_scale None multiplies by None
            m.train()
            y_train = m(t(x)).numpy()
            m.eval()
            y_eval = m(t(x)).numpy()
            ws = m.binarize_except_outliers().numpy()
            m2 = quant.BinaryXnorExceptOutliersLinearHessian(t(W).clone(), None, 0.1)
            m2.global_name = "synthetic/missing"
            m2.eval()
            y_fallback = quiet(m2, t(x)).numpy()
            m3 = quant.BinaryXnorExceptOutliersLinear(t(W).clone(), None, 0.1)
            m3.eval()
            assert np.array_equal(y_fallback, quiet(m3, t(x)).numpy())
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(GOLD, "hessian_mask.npz"), W=W, x=x, low_mask=low_mask, y_train=y_train,
                        y_eval=y_eval, wsim=ws, binary_scale=float(m.binary_scale), y_fallback=y_fallback,
                        outlier_mask=m.outlier_mask.numpy())

    # ---- 5. GPTQ-PB output format: the reference's own LowHighGPT on tiny layers -----------
    import torch.nn as nn
    from gptq import LowHighGPT
    from low_quant import LowQuantizer
    from high_quant import HighQuantizer
    N, K = 48, 256
    for tag, gs, metric, disable in [("gptqpb_rtn_g-1_mag", -1, "magnitude", True),
                                     ("gptqpb_rtn_g128_hes", 128, "hessian", True),
                                     ("gptqpb_gptq_g-1_hes", -1, "hessian", False),
                                     ("gptqpb_gptq_g128_mag", 128, "magnitude", False)]:
        W = make_weight(51, N, K, "heavy", np.float16)
        calib = make_x(52, (2, 64, K)) * (1.0 + np.arange(K, dtype=np.float32) / 64.0)
        with tempfile.TemporaryDirectory() as td:
            os.chdir(td)
            try:
                os.mkdir("outputs")
                layer = nn.Linear(K, N, bias=False).half()
                layer.weight.data = t(W).clone()
                layer.global_name = "synthetic/" + tag
                lowq = LowQuantizer(layer.weight, method="xnor", groupsize=gs)
                highq = HighQuantizer(8, True, False, False)
                g = LowHighGPT(layer, lowq, highq, salient_metric=metric, disable_gptq=disable)
                g.add_batch(t(calib[0]), None)
                g.add_batch(t(calib[1]), None)
                quiet(g.fasterquant, 0.9, blocksize=128, percdamp=0.01)
                mask = torch.load(f"outputs/mask/mask_0.9_synthetic_{tag}.pkl").numpy()
                Wq = layer.weight.data.clone()
                assert Wq.dtype == torch.float16
                lo_mean = lowq.mean.squeeze(-1).numpy()   # [G, N]
                lo_scale = lowq.scale.squeeze(-1).numpy()
                hs, hz = highq.scale.flatten().numpy(), highq.zero.flatten().numpy()
            finally:
                os.chdir(cwd)
        x = make_x(53, (4, K), np.float16)
        with torch.no_grad():
            y = torch.nn.functional.linear(t(x).float(), Wq.float()).numpy()
        np.savez_compressed(os.path.join(GOLD, tag + ".npz"), W=W, Wq=Wq.numpy(), low_mask=mask, x=x, y=y,
                            groupsize=gs, low_mean=lo_mean, low_scale=lo_scale, high_scale=hs, high_zero=hz)

    print("golden fixtures written to", GOLD)
    for f in sorted(os.listdir(GOLD)):
        print(f"  {f:32s} {os.path.getsize(os.path.join(GOLD, f)):9d} B")


if __name__ == "__main__":
    main()
