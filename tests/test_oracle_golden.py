"""Pins the CPU oracle (oracle/pbllm_oracle.c) against fixtures produced by executing the
unmodified reference (oracle/gen_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from oracle.gen_golden import make_weight, make_x, sha

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(G, name + ".npz"), allow_pickle=False)


def relmax(a, b):
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / max(np.abs(b).max(), 1e-30))


def test_quantizer_small_all_classes():
    g = load("quantizer_small")
    W, b, x = g["W"], g["b"], g["x"]
    assert np.array_equal(orc.binary_wsim(W), g["wsim_Binary"])
    ws, mu, al = orc.xnor_wsim(W)
    # sign pattern must be identical; alpha within an ulp or two of torch's fp32 reduction
    assert np.array_equal(np.sign(ws), np.sign(g["wsim_Xnor"]))
    assert relmax(ws, g["wsim_Xnor"]) < 5e-7
    assert relmax(orc.forward_binary(x, W, b), g["y_BinaryLinear"]) < 2e-6
    assert relmax(orc.forward_xnor(x, W, b), g["y_XnorBinaryLinear"]) < 2e-6
    assert relmax(orc.forward_xnor(x, W, None), g["y_Xnor_nobias"]) < 2e-6
    # forward-equivalence classes: Ir == Xnor, Fda == Binary, exactly
    assert np.array_equal(g["y_IrBinaryLinear"], g["y_XnorBinaryLinear"])
    assert np.array_equal(g["y_FdaBinaryLinear"], g["y_BinaryLinear"])
    assert relmax(orc.bireal_forward(x, W), g["y_BiRealLinear"]) < 2e-6
    assert orc.forward_xnor(x, W, b).shape == (2, 3, W.shape[0])


def test_cfg1_xnor_768_regenerated_inputs():
    g = load("cfg1_xnor_768")
    seed = int(g["seed"])
    W, b, x = make_weight(seed, 768, 768), make_x(seed + 1, (768,)) * 0.1, make_x(seed + 2, (4, 768))
    assert sha(W) == str(g["w_sha"]) and sha(x) == str(g["x_sha"]) and sha(b) == str(g["b_sha"])
    ws, mu, al = orc.xnor_wsim(W)
    assert np.array_equal(np.packbits(ws > 0, axis=-1, bitorder="little"), g["sign_bits"])
    assert relmax(al, g["alpha"]) < 5e-7
    assert relmax(orc.linear(x, ws, b), g["y"]) < 2e-6
    assert relmax(orc.forward_binary(x, W, b), g["y_binary"]) < 2e-6


@pytest.mark.parametrize("tag,half", [("outlier_f32_small", False), ("outlier_f16_small", True),
                                      ("outlier_f32_heavy", False), ("outlier_f16_heavy", True)])
def test_outlier_small(tag, half):
    g = load(tag)
    W, b, x = g["W"].astype(np.float32), g["b"].astype(np.float32), g["x"].astype(np.float32)
    st = orc.outlier_state(W, float(g["frac"]), half_mode=half)
    assert np.array_equal(st["mask"], g["mask"])
    assert tuple(g["binary_scale_shape"]) == (1, 1)            # scalar alpha, SURVEY 8a row a6
    assert np.array_equal(st["w8"], g["w8"])                    # 8-bit fake-quant incl. the uint8 wrap, bit-exact
    ws = orc.outlier_wsim(st, half_mode=half)
    sal = g["mask"]
    assert np.array_equal(ws[sal], g["wsim"][sal])
    assert np.array_equal(ws == 0, g["wsim"] == 0)
    tol = 1e-3 if half else 5e-7                                # alpha: one fp16 ulp / fp32 reduction order
    assert relmax(ws, g["wsim"]) < tol
    assert abs(st["nbits"] - float(g["nbits"])) < 1e-12
    ws_t = orc.outlier_wsim(st, training=True, half_mode=half)  # train-mode alpha differs
    assert relmax(ws_t, g["wsim_train"]) < tol
    y = orc.linear(x, g["wsim"], b)                             # forward from the reference's own w_sim
    assert relmax(y, g["y"]) < (2e-3 if half else 2e-6)
    y2 = orc.forward_outlier(x, W, b, float(g["frac"]), half_mode=half)
    assert relmax(y2, g["y"]) < (3e-3 if half else 2e-6)


def test_outlier_768_known_answers():
    g = load("outlier_768_kat")
    W = make_weight(int(g["seed"]), 768, 768)
    assert sha(W) == str(g["w_sha"])
    st = orc.outlier_state(W, 0.1)
    assert st["count"] == int(g["count"]) == 58982              # int(n*f/2) ranks 29491 / 560332
    assert np.array_equal(np.packbits(st["mask"], axis=-1, bitorder="little"), g["mask_bits"])
    assert abs(st["binary_scale"] - float(g["binary_scale"])) < 1e-8
    assert abs(st["nbits"] - float(g["nbits"])) < 1e-12
    ws = orc.outlier_wsim(st)
    assert int((ws < 0).sum()) == int(g["n_neg"]) == 0          # zero-point wrap: every weight >= 0 (fact 5)
    assert int((ws == 0).sum()) == int(g["n_zero"])
    lev = np.unique(ws[~st["mask"]])
    assert lev.size == 2 and lev[0] == 0.0 and relmax(lev, g["nonsalient_levels"]) < 5e-7
    x = make_x(33, (4, 768))
    assert relmax(orc.linear(x, ws, None), g["y"]) < 2e-6


def test_hessian_mask_file_semantics():
    g = load("hessian_mask")
    W, x = g["W"], g["x"]
    assert np.array_equal(g["outlier_mask"], ~g["low_mask"])   # outlier_quantizer.py:138
    w8, _ = orc.weight_quant_8bit(W)
    st = dict(mask=g["outlier_mask"], w8=w8, binary_scale=0.0)
    ws = orc.outlier_wsim(st, training=True)                    # alpha only exists after a train-mode forward
    assert relmax(ws, g["wsim"]) < 5e-7
    assert relmax(orc.linear(x, ws), g["y_train"]) < 2e-6
    assert np.array_equal(g["y_train"], g["y_eval"])
    assert relmax(orc.forward_outlier(x, W, None, 0.1), g["y_fallback"]) < 2e-6


@pytest.mark.parametrize("tag", ["gptqpb_rtn_g-1_mag", "gptqpb_rtn_g128_hes", "gptqpb_gptq_g-1_hes",
                                 "gptqpb_gptq_g128_mag"])
def test_gptqpb_format(tag):
    g = load(tag)
    Wq, mask, gs = g["Wq"].astype(np.float32), g["low_mask"], int(g["groupsize"])
    N, K = Wq.shape
    gs_eff = K if gs <= 0 else gs
    assert abs(mask.mean() - 0.9) < 2e-3                        # low_frac of each column group (gptq.py:83-99)
    # format property a10: <= 2 dominant low values per (row, group) (+ a rare third, mu, when sign(0))
    for gi in range(K // gs_eff):
        sl = slice(gi * gs_eff, (gi + 1) * gs_eff)
        mean, scale = g["low_mean"][gi], g["low_scale"][gi]
        lo = (mean - scale).astype(np.float16).astype(np.float32)
        hi = (mean + scale).astype(np.float16).astype(np.float32)
        mid = mean.astype(np.float16).astype(np.float32)
        blk, m = Wq[:, sl], mask[:, sl]
        ok = (blk == lo[:, None]) | (blk == hi[:, None]) | (blk == mid[:, None])
        assert ok[m].all()
    # salient entries sit on the per-row 8-bit grid s*(q - z), q in [0,255] (high_quant.py:6-8)
    s, z = g["high_scale"], g["high_zero"]
    q = np.rint(Wq / s[:, None] + z[:, None])
    sal = ~mask
    lev = (s[:, None] * (q - z[:, None])).astype(np.float16).astype(np.float32)   # fp16 cast: gptq.py:180-184
    assert np.array_equal(lev[sal], Wq[sal]) and (q[sal] >= 0).all() and (q[sal] <= 255).all()
    if "rtn" in tag:                                             # RTN has no error feedback: restatement is exact
        out, lo, hi = orc.gptqpb_rtn(g["W"].astype(np.float32), mask, gs, 8, True)
        assert np.array_equal(out, Wq)
    assert relmax(orc.linear(g["x"].astype(np.float32), Wq), g["y"]) < 2e-6
