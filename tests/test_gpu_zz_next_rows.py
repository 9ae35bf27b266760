"""GPU parity tests of the SURVEY 8(f) "next" rows (BiRealLinear XNOR-popcount, ...). Kept in a file that sorts AFTER
the hot-path parity tests so that a failure here can never hide hot-path rows behind `pytest -x`."""
import numpy as np
import pytest
import torch

import pbllm_b200 as pb
from pbllm_b200 import _lib
from oracle import oracle as orc
from oracle.gen_golden import make_x
from test_gpu_parity import DEV, load, t, relmax, rounded

pytestmark = pytest.mark.gpu


# ---- BiRealLinear: the XNOR-popcount layer ------------------------------------------------
def test_golden_bireal_xnor_popcount():
    g = load("quantizer_small")
    m = pb.BiRealLinear(torch.from_numpy(g["W"]), torch.from_numpy(g["b"])).to(DEV)
    y = m(t(g["x"]))
    assert y.dtype == torch.float32 and y.shape == (2, 3, 96)
    assert relmax(y, g["y_BiRealLinear"]) <= 2e-6                   # the executed reference; bias is dropped (:168)
    assert relmax(m(t(g["x"]).half()), orc.bireal_forward(rounded(g["x"], torch.float16), g["W"])) <= 2e-6


@pytest.mark.parametrize("xdtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("N,K,M", [(128, 64, 1), (100, 70, 3), (300, 520, 8), (768, 768, 4), (4096, 4096, 8), (1024, 11008, 19),
                                   (300, 520, 9), (768, 768, 16), (768, 768, 17), (1024, 4096, 33), (300, 520, 64), (300, 520, 65)])
def test_bireal_matches_oracle_and_popcount_identity(xdtype, N, K, M):
    rs = np.random.RandomState(N + K + M)
    W = (rs.standard_normal((N, K)) * 0.02).astype(np.float32)
    W[rs.rand(N, K) < 0.01] = 0.0                                   # sign(0) = 0 weights
    x = rounded(make_x(N + M, (M, K)), xdtype)
    x[rs.rand(M, K) < 0.05] = 0.0                                   # and sign(0) = 0 activations (e.g. after ReLU)
    m = pb.BiRealLinear(torch.from_numpy(W), None).to(DEV)
    y = m(t(x, xdtype))
    ref = orc.bireal_forward(x, W) if M * N * K <= 4e8 else \
        (torch.sign(t(x)).double() @ (t(W).abs().mean(1, keepdim=True) * torch.sign(t(W))).double().t()).cpu().numpy()
    assert relmax(y, ref) <= 2e-6
    # the north star's literal formula: alpha_i * (2*popcount(xnor) - K) when no operand is zero
    Wn, xn = np.where(W == 0, 0.01, W).astype(np.float32), np.where(x == 0, 1.0, x).astype(np.float32)
    yn = pb.BiRealLinear(torch.from_numpy(Wn), None).to(DEV)(t(xn, xdtype)).cpu().numpy().astype(np.float64)
    alpha = np.abs(Wn).astype(np.float64).mean(1)
    agree = ((xn > 0)[:, None, :] == (Wn > 0)[None, :, :]).sum(-1) if M * N * K <= 2e7 else None
    if agree is not None:
        assert np.abs(yn - alpha[None, :] * (2.0 * agree - K)).max() <= 1e-6 * np.abs(yn).max()



def test_bireal_stream_k_and_row_group_kernels_agree():
    """pbl_bireal_forward_ws (stream-K XNOR kernel, zeroed reduction workspace) against pbl_bireal_forward (one CTA per
    row group): same integer counts, fp32 folding in a different order; the stream-K result is repeatable bit for bit."""
    lib = _lib.load()
    for (N, K, M, zeros) in [(768, 768, 8, 0.0), (3072, 768, 5, 0.01), (130, 200, 19, 0.02), (4096, 4096, 8, 0.0), (768, 768, 9, 0.0),
                             (768, 768, 16, 0.01), (1024, 2048, 17, 0.0), (512, 1024, 64, 0.0)]:
        rs = np.random.RandomState(N + K)
        W = (rs.standard_normal((N, K)) * 0.02).astype(np.float32)
        W[rs.rand(N, K) < zeros] = 0.0
        m = pb.BiRealLinear(torch.from_numpy(W), None).to(DEV)
        p = m.packed()
        assert (p.sign_planes is not None) == (zeros == 0.0)
        x = t(rounded(make_x(N + M, (M, K)), torch.float16), torch.float16)
        y_new = p.bireal_forward(x)
        assert int(lib.pbl_bireal_fixup_workspace(p.handle, M)) > 0
        for _ in range(3):
            assert torch.equal(y_new, p.bireal_forward(x))
        y_old = torch.empty_like(y_new)
        ws = torch.empty(p.bireal_workspace_bytes(M), dtype=torch.uint8, device=DEV)
        rc = lib.pbl_bireal_forward(p.handle, x.data_ptr(), K, 0, y_old.data_ptr(), N, M, ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert rc == 0, _lib.last_error()
        assert relmax(y_new, y_old.cpu().numpy()) <= 2e-6


# ---- GPTQ-PB calibration on the device: pb.gptq_pb against the executed reference's LowHighGPT ------------------
@pytest.mark.parametrize("tag,gs,metric,disable", [("gptqpb_rtn_g-1_mag", -1, "magnitude", True), ("gptqpb_rtn_g128_hes", 128, "hessian", True),
                                                   ("gptqpb_gptq_g-1_hes", -1, "hessian", False), ("gptqpb_gptq_g128_mag", 128, "magnitude", False)])
def test_gptqpb_calibration_matches_reference_fixture(tag, gs, metric, disable, tmp_path, monkeypatch):
    """Same inputs as oracle/gen_golden.py section 5 (the reference's own LowHighGPT, run on the CPU): weights, two calibration
    batches, low_frac 0.9.  The mask and the quantiser tables must match; the fake-quant weights must match up to the few
    rounding decisions that a Cholesky factor computed by another LAPACK can flip, and the packed layer built from them
    must reproduce the reference layer's outputs."""
    from pbllm_b200 import gptq_pb
    from oracle.gen_golden import make_weight
    g = load(tag)
    N, K = 48, 256
    W = make_weight(51, N, K, "heavy", np.float16)
    assert np.array_equal(W, g["W"])
    calib = make_x(52, (2, 64, K)) * (1.0 + np.arange(K, dtype=np.float32) / 64.0)
    monkeypatch.chdir(tmp_path)
    layer = torch.nn.Linear(K, N, bias=False).half().to(DEV)
    layer.weight.data = t(W).clone()
    layer.global_name = "synthetic/" + tag
    lowq = gptq_pb.LowQuantizer(layer.weight, method="xnor", groupsize=gs)
    highq = gptq_pb.HighQuantizer(8, True, False, False)
    gp = gptq_pb.LowHighGPT(layer, lowq, highq, salient_metric=metric, disable_gptq=disable)
    gp.add_batch(t(calib[0]), None)
    gp.add_batch(t(calib[1]), None)
    res = gp.fasterquant(0.9, blocksize=128, percdamp=0.01)
    assert layer.weight.dtype == torch.float16 and res["error"] >= 0
    mask = torch.load(f"outputs/mask/mask_0.9_synthetic_{tag}.pkl").cpu().numpy()
    assert (mask == g["low_mask"]).mean() >= 0.999 and abs(mask.mean() - 0.9) < 1e-3
    assert relmax(highq.scale.flatten(), g["high_scale"]) <= 1e-6 and np.array_equal(highq.zero.flatten().cpu().numpy(), g["high_zero"])
    assert relmax(lowq.mean.squeeze(-1), g["low_mean"]) <= 1e-4 and relmax(lowq.scale.squeeze(-1), g["low_scale"]) <= 1e-4
    Wq, ref = layer.weight.data.float().cpu().numpy(), g["Wq"].astype(np.float32)
    same = (Wq == ref).mean()
    if disable:
        assert same >= 0.999, same                                # RTN: no error feedback, nothing to amplify
    else:
        assert same >= 0.97 and np.linalg.norm(Wq - ref) <= 3e-2 * np.linalg.norm(ref), (same, np.linalg.norm(Wq - ref) / np.linalg.norm(ref))
        # calibration quality: the layer-output error on the calibration data equals the reference's to a few percent
        X = calib.reshape(-1, K).astype(np.float64)
        e_ours = np.linalg.norm(X @ (Wq - W.astype(np.float32)).T.astype(np.float64))
        e_ref = np.linalg.norm(X @ (ref - W.astype(np.float32)).T.astype(np.float64))
        assert abs(e_ours - e_ref) <= 0.05 * e_ref, (e_ours, e_ref)
    # the calibrated layer + its mask file are exactly what replace_from_fakequant consumes
    m = pb.PackedFakeQuantLinear.from_linear(layer, torch.load(f"outputs/mask/mask_0.9_synthetic_{tag}.pkl"), gs)
    assert torch.equal(m.dense_weight(), layer.weight.data)
    y = m(t(g["x"]))
    ref_y = orc.linear(g["x"].astype(np.float32), Wq)
    assert relmax(y, ref_y) <= 1e-3
    assert m.packed().salient_count() <= int((~mask).sum()) + 0.01 * mask.size


# ---- magnitude thresholds on the device: radix select against torch.kthvalue ------------------------------------
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("n", [1, 7, 1000, 4096 * 4096 + 3])
def test_kth_value_is_torch_kthvalue(dtype, n):
    from pbllm_b200.packing import kth_value
    gen = torch.Generator(device=DEV).manual_seed(n)
    x = (torch.randn(n, device=DEV, generator=gen) * 0.02).to(dtype)
    x[::5] = x[::5].abs()
    if n > 10:
        x[3], x[4] = 0.0, -0.0
    ks = sorted({1, n, max(1, n // 2), max(1, int(n * 0.05)), max(1, int(n * 0.95))})
    for k in ks:
        got = kth_value(x, k)
        ref = torch.kthvalue(x.float(), k)[0]              # exact in fp32 for 16-bit inputs
        assert got.dtype == dtype and float(got) == float(ref), (k, float(got), float(ref))
    xv = x[1:] if n > 8 else x                              # unaligned base pointer: scalar tail path
    assert float(kth_value(xv, 1)) == float(xv.float().min())
    with pytest.raises(IndexError):
        kth_value(x, n + 1)
