"""Parity tests proper (need a B200): the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs, against the golden fixtures made from the executed reference,
and -- at BASELINE.json's full layer sizes -- through size-independent properties.

Tolerance (BASELINE.json north_star): 1e-3 relative, fp16. Written here as
    max|y - y_ref| / max|y_ref| <= 1e-3      (max-normalised; SURVEY.md 7 "Tolerance definition")
for fp16 I/O; 6e-3 for bf16 I/O (not a north-star dtype: its output rounds at 2^-9 and the salient weights are held in
the layer's own 16-bit type in the kernel's +-1 units, a second 2^-9 rounding); 2e-5 for fp32 I/O.
Packing is integer/bit work: unpack(pack(w)) must be bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import pbllm_b200 as pb
from pbllm_b200 import _lib
from oracle import oracle as orc
from oracle.gen_golden import make_weight, make_x

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"
TOL = {torch.float16: 1e-3, torch.bfloat16: 6e-3, torch.float32: 2e-5}


def load(name):
    return np.load(os.path.join(G, name + ".npz"), allow_pickle=False)


def t(a, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return x if dtype is None else x.to(dtype)


def relmax(y, ref):
    y = y.detach().float().cpu().numpy().astype(np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.abs(y - ref).max() / max(np.abs(ref).max(), 1e-30))


def rms_rel(y, ref):
    y = y.detach().float().cpu().numpy().astype(np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.sqrt(((y - ref) ** 2).mean()) / max(np.sqrt((ref ** 2).mean()), 1e-30))


def forced(p, x, kernel):
    """Run p.forward with PBL_FORCE_KERNEL=kernel (0 CUDA cores (fp32 layers), 1 two-phase prefill: expansion + tcgen05
    GEMM, 4 decode kernel in token passes)."""
    os.environ["PBL_FORCE_KERNEL"] = str(kernel)
    try:
        return p.forward(x)
    finally:
        os.environ.pop("PBL_FORCE_KERNEL", None)


def rounded(a, dtype):
    """numpy fp32 array holding dtype-representable values (what the device tensor really holds)."""
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).float().numpy()


def test_native_library_is_the_one_running():
    lib = _lib.load()
    assert lib.pbl_device_check() == 0, _lib.last_error()
    n0 = lib.pbl_launch_count()
    p = pb.PackedLinear.from_dense(torch.randn(64, 64, device=DEV, dtype=torch.float16).sign())
    p.forward(torch.randn(1, 64, device=DEV, dtype=torch.float16))
    torch.cuda.synchronize()
    assert lib.pbl_launch_count() >= n0 + 5   # affine, count, scan, fill, forward


# ---- packing: bit-exact ------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("N,K,gs", [(128, 64, -1), (96, 160, -1), (100, 70, -1), (1, 1, -1), (300, 520, -1),
                                    (256, 512, 128), (130, 384, 64), (4096, 4096, -1)])
def test_pack_unpack_roundtrip_bit_exact(dtype, N, K, gs):
    gen = torch.Generator(device="cpu").manual_seed(N * 131 + K)
    groups = 1 if gs <= 0 else (K + gs - 1) // gs
    gse = K if gs <= 0 else gs
    mu = torch.randn(N, groups, generator=gen) * 0.01
    al = torch.rand(N, groups, generator=gen) * 0.02 + 0.005
    sign = (torch.rand(N, K, generator=gen) < 0.5).float() * 2 - 1
    gi = torch.arange(K) // gse
    w = (mu[:, gi] + al[:, gi] * sign).to(dtype)
    low = torch.rand(N, K, generator=gen) < 0.9
    sal_vals = (torch.randn(N, K, generator=gen) * 0.05).to(dtype)
    w = torch.where(low, w, sal_vals)
    w[torch.rand(N, K, generator=gen) < 0.002] = 0               # third values (sign(0) zeros)
    wd = w.to(DEV)
    p = pb.PackedLinear.from_dense(wd, None, low_mask=low.to(DEV), groupsize=gs)
    assert torch.equal(p.unpack(), wd)
    frac = p.salient_count() / (N * K)
    assert 0.05 < frac < 0.2 or N * K < 4096
    assert p.stream_layout == (dtype != torch.float32)
    if p.stream_layout and N * K >= 4096:
        # salient values far smaller than both levels cannot be written as level + 16-bit correction within 7 ulps: they
        # must show up in the exception list (and still round-trip, asserted above)
        assert p.n_exc > 0 and p.n_exc < 0.2 * p.salient_count()
    p2 = pb.PackedLinear.from_dense(wd, None, low_mask=None, groupsize=gs)   # mask-free: min/max levels only
    assert torch.equal(p2.unpack(), wd)


def test_pack_planes_levels_and_counts_against_numpy():
    """fp32 layers: the planes layout."""
    rs = np.random.RandomState(5)
    N, K = 200, 330
    w = np.where(rs.rand(N, K) < 0.5, 0.25, -0.5).astype(np.float32)
    sal = rs.rand(N, K) < 0.1
    w[sal] = rs.standard_normal(sal.sum()).astype(np.float32)
    p = pb.PackedLinear.from_dense(t(w), None, low_mask=t(~sal))
    assert p.nnz == int(sal.sum()) and not p.stream_layout
    aff = p.affine.view(-1, 2).cpu().numpy()
    assert np.array_equal(aff[:N, 0], np.full(N, -0.5, np.float32)) and np.array_equal(aff[:N, 1], np.full(N, 0.25, np.float32))
    assert np.array_equal(aff[N:], np.zeros_like(aff[N:]))
    planes = p.planes.view(-1, 4).cpu().numpy().view(np.uint32)
    # tile (0,0), row 3: sign / salient bits of columns 0..63
    row = planes[3]
    bits = np.array([(row[j // 32] >> (j % 32)) & 1 for j in range(64)])
    salb = np.array([(row[2 + j // 32] >> (j % 32)) & 1 for j in range(64)])
    assert np.array_equal(salb, sal[3, :64].astype(int))
    assert np.array_equal(bits, ((w[3, :64] == 0.25) & ~sal[3, :64]).astype(int))
    vptr = p.vptr.cpu().numpy().view(np.uint32)
    assert vptr[0] == 0 and vptr[-1] == p.nnz and np.all(np.diff(vptr.astype(np.int64)) >= 0)
    v0 = p.vals[: int(vptr[1])].cpu().numpy()
    exp = np.concatenate([w[r, :64][sal[r, :64]] for r in range(32)])
    assert np.array_equal(v0, exp)                                 # (tile, row-group, row, column) order


# ---- forward vs oracle ---------------------------------------------------------------------------
def synth_wsim(N, K, gs, dtype, seed, sal_frac=0.1):
    rs = np.random.RandomState(seed)
    groups = 1 if gs <= 0 else (K + gs - 1) // gs
    gse = K if gs <= 0 else gs
    gi = np.arange(K) // gse
    mu = rs.standard_normal((N, groups)) * 0.004
    al = rs.rand(N, groups) * 0.02 + 0.005
    w = mu[:, gi] + al[:, gi] * np.where(rs.rand(N, K) < 0.5, 1.0, -1.0)
    low = rs.rand(N, K) >= sal_frac
    w = np.where(low, w, rs.standard_t(3, (N, K)) * 0.03)
    w = rounded(w.astype(np.float32), dtype)
    return w, low


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("N,K,gs,M", [(128, 64, -1, 1), (96, 160, -1, 3), (100, 70, -1, 5), (300, 520, -1, 8),
                                      (256, 512, 128, 2), (768, 768, -1, 1), (768, 768, -1, 4), (768, 3072, -1, 16),
                                      (33, 2048, -1, 1), (512, 1024, 256, 37)])
def test_forward_matches_oracle(dtype, N, K, gs, M):
    w, low = synth_wsim(N, K, gs, dtype, seed=N + K + M)
    b = rounded(np.random.RandomState(1).standard_normal(N).astype(np.float32) * 0.1, dtype)
    x = rounded(make_x(N * 7 + M, (M, K)), dtype)
    ref = orc.linear(x, w, b)
    for mask in (low, None):
        p = pb.PackedLinear.from_dense(t(w, dtype), t(b, dtype), None if mask is None else t(mask), gs)
        y = p.forward(t(x, dtype))
        assert y.shape == (M, N) and y.dtype == dtype
        assert relmax(y, ref) <= TOL[dtype], (relmax(y, ref), rms_rel(y, ref))
        assert rms_rel(y, ref) <= TOL[dtype]


def test_forward_shapes_strides_and_errors():
    w, low = synth_wsim(96, 160, -1, torch.float16, 3)
    p = pb.PackedLinear.from_dense(t(w, torch.float16), None, t(low))
    x = t(make_x(9, (2, 3, 160)), torch.float16)
    y = p.forward(x)
    assert y.shape == (2, 3, 96)
    ref = orc.linear(rounded(make_x(9, (2, 3, 160)), torch.float16), w)
    assert relmax(y, ref) <= 1e-3
    big = t(make_x(10, (6, 320)), torch.float16)
    xs = big[:, :160]                                             # row stride 320 (ldx > K)
    assert relmax(p.forward(xs), orc.linear(xs.float().cpu().numpy(), w)) <= 1e-3
    xt = t(make_x(11, (160, 6)), torch.float16).t()              # non-unit inner stride -> made contiguous
    assert relmax(p.forward(xt), orc.linear(xt.float().cpu().numpy(), w)) <= 1e-3
    assert p.forward(torch.empty(0, 160, device=DEV, dtype=torch.float16)).shape == (0, 96)   # empty input
    with pytest.raises(RuntimeError, match="dtype"):
        p.forward(x.float())                                       # mixed dtypes raise, as in the reference
    with pytest.raises(RuntimeError, match="in_features"):
        p.forward(x[..., :100])
    with pytest.raises(RuntimeError, match="CUDA"):
        p.forward(x.cpu())
    lib = _lib.load()
    assert lib.pbl_linear_forward(p.handle, None, 160, None, 96, 4, None) == -1
    assert lib.pbl_linear_forward(p.handle, C.c_void_p(x.data_ptr()), 100, C.c_void_p(y.data_ptr()), 96, 4, None) == -3


# ---- two-phase prefill path (M above PBL_DECODE_MAX_M): expansion + tcgen05 GEMM ---------------------------------
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("N,K,gs,M,bias", [(256, 64, -1, 16, False), (256, 256, -1, 128, True), (256, 256, -1, 129, False),
                                           (264, 520, -1, 300, True), (768, 768, -1, 1000, True), (512, 1024, 256, 37, False),
                                           (128, 4096, -1, 256, False), (1024, 11008, -1, 64, False),
                                           (4096, 4096, -1, 512, False), (264, 520, -1, 700, True), (512, 1024, 256, 513, False),
                                           (1024, 4096, -1, 2048, False), (4096, 4096, -1, 2560, False), (1024, 11008, -1, 600, True)])
def test_prefill_gemm_matches_oracle(dtype, N, K, gs, M, bias):
    """expand-once + TMA/TMA tcgen05 GEMM (cta_group::2), incl. ragged M/N/K, groups, single-CTA-valid tiles."""
    w, low = synth_wsim(N, K, gs, dtype, seed=N + K + M)
    b = rounded(np.random.RandomState(2).standard_normal(N).astype(np.float32) * 0.1, dtype) if bias else None
    x = rounded(make_x(N * 3 + M, (M, K)), dtype)
    p = pb.PackedLinear.from_dense(t(w, dtype), None if b is None else t(b, dtype), t(low), gs)
    assert p.select_kernel(M) == (4 if M <= 64 else 1)               # decode kernel in token passes / prefill GEMM
    y = forced(p, t(x, dtype), 1)
    if M * N * K <= 4e8:
        ref = orc.linear(x, w, b)                                   # CPU oracle (double accumulate)
    else:                                                           # large: fp64 on device over the same w_sim
        ref = (t(x).double() @ t(w).double().t() + (0 if b is None else t(b).double())).cpu().numpy()
    assert relmax(y, ref) <= TOL[dtype], (relmax(y, ref), rms_rel(y, ref))
    assert rms_rel(y, ref) <= TOL[dtype]
    assert torch.equal(y, forced(p, t(x, dtype), 1))                # deterministic
    if M <= 300:                                                    # both kernels agree on the same packed layer
        y4 = forced(p, t(x, dtype), 4)
        assert relmax(y4, ref) <= TOL[dtype]


def test_prefill_one_hot_activations_reproduce_w_sim_exactly():
    """x = I  =>  y = w_sim^T bit-for-bit: the expanded scratch IS the reference's tensor."""
    w, low = synth_wsim(512, 256, -1, torch.float16, 77)
    p = pb.PackedLinear.from_dense(t(w, torch.float16), None, t(low))
    y = p.forward(torch.eye(256, device=DEV, dtype=torch.float16))
    assert p.select_kernel(256) == 1
    assert torch.equal(y, t(w, torch.float16).t())
    # the decode kernel sums level + correction in fp32: within an ulp or two of w_sim, not bit-identical
    y4 = forced(p, torch.eye(256, device=DEV, dtype=torch.float16), 4)
    assert relmax(y4, w.T) <= 1e-3


def test_prefill_back_to_back_calls_share_no_scratch():
    """Consecutive prefill calls on one stream: the expansion of call i+1 is launched early (programmatic dependent
    launch) into the scratch buffer call i's GEMM is NOT reading.  Layers of different sizes, no synchronisation in
    between, every output bit-identical to the same call made alone."""
    shapes = [(512, 256), (256, 1024), (768, 512), (512, 256)]
    layers, xs, alone = [], [], []
    for i, (N, K) in enumerate(shapes):
        w, low = synth_wsim(N, K, -1, torch.float16, 300 + i)
        layers.append(pb.PackedLinear.from_dense(t(w, torch.float16), None, t(low)))
        xs.append(t(rounded(make_x(310 + i, (384, K)), torch.float16), torch.float16))
    for p, x in zip(layers, xs):
        assert p.select_kernel(x.shape[0]) == 1
        alone.append(p.forward(x).clone())
        torch.cuda.synchronize()
    for rep in range(6):
        outs = [p.forward(x) for p, x in zip(layers, xs)] + [p.forward(x) for p, x in zip(reversed(layers), reversed(xs))]
        torch.cuda.synchronize()
        for y, ref in zip(outs, alone + alone[::-1]):
            assert torch.equal(y, ref), rep


def test_prefill_inside_cuda_graph_uses_the_pool_path():
    """Under stream capture the scratch comes from the stream-ordered pool (allocation nodes): replay reproduces the eager result."""
    w, low = synth_wsim(512, 256, -1, torch.float16, 321)
    p = pb.PackedLinear.from_dense(t(w, torch.float16), None, t(low))
    x = t(rounded(make_x(322, (256, 256)), torch.float16), torch.float16)
    ref = p.forward(x).clone()
    out = torch.empty_like(ref)
    side, graph = torch.cuda.Stream(), torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        p.forward(x, out=out)
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=side):
            p.forward(x, out=out)
    out.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref)


# ---- golden fixtures from the executed reference --------------------------------------------------
def test_golden_cfg1_xnor_768_drop_in_ctor():
    """BASELINE config 1: OPT-125m-shaped 768x768 XnorBinaryLinear, fp32 as constructed."""
    g = load("cfg1_xnor_768")
    seed = int(g["seed"])
    W, b, x = make_weight(seed, 768, 768), make_x(seed + 1, (768,)) * 0.1, make_x(seed + 2, (4, 768))
    m = pb.XnorBinaryLinear(torch.from_numpy(W), torch.from_numpy(b)).to(DEV)
    y = m(t(x))
    assert relmax(y, g["y"]) <= 2e-5
    ws = m.dense_weight()
    bits = np.packbits((ws > 0).cpu().numpy(), axis=-1, bitorder="little")
    assert np.array_equal(bits, g["sign_bits"])                    # the reference's own sign bits
    yb = pb.BinaryLinear(torch.from_numpy(W), torch.from_numpy(b)).to(DEV)(t(x))
    assert relmax(yb, g["y_binary"]) <= 2e-5
    for M in (1, 4, 2048):                                          # SURVEY 7 minimum slice, fp32 and fp16
        xm = make_x(77 + M, (M, 768))
        ref = orc.forward_xnor(xm, W, b)
        assert relmax(m(t(xm)), ref) <= 2e-5
    mh = pb.XnorBinaryLinear(torch.from_numpy(W), torch.from_numpy(b)).to(DEV)
    ph = pb.PackedLinear.from_dense(mh.quant_weight().half(), mh.bias.data)
    xm = rounded(make_x(5, (4, 768)), torch.float16)
    ref = orc.linear(xm, mh.quant_weight().half().float().cpu().numpy(), b)
    assert relmax(ph.forward(t(xm, torch.float16)), ref) <= 1e-3


def test_golden_quantizer_small_all_classes():
    g = load("quantizer_small")
    W, b, x = torch.from_numpy(g["W"]), torch.from_numpy(g["b"]), t(g["x"])
    for name in ["BinaryLinear", "XnorBinaryLinear", "IrBinaryLinear", "FdaBinaryLinear"]:
        m = getattr(pb, name)(W, b).to(DEV)
        y = m(x)
        assert y.shape == (2, 3, 96)
        assert relmax(y, g["y_" + name]) <= 2e-5, name
    assert relmax(pb.XnorBinaryLinear(W, None).to(DEV)(x), g["y_Xnor_nobias"]) <= 2e-5
    m = pb.XnorBinaryLinear(W, b).to(DEV)
    assert torch.equal(m.dense_weight().cpu(), torch.from_numpy(g["wsim_Xnor"])) or \
        relmax(m.dense_weight(), g["wsim_Xnor"]) < 5e-7            # GPU reduction order may move alpha by an ulp
    lin = m.to_regular_linear()
    assert relmax(lin(x), g["y_XnorBinaryLinear"]) <= 2e-5


@pytest.mark.parametrize("tag", ["outlier_f32_small", "outlier_f16_small", "outlier_f32_heavy", "outlier_f16_heavy"])
def test_golden_outlier_drop_in_ctor(tag):
    g = load(tag)
    half = g["W"].dtype == np.float16
    dtype = torch.float16 if half else torch.float32
    m = pb.BinaryXnorExceptOutliersLinear(torch.from_numpy(g["W"]).clone(), torch.from_numpy(g["b"]), float(g["frac"]))
    m = m.to(DEV).eval()
    y = m(t(g["x"]))
    assert torch.equal(m.outlier_mask.cpu(), torch.from_numpy(g["mask"]))
    if half:   # 8-bit codes incl. the uint8 wrap: bit-exact once rounded to fp16
        assert torch.equal(m.weight.data.float().cpu(), torch.from_numpy(g["w8"]))
    else:      # fp32 dequant q*(range/255)+zp is FMA-contracted by torch's CUDA kernel: <= 1 ulp off the CPU fixture
        assert relmax(m.weight.data, g["w8"]) <= 2e-7
    ws = m.dense_weight().float().cpu().numpy()
    assert relmax(torch.from_numpy(ws[g["mask"]]), g["wsim"][g["mask"]]) <= (0 if half else 2e-7)
    assert relmax(torch.from_numpy(ws), g["wsim"]) <= (1e-3 if half else 5e-7)   # alpha: reduction order
    assert relmax(y, g["y"]) <= (3e-3 if half else 2e-5)            # golden y is the reference's CPU fp16 GEMM
    ref = orc.linear(g["x"].astype(np.float32), ws, g["b"].astype(np.float32))
    assert relmax(y, ref) <= TOL[dtype]
    assert abs(m.outlier_nbits - float(g["nbits"])) < 1e-12
    p = m.packed()
    assert p.salient_count() >= int(g["mask"].sum())
    reg = m.to_regular_linear()
    assert torch.equal(reg.weight.data, m.dense_weight())
    m.pack(keep_latent=False)
    assert m.weight.numel() == 0 and relmax(m(t(g["x"])), ref) <= TOL[dtype]


def test_golden_outlier_768_known_answers():
    g = load("outlier_768_kat")
    W = make_weight(int(g["seed"]), 768, 768)
    m = pb.BinaryXnorExceptOutliersLinear(torch.from_numpy(W).clone(), None, 0.1).to(DEV).eval()
    y = m(t(make_x(33, (4, 768))))
    assert int(m.outlier_mask.sum()) == 58982
    assert np.array_equal(np.packbits(m.outlier_mask.cpu().numpy(), axis=-1, bitorder="little"), g["mask_bits"])
    assert abs(m.outlier_nbits - 1.6104193793402777) < 1e-12
    ws = m.dense_weight()
    assert int((ws < 0).sum()) == 0 and int((ws == 0).sum()) == int(g["n_zero"])
    assert relmax(y, g["y"]) <= 2e-5
    p = m.packed()
    aff = p.affine.view(-1, 2)[:768].cpu().numpy()
    assert np.all(aff[:, 0] == 0.0) and np.allclose(aff[:, 1], float(g["binary_scale"]), rtol=1e-6)  # levels {0, alpha}
    assert p.nnz == 58982 and not p.stream_layout                  # fp32 module: planes layout; zeros are a LEVEL here


def test_golden_hessian_mask(tmp_path, monkeypatch):
    g = load("hessian_mask")
    monkeypatch.chdir(tmp_path)
    os.makedirs("gptq_pb/outputs/mask")
    torch.save(torch.from_numpy(g["low_mask"]), "gptq_pb/outputs/mask/mask_0.9_synthetic_model.layers.0.q_proj.pkl")
    m = pb.BinaryXnorExceptOutliersLinearHessian(torch.from_numpy(g["W"]).clone(), None, 0.1).to(DEV)
    m.global_name = "synthetic/model.layers.0.q_proj"
    m.eval()
    m.gen_outlier_mask()
    with pytest.raises(TypeError):
        m(t(g["x"]))
    m.train()
    assert relmax(m(t(g["x"])), g["y_train"]) <= 2e-5
    m.eval()
    assert relmax(m(t(g["x"])), g["y_eval"]) <= 2e-5
    m2 = pb.BinaryXnorExceptOutliersLinearHessian(torch.from_numpy(g["W"]).clone(), None, 0.1).to(DEV).eval()
    m2.global_name = "synthetic/missing"
    assert relmax(m2(t(g["x"])), g["y_fallback"]) <= 2e-5


@pytest.mark.parametrize("tag", ["gptqpb_rtn_g-1_mag", "gptqpb_rtn_g128_hes", "gptqpb_gptq_g-1_hes",
                                 "gptqpb_gptq_g128_mag"])
def test_golden_gptqpb_fakequant_layers(tag):
    g = load(tag)
    gs = int(g["groupsize"])
    lin = torch.nn.Linear(256, 48, bias=False, device=DEV, dtype=torch.float16)
    lin.weight.data = t(g["Wq"])
    m = pb.PackedFakeQuantLinear.from_linear(lin, t(g["low_mask"]), gs)
    y = m(t(g["x"]))
    assert relmax(y, g["y"]) <= 1e-3
    assert torch.equal(m.dense_weight(), lin.weight.data)           # exact weights, only the sum order differs
    p = m.packed()
    sal = int((~g["low_mask"]).sum())
    assert sal <= p.salient_count() <= sal + 0.01 * g["Wq"].size   # + rare third value mu (sign(0))
    m2 = pb.PackedFakeQuantLinear.from_linear(lin, None, gs)       # no mask file: levels from min/max
    assert relmax(m2(t(g["x"])), g["y"]) <= 1e-3


# ---- end-to-end host-buffer entry point ---------------------------------------------------------
def test_forward_host_buffers():
    w, low = synth_wsim(256, 512, -1, torch.float16, 8)
    p = pb.PackedLinear.from_dense(t(w, torch.float16), None, t(low))
    xh = torch.from_numpy(make_x(2, (6, 512))).half().pin_memory()
    yh = torch.empty(6, 256, dtype=torch.float16).pin_memory()
    ws = torch.empty(p.host_workspace_bytes(6), dtype=torch.uint8, device=DEV)
    p.forward_host(xh, yh, ws)
    assert relmax(yh, orc.linear(xh.float().numpy(), w)) <= 1e-3


# ---- BASELINE.json full layer sizes: size-independent properties --------------------------------
@pytest.mark.parametrize("N,K", [(4096, 4096), (11008, 4096), (4096, 11008)])
@pytest.mark.parametrize("M", [1, 8])
def test_full_size_properties_llama7b_shapes(N, K, M):
    gen = torch.Generator(device=DEV).manual_seed(N + K)
    al = torch.rand(N, 1, device=DEV, generator=gen) * 0.02 + 0.005
    mu = torch.randn(N, 1, device=DEV, generator=gen) * 0.003
    w = (mu + al * (torch.rand(N, K, device=DEV, generator=gen) < 0.5).float().mul(2).sub(1)).half()
    low = torch.rand(N, K, device=DEV, generator=gen) < 0.9
    w = torch.where(low, w, (torch.randn(N, K, device=DEV, generator=gen) * 0.03).half())
    p = pb.PackedLinear.from_dense(w, None, low)
    assert torch.equal(p.unpack(), w)                               # round trip at full size
    assert abs(p.salient_count() / (N * K) - 0.1) < 0.005
    assert p.bits_per_weight() < 4.5                                # every resident byte: sign plane + 32-bit entries + tables
    x1 = torch.randn(M, K, device=DEV, generator=gen).half()
    x2 = torch.randn(M, K, device=DEV, generator=gen).half()
    y1, y2 = p.forward(x1).float(), p.forward(x2).float()
    y12 = p.forward((x1.float() + x2.float()).half()).float()       # linearity (up to fp16 rounding of x1+x2, y)
    scale = y12.abs().max()
    assert ((y1 + y2 - y12).abs().max() / scale) < 4e-3
    ref = x1.double() @ w.double().t()                              # dense fp64 check of the same w_sim
    assert ((y1.double() - ref).abs().max() / ref.abs().max()) <= 1e-3
    assert torch.equal(p.forward(x1), p.forward(x1))                # deterministic (no atomics)


# ---- BASELINE configs[1]: OPT-1.3b shapes through the drop-in constructor (magnitude mask, f = 0.1) --------------
@pytest.mark.parametrize("N,K", [(2048, 2048), (8192, 2048), (2048, 8192)])
def test_full_size_outlier_module_opt13b_shapes(N, K):
    gen = torch.Generator(device=DEV).manual_seed(N * 3 + K)
    W = (torch.empty(N, K, device=DEV).normal_(0, 0.02, generator=gen)
         * (1 + 4 * (torch.rand(N, K, device=DEV, generator=gen) < 0.02))).half()      # heavy-tailed
    b = (torch.randn(N, device=DEV, generator=gen) * 0.1).half()
    m = pb.BinaryXnorExceptOutliersLinear(W.clone(), b, 0.1).eval()
    x = torch.randn(1, 2048, K, device=DEV, generator=gen).half()                       # seq_len 2048, batch 1
    y = m(x)
    assert m.packed().select_kernel(2048) == 1
    frac = float(m.outlier_mask.float().mean())
    assert abs(frac - 0.1) < 2e-3 and tuple(m.binary_scale.shape) == (1, 1)
    w_sim = m.dense_weight()
    assert torch.equal(w_sim, m.binarize_except_outliers())                             # packed form == reference tensor
    ref = x.double().view(-1, K) @ w_sim.double().t() + b.double()
    assert float((y.double().view(-1, N) - ref).abs().max() / ref.abs().max()) <= 1e-3
    y1 = m(x[:, :1])                                                                    # decode-sized call, decode kernel
    assert float((y1.double().view(-1, N) - ref[:1]).abs().max() / ref[:1].abs().max()) <= 1e-3
    assert 1.0 < m.outlier_nbits < 2.0 and m.packed().bits_per_weight() < 4.6


# ---- stress the rarely-taken paths of the expansion / packing -----------------------------------------------------
@pytest.mark.parametrize("sal_frac", [0.0, 0.35, 0.6, 1.0])
@pytest.mark.parametrize("M", [4, 40, 300, 700])
def test_dense_salient_blocks_all_kernels(sal_frac, M):
    """Salient density from none to every position: blocks with more than 256 entries take the in-loop entry loads of the
    decode kernel (patch and un-patch); every kernel must still be within tolerance and unpack exact."""
    N, K, dtype = 320, 704, torch.float16
    w, low = synth_wsim(N, K, -1, dtype, seed=int(sal_frac * 100) + M, sal_frac=sal_frac)
    x = rounded(make_x(M + 17, (M, K)), dtype)
    p = pb.PackedLinear.from_dense(t(w, dtype), None, t(low))
    assert torch.equal(p.unpack(), t(w, dtype))
    assert abs(p.salient_count() / (N * K) - sal_frac) < 0.02
    ref = orc.linear(x, w)
    y = p.forward(t(x, dtype))
    assert relmax(y, ref) <= 1e-3, (p.select_kernel(M), relmax(y, ref))
    assert relmax(forced(p, t(x, dtype), 1), ref) <= 1e-3
    if M <= 300:
        assert relmax(forced(p, t(x, dtype), 4), ref) <= 1e-3


def test_degenerate_rows_and_levels():
    """Rows with a single level (lo == hi), all-zero rows, a row that is entirely salient, groupsize 64."""
    N, K = 136, 256                  # a multiple of 8 (tcgen05 store granularity) but not of the 128-row tile
    rs = np.random.RandomState(9)
    w = np.where(rs.rand(N, K) < 0.5, 0.0625, -0.03125).astype(np.float32)
    w[0] = 0.25                      # one level only
    w[1] = 0.0                       # all zeros (also one level)
    w[2] = rs.standard_normal(K)     # nothing binarizable
    w[3, ::2] = 0.5                  # three values in a row -> third value goes to the salient list
    w = rounded(w, torch.float16)
    x = rounded(make_x(3, (9, K)), torch.float16)
    for gs in (-1, 64, 128):
        p = pb.PackedLinear.from_dense(t(w, torch.float16), None, None, gs)
        assert torch.equal(p.unpack(), t(w, torch.float16))
        ref = orc.linear(x, w)
        for kern in (1, 4):
            assert relmax(forced(p, t(x, torch.float16), kern), ref) <= 1e-3, (gs, kern)


# ---- the block-stream layout and the decode kernel (pbl_select_kernel == 4) ------------------------------------------
def bits16_to_f32(b, dtype):
    b = np.asarray(b, np.uint32)
    if dtype == torch.float16:
        return b.astype(np.uint16).view(np.float16).astype(np.float32)
    return (b << 16).view(np.float32)


def f32_to_bits16(v, dtype):
    tt = torch.from_numpy(np.ascontiguousarray(v, np.float32)).to(dtype)
    return tt.view(torch.int16).numpy().astype(np.uint32) & 0xFFFF


def stream_to_dense(p):
    """Rebuild w_sim from the block stream alone (fsign + eptr + ent + exc + affine) on the host, following
    csrc/pbllm_stream.cuh: level from the fragment-ordered sign bit, salient value = step(fl16(mid + half * tau), k)."""
    lib = _lib.load()
    out4 = (C.c_uint32 * 4)()
    pos = np.zeros((32, 64, 4), np.int64)
    for r in range(32):
        for c in range(64):
            lib.pbl_stream_position(r, c, out4)
            pos[r, c] = list(out4)
    slot_r, slot_c = np.zeros(2048, np.int64), np.zeros(2048, np.int64)
    slot_r[pos[..., 3].ravel()] = np.repeat(np.arange(32), 64)
    slot_c[pos[..., 3].ravel()] = np.tile(np.arange(64), 32)
    TC, rgs = int(p.sizes.tiles_c), int(p.sizes.n_pad) // 32
    G = int(p.sizes.groups)
    tpg = TC if G == 1 else p.groupsize // 64
    fsign = p.fsign.cpu().numpy().view(np.uint32).reshape(rgs * TC, 32, 2)
    eptr = p.eptr.cpu().numpy().view(np.uint32)
    ent = p.ent.cpu().numpy().view(np.uint32)
    aff = p.affine.cpu().numpy().reshape(int(p.sizes.n_pad), G, 2)
    out = np.zeros((rgs * 32, TC * 64), np.float32)

    def ord16(b):
        b = b.astype(np.int64)
        return np.where(b & 0x8000, -(b & 0x7FFF), b & 0x7FFF)

    for blk in range(rgs * TC):
        rg, kb = divmod(blk, TC)
        g = kb // tpg
        low = (fsign[blk][pos[..., 0], pos[..., 1]] >> pos[..., 2].astype(np.uint32)) & 1          # [32][64], 1 = LOW level
        lo, hi = aff[rg * 32:rg * 32 + 32, g, 0:1], aff[rg * 32:rg * 32 + 32, g, 1:2]
        tile = np.where(low == 1, lo, hi).astype(np.float32)
        e = ent[eptr[blk] * 4:eptr[blk + 1] * 4]
        if len(e):
            slot, k4, tau16 = (e >> 21).astype(np.int64), ((e >> 16) & 15).astype(np.int64), e & 0xFFFF
            k = np.where(k4 & 8, k4 - 16, k4)
            r, c = slot_r[slot], slot_c[slot]
            lo32, hi32 = lo[r, 0].astype(np.float32), hi[r, 0].astype(np.float32)
            mid, half = np.float32(0.5) * (lo32 + hi32), np.float32(0.5) * (hi32 - lo32)
            s = (half.astype(np.float64) * bits16_to_f32(tau16, p.dtype).astype(np.float64) + mid.astype(np.float64)).astype(np.float32)  # fma
            n = ord16(f32_to_bits16(s, p.dtype)) + k
            vb = np.where(n < 0, 0x8000 | (-n), n).astype(np.uint32)
            keep = k != -8
            tile[r[keep], c[keep]] = bits16_to_f32(vb[keep], p.dtype)
        out[rg * 32:rg * 32 + 32, kb * 64:kb * 64 + 64] = tile
    if p.n_exc:
        exc = p.exc.cpu().numpy().view(np.uint32).reshape(-1, 2)[: p.n_exc]
        rg, kb = np.divmod(exc[:, 0].astype(np.int64), TC)
        sl = (exc[:, 1] >> 16).astype(np.int64)
        out[rg * 32 + slot_r[sl], kb * 64 + slot_c[sl]] = bits16_to_f32(exc[:, 1] & 0xFFFF, p.dtype)
    return out[:p.N, :p.K]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("N,K,gs", [(128, 64, -1), (100, 70, -1), (300, 520, -1), (256, 512, 128), (33, 2048, -1)])
def test_stream_layout_reproduces_w_sim_bit_exactly(dtype, N, K, gs):
    w, low = synth_wsim(N, K, gs, dtype, seed=N + K)
    w[::7, ::5] = rounded(np.float32(w[::7, ::5]) * 1e-3, dtype)          # tiny salient values: ulp corrections and exceptions
    low[::7, ::5] = False
    p = pb.PackedLinear.from_dense(t(w, dtype), None, t(low), gs)
    assert p.stream_layout and p.packed_bytes() > 0
    eptr = p.eptr.cpu().numpy().view(np.uint32)
    nnz = p.salient_count()
    assert eptr[0] == 0 and np.all(np.diff(eptr.astype(np.int64)) >= 0)
    assert int(eptr[-1]) * 4 >= nnz and int(eptr[-1]) * 4 <= nnz + 3 * (len(eptr) - 1)          # padded to 4 per block
    assert nnz == int((~low).sum()) and p.n_exc > 0
    assert np.array_equal(stream_to_dense(p), w)                                               # integer / bit work: exact
    assert torch.equal(p.unpack(), t(w, dtype))
    assert torch.equal(p.low_mask_dense().cpu(), torch.from_numpy(low))


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("N,K,gs,M,bias", [(128, 64, -1, 1, False), (96, 160, -1, 3, True), (100, 70, -1, 5, False),
                                           (300, 520, -1, 8, True), (256, 512, 128, 2, False), (768, 768, -1, 1, True),
                                           (33, 2048, -1, 9, False), (512, 1024, 256, 16, True), (4096, 4096, -1, 8, False),
                                           (1024, 11008, -1, 7, False), (264, 1030, -1, 13, True), (2048, 128, -1, 8, True),
                                           (5000, 64, -1, 4, False), (64, 8192, -1, 16, True), (11008, 4096, -1, 8, False)])
def test_decode_kernel_matches_oracle(dtype, N, K, gs, M, bias):
    """Decode kernel incl. ragged N/K (generic activation loads), groups, two token passes (M > 8), layers with fewer
    blocks than CTAs, row groups finished by one warp (tiny K), row groups split across many CTAs (large K)."""
    w, low = synth_wsim(N, K, gs, dtype, seed=N + K + M)
    b = rounded(np.random.RandomState(3).standard_normal(N).astype(np.float32) * 0.1, dtype) if bias else None
    x = rounded(make_x(N * 5 + M, (M, K)), dtype)
    p = pb.PackedLinear.from_dense(t(w, dtype), None if b is None else t(b, dtype), t(low), gs)
    assert p.select_kernel(M) == 4
    y = p.forward(t(x, dtype))
    ref = orc.linear(x, w, b) if M * N * K <= 4e8 else \
        (t(x).double() @ t(w).double().t() + (0 if b is None else t(b).double())).cpu().numpy()
    assert relmax(y, ref) <= TOL[dtype], (relmax(y, ref), rms_rel(y, ref))
    assert rms_rel(y, ref) <= TOL[dtype]
    for _ in range(3):
        assert torch.equal(y, p.forward(t(x, dtype)))                # deterministic: slots are summed in CTA order


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("N,K", [(32, 128), (96, 256), (320, 1152), (64, 4096), (1000, 384), (2048, 2048)])
@pytest.mark.parametrize("M,bias", [(1, False), (3, True), (8, False)])
def test_decode_pair_kernel_matches_oracle(dtype, N, K, M, bias):
    """The pair kernel (K % 128 == 0, one group per row, M <= 8, aligned activations): one pair per row group (K = 128), fewer
    pairs than warps, ragged N, row groups split across many CTAs (K = 4096), and the same call through the block kernel."""
    w, low = synth_wsim(N, K, -1, dtype, seed=N + K + M + 1)
    b = rounded(np.random.RandomState(5).standard_normal(N).astype(np.float32) * 0.1, dtype) if bias else None
    x = rounded(make_x(N * 3 + M, (M, K)), dtype)
    p = pb.PackedLinear.from_dense(t(w, dtype), None if b is None else t(b, dtype), t(low))
    xd = t(x, dtype)
    assert p.select_kernel(M) == 4 and p.decode_variant(xd) == 2
    y = p.forward(xd)
    ref = orc.linear(x, w, b)
    assert relmax(y, ref) <= TOL[dtype], relmax(y, ref)
    assert rms_rel(y, ref) <= TOL[dtype]
    for _ in range(3):
        assert torch.equal(y, p.forward(xd))                         # deterministic
    buf = torch.zeros((M, K + 8), dtype=dtype, device=DEV)            # 16-byte aligned rows only: the block kernel
    buf[:, 8:] = xd
    xv = buf[:, 8:]
    assert p.decode_variant(xv) == 1
    yb = p.forward(xv)
    ulp = 2.0 ** -9 if dtype == torch.float16 else 2.0 ** -6          # same arithmetic, another summation order
    assert float((y.float() - yb.float()).abs().max()) <= ulp * float(yb.float().abs().max())


@pytest.mark.parametrize("sal_frac", [0.0, 0.25, 0.6, 1.0])
def test_decode_pair_kernel_dense_salient(sal_frac):
    """No salient weights at all, and pairs with more entry units than a ring stage holds (the in-loop global loads)."""
    N, K, M, dtype = 320, 768, 5, torch.float16
    w, low = synth_wsim(N, K, -1, dtype, seed=int(sal_frac * 100) + 9, sal_frac=sal_frac)
    x = rounded(make_x(23, (M, K)), dtype)
    p = pb.PackedLinear.from_dense(t(w, dtype), None, t(low))
    xd = t(x, dtype)
    assert p.decode_variant(xd) == 2
    assert relmax(p.forward(xd), orc.linear(x, w)) <= 1e-3


def test_decode_pair_kernel_symmetric_levels():
    """lo == -hi in every row (XnorBinaryLinear without a mean shift): the sum-of-x MMAs are skipped (no PBL_LAYER_HAS_MID)."""
    N, K, M = 256, 512, 8
    rs = np.random.RandomState(11)
    alpha = (0.01 + 0.05 * rs.rand(N, 1)).astype(np.float32)
    w = rounded(alpha * np.where(rs.rand(N, K) < 0.5, -1.0, 1.0).astype(np.float32), torch.float16)
    w = np.where(w > 0, np.abs(w).max(1, keepdims=True), -np.abs(w).max(1, keepdims=True)).astype(np.float32)   # exactly two levels per row
    x = rounded(make_x(31, (M, K)), torch.float16)
    p = pb.PackedLinear.from_dense(t(w, torch.float16), None, None)
    xd = t(x, torch.float16)
    assert p.decode_variant(xd) == 2 and p.salient_count() == 0 and (p.flags & 1) == 0
    assert torch.equal(p.unpack(), t(w, torch.float16))
    assert relmax(p.forward(xd), orc.linear(x, w)) <= 1e-3


def test_decode_kernel_workspace_sharing_and_graph():
    """The library-allocated
    workspace path (plain pbl_linear_forward) agrees with the persistent one; one workspace serves layers of
    different shapes back to back; the launch is CUDA-graph capturable."""
    dtype = torch.float16
    layers = []
    for (N, K) in [(768, 768), (3072, 768), (768, 3072), (130, 200)]:
        w, low = synth_wsim(N, K, -1, dtype, seed=N + K)
        layers.append((pb.PackedLinear.from_dense(t(w, dtype), None, t(low)), w))
    xs = {K: t(rounded(make_x(K, (8, K)), dtype), dtype) for K in (768, 3072, 200)}
    outs = [p.forward(xs[p.K]) for p, _ in layers]
    for (p, w), y in zip(layers, outs):
        ref = orc.linear(xs[p.K].float().cpu().numpy(), w)
        assert relmax(y, ref) <= 1e-3
    # interleaved replays through the shared workspace
    for _ in range(3):
        for (p, _), y in zip(layers, outs):
            assert torch.equal(p.forward(xs[p.K]), y)
    # plain pbl_linear_forward: transient workspace from the stream-ordered pool
    lib = _lib.load()
    for (p, _), y in zip(layers, outs):
        x = xs[p.K]
        y2 = torch.empty_like(y)
        rc = lib.pbl_linear_forward(p.handle, x.data_ptr(), p.K, y2.data_ptr(), p.N, 8, torch.cuda.current_stream().cuda_stream)
        assert rc == 0, _lib.last_error()
        assert torch.equal(y2, y)
    # CUDA graph capture + replay
    side = torch.cuda.Stream()
    p, _ = layers[1]
    yg = torch.empty_like(outs[1])
    with torch.cuda.stream(side):
        p.forward(xs[p.K], out=yg)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(4):
                p.forward(xs[p.K], out=yg)
    yg.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(yg, outs[1])
    # a too-small caller workspace is refused
    small = torch.zeros(1024, dtype=torch.uint8, device=DEV)
    rc = lib.pbl_linear_forward_ws(p.handle, xs[p.K].data_ptr(), p.K, yg.data_ptr(), p.N, 8, small.data_ptr(), small.numel(),
                                   torch.cuda.current_stream().cuda_stream)
    assert rc == -3 and "workspace" in _lib.last_error()


def test_stream_entry_order_spreads_banks():
    """The packer deals a block's entries to the kernel's 8 patch stores by shared-memory bank: with <= 8
    entries per bank every store is conflict free; at 10 % density the average store must need well under the ~3
    wavefronts of the row-major order."""
    w, low = synth_wsim(256, 512, -1, torch.float16, seed=5)
    p = pb.PackedLinear.from_dense(t(w, torch.float16), None, t(low))
    eptr = p.eptr.cpu().numpy().view(np.uint32)
    ent = p.ent.cpu().numpy().view(np.uint32)
    waves, stores = 0, 0
    for blk in range(len(eptr) - 1):
        n4 = int(eptr[blk + 1] - eptr[blk])
        if n4 == 0 or n4 > 64:
            continue
        e = ent[eptr[blk] * 4:eptr[blk + 1] * 4].reshape(n4, 4)
        h1 = (n4 + 1) // 2
        for units in (e[:h1], e[h1:]):
            for j in range(4):
                slot = units[:, j] >> 21                         # 16-bit slot; two slots per 32-bit word
                words = np.unique(slot // 2)                     # lanes writing the same word do not conflict
                waves += np.bincount(words % 32).max()
                stores += 1
    assert stores > 0 and waves / stores < 2.4, waves / stores


def test_decode_kernel_activation_alignment_paths():
    """The three activation-load paths of the decode kernel: one 32 B load (32 B aligned rows), two 16 B loads
    (16 B aligned only) and bounds-checked elements (anything else), through strided views of one buffer."""
    N, K, M = 256, 512, 8
    w, low = synth_wsim(N, K, -1, torch.float16, seed=21)
    p = pb.PackedLinear.from_dense(t(w, torch.float16), None, t(low))
    buf = t(rounded(make_x(77, (M, K + 64)), torch.float16), torch.float16)
    for off in (0, 16, 8, 24, 1, 3):                      # element offsets: 0/16 -> 32 B aligned, 8/24 -> 16 B, 1/3 -> 2 B
        xv = buf[:, off:off + K]
        assert xv.data_ptr() % 2 == 0 and xv.stride(0) == K + 64
        y = p.forward(xv)
        assert p.select_kernel(M) == 4
        ref = orc.linear(xv.float().cpu().numpy(), w)
        assert relmax(y, ref) <= 1e-3, off
        # the aligned call takes the pair kernel, the others the block kernel: same arithmetic, another summation order
        yc = p.forward(xv.contiguous())
        assert float((y.float() - yc.float()).abs().max()) <= 2.0 ** -9 * float(yc.float().abs().max()), off
