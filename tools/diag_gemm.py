"""GPU diagnostic for the tcgen05 path: structured inputs that localise layout / descriptor /
pipeline bugs from one run.  Prints compact summaries; not a test."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pbllm_b200 as pb  # noqa: E402

DEV = "cuda:0"


def run(p, x, kernel):
    os.environ["PBL_FORCE_KERNEL"] = str(kernel)
    try:
        y = p.forward(x)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("PBL_FORCE_KERNEL", None)
    return y


def summarize(name, y, ref):
    d = (y.float() - ref.float()).abs()
    scale = ref.float().abs().max().clamp_min(1e-30)
    bad = d > (6e-3 if y.dtype == torch.bfloat16 else 2e-3) * scale
    nb = int(bad.sum())
    msg = f"[{name}] shape {tuple(y.shape)} relmax {float(d.max() / scale):.3e} bad {nb}/{bad.numel()}"
    if nb:
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        msg += f" | bad rows {rows.numel()} (first {rows[:6].tolist()}, last {rows[-3:].tolist()})"
        msg += f" bad cols {cols.numel()} (first {cols[:6].tolist()}, last {cols[-3:].tolist()})"
        i = bad.nonzero()[0]
        msg += f" | first bad y[{int(i[0])},{int(i[1])}]={float(y[i[0], i[1]]):.4f} ref={float(ref[i[0], i[1]]):.4f}"
        msg += f" | nan {int(torch.isnan(y.float()).sum())}"
    print(msg, flush=True)
    return nb == 0


def mk(N, K, dtype, sal=0.1, seed=0, gs=-1, bias=False):
    g = torch.Generator(device=DEV).manual_seed(seed)
    groups = 1 if gs <= 0 else (K + gs - 1) // gs
    gi = torch.arange(K, device=DEV) // (K if gs <= 0 else gs)
    mu = torch.randn(N, groups, device=DEV, generator=g) * 0.004
    al = torch.rand(N, groups, device=DEV, generator=g) * 0.02 + 0.005
    sgn = (torch.rand(N, K, device=DEV, generator=g) < 0.5).float() * 2 - 1
    w = (mu[:, gi] + al[:, gi] * sgn).to(dtype)
    low = torch.rand(N, K, device=DEV, generator=g) >= sal
    w = torch.where(low, w, (torch.randn(N, K, device=DEV, generator=g) * 0.03).to(dtype))
    b = (torch.randn(N, device=DEV, generator=g) * 0.1).to(dtype) if bias else None
    return w, low, b


def main():
    torch.manual_seed(0)
    ok = True
    dt = torch.float16
    # 1. one-hot activations: y[m, n] = w[n, m] exactly -> pinpoints B-tile (n,k) placement
    for N, K in [(256, 64), (256, 128), (128, 64), (512, 256)]:
        w, low, _ = mk(N, K, dt, sal=0.1, seed=N + K)
        p = pb.PackedLinear.from_dense(w, None, low)
        x = torch.eye(K, device=DEV, dtype=dt)
        y = run(p, x, 1)
        ok &= summarize(f"onehot N={N} K={K}", y, w.t())
    # 2. no salient entries, constant rows: isolates MMA/TMA/epilogue from the patch path
    w = torch.full((256, 64), 0.5, device=DEV, dtype=dt)
    p = pb.PackedLinear.from_dense(w)
    x = torch.randn(256, 64, device=DEV, dtype=dt)
    ok &= summarize("const w, M=256", run(p, x, 1), x.float() @ w.float().t())
    # 3. general vs fp64 dense, sweep of shapes
    for (M, N, K, gs, bias) in [(16, 256, 64, -1, False), (128, 256, 256, -1, False), (129, 256, 256, -1, True),
                                (256, 512, 4096, -1, False), (300, 264, 520, -1, True), (1000, 768, 768, -1, True),
                                (37, 512, 1024, 256, False), (2048, 4096, 4096, -1, False), (512, 11008, 4096, -1, False),
                                (4096, 4096, 11008, -1, False)]:
        w, low, b = mk(N, K, dt, seed=M + N + K, gs=gs, bias=bias)
        p = pb.PackedLinear.from_dense(w, b, low, gs)
        x = torch.randn(M, K, device=DEV, dtype=dt)
        ref = x.double() @ w.double().t()
        if b is not None:
            ref = ref + b.double()
        ok &= summarize(f"gemm M={M} N={N} K={K} gs={gs} bias={bias}", run(p, x, 1), ref)
        if M <= 300:
            ok &= summarize(f"  gemv same", run(p, x, 0), ref)
    # 4. bf16
    w, low, b = mk(512, 512, torch.bfloat16, seed=5, bias=True)
    p = pb.PackedLinear.from_dense(w, b, low)
    x = torch.randn(384, 512, device=DEV, dtype=torch.bfloat16)
    ref = x.double() @ w.double().t() + b.double()
    ok &= summarize("bf16 gemm M=384", run(p, x, 1), ref)
    # 5. timing of the big shapes (CUDA events)
    for (M, N, K) in [(16384, 4096, 4096), (16384, 11008, 4096), (16384, 4096, 11008)]:
        w, low, _ = mk(N, K, dt, seed=1)
        p = pb.PackedLinear.from_dense(w, None, low)
        x = torch.randn(M, K, device=DEV, dtype=dt)
        out = torch.empty(M, N, device=DEV, dtype=dt)
        for _ in range(2):
            p.forward(x, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            p.forward(x, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"[time] M={M} N={N} K={K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
        ref = x[:256].double() @ w.double().t()
        ok &= summarize("   check rows 0..255", out[:256], ref)
    print("DIAG", "OK" if ok else "FAILED")


if __name__ == "__main__":
    main()
