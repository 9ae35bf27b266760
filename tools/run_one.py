"""Run one packed linear a few times (for ncu captures): python tools/run_one.py M N K [reps] [sal]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pbllm_b200 as pb  # noqa: E402
from tools.diag_gemm import mk  # noqa: E402

M, N, K = (int(a) for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
sal = float(sys.argv[5]) if len(sys.argv) > 5 else 0.1
w, low, _ = mk(N, K, torch.float16, sal=sal, seed=1)
p = pb.PackedLinear.from_dense(w, None, low)
x = torch.randn(M, K, device="cuda:0", dtype=torch.float16)
out = torch.empty(M, N, device="cuda:0", dtype=torch.float16)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
p.forward(x, out=out)
e0.record()
for _ in range(reps):
    p.forward(x, out=out)
e1.record()
torch.cuda.synchronize()
ms_eager = e0.elapsed_time(e1) / reps
# CUDA-graph replay: GPU time without the Python launch overhead
g = torch.cuda.CUDAGraph()
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    p.forward(x, out=out)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=st):
        for _ in range(reps):
            p.forward(x, out=out)
g.replay()
torch.cuda.synchronize()
e0.record()
g.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"eager {ms_eager:.4f} ms/call;", end=" ")
print(f"M={M} N={N} K={K} kernel={p.select_kernel(M)} {ms:.4f} ms {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s "
      f"packed {p.packed_bytes() / 1e6:.1f} MB -> {p.packed_bytes() / ms / 1e6:.1f} GB/s")
