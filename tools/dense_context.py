"""Context numbers (NOT our path): the reference's own GPU evaluation is a dense fp16 nn.Linear over the
fake-quant weights (cuBLAS). Times the 224 Llama-7B decoder linears as dense fp16 for prefill (M=16384)
and decode (M=8, M=1) on this GPU, CUDA events / CUDA graph replay."""
import json
import sys

import torch
import torch.nn.functional as F

dev = torch.device("cuda:0")
HID, FFN, NL = 4096, 11008, 32
SHAPES = [(HID, HID, "h"), (HID, HID, "h"), (HID, HID, "h"), (HID, HID, "a"), (FFN, HID, "h"), (FFN, HID, "h"), (HID, FFN, "f")]
ws = [[(torch.randn(N, K, device=dev, dtype=torch.float16) * 0.02) for N, K, _ in SHAPES] for _ in range(NL)]
out = {}
for M in (16384, 8, 1):
    xin = {"h": torch.randn(M, HID, device=dev, dtype=torch.float16), "a": torch.randn(M, HID, device=dev, dtype=torch.float16),
           "f": torch.randn(M, FFN, device=dev, dtype=torch.float16)}

    def step():
        for row in ws:
            for (N, K, src), w in zip(SHAPES, row):
                F.linear(xin[src], w)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if M <= 8:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            step()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                step()
        fn, reps = g.replay, 30
    else:
        fn, reps = step, 3
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out[f"M={M}"] = {"ms_per_step": ms, "tokens_per_s": M / ms * 1e3, "weight_bytes": sum(w.numel() * 2 for r in ws for w in r)}
print(json.dumps({"dense_fp16_cublas_context": out}))
