"""Per-warp timeline of the decode kernel (pbl_decode_set_trace): where the microseconds of a launch go.
One decoder layer's 7 linears, batch 8, replayed as a CUDA graph; prints per launch the spread of the phase stamps."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pbllm_b200 as pb  # noqa: E402
from bench import CONFIGS, layer_shapes, synth_layer_gpu  # noqa: E402

SHAPES = layer_shapes(CONFIGS["llama7b"])

STRIDE = 16 * 148 * 8 * 8


def main():
    dev = torch.device("cuda", 0)
    lib = pb._lib.load()
    nlayers = 2
    layers = []
    for li in range(nlayers):
        for si, (name, N, K, src) in enumerate(SHAPES):
            w, low = synth_layer_gpu(N, K, 0.9, 1000 * li + si, dev)
            layers.append((pb.PackedLinear.from_dense(w, None, low), src, name))
    M = 8
    xin = {"h": torch.randn(M, 4096, device=dev).half(), "a": torch.randn(M, 4096, device=dev).half(), "f": torch.randn(M, 11008, device=dev).half()}
    outs = [torch.empty(M, p.N, device=dev, dtype=torch.float16) for p, _, _ in layers]

    def step():
        for (p, src, _), o in zip(layers, outs):
            p.forward(xin[src], out=o)

    nl = len(layers)
    buf = torch.zeros(nl * STRIDE, dtype=torch.int64, device=dev)
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        step()
        torch.cuda.synchronize()
        lib.pbl_decode_set_trace(buf.data_ptr(), buf.numel() * 8)
        with torch.cuda.graph(graph, stream=side):
            step()
        lib.pbl_decode_set_trace(None, 0)
    for _ in range(5):
        graph.replay()
    torch.cuda.synchronize()
    t = buf.cpu().numpy().view(np.uint64).reshape(nl, -1, 8)
    rows = []
    t_first = None
    for i, (p, src, name) in enumerate(layers):
        a = t[i]
        a = a[a[:, 0] != 0]
        st = a[:, :7].astype(np.float64)
        k0 = st[:, 0].min()
        if t_first is None:
            t_first = k0
        nb = (a[:, 7] & np.uint64(0xffffffff)).astype(np.int64)
        r = {"layer": f"{i}:{name}", "warps": int(len(a)), "blocks_per_warp": [int(nb.min()), int(nb.max())],
             "kernel_start_us": (k0 - t_first) / 1e3, "kernel_span_us": (st[:, 6].max() - k0) / 1e3}
        names = ["start", "before_wait", "after_wait", "loop_end", "after_barrier", "after_reduce", "end"]
        for j, nme in enumerate(names):
            v = (st[:, j] - k0) / 1e3
            r[nme] = [round(float(np.percentile(v, q)), 2) for q in (0, 50, 100)]
        rows.append(r)
        print(json.dumps(r))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "decode_trace.json"), "w"), indent=1)
    # raw stamps of the second decoder layer (steady state) for offline analysis: [launch][cta*8+warp][8] uint64
    np.save(os.path.join(ROOT, "gpurun_out", "decode_trace_raw.npy"), t[7:14].copy())


if __name__ == "__main__":
    main()
