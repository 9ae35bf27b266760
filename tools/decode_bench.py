"""Decode-regime micro-benchmark: a few decoder layers' worth of Llama-7B-shaped PB linears (distinct weights per
layer, working set > L2), batch M tokens, replayed as one CUDA graph.  Run once per variant; the variant knobs are
environment variables read by libpbllm.so at first use:
    PBL_DK_CTAS=1..4   decode grid = SMs x this        PBL_DK_OCC=2|3   register budget (128 / 80 regs)
    PBL_PDL=0|1        programmatic dependent launch
Prints one JSON line."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pbllm_b200 as pb  # noqa: E402
from bench import CONFIGS, layer_shapes, synth_layer_gpu  # noqa: E402

SHAPES = layer_shapes(CONFIGS["llama7b"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--low-frac", type=float, default=0.9)
    ap.add_argument("--reps", type=int, default=50)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    layers = []
    for li in range(a.layers):
        for si, (name, N, K, src) in enumerate(SHAPES):
            w, low = synth_layer_gpu(N, K, a.low_frac, 1000 * li + si, dev)
            layers.append((pb.PackedLinear.from_dense(w, None, low), src))
            del w, low
    M = a.batch
    g = torch.Generator(device=dev).manual_seed(1)
    xin = {"h": torch.randn(M, 4096, device=dev, generator=g).half(), "a": torch.randn(M, 4096, device=dev, generator=g).half(),
           "f": torch.randn(M, 11008, device=dev, generator=g).half()}
    outs = [torch.empty(M, p.N, device=dev, dtype=torch.float16) for p, _ in layers]

    def step():
        for (p, src), o in zip(layers, outs):
            p.forward(xin[src], out=o)

    # parity of this variant against fp64 dense math, three shapes
    err = 0.0
    for i in (0, 4, 6):
        p, src = layers[i]
        ref = xin[src].double() @ p.unpack().double().t()
        err = max(err, float((p.forward(xin[src]).double() - ref).abs().max() / ref.abs().max()))
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        step()
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=side):
            step()
    for _ in range(5):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    nk = sum(p.N * p.K for p, _ in layers)
    nnz = sum(p.salient_count() for p, _ in layers)
    alg = nk / 8 + 4 * sum(p.N for p, _ in layers) + 2 * M * sum(p.K + p.N for p, _ in layers) + 2 * nnz + sum(p.N + 1 for p, _ in layers)
    print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("PBL_")}, "kernel": layers[0][0].select_kernel(M),
                      "layers": a.layers, "batch": M, "ms": ms, "ms_per_32_layers": ms * 32 / a.layers,
                      "us_per_launch": ms * 1e3 / len(layers), "alg_gbs": alg / ms / 1e6, "max_rel_err": err,
                      "packed_gbs": sum(p.packed_bytes() for p, _ in layers) / ms / 1e6}))


if __name__ == "__main__":
    main()
