"""BiRealLinear XNOR-popcount micro-benchmark: a few decoder layers of Llama-7B-shaped pure sign layers, batch M,
one CUDA graph.  PBL_BIREAL_SK=0 selects the row-group-per-CTA kernel, default the stream-K kernel."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pbllm_b200 as pb  # noqa: E402
from bench import CONFIGS, layer_shapes  # noqa: E402

SHAPES = layer_shapes(CONFIGS["llama7b"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--reps", type=int, default=50)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    layers = []
    for li in range(a.layers):
        for si, (name, N, K, src) in enumerate(SHAPES):
            g = torch.Generator(device=dev).manual_seed(7000 * li + si)
            w = torch.randn(N, K, device=dev, generator=g)
            layers.append((pb.PackedLinear.from_dense(w.abs().mean(1, keepdim=True) * torch.sign(w)), src))   # fp32: planes layout, as BiRealLinear packs
            del w
    M = a.batch
    xin = {"h": torch.randn(M, 4096, device=dev).half(), "a": torch.randn(M, 4096, device=dev).half(), "f": torch.randn(M, 11008, device=dev).half()}
    outs = [torch.empty(M, p.N, device=dev, dtype=torch.float32) for p, _ in layers]
    ws = torch.zeros(max(p.bireal_workspace_bytes(M) for p, _ in layers), dtype=torch.uint8, device=dev)

    def step():
        for (p, src), o in zip(layers, outs):
            p.bireal_forward(xin[src], out=o, workspace=ws)

    err = 0.0
    for i in (0, 4, 6):
        p, src = layers[i]
        ref = torch.sign(xin[src].double()) @ p.unpack().double().t()
        err = max(err, float((p.bireal_forward(xin[src]).double() - ref).abs().max() / ref.abs().max()))
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
    print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("PBL_")}, "layers": a.layers, "batch": M, "ms": ms,
                      "ms_per_32_layers": ms * 32 / a.layers, "us_per_linear": ms * 1e3 / len(layers), "max_rel_err": err,
                      "fixup_ws": int(pb._lib.load().pbl_bireal_fixup_workspace(layers[0][0].handle, M)),
                      "sign_planes": layers[0][0].sign_planes is not None}))


if __name__ == "__main__":
    main()
