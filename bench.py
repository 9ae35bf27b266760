#!/usr/bin/env python
"""bench.py -- tokens/sec through the partially-binarized linears of a Llama-7B-shaped model.

Workload (BASELINE.json configs[2]): huggyllama/llama-7b shapes, GPTQ-PB format (xnor low part,
low_frac=0.9, 8-bit salient part, Hessian-like column-skewed salient mask), batch 8 x seq_len
2048 = 16384 tokens per step, fp16. A "step" is one pass of those tokens through all 224 decoder
linears (32 x [q,k,v,o 4096x4096; gate,up 11008x4096; down 4096x11008]) -- the hot path of
SURVEY.md section 8, i.e. every F.linear(x, w_sim) the reference's forward executes for them.
Weights and activations are synthetic (no checkpoints / network); every layer has its own
weights, so each step streams 2.3 GB of packed weights plus the activations: far larger than
the 126 MB L2, no flush needed between iterations.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # our CUDA path
    python bench.py --impl reference ...                          # the reference algorithm on host cores

Prints ONE JSON line (rank 0). See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HID, FFN, NLAYERS = 4096, 11008, 32
SHAPES = [("q_proj", HID, HID, "h"), ("k_proj", HID, HID, "h"), ("v_proj", HID, HID, "h"), ("o_proj", HID, HID, "a"),
          ("gate_proj", FFN, HID, "h"), ("up_proj", FFN, HID, "h"), ("down_proj", HID, FFN, "f")]
METRIC = "tokens/sec Llama-7B PB low_frac=0.9 (decoder linears forward, batch 8 x seq 2048)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback (B200_PROFILING.md)")


# ---- synthetic GPTQ-PB-format weights (SURVEY.md 8d "Synthetic inputs") -----------------------------
def synth_layer_gpu(N, K, low_frac, seed, dev):
    """Dense fp16 fake-quant weight in the format gptq_pb/gptq.py:149-155,180-184 produces, built on
    the GPU: low positions mu_i +- alpha_i (fp16), salient positions on the row's 8-bit grid
    s_i*(q - z_i) (high_quant.py:6-8), salient mask column-skewed like a Hessian metric."""
    g = torch.Generator(device=dev).manual_seed(seed)
    w = torch.randn(N, K, device=dev, generator=g) * 0.02
    col = 1.0 + 3.0 * torch.rand(1, K, device=dev, generator=g) ** 4          # column skew of the saliency
    sal_score = w.abs() * col
    thr = torch.quantile(sal_score.flatten()[:: max(1, (N * K) // 2_000_000)].float(), low_frac)
    low = sal_score <= thr
    wl = w * low
    mu = wl.mean(-1, keepdim=True)
    al = (wl - mu).abs().mean(-1, keepdim=True)
    q_low = mu + al * torch.sign(w - mu)
    mn = torch.minimum(w.amin(-1, keepdim=True), torch.zeros((), device=dev))
    mx = torch.maximum(w.amax(-1, keepdim=True), torch.zeros((), device=dev))
    s = (mx - mn) / 255.0
    z = torch.round(-mn / s)
    q_high = s * (torch.clamp(torch.round(w / s) + z, 0, 255) - z)
    return torch.where(low, q_low, q_high).half(), low


def clocks_sampler(stop, out, dev_index):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    try:
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                              "-i", str(dev_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:  # noqa: BLE001
        return
    try:
        while not stop.is_set():
            line = p.stdout.readline()
            if not line:
                break
            f = [s.strip() for s in line.split(",")]
            if len(f) >= 6 and f[0].isdigit():
                out.append(f)
    finally:
        p.terminate()


def clocks_summary(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    sm = sorted(int(s[0]) for s in samples)
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(samples[0][1]), "reasons": reasons, "samples": len(sm)}


# ---- CPU baseline: the reference algorithm on host cores ---------------------------------------------
def cpu_baseline(seconds_budget=20.0, tokens=2048, steps=None, warmup=1):
    """The reference's CPU path for this config: the GPTQ-PB checkpoint is a plain nn.Linear holding
    fake-quant weights (gptq_pb/gptq.py:180-184), evaluated as dense F.linear in fp32 on all host
    cores. Sample: ONE decoder layer (its 7 linears) x `tokens` tokens; the
    tokens/sec figure divides by the 32 identical layers. Weights come from the oracle's GPTQ-PB
    RTN restatement (oracle.gptqpb_rtn) at reduced rows to keep setup short, tiled to full size."""
    import torch.nn.functional as F
    from oracle import oracle as orc
    ncores = os.cpu_count() or 1
    rs = np.random.RandomState(0)
    ws = []
    for _, N, K, _src in SHAPES:
        base = (rs.standard_normal((256, K)) * 0.02).astype(np.float32)
        low = rs.rand(256, K) < 0.9
        wq, _, _ = orc.gptqpb_rtn(base, low, -1, 8, True)
        ws.append(torch.from_numpy(np.tile(wq, (N // 256, 1))).contiguous())
    xs = {"h": torch.randn(tokens, HID), "a": torch.randn(tokens, HID), "f": torch.randn(tokens, FFN)}

    def layer_step():
        with torch.no_grad():
            for (_, N, K, src), w in zip(SHAPES, ws):
                F.linear(xs[src], w)

    # give the reference its best thread count on this host (oversubscribing a big NUMA box is slower)
    best_t, best_n = None, ncores
    for n in sorted({ncores, max(1, ncores // 2), max(1, ncores // 4), min(ncores, 32), min(ncores, 16)}, reverse=True):
        torch.set_num_threads(n)
        layer_step()
        t0 = time.perf_counter()
        layer_step()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best_t, best_n = dt, n
    torch.set_num_threads(best_n)
    threads = best_n

    for _ in range(warmup):
        layer_step()
    times = []
    t_end = time.perf_counter() + seconds_budget
    while (steps is None and time.perf_counter() < t_end) or (steps is not None and len(times) < steps):
        t0 = time.perf_counter()
        layer_step()
        times.append(time.perf_counter() - t0)
        if steps is None and len(times) >= 50:
            break
    t_layer = float(np.mean(times))
    return dict(value=tokens / (t_layer * NLAYERS), unit="tokens/s", cores=threads, kind="port",
                sample=f"1 of {NLAYERS} decoder layers (7 dense fp32 F.linear over GPTQ-PB fake-quant weights, "
                       f"torch CPU, best of several thread counts = {threads} threads on {ncores} logical cores) x {tokens} "
                       f"tokens, {len(times)} reps, {t_layer * 1e3:.1f} ms/layer-sample; tokens/s = {tokens}/(t_layer*{NLAYERS})",
                ms_per_step=t_layer * 1e3, steps=len(times))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(steps=args.steps, warmup=max(1, min(args.warmup, 2)))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": cb["steps"], "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "parallelism": "host cores only"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_name(args):
    return (f"llama-7b decoder linears ({args.layers}x7 PB linears, GPTQ-PB format low_frac={args.low_frac} high_bit=8, "
            f"synthetic Hessian-skewed mask), batch {args.batch} x seq_len {args.seq} = {args.batch * args.seq} tokens/step")


# ---- our arm --------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout: point fd 1 at stderr while the run is going (NCCL prints its version
    banner to stdout when NCCL_DEBUG is set, C libraries may print too) and keep the real stdout for the final line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layers", type=int, default=NLAYERS, help="decoder layers (32 = the named config)")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--low-frac", type=float, default=0.9, dest="low_frac")
    ap.add_argument("--parallel", default="replica", choices=["replica", "rowshard"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    import pbllm_b200 as pb
    from pbllm_b200.sharding import shard_rows, gather_rows

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (impl ours) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = pb._lib.load()
    pb._lib.check(lib.pbl_device_check(), "device check")
    M = args.batch * args.seq
    rowshard = args.parallel == "rowshard" and world > 1

    # -- build the packed model (one-time, untimed) ----------------------------------------------------------
    layers = []
    t0 = time.time()
    for li in range(args.layers):
        row = []
        for si, (name, N, K, src) in enumerate(SHAPES):
            w, low = synth_layer_gpu(N, K, args.low_frac, 1000 * li + si, dev)
            if rowshard:
                r0, r1, _ = shard_rows(N, world, rank)
                w, low = w[r0:r1], low[r0:r1]
            row.append(pb.PackedLinear.from_dense(w, None, low))
            del w, low
        layers.append(row)
    torch.cuda.synchronize()
    build_s = time.time() - t0
    packed_bytes = sum(p.packed_bytes() for row in layers for p in row)
    dindex_bytes = 0
    nnz = sum(p.salient_count() for row in layers for p in row)
    nk = sum(N * K for _, N, K, _ in SHAPES) * args.layers

    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    xin = {"h": torch.randn(M, HID, device=dev, generator=g).half(), "a": torch.randn(M, HID, device=dev, generator=g).half(),
           "f": torch.randn(M, FFN, device=dev, generator=g).half()}
    outs = [torch.empty(M, p.N, device=dev, dtype=torch.float16) for p in layers[0]]
    full_outs = [torch.empty(M, N, device=dev, dtype=torch.float16) for _, N, _, _ in SHAPES] if rowshard else None

    # -- parity in the same run: CUDA path vs fp64 dense math over the bit-exact unpacked w_sim -------
    parity = {}
    if rank == 0:
        for i in (0, 4, 6):                                   # the three distinct layer shapes
            p = layers[0][i]
            xs_ = xin[SHAPES[i][3]]
            w64 = p.unpack().double()
            for tag, rows in (("prefill", slice(0, 1024)), ("decode", slice(0, 8))):
                yk = p.forward(xs_[rows].contiguous()).double()
                ref = xs_[rows].double() @ w64.t()
                parity[f"{SHAPES[i][0]}:{tag}"] = float((yk - ref).abs().max() / ref.abs().max())
            del w64
        parity["max_rel_err"] = max(parity.values())
        parity["tolerance"] = 1e-3
        assert parity["max_rel_err"] <= 1e-3, parity

    def step(x_h=None):
        for row in layers:
            for i, p in enumerate(row):
                src = SHAPES[i][3]
                x = x_h if (x_h is not None and src == "h") else xin[src]
                p.forward(x, out=outs[i])
                if rowshard:
                    gather_rows(outs[i], SHAPES[i][1], out=full_outs[i])
        return (full_outs if rowshard else outs)[-1]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.pbl_launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.pbl_launch_count() - n0
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms / steps, launches

    # -- clocks sampled during the timed regions -----------------------------------------------------------------
    stop, samples = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, local), daemon=True)
    if rank == 0:
        th.start()

    ms_step, launches = timed(step, args.steps, args.warmup)
    tokens_per_step = M * (1 if rowshard else world)
    value = tokens_per_step / (ms_step * 1e-3)

    # -- per-launch durations of the dominant kernel (events around every launch, separate pass) -----------------
    per = []
    if rank == 0:
        evs = []
        for row in layers:
            for i, p in enumerate(row):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                p.forward(xin[SHAPES[i][3]], out=outs[i])
                b.record()
                evs.append((a, b, p))
        torch.cuda.synchronize()
        per = [(a.elapsed_time(b), p) for a, b, p in evs]
    kern_ms = sum(t for t, _ in per)
    flops = sum(2.0 * M * p.N * p.K for _, p in per)
    pk = peaks()
    kernel_id = layers[0][0].select_kernel(M)
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (profiles/
    # r01_gemm_two_phase_M16384_N4096_K4096.md): gemm_tt_kernel dram read 264.7 MB + write 109.3 MB for the
    # q/k/v/o-shaped launch (algorithmic: 134 MB x + 134 MB y + 7.6 MB packed weights; the 33.5 MB dense
    # scratch written by expand_dense_kernel stays in L2 -- its own dram write is 0.02 MB -- but is partly
    # re-fetched by the GEMM waves, as is x).
    ncu_traffic = {"launch": "M=16384 N=4096 K=4096 (gemm_tt_kernel)", "bytes": 264.709376e6 + 109.28e6,
                   "algorithmic_bytes": 2 * 16384 * 4096 * 2 + 7.6e6}
    roofline = None
    if per:
        ach = flops / (kern_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": ach, "peak": pk["tf_sus"], "unit": "TFLOP/s", "frac": ach / pk["tf_sus"],
                    "traffic": ncu_traffic["bytes"] if (kernel_id == 1 and M >= 2048) else None, "traffic_note": ncu_traffic, "kernel": {0: "pbl CUDA-core bit-plane kernel", 1: "pbl two-phase prefill: expand_dense_kernel + gemm_tt_kernel (tcgen05 cta_group::2, TMA both operands); times include both launches",
                               2: "pbl mma.sync bit-plane skinny kernel", 3: "pbl tcgen05 split-K cluster kernel",
                               4: "pbl decode kernel"}[kernel_id],
                    "launches": len(per), "avg_launch_ms": kern_ms / len(per), "peak_source": pk["src"] + ", sustained bf16",
                    "frac_of_burst_peak": ach / pk["tf_burst"],
                    "algorithmic_flops_per_launch": "2*M*N*K (M=tokens/step, N,K of the linear)"}

    # -- end-to-end: host buffers in, host buffers out, through the module API ---------------------------------------
    e2e = None
    if not args.no_e2e:
        xh = torch.randn(M, HID).half().pin_memory()
        yh = torch.empty(M, SHAPES[-1][1], dtype=torch.float16).pin_memory()
        xd = torch.empty(M, HID, device=dev, dtype=torch.float16)

        def step_e2e():
            xd.copy_(xh, non_blocking=True)
            y = step(xd)
            yh.copy_(y, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        ms_e2e, _ = timed(step_e2e, max(2, args.steps // 2), 1)
        e2e = {"value": tokens_per_step / (ms_e2e * 1e-3), "unit": "tokens/s", "h2d_bytes_per_step": xh.numel() * 2,
               "d2h_bytes_per_step": yh.numel() * 2, "ms_per_step": ms_e2e,
               "api": "PackedLinear.forward (pbl_linear_forward) per linear; inputs from pinned host memory each step"}

    # -- decode regime (batch x 1 token): the HBM-bound bit-plane kernel ------------------------------------------------
    decode = None
    if not args.no_decode:
        Md = args.batch
        xd_in = {k: v[:Md].contiguous() for k, v in xin.items()}
        douts = [torch.empty(Md, p.N, device=dev, dtype=torch.float16) for p in layers[0]]

        def dstep():
            for row in layers:
                for i, p in enumerate(row):
                    p.forward(xd_in[SHAPES[i][3]], out=douts[i])

        ms_d_eager, _ = timed(dstep, 20, 3)
        # the decode step is launch-bound from Python (224 kernels of a few microseconds): replay it as a CUDA graph
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            dstep()
            torch.cuda.synchronize()
            with torch.cuda.graph(graph, stream=side):
                dstep()
        ms_d, _ = timed(graph.replay, 50, 5)
        # batch-1 decode (the classic GEMV regime) for reference
        x1_in = {k: v[:1].contiguous() for k, v in xin.items()}
        d1outs = [torch.empty(1, p.N, device=dev, dtype=torch.float16) for p in layers[0]]

        def d1step():
            for row in layers:
                for i, p in enumerate(row):
                    p.forward(x1_in[SHAPES[i][3]], out=d1outs[i])

        graph1 = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            d1step()
            torch.cuda.synchronize()
            with torch.cuda.graph(graph1, stream=side):
                d1step()
        ms_d1, _ = timed(graph1.replay, 50, 5)
        # sibling fusion (what a model-level integration can do on top of the drop-in modules): q/k/v and gate/up
        # share their input, so their weights can be packed as ONE layer each (rows concatenated before packing):
        # 4 launches per decoder layer instead of 7.
        ms_d_fused = None
        if not rowshard and args.layers <= NLAYERS:
            flayers = []
            for li in range(args.layers):
                ws_, ms_ = [], []
                for si, (name, N, K, src) in enumerate(SHAPES):
                    w, low = synth_layer_gpu(N, K, args.low_frac, 1000 * li + si, dev)
                    ws_.append(w)
                    ms_.append(low)
                grp = [(0, 1, 2), (3,), (4, 5), (6,)]
                flayers.append([(pb.PackedLinear.from_dense(torch.cat([ws_[i] for i in gidx]), None, torch.cat([ms_[i] for i in gidx])),
                                 SHAPES[gidx[0]][3]) for gidx in grp])
                del ws_, ms_
            fouts = [torch.empty(Md, p.N, device=dev, dtype=torch.float16) for p, _ in flayers[0]]

            def fstep():
                for row in flayers:
                    for i, (p, src) in enumerate(row):
                        p.forward(xd_in[src], out=fouts[i])

            fgraph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                fstep()
                torch.cuda.synchronize()
                with torch.cuda.graph(fgraph, stream=side):
                    fstep()
            ms_d_fused, _ = timed(fgraph.replay, 50, 5)
            del flayers
        G = 1
        b_bin = nk / 8 + 4 * sum(p.N for row in layers for p in row) * G + 2 * Md * sum(p.K + p.N for row in layers for p in row)
        b_sal = 2 * nnz + sum(p.N + 1 for row in layers for p in row)
        ach = (b_bin + b_sal) / (ms_d * 1e-3) / 1e9
        decode = {"tokens_per_s": Md * (1 if rowshard else world) / (ms_d * 1e-3), "ms_per_step": ms_d,
                  "ms_per_step_eager_python_launch": ms_d_eager, "batch": Md, "ms_per_step_batch1": ms_d1,
                  "batch1_actual_bytes_gbs": packed_bytes / (ms_d1 * 1e-3) / 1e9,
                  "ms_per_step_fused_siblings": ms_d_fused, "launch": "CUDA graph replay of the 224 launches",
                  "kernel": {4: "pbl decode kernel (positioned salient entries, warp-granular stream-K, mma.sync)",
                             2: "pbl mma.sync bit-plane skinny kernel"}.get(layers[0][0].select_kernel(Md), "pbl CUDA-core bit-plane kernel"),
                  "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                               # dram__bytes_read+write per launch from the committed ncu --set full capture of the same
                               # launches (profiles/r01_decode_kernel_M8_llama7b_shapes.md): 9.24 MB (4096x4096), 24.53 MB
                               # (11008x4096), 24.35 MB (4096x11008) -> average over the 7 linears of a decoder layer
                               "traffic": ((4 * 9.2352e6 + 2 * 24.526848e6 + 24.348672e6) / 7.0) if (layers[0][0].select_kernel(Md) == 4 and Md == 8 and args.low_frac == 0.9) else None,
                               "algorithmic_bytes_per_launch": (b_bin + b_sal) / max(1, sum(len(r) for r in layers)),
                               "algorithmic_bytes_per_step": b_bin + b_sal,
                               "actual_packed_bytes": packed_bytes, "decode_index_bytes": dindex_bytes,
                               "actual_bytes_gbs": packed_bytes / (ms_d * 1e-3) / 1e9,
                               "peak_source": pk["src"]}}

    # -- the literal XNOR-popcount kernel (BiRealLinear format: alpha*sign(W), binarized activations) ---------------
    xnor = None
    if not args.no_decode and not rowshard:
        Md = args.batch
        blayers = []
        for li in range(args.layers):
            row = []
            for si, (name, N, K, src) in enumerate(SHAPES):
                gq = torch.Generator(device=dev).manual_seed(7000 * li + si)
                w = torch.randn(N, K, device=dev, generator=gq)
                row.append(pb.PackedLinear.from_dense((w.abs().mean(1, keepdim=True) * torch.sign(w)).half()))
                del w
            blayers.append(row)
        xb_in = {k: v[:Md].contiguous() for k, v in xin.items()}
        bouts = [torch.empty(Md, p.N, device=dev, dtype=torch.float32) for p in blayers[0]]
        bws = torch.zeros(max(p.bireal_workspace_bytes(Md) for p in blayers[0]), dtype=torch.uint8, device=dev)   # zeroed once

        def bstep():
            for row in blayers:
                for i, p in enumerate(row):
                    p.bireal_forward(xb_in[SHAPES[i][3]], out=bouts[i], workspace=bws)

        bgraph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            bstep()
            torch.cuda.synchronize()
            with torch.cuda.graph(bgraph, stream=side):
                bstep()
        ms_b, _ = timed(bgraph.replay, 50, 5)
        nb = sum(p.N for row in blayers for p in row)
        kb_ = sum(p.K for row in blayers for p in row)
        b_alg = nk / 8 + 8 * nb + Md * (2 * kb_ + 4 * nb)       # sign plane + {lo,hi} + fp16 x in + fp32 y out
        ach = b_alg / (ms_b * 1e-3) / 1e9
        xnor = {"what": "BiRealLinear forward (quant/quantizer.py:151-169) as XNOR-popcount over the packed sign plane, "
                        "Llama-7B shapes, 224 launches + 224 activation-binarize launches per step, CUDA graph replay",
                "tokens_per_s": Md * world / (ms_b * 1e-3), "ms_per_step": ms_b, "batch": Md,
                "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                             "traffic": None, "algorithmic_bytes_per_step": b_alg, "peak_source": pk["src"]}}
        del blayers

    stop.set()
    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline()
        cb = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if rowshard else "weak",
                "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": {"workload": workload_name(args), "tokens_per_step_per_gpu": M,
                           "parallelism": (f"rowshard{world}+allgather" if rowshard else f"dp{world} (replicas, no collective)"),
                           "l2": "per-step working set (2.3 GB packed weights + activations) >> 126 MB L2; no flush needed",
                           "packed_bytes": packed_bytes, "decode_index_bytes": dindex_bytes, "bits_per_weight": 8.0 * packed_bytes / nk * (world if rowshard else 1),
                           "salient_fraction": nnz / nk * (world if rowshard else 1), "model_build_s": build_s},
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks_summary(samples), "roofline": roofline,
                "cpu_baseline": cb, "decode": decode, "xnor_popcount": xnor, "parity": parity}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
