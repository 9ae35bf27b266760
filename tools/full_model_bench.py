"""Full-model context run (needs one B200): a random-init HF LlamaForCausalLM with the Llama-7B config, first as
it is (dense fp16 cuBLAS linears = how the reference evaluates a GPTQ-PB checkpoint on a GPU), then with every
decoder linear replaced by a packed partially-binarized layer (GPTQ-PB format, low_frac 0.9, synthetic weights).
Measures prefill (batch x seq) and cached single-token decode through the unmodified HF forward.

    python tools/full_model_bench.py [--batch 8] [--seq 2048]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pbllm_b200 as pb  # noqa: E402
from bench import synth_layer_gpu  # noqa: E402


def timed(fn, reps, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def measure(model, ids, tag):
    B, S = ids.shape
    with torch.no_grad():
        ms_prefill = timed(lambda: model(ids, use_cache=False).logits, 3, 1)
        out = model(ids, use_cache=True)
        cache = out.past_key_values
        nxt = out.logits[:, -1:].argmax(-1)
        del out
        t = []
        for _ in range(12):   # cached decode: the cache grows by one token per call
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            o = model(nxt, past_key_values=cache, use_cache=True)
            nxt = o.logits[:, -1:].argmax(-1)
            torch.cuda.synchronize()
            t.append((time.perf_counter() - t0) * 1e3)
        ms_decode = sorted(t[2:])[len(t[2:]) // 2]
    mem = torch.cuda.max_memory_allocated() / 2**30
    return {"tag": tag, "prefill_ms": ms_prefill, "prefill_tokens_per_s": B * S / ms_prefill * 1e3, "decode_ms_per_step": ms_decode,
            "decode_tokens_per_s": B / ms_decode * 1e3, "peak_mem_GiB": mem}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--layers", type=int, default=32)
    args = ap.parse_args()
    from transformers import LlamaConfig, LlamaForCausalLM
    dev = torch.device("cuda:0")
    cfg = LlamaConfig(num_hidden_layers=args.layers)          # defaults == Llama-7B
    torch.set_default_dtype(torch.float16)
    with torch.device(dev):
        model = LlamaForCausalLM(cfg)
    torch.set_default_dtype(torch.float32)
    model.eval()
    ids = torch.randint(0, cfg.vocab_size, (args.batch, args.seq), device=dev)
    res = [measure(model, ids, "dense fp16 (reference GPU evaluation path: cuBLAS nn.Linear over fake-quant weights)")]
    torch.cuda.reset_peak_memory_stats()

    n, t0 = 0, time.time()
    for li, layer in enumerate(model.model.layers):
        for si, (parent, name) in enumerate([(layer.self_attn, "q_proj"), (layer.self_attn, "k_proj"), (layer.self_attn, "v_proj"),
                                             (layer.self_attn, "o_proj"), (layer.mlp, "gate_proj"), (layer.mlp, "up_proj"),
                                             (layer.mlp, "down_proj")]):
            lin = getattr(parent, name)
            N, K = lin.weight.shape
            w, low = synth_layer_gpu(N, K, 0.9, 1000 * li + si, dev)   # GPTQ-PB format weights (gptq_pb/gptq.py:149-155)
            q = pb.PackedFakeQuantLinear(w, None, low)
            q.pack(keep_latent=False)
            setattr(parent, name, q)
            del lin, w, low
            n += 1
    torch.cuda.empty_cache()
    pack_s = time.time() - t0
    res.append(measure(model, ids, f"packed PB linears (libpbllm): {n} layers, lm_head dense"))
    res[-1]["pack_seconds"] = pack_s
    print(json.dumps({"model": "LlamaForCausalLM(LlamaConfig()) random init", "batch": args.batch, "seq": args.seq, "results": res}))


if __name__ == "__main__":
    main()
