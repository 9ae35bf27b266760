"""Minimal drop-in demo (needs a B200): a random-init HF LLaMA model, every nn.Linear swapped for a packed
partially-binarized layer the way the reference's qat/run_qat.py does, served from ~3.6 bit/weight.

    python examples/hf_dropin.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pbllm_b200 as pb  # noqa: E402
from transformers import LlamaConfig, LlamaForCausalLM  # noqa: E402

torch.manual_seed(0)
cfg = LlamaConfig(hidden_size=1024, intermediate_size=2816, num_hidden_layers=4, num_attention_heads=8,
                  num_key_value_heads=8, vocab_size=4096)
model = LlamaForCausalLM(cfg).to("cuda:0").half().eval()
dense_bytes = sum(m.weight.numel() * 2 for m in model.modules() if isinstance(m, torch.nn.Linear))

# the reference's surgery (qat/run_qat.py:45-66): every nn.Linear -> BinaryXnorExceptOutliersLinear(w, b, 0.1)
pb.replace_with_qlinear(model, "xnor_outlier", outlier_fraction=0.1, model_id="demo/")
n = pb.pack_model(model, keep_latent=False)          # pack now, free the latent weights
packed_bytes = sum(m.packed().packed_bytes() for m in model.modules() if isinstance(m, pb.BinaryInterface))
print(f"{n} linears packed: {dense_bytes / 1e6:.1f} MB fp16 -> {packed_bytes / 1e6:.1f} MB "
      f"({8 * packed_bytes / (dense_bytes / 2):.2f} bit/weight)")

ids = torch.randint(0, cfg.vocab_size, (2, 512), device="cuda:0")
with torch.no_grad():
    logits = model(ids).logits                        # prefill: tcgen05 CTA-pair kernel
    step = model(ids[:, -1:]).logits                  # one decode step: mma.sync skinny kernel
print("prefill logits", tuple(logits.shape), "decode logits", tuple(step.shape))
pb.save_packed(model, "/tmp/pbllm_demo_ckpt")
print("packed checkpoint:", sum(os.path.getsize(os.path.join("/tmp/pbllm_demo_ckpt", f)) for f in os.listdir("/tmp/pbllm_demo_ckpt")) / 1e6, "MB")
