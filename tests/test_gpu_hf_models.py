"""Drop-in check on real HF model classes (random-init, tiny configs; no checkpoints/network):
existing OPT / LLaMA models load unchanged, the surgery swaps every nn.Linear for a packed
module (reference qat/run_qat.py:45-66), and the logits equal those of the same model with the
modules baked back into dense nn.Linear layers (reference to_regular_linear, qat/run_qat.py:69-80),
i.e. the reference's own simulated-quant forward."""
import copy

import pytest
import torch

import pbllm_b200 as pb

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def tiny_llama():
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=4, vocab_size=512, max_position_embeddings=128)
    return LlamaForCausalLM(cfg)


def tiny_opt():
    from transformers import OPTConfig, OPTForCausalLM
    cfg = OPTConfig(hidden_size=256, ffn_dim=512, num_hidden_layers=2, num_attention_heads=4, vocab_size=512,
                    max_position_embeddings=128, word_embed_proj_dim=256)
    return OPTForCausalLM(cfg)


@pytest.mark.parametrize("make,method", [(tiny_llama, "xnor_outlier"), (tiny_opt, "xnor_outlier"), (tiny_llama, "xnor")])
def test_hf_model_drop_in(make, method):
    torch.manual_seed(0)
    model = make().to(DEV).half().eval()
    pb.replace_with_qlinear(model, method, 0.1, model_id="tiny/")
    if method == "xnor":
        model.half()                                        # quantizer.py classes cast to fp32 in the ctor
    n_q = sum(isinstance(m, pb.BinaryInterface) for m in model.modules())
    assert n_q >= 2 * 6 + 1                                 # every nn.Linear incl. lm_head (tied or not)
    ids = torch.randint(0, 512, (2, 64), device=DEV)
    with torch.no_grad():
        logits = model(ids).logits                          # M = 128 tokens: tcgen05 path
        short = model(ids[:, :3]).logits                    # M = 6 tokens: mma.sync skinny path
        dense = copy.deepcopy(model)
        pb.to_regular_linear(dense)                         # reference's simulated-quant model
        assert not any(isinstance(m, pb.BinaryInterface) for m in dense.modules())
        ref = dense(ids).logits
        ref_short = dense(ids[:, :3]).logits
    scale = ref.float().abs().max()
    assert ((logits.float() - ref.float()).abs().max() / scale) < 5e-3     # fp16 model, 2 layers of accumulation
    assert ((short.float() - ref_short.float()).abs().max() / ref_short.float().abs().max()) < 5e-3
    assert pb.pack_model(model, keep_latent=False) == n_q  # free the latent weights, serve from packed form only
    with torch.no_grad():
        again = model(ids).logits
    assert torch.equal(again, logits)
